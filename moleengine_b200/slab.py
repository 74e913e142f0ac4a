"""One large world cut into x-slabs, one per GPU (north_star's second partition).

Every rank uploads the whole world (identical slots) and owns the bodies whose x lies in its interval. Bodies within
`halo` of a cut are mirrored on the neighbour as ghosts: the owner's copy is authoritative, the neighbour solves the
ghost too (so that pairs across the cut see both bodies move inside a substep) and has it overwritten by the owner's
state after every substep. The only data that crosses ranks are 64-byte body records between x-neighbours:

    per tick:  refresh (who is near the cut, by position)          1 exchange
               after every substep, the last included (owner -> ghost)   substeps exchanges

The exchange after the LAST substep is what keeps ownership well defined: the next tick's refresh decides by position, each
rank from its own copy, so the owner's and the ghost's copy of a body near a cut must be bitwise equal at that moment
(otherwise a body whose two copies ended on different sides of the cut would be sent by nobody, or by both).

The C library packs / unpacks device buffers (sfg_slab_pack / sfg_slab_unpack); this module owns the buffers (torch
tensors) and the transport: torch.distributed point-to-point ops (NCCL over NVLink on GPUs, gloo in the CPU tests) or
plain device copies when several slabs live in one process (tests). Results next to a cut are only statistically equal
to the single-GPU run: the Gauss-Seidel order differs there.
"""
import math

REC_WORDS = 16   # one record = 64 bytes = 16 float32 words; record 0 of a buffer is the header (word 0 = count)


def slab_bounds(rank, n_ranks, x_min, x_max):
    """Equal-width x-interval of `rank`; the outermost slabs extend to infinity so that no body is ever unowned."""
    if n_ranks <= 0 or not (0 <= rank < n_ranks):
        raise ValueError("bad rank / n_ranks")
    w = (x_max - x_min) / n_ranks
    lo = -math.inf if rank == 0 else x_min + rank * w
    hi = math.inf if rank == n_ranks - 1 else x_min + (rank + 1) * w
    return lo, hi


def neighbours(rank, n_ranks):
    """(left, right) ranks or None at the ends; side 0 = towards lower x."""
    return (rank - 1 if rank > 0 else None, rank + 1 if rank + 1 < n_ranks else None)


class DistTransport:
    """torch.distributed point-to-point exchange with both x-neighbours in one batch (no collective over all ranks)."""
    name = "torch.distributed send/recv between x-neighbours (NCCL P2P over NVLink on GPUs), stream-ordered: no host synchronisation inside a tick"

    def __init__(self, group=None):
        self.group = group

    def exchange(self, rank, n_ranks, send, recv):
        import torch
        import torch.distributed as dist
        ops = []
        for side, peer in enumerate(neighbours(rank, n_ranks)):
            if peer is None:
                continue
            ops.append(dist.P2POp(dist.isend, send[side], peer, self.group))
            ops.append(dist.P2POp(dist.irecv, recv[side], peer, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            if send[0].is_cuda:
                torch.cuda.current_stream().synchronize()


class LocalTransport:
    """All slabs in one process (tests): the 'wire' is a copy between the drivers' buffers."""
    name = "in-process device copies"

    def __init__(self):
        self.drivers = {}

    def register(self, driver):
        self.drivers[driver.rank] = driver

    def exchange_all(self):
        for d in self.drivers.values():
            for side, peer in enumerate(neighbours(d.rank, d.n_ranks)):
                if peer is not None:
                    d.recv[side].copy_(self.drivers[peer].send[1 - side])


class SlabDriver:
    """Runs the ticks of one slab. `world` needs: slab_configure, slab_begin, slab_pack(mode, side, ptr, cap) -> count,
    slab_unpack(ptr, cap), tick_begin(frame_dt, time_scale, ff), tick_substeps(n), tick_end() — GpuWorld has them."""

    def __init__(self, world, rank, n_ranks, x_min, x_max, halo, substeps, capacity=16384, device=None, transport=None):
        import torch
        self.world, self.rank, self.n_ranks, self.substeps, self.capacity = world, rank, n_ranks, substeps, capacity
        self.lo, self.hi = slab_bounds(rank, n_ranks, x_min, x_max)
        self.halo = halo
        self.transport = transport if transport is not None else DistTransport()
        self.transport_name = getattr(self.transport, "name", type(self.transport).__name__)
        f32max = 3.0e38
        world.slab_configure(max(self.lo, -f32max), min(self.hi, f32max), halo)
        words = (capacity + 1) * REC_WORDS
        self.send = [torch.zeros(words, dtype=torch.float32, device=device) for _ in range(2)]
        self.recv = [torch.zeros(words, dtype=torch.float32, device=device) for _ in range(2)]
        self.sent_records = 0      # bodies sent in the last tick (both sides, all exchanges)
        self.exchanges = 0
        # asynchronous exchanges need a CUDA world (sfg_slab_pack_async / sfg_stream_fence) and a stream-ordered transport (NCCL)
        self.use_async = (isinstance(self.transport, DistTransport) and hasattr(world, "slab_pack_async") and self.send[0].is_cuda)
        if self.use_async:
            self.counts = torch.zeros((substeps + 2, 2), dtype=torch.int32, device=device)
        if isinstance(self.transport, LocalTransport):
            self.transport.register(self)

    def pack(self, mode):
        n = 0
        for side, peer in enumerate(neighbours(self.rank, self.n_ranks)):
            if peer is not None:
                n += self.world.slab_pack(mode, side, self.send[side].data_ptr(), self.capacity)
        self.sent_records += n

    def unpack(self):
        for side, peer in enumerate(neighbours(self.rank, self.n_ranks)):
            if peer is not None:
                self.world.slab_unpack(self.recv[side].data_ptr(), self.capacity)

    def _exchange(self, mode):
        if self.use_async:
            return self._exchange_async(mode)
        self.pack(mode)
        self.transport.exchange(self.rank, self.n_ranks, self.send, self.recv)
        self.unpack()
        self.exchanges += 1

    def _exchange_async(self, mode):
        """The same exchange without the host ever waiting: packs queued on the world's stream, the transport's stream made to wait
        for them (sfg_stream_fence), NCCL send / receive, the world's stream made to wait for the received records, unpacks queued.
        Record counts stay on the device (one small copy per exchange into self.counts, read once after the tick)."""
        import torch
        import torch.distributed as dist
        w = self.world
        sides = [(side, peer) for side, peer in enumerate(neighbours(self.rank, self.n_ranks)) if peer is not None]
        for side, _ in sides:
            w.slab_pack_async(mode, side, self.send[side].data_ptr(), self.capacity)
        ts = torch.cuda.current_stream(self.send[0].device)
        w.stream_fence(ts.cuda_stream, 0)
        k = self.exchanges
        for side, _ in sides:
            self.counts[k, side].copy_(self.send[side][:1].view(torch.int32)[0], non_blocking=True)
        ops = []
        for side, peer in sides:
            ops.append(dist.P2POp(dist.isend, self.send[side], peer, self.transport.group))
            ops.append(dist.P2POp(dist.irecv, self.recv[side], peer, self.transport.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()                       # NCCL: the current stream waits for the transfer, the host does not
        w.stream_fence(ts.cuda_stream, 1)
        for side, _ in sides:
            w.slab_unpack_async(self.recv[side].data_ptr(), self.capacity)
        self.exchanges += 1

    def _collect_counts(self):
        """After sfg_tick_end (the device is idle): how many records went out in this tick, and did any exchange overflow its buffer?"""
        c = self.counts[: self.exchanges].cpu()
        if int(c.max()) > self.capacity:
            raise RuntimeError(f"slab halo buffer too small: {int(c.max())} records for a capacity of {self.capacity}")
        self.sent_records = int(c.sum())

    def tick(self, frame_dt=1.0 / 60.0, forcefield=None):
        """One frame: refresh the halo, then `substeps` substeps with an owner -> ghost exchange after each of them."""
        self.sent_records, self.exchanges = 0, 0
        w = self.world
        w.slab_begin()
        self._exchange(0)
        w.tick_begin(frame_dt, None, forcefield)
        for s in range(self.substeps):
            w.tick_substeps(1)
            self._exchange(1)
        w.tick_end()
        if self.use_async:
            self._collect_counts()


def tick_local_async(drivers, transport, frame_dt=1.0 / 60.0, forcefield=None):
    """tick_local with the asynchronous pack / fence / unpack calls of SlabDriver._exchange_async (all slabs in one process on one
    GPU, the 'wire' = copies on torch's current stream): exercises the stream fences without NCCL."""
    import torch
    ts = torch.cuda.current_stream()

    def exchange(mode):
        for d in drivers:
            for side, peer in enumerate(neighbours(d.rank, d.n_ranks)):
                if peer is not None:
                    d.world.slab_pack_async(mode, side, d.send[side].data_ptr(), d.capacity)
            d.world.stream_fence(ts.cuda_stream, 0)
        transport.exchange_all()
        for d in drivers:
            d.world.stream_fence(ts.cuda_stream, 1)
            for side, peer in enumerate(neighbours(d.rank, d.n_ranks)):
                if peer is not None:
                    d.world.slab_unpack_async(d.recv[side].data_ptr(), d.capacity)
    for d in drivers:
        d.world.slab_begin()
    exchange(0)
    for d in drivers:
        d.world.tick_begin(frame_dt, None, forcefield)
    for s in range(drivers[0].substeps):
        for d in drivers:
            d.world.tick_substeps(1)
        exchange(1)
    for d in drivers:
        d.world.tick_end()


def tick_local(drivers, transport, frame_dt=1.0 / 60.0, forcefield=None):
    """Lock-step tick of several slabs living in one process (tests): same call order as SlabDriver.tick on every rank."""
    def exchange(mode):
        for d in drivers:
            d.pack(mode)
        transport.exchange_all()
        for d in drivers:
            d.unpack()
    for d in drivers:
        d.sent_records = 0
        d.world.slab_begin()
    exchange(0)
    for d in drivers:
        d.world.tick_begin(frame_dt, None, forcefield)
    for s in range(drivers[0].substeps):
        for d in drivers:
            d.world.tick_substeps(1)
        exchange(1)
    for d in drivers:
        d.world.tick_end()


# ---------------------------------------------------------------- island sharding (second row)
class IslandShardDriver:
    """One world replicated on every rank; each rank solves the islands it owns (sfg_island_shard_configure) and, after the
    tick, all-gathers the bodies it solved so that every copy is whole again. Islands share no body, pair or constraint
    (physics.rs:155-172), so the state after the exchange is bit-identical to a single-GPU tick. The exchange is the one
    real collective of this partition: an all-gather of (capacity + 1) 64-byte records per rank."""

    def __init__(self, world, rank, n_ranks, capacity, device=None, group=None):
        import torch
        self.world, self.rank, self.n_ranks, self.capacity, self.group = world, rank, n_ranks, capacity, group
        world.island_shard_configure(rank, n_ranks)
        words = (capacity + 1) * REC_WORDS
        self.send = torch.zeros(words, dtype=torch.float32, device=device)
        self.recv = torch.zeros(words * n_ranks, dtype=torch.float32, device=device)
        self.sent_records = 0

    def pack(self):
        self.sent_records = self.world.slab_pack(2, 0, self.send.data_ptr(), self.capacity)

    def unpack(self):
        words = (self.capacity + 1) * REC_WORDS
        for r in range(self.n_ranks):
            if r != self.rank:
                self.world.slab_unpack(self.recv[r * words:(r + 1) * words].data_ptr(), self.capacity)

    def tick(self, frame_dt=1.0 / 60.0, forcefield=None):
        import torch
        import torch.distributed as dist
        self.world.tick(frame_dt, None, forcefield)
        self.pack()
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        if self.send.is_cuda:
            torch.cuda.current_stream().synchronize()
        self.unpack()


def tick_island_shards_local(drivers, frame_dt=1.0 / 60.0, forcefield=None):
    """All shards in one process (tests): the all-gather is a copy of every driver's send buffer into every recv buffer."""
    for d in drivers:
        d.world.tick(frame_dt, None, forcefield)
        d.pack()
    words = (drivers[0].capacity + 1) * REC_WORDS
    for d in drivers:
        for src in drivers:
            d.recv[src.rank * words:(src.rank + 1) * words].copy_(src.send)
    for d in drivers:
        d.unpack()


# ---------------------------------------------------------------- ray batches over slabs (physics.rs:1255-1311)
# Every rank knows the colliders of its own bodies, of the ghosts within `halo` of its cuts and every static collider, all at
# their true poses. A ray is therefore sent to the slabs its swept segment crosses; each of them reports its first hit, and the
# first hit of the world is the minimum over those reports: min t, ties to the lowest collider slot — the rule of the
# single-world cast (sfg_cast.cuh). A ghost and its owner report the same hit; the lower rank's copy is kept.
def route_rays(rays, lo, hi):
    """Boolean mask: rays whose swept x-interval [min(x0, x1) - radius, max(x0, x1) + radius] meets the slab [lo, hi)."""
    import numpy as np
    x0 = rays["sx"].astype(np.float64)
    x1 = x0 + rays["dx"].astype(np.float64) * rays["max_distance"].astype(np.float64)
    r = rays["radius"].astype(np.float64)
    return (np.minimum(x0, x1) - r < hi) & (np.maximum(x0, x1) + r >= lo)


def hit_keys(hits):
    """Sort key of a hit: (t as ordered bits) << 32 | collider slot; a miss is the largest key."""
    import numpy as np
    t = np.ascontiguousarray(hits["t"], np.float32).view(np.uint32).astype(np.uint64)       # non-negative floats order like their bits
    key = (t << np.uint64(32)) | hits["collider"].astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFF)
    return np.where(hits["collider"] >= 0, key, np.uint64(0xFFFFFFFFFFFFFFFF))


def merge_hits(per_rank_hits):
    """First hit of the world from every slab's report (numpy; the in-process form of SlabDriver.cast's reduction)."""
    import numpy as np
    keys = np.stack([hit_keys(h) for h in per_rank_hits])
    win = keys.argmin(0)                         # argmin returns the FIRST minimum: the lowest rank among equal reports
    out = per_rank_hits[0].copy()
    for r, h in enumerate(per_rank_hits):
        sel = win == r
        out[sel] = h[sel]
    return out


def cast_local(drivers, rays):
    """All slabs in one process (tests): route, cast on every slab's world, merge."""
    import numpy as np
    from . import abi
    reports = []
    for d in drivers:
        hits = np.zeros(len(rays), abi.hit_dtype)
        hits["collider"] = -1
        m = route_rays(rays, d.lo, d.hi)
        if m.any():
            hits[m] = d.world.cast_batch(np.ascontiguousarray(rays[m]))
        reports.append(hits)
    return merge_hits(reports), [int(route_rays(rays, d.lo, d.hi).sum()) for d in drivers]


def cast_distributed(driver, rays, group=None, device=None):
    """One rank of a slab-decomposed world: cast the rays routed to this slab, then reduce over the ranks (three small
    all-reduces: the winning key, the winning rank, the winner's payload). Every rank passes the same `rays` and gets all hits."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from . import abi
    hits = np.zeros(len(rays), abi.hit_dtype)
    hits["collider"] = -1
    m = route_rays(rays, driver.lo, driver.hi)
    if m.any():
        hits[m] = driver.world.cast_batch(np.ascontiguousarray(rays[m]))
    key = torch.from_numpy(hit_keys(hits).view(np.int64) ^ np.int64(-2 ** 63)).to(device)      # unsigned order -> signed order
    best = key.clone()
    dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
    who = torch.where(key == best, torch.full_like(key, driver.rank), torch.full_like(key, 1 << 30))
    dist.all_reduce(who, op=dist.ReduceOp.MIN, group=group)
    mine = (who == driver.rank).cpu().numpy()
    payload = np.zeros((len(rays), 6), np.float32)
    for k, f in enumerate(("t", "px", "py", "nx", "ny")):
        payload[mine, k] = hits[f][mine]
    payload[mine, 5] = hits["collider"][mine].astype(np.float32)        # exact below 2^24 slots; the key carries the exact slot
    p = torch.from_numpy(payload).to(device)
    dist.all_reduce(p, op=dist.ReduceOp.SUM, group=group)
    p = p.cpu().numpy()
    out = np.zeros(len(rays), abi.hit_dtype)
    bestk = (best.cpu().numpy() ^ np.int64(-2 ** 63)).view(np.uint64)
    out["collider"] = np.where(bestk == np.uint64(0xFFFFFFFFFFFFFFFF), -1, (bestk & np.uint64(0xFFFFFFFF)).astype(np.int64)).astype(np.int32)
    for k, f in enumerate(("t", "px", "py", "nx", "ny")):
        out[f] = p[:, k]
    return out

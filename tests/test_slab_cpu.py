"""Host side of the slab decomposition on CPU: interval partition, neighbour schedule, and the per-tick call order and
delivery of SlabDriver over 2 gloo ranks (the pack / unpack kernels themselves need a GPU: tests/test_gpu_slab.py)."""
import math
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from moleengine_b200 import slab


def test_slab_bounds_tile_the_line():
    for n in (1, 2, 3, 8):
        b = [slab.slab_bounds(r, n, -1000.0, 1000.0) for r in range(n)]
        assert b[0][0] == -math.inf and b[-1][1] == math.inf
        assert all(b[i][1] == b[i + 1][0] for i in range(n - 1))
        widths = [hi - lo for lo, hi in b[1:-1]]
        assert all(abs(w - 2000.0 / n) < 1e-9 for w in widths)
    assert slab.neighbours(0, 1) == (None, None)
    assert slab.neighbours(0, 3) == (None, 1) and slab.neighbours(1, 3) == (0, 2) and slab.neighbours(2, 3) == (1, None)


class FakeWorld:
    """Stands in for GpuWorld: records the call order and packs recognisable records into the caller's buffer."""

    def __init__(self, rank):
        self.rank, self.calls, self.got, self.bufs = rank, [], [], {}

    def slab_configure(self, lo, hi, halo):
        self.calls.append(("configure", lo, hi, halo))

    def slab_begin(self):
        self.calls.append("begin")

    def bind(self, driver):
        for side in range(2):
            self.bufs[driver.send[side].data_ptr()] = driver.send[side]
            self.bufs[driver.recv[side].data_ptr()] = driver.recv[side]

    def slab_pack(self, mode, side, ptr, cap):
        self.calls.append(("pack", mode, side))
        buf = self.bufs[ptr]
        n = 3 + self.rank + side                      # distinct counts per rank and side
        buf.zero_()
        buf[0] = float(n)
        for i in range(n):
            buf[(1 + i) * slab.REC_WORDS] = 1000.0 * self.rank + 100.0 * side + 10.0 * mode + i
        return n

    def slab_unpack(self, ptr, cap):
        buf = self.bufs[ptr]
        n = int(buf[0].item())
        self.got.append([float(buf[(1 + i) * slab.REC_WORDS].item()) for i in range(n)])
        self.calls.append("unpack")

    def tick_begin(self, frame_dt, time_scale, ff):
        self.calls.append("tick_begin")

    def tick_substeps(self, n):
        self.calls.append(("substeps", n))

    def tick_end(self):
        self.calls.append("tick_end")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    w = FakeWorld(rank)
    d = slab.SlabDriver(w, rank, world_size, -100.0, 100.0, halo=4.0, substeps=3, capacity=32, device="cpu")
    w.bind(d)
    d.tick()
    out.put((rank, w.calls, w.got, d.sent_records, d.exchanges, (d.lo, d.hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_exchange_and_call_order():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r = out.get(timeout=120)
        res[r[0]] = r
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    # rank 0 only has a right neighbour (side 1), rank 1 only a left one (side 0)
    _, calls0, got0, sent0, ex0, b0 = res[0]
    _, calls1, got1, sent1, ex1, b1 = res[1]
    assert b0 == (-math.inf, 0.0) and b1 == (0.0, math.inf)
    want_order = ["begin", ("pack", 0, 1), "unpack", "tick_begin", ("substeps", 1), ("pack", 1, 1), "unpack", ("substeps", 1), ("pack", 1, 1), "unpack",
                  ("substeps", 1), ("pack", 1, 1), "unpack", "tick_end"]
    assert calls0[1:] == want_order
    assert calls1[1:] == [c if not (isinstance(c, tuple) and c[0] == "pack") else ("pack", c[1], 0) for c in want_order]
    assert ex0 == ex1 == 4                      # refresh + one per substep (the last included: ownership is decided from equal copies)
    # what rank 0 received is what rank 1 packed for its side 0 (count 3 + 1 + 0 = 4), refresh then one sync per substep; and vice versa
    assert got0 == [[1000.0 + 10.0 * m + i for i in range(4)] for m in (0, 1, 1, 1)]
    assert got1 == [[100.0 + 10.0 * m + i for i in range(4)] for m in (0, 1, 1, 1)]
    assert sent0 == 4 * 4 and sent1 == 4 * 4


def _shard_worker(rank, world_size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)

    class W(FakeWorld):
        def island_shard_configure(self, r, n):
            self.calls.append(("shard", r, n))

        def tick(self, frame_dt, time_scale, ff):
            self.calls.append("tick")

    w = W(rank)
    d = slab.IslandShardDriver(w, rank, world_size, capacity=16, device="cpu")
    w.bufs[d.send.data_ptr()] = d.send
    words = (d.capacity + 1) * slab.REC_WORDS
    for r in range(world_size):
        w.bufs[d.recv[r * words:(r + 1) * words].data_ptr()] = d.recv[r * words:(r + 1) * words]
    d.tick()
    out.put((rank, w.calls, w.got, d.sent_records))
    dist.barrier()
    dist.destroy_process_group()


def test_island_shard_driver_all_gather_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r = out.get(timeout=120)
        res[r[0]] = r
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for rank in (0, 1):
        _, calls, got, sent = res[rank]
        other = 1 - rank
        assert calls == [("shard", rank, 2), "tick", ("pack", 2, 0), "unpack"]
        assert sent == 3 + rank                                  # FakeWorld packs 3 + rank + side records
        assert got == [[1000.0 * other + 10.0 * 2 + i for i in range(3 + other)]]   # exactly the other rank's buffer, once


def _cast_worker(rank, world_size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from moleengine_b200 import abi

    class CastWorld(FakeWorld):
        def cast_batch(self, rays):
            # rank 0 sees nearer hits for even rays, rank 1 for odd rays; ray 4 is an exact tie (a ghost and its owner); ray 6 misses everywhere
            h = np.zeros(len(rays), abi.hit_dtype)
            idx = np.round(rays["sy"]).astype(int)                     # the test stores the ray number in sy
            h["collider"] = 10 * idx + self.rank
            h["t"] = np.where(idx % 2 == self.rank, 1.0, 2.0) + idx
            tie = idx == 4
            h["collider"][tie], h["t"][tie] = 77, 3.25
            h["collider"][idx == 6] = -1
            h["px"] = 100.0 * self.rank + idx
            return h

    w = CastWorld(rank)
    d = slab.SlabDriver(w, rank, world_size, -100.0, 100.0, halo=4.0, substeps=1, capacity=8, device="cpu")
    rays = np.zeros(8, abi.ray_dtype)
    rays["sy"] = np.arange(8)
    rays["sx"] = [-50, -50, -50, -50, -1, -1, -1, 50]                  # rays 0..3 stay in slab 0, 4..6 cross the cut at 0, 7 stays in slab 1
    rays["dx"], rays["max_distance"] = 1.0, [10, 10, 10, 10, 5, 5, 5, 10]
    hits = slab.cast_distributed(d, rays, device="cpu")
    out.put((rank, hits["collider"].tolist(), hits["t"].tolist(), hits["px"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_cast_over_slabs_routes_and_min_reduces_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cast_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r = out.get(timeout=120)
        res[r[0]] = r
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert res[0][1:] == res[1][1:]                                     # every rank ends up with all hits
    coll, t, px = res[0][1], res[0][2], res[0][3]
    # rays 0..3: only slab 0 was asked (rank 0's answer, whatever rank 1 would have said); ray 7: only slab 1
    assert coll[:4] == [0, 10, 20, 30] and t[:4] == [1.0, 3.0, 3.0, 5.0] and px[:4] == [0.0, 1.0, 2.0, 3.0]
    assert coll[7] == 71 and t[7] == 8.0 and px[7] == 107.0
    # ray 4: both slabs report the same hit (ghost + owner) -> rank 0's copy; ray 5: rank 1 is nearer; ray 6: a miss everywhere
    assert coll[4] == 77 and t[4] == 3.25 and px[4] == 4.0
    assert coll[5] == 51 and t[5] == 6.0 and px[5] == 105.0
    assert coll[6] == -1

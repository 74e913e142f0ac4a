#!/usr/bin/env python
"""bench.py — body-substeps/s of the Starframe physics tick on B200 (BASELINE.json metric).

A "step" is one physics tick (frame_dt = 1/60, 4 XPBD substeps) of the C3 workload: 1 M mixed-collider bodies
in one large world. With --gpus N > 1 every rank steps its own independent 1 M-body world
(different seed, no data-path collective): weak scaling over independent worlds, the partition north_star
names first. `value` is device-resident throughput (state already in HBM); `e2e` goes through the C ABI with
pinned HOST buffers: poses in -> tick -> poses out every step in one sfg_tick_poses call, which is
what the reference's Game::physics_tick does around PhysicsWorld::tick (game.rs:297-303).

--impl reference times the CPU restatement of the reference tick (oracle/, MODE_REFERENCE, all host threads) on a
bounded sample of the same workload (the reference itself is Rust and cannot be built in this image).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SUBSTEPS = 4
FRAME_DT = 1.0 / 60.0


def workload(scale, seed):
    from moleengine_b200 import scenes
    nx, ny = {"full": (1250, 800), "tenth": (400, 250), "tiny": (100, 50)}[scale]
    return scenes.c3_mixed(nx, ny, seed=seed)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region: the poller runs from before the warm-up (it takes
    tens of milliseconds to start) and every row is stamped on arrival; only rows between mark_begin() and mark_end() count."""

    def __init__(self, device):
        self.rows, self.proc, self.device, self.t0, self.t1 = [], None, device, None, None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc:
            time.sleep(0.03)          # let the rows of the last milliseconds arrive
            self.proc.terminate()
        slack = 0.015                 # a row describes the ~10 ms before it was printed
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + slack]
        rows = inside if inside else [r for _, r in self.rows]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "samples_in_timed_region": len(inside)}


def algorithmic_bytes_per_tick(st, n_bodies, substeps):
    """Compulsory HBM bytes of the solver kernel for one tick, from the data layout in DESIGN.md §3:
    per substep, 384 B per body (integrate 136, contact position gather/scatter 80, derive 80, contact velocity 88),
    112 B per pair unit (H, G0, G1, L0, L1 once + H, S once) and 72 B per contact entry (manifold + latch written,
    manifold read back); constraints 96 B per unit, rope particles 64 B."""
    n_units = int(st["n_pairs"])
    n_entries = 2 * n_units   # upper bound is exact for body-body pairs; static pairs have one entry
    per_sub = 384 * n_bodies + 112 * n_units + 72 * n_entries + 96 * int(st["n_constraint_colours"] > 0) * 0 + 64 * int(st["n_rope_particles"])
    return per_sub * substeps


def profiled_traffic():
    """DRAM bytes of one k_solve_persistent launch from the committed ncu --set full capture of this workload (profiles/),
    or None. bench.py never runs under a profiler; this only quotes the capture next to the algorithmic byte count."""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_solver_full_1M.txt"))):
        tot, n = 0.0, 0
        for line in open(path):
            m = re.match(r"dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
            if m:
                tot += float(m.group(2)) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(3)]
                n += 1
        if n == 2:
            best = (tot, os.path.basename(path))
    return best


def run_reference(args, rank, world_size):
    """CPU arm: the oracle port of the reference tick (DFS islands, BVH broadphase, sequential Gauss-Seidel, island-group threads)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    from moleengine_b200 import abi
    scale = "tenth" if args.scale == "full" else args.scale          # the bounded sample: 1/10 of the 1 M-body workload per step
    sc = workload(scale, 0xC3)
    tun = abi.default_tuning(substeps=SUBSTEPS)
    cores = os.cpu_count() or 1
    w = O.OracleWorld(tuning=tun)
    w.load_scene(sc)
    nb = len(sc["bodies"])
    for _ in range(args.warmup):
        w.tick(FRAME_DT, mode=O.MODE_REFERENCE, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        w.tick(FRAME_DT, mode=O.MODE_REFERENCE, threads=cores)
    dt = time.perf_counter() - t0
    val = nb * SUBSTEPS * args.steps / dt
    sample = f"c3_mixed sample ({scale}: {nb} bodies), {args.steps} ticks of {SUBSTEPS} substeps, f64 oracle port of the reference tick"
    print(json.dumps({
        "impl": "reference", "metric": "body-substeps/sec", "value": val, "unit": "body-substeps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": "c3_mixed 1M bodies (reference arm: bounded 1/10-scale sample per step)", "substeps": SUBSTEPS},
        "cpu_baseline": {"value": val, "unit": "body-substeps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "body-substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", default="full", choices=["full", "tenth", "tiny"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-settled", action="store_true", help="skip the extra measurement of the same world after the bodies have come to rest (N = 1 only)")
    ap.add_argument("--no-slab", action="store_true", help="skip the slab-decomposed strong-scaling arm (N > 1 only)")
    ap.add_argument("--island-arm", action="store_true", help="also time island sharding of the single world (N > 1 only; bit-exact partition)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world_size)
        return

    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner, for one) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    from moleengine_b200 import GpuWorld, abi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: moleengine_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world_size == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sc = workload(args.scale, 0xC3 + rank)
    nb = len(sc["bodies"])
    tun = abi.default_tuning(substeps=SUBSTEPS)
    g = GpuWorld(tuning=tun, device=local_rank)
    g.load_scene(sc)
    ff = abi.gravity_field()

    # ---------------- device-resident arm: `value`
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        g.tick(FRAME_DT, None, ff)
    barrier()
    solver_ms, total_ms, broad_ms, graph_ms, launches = 0.0, 0.0, 0.0, 0.0, 0
    sampler.mark_begin()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g.tick(FRAME_DT, None, ff)      # synchronous: returns after the tick's stream has drained
        st = g.stats()
        solver_ms += float(st["ms_solver"]); total_ms += float(st["ms_total"]); broad_ms += float(st["ms_broadphase"]); graph_ms += float(st["ms_graph"])
        launches += int(st["kernel_launches"])
    barrier()
    wall = time.perf_counter() - t0
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    step_s = max_over_ranks(wall) / args.steps
    value = world_size * nb * SUBSTEPS / step_s
    st = g.stats()

    # ---------------- end-to-end arm: pinned host poses in -> tick -> poses out, through the C ABI.
    # This is Game::physics_tick (game.rs:297-303): sync_hecs_to_physics writes f32 poses into the world, tick,
    # sync_physics_to_hecs reads them back (hecs_sync.rs:121-227) -> sfg_set_poses / sfg_tick / sfg_get_poses.
    # The scene is reloaded and warmed up again so that both arms time the SAME ticks of the scene (bodies land on the
    # shelves as it runs: later ticks carry more contacts and cost more than earlier ones).
    host = torch.empty(nb * 16, dtype=torch.uint8).pin_memory()
    host_np = host.numpy().view(np.float32).reshape(nb, 4)
    lib, h = g.lib, g.h
    assert lib.sfg_world_clear(h) == 0
    g.load_scene(sc)
    assert lib.sfg_get_poses(h, 0, nb, abi.ptr(host_np)) == 0
    # sfg_tick_poses = sfg_set_poses + sfg_tick + sfg_get_poses in one call (the upload queued in front of the tick, the download
    # overlapping its last kernels); both host buffers are pinned
    host_out = torch.empty(nb * 16, dtype=torch.uint8).pin_memory()
    host_out_np = host_out.numpy().view(np.float32).reshape(nb, 4)
    ffp = abi.ptr(ff)
    for _ in range(args.warmup):
        assert lib.sfg_tick_poses(h, abi.ptr(host_np), FRAME_DT, 0, 0.0, ffp, abi.ptr(host_out_np)) == 0
        host_np, host_out_np = host_out_np, host_np
    barrier()
    e2e_steps = args.steps
    per_step = []
    for _ in range(e2e_steps):
        t0 = time.perf_counter()
        rc = lib.sfg_tick_poses(h, abi.ptr(host_np), FRAME_DT, 0, 0.0, ffp, abi.ptr(host_out_np))   # H2D from pinned memory, tick, D2H into pinned memory, synchronous
        assert rc == 0
        per_step.append(time.perf_counter() - t0)
        host_np, host_out_np = host_out_np, host_np      # the poses read back are the ones written in next step
    barrier()
    # every call above returns only after its stream has drained, so the wall clock of one step is its end-to-end time;
    # the median keeps one host hiccup (a fresh box paging in) from deciding the number, the worst step is reported too
    e2e_s = max_over_ranks(float(np.median(per_step)))
    e2e_worst = max_over_ranks(max(per_step))
    e2e_value = world_size * nb * SUBSTEPS / e2e_s
    checksum = float(np.abs(host_np[:1000, 1]).sum())   # the step's result is read on the host

    # ---------------- the same world once the bodies have come to rest on the shelves (N = 1): every candidate pair touches, twice
    # the pairs and colours of the landing phase the headline is quoted on. Reported next to it, not instead of it.
    settled = None
    if world_size == 1 and not args.no_settled:
        ticks_done = args.warmup + e2e_steps          # g was reloaded for the end-to-end arm
        for _ in range(max(0, 300 - ticks_done)):
            g.tick(FRAME_DT, None, ff)
        t0 = time.perf_counter()
        n_s = 10
        sph = {"broadphase": 0.0, "graph": 0.0, "solver": 0.0, "device_total": 0.0}
        for _ in range(n_s):
            g.tick(FRAME_DT, None, ff)
            s2 = g.stats()
            sph["broadphase"] += float(s2["ms_broadphase"]) / n_s; sph["graph"] += float(s2["ms_graph"]) / n_s
            sph["solver"] += float(s2["ms_solver"]) / n_s; sph["device_total"] += float(s2["ms_total"]) / n_s
        s_s = (time.perf_counter() - t0) / n_s
        settled = {"after_ticks": max(300, ticks_done), "steps": n_s, "ms_per_step": s_s * 1e3, "value": nb * SUBSTEPS / s_s, "unit": "body-substeps/s",
                   "pairs": int(s2["n_pairs"]), "contact_colours": int(s2["n_contact_colours"]), "sleeping_bodies": int(s2["n_sleeping_bodies"]), "phases_ms": sph}

    # ---------------- slab arm (N > 1): ONE 1 M-body world cut into N x-slabs, halo bodies exchanged between x-neighbours
    # over NCCL point-to-point after every substep (north_star's second partition; strong scaling, reported next to `value`)
    slab_info = None
    if world_size > 1 and not args.no_slab:
        from moleengine_b200 import slab
        scs = sc if rank == 0 else workload(args.scale, 0xC3)        # every rank uploads the SAME world (seed 0xC3)
        gs = GpuWorld(tuning=abi.default_tuning(substeps=SUBSTEPS, flags=abi.TUNE_NO_SLEEPING), device=local_rank)
        gs.load_scene(scs)
        xs = scs["bodies"]["x"]
        halo = 4.0
        drv = slab.SlabDriver(gs, rank, world_size, float(xs.min()), float(xs.max()), halo=halo, substeps=SUBSTEPS, capacity=65536,
                              device=torch.device("cuda", local_rank))
        for _ in range(args.warmup):
            drv.tick(FRAME_DT, ff)
        barrier()
        t0 = time.perf_counter()
        dev_ms, recs = 0.0, 0
        for _ in range(args.steps):
            drv.tick(FRAME_DT, ff)
            dev_ms += float(gs.stats()["ms_total"]); recs += drv.sent_records
        barrier()
        slab_s = max_over_ranks(time.perf_counter() - t0) / args.steps
        recs_all = max_over_ranks(recs / args.steps)
        slab_info = {"value": len(scs["bodies"]) * SUBSTEPS / slab_s, "unit": "body-substeps/s", "scaling": "strong", "ms_per_step": slab_s * 1e3,
                     "device_ms_per_step_rank0": dev_ms / args.steps, "slabs": world_size, "halo_m": halo, "exchanges_per_tick": drv.exchanges,
                     "halo_records_per_tick_max_rank": recs_all, "halo_bytes_per_tick_max_rank": recs_all * 64,
                     "active_colliders_rank0": int(gs.stats()["n_colliders"]), "transport": "torch.distributed NCCL send/recv between x-neighbours"}
        gs.close()

    # ---------------- island-shard arm (N > 1, opt-in): the same single world, every rank solves the islands it owns and
    # all-gathers the bodies it solved (exact parity with one GPU; the broadphase and the island pass are replicated)
    island_info = None
    if world_size > 1 and args.island_arm:
        from moleengine_b200 import slab
        sci = sc if rank == 0 else workload(args.scale, 0xC3)
        gi = GpuWorld(tuning=abi.default_tuning(substeps=SUBSTEPS), device=local_rank)
        gi.load_scene(sci)
        drv = slab.IslandShardDriver(gi, rank, world_size, capacity=len(sci["bodies"]), device=torch.device("cuda", local_rank))
        for _ in range(args.warmup):
            drv.tick(FRAME_DT, ff)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            drv.tick(FRAME_DT, ff)
        barrier()
        isl_s = max_over_ranks(time.perf_counter() - t0) / args.steps
        island_info = {"value": len(sci["bodies"]) * SUBSTEPS / isl_s, "unit": "body-substeps/s", "scaling": "strong", "ms_per_step": isl_s * 1e3,
                       "ranks": world_size, "islands": int(gi.stats()["n_islands"]), "bodies_solved_rank0": drv.sent_records,
                       "collective": "all_gather_into_tensor of (n_bodies + 1) x 64 B per rank (NCCL)"}
        gi.close()

    if rank != 0:
        g.close()
        if world_size > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (persistent solver)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = algorithmic_bytes_per_tick(st, nb, SUBSTEPS)
    solver_s = solver_ms / args.steps / 1e3
    achieved = alg_bytes / solver_s / 1e9 if solver_s > 0 else 0.0

    # ---------------- CPU baseline: the oracle port of the reference tick on a bounded sample
    cpu = None
    if not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py as O
        from moleengine_b200 import scenes  # noqa: F401
        scs = workload("tenth", 0xC3)
        cores = os.cpu_count() or 1
        ow = O.OracleWorld(tuning=tun)
        ow.load_scene(scs)
        ow.tick(FRAME_DT, mode=O.MODE_REFERENCE, threads=cores)
        t0 = time.perf_counter()
        n_t = 4
        for _ in range(n_t):
            ow.tick(FRAME_DT, mode=O.MODE_REFERENCE, threads=cores)
        cdt = time.perf_counter() - t0
        cpu = {"value": len(scs["bodies"]) * SUBSTEPS * n_t / cdt, "unit": "body-substeps/s", "cores": cores, "kind": "port",
               "sample": f"c3_mixed at 1/10 scale ({len(scs['bodies'])} bodies), {n_t} ticks x {SUBSTEPS} substeps, f64 oracle restatement of the reference tick "
                         f"(BVH broadphase + DFS islands + sequential GS, island groups on {cores} threads)"}

    out = {
        "metric": "body-substeps/sec", "value": value, "unit": "body-substeps/s", "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"c3_mixed: {nb} mixed-collider bodies per GPU in one large world ({sc['name']})", "substeps": SUBSTEPS, "frame_hz": 60,
                   "pairs": int(st["n_pairs"]), "contact_colours": int(st["n_contact_colours"]), "l2": "working set (>1 GB of body + pair columns) exceeds the 126 MB L2",
                   "parallelism": f"independent worlds x{world_size}", "ticks_hz": 1.0 / step_s},
        "e2e": {"value": e2e_value, "unit": "body-substeps/s", "h2d_bytes_per_step": nb * 16, "d2h_bytes_per_step": nb * 16,
                "ms_per_step": e2e_s * 1e3, "worst_step_ms": e2e_worst * 1e3, "steps": e2e_steps, "statistic": "median step", "checksum": checksum},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_solve_persistent", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (profiled_traffic() or (None, None))[0], "traffic_source": (profiled_traffic() or (None, None))[1], "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": solver_s * 1e3},
        "phases_ms": {"broadphase": broad_ms / args.steps, "graph": graph_ms / args.steps, "solver": solver_ms / args.steps, "device_total": total_ms / args.steps},
        "cpu_baseline": cpu,
    }
    if settled is not None:
        out["settled"] = settled
    if slab_info is not None:
        out["slab"] = slab_info
    if island_info is not None:
        out["island_shards"] = island_info
    os.write(real_stdout, (json.dumps(out) + "\n").encode())
    g.close()
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

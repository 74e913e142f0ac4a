"""Throughput of every BASELINE.json config on one GPU (bench.py only times the headline config, C3).
One JSON line per config: bodies, candidate pairs, colours, per-phase device milliseconds (CUDA events inside sfg_tick),
body-substeps/s from the wall clock around synchronous ticks; C5 also times the batched casts of one step.
Usage: python tools/bench_configs.py [c1 c2 c4 c5 c3] [--ticks K] [--warmup W]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from moleengine_b200 import GpuWorld, abi, scenes  # noqa: E402


def build(name):
    if name == "c1":
        return scenes.c1_sandbox_stack(), 10, None
    if name == "c2":
        return scenes.c2_circles(500, 200), 4, None
    if name == "c3":
        return scenes.c3_mixed(1250, 800), 4, None
    if name == "c4":
        return scenes.c4_ropes(10000, 64, 30000), 4, None
    if name == "c5":
        sc = scenes.c5_worlds(8192, 16)
        return sc, 4, scenes.c5_rays(sc["bodies"], 8192, 256, 64)
    raise SystemExit(f"unknown config {name}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c1", "c2", "c4", "c5"])
    ap.add_argument("--ticks", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--cell-size", type=float, default=0.0, help="level-0 cell size override (SfgTuning.cell_size); 0 = automatic")
    args = ap.parse_args()
    for name in args.configs:
        t0 = time.perf_counter()
        sc, substeps, rays = build(name)
        t_build = time.perf_counter() - t0
        g = GpuWorld(tuning=abi.default_tuning(substeps=substeps, cell_size=args.cell_size))
        g.load_scene(sc)
        for _ in range(args.warmup):
            g.tick()
        ph = {"broadphase": 0.0, "graph": 0.0, "solver": 0.0, "device_total": 0.0}
        t0 = time.perf_counter()
        for _ in range(args.ticks):
            g.tick()
            st = g.stats()
            ph["broadphase"] += float(st["ms_broadphase"]); ph["graph"] += float(st["ms_graph"]); ph["solver"] += float(st["ms_solver"])
            ph["device_total"] += float(st["ms_total"])
        wall = time.perf_counter() - t0
        nb = int((sc["bodies"]["flags"] & abi.BODY_ALIVE != 0).sum())
        out = {"config": name, "scene": sc["name"], "bodies": nb, "colliders": int(st["n_colliders"]), "substeps": substeps, "pairs": int(st["n_pairs"]),
               "contact_colours": int(st["n_contact_colours"]), "constraint_colours": int(st["n_constraint_colours"]), "rope_particles": int(st["n_rope_particles"]),
               "islands": int(st["n_islands"]), "sleeping_bodies": int(st["n_sleeping_bodies"]), "ticks": args.ticks, "warmup": args.warmup,
               "ms_per_tick": wall / args.ticks * 1e3, "ticks_hz": args.ticks / wall, "body_substeps_per_s": nb * substeps * args.ticks / wall,
               "phases_ms": {k: v / args.ticks for k, v in ph.items()}, "launches_per_tick": int(st["kernel_launches"]), "host_scene_build_s": round(t_build, 2)}
        if rays is not None:
            hits = g.cast_batch(rays)
            t0 = time.perf_counter()
            for _ in range(5):
                hits = g.cast_batch(rays)
            dt = (time.perf_counter() - t0) / 5
            out["casts"] = {"rays": len(rays), "ms_per_batch_incl_pcie": dt * 1e3, "rays_per_s": len(rays) / dt, "hit_fraction": float((hits["collider"] >= 0).mean())}
        print(json.dumps(out), flush=True)
        g.close()


if __name__ == "__main__":
    main()

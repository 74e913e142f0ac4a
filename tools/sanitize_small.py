"""A small tour of every kernel family, written for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
(compute-sanitizer is closed on the pool this repository is developed on; tests/test_gpu_tour.py runs the same tour plainly, so a
crash or a CUDA error in any of these paths still fails the GPU suite)
ticks of a mixed scene (persistent solver and one launch per stage), rope topology edits, casts and queries, batched worlds,
island shards and slabs exchanged in-process (synchronous and stream-ordered), pose sync."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moleengine_b200 import GpuWorld, abi, scenes, slab  # noqa: E402

flags_list = [0, abi.TUNE_SEPARATE_LAUNCHES, abi.TUNE_COUNT_ENTRIES]
sc = scenes.parity_mix(150, 5)
for fl in flags_list:
    g = GpuWorld(tuning=abi.default_tuning(substeps=2, flags=fl))
    g.load_scene(sc)
    for _ in range(3):
        g.tick()
    g.get_contacts(); g.debug_manifolds(); g.debug_islands(); g.debug_pair_hints()
    rays = scenes.c5_rays(sc["bodies"], 1, len(sc["bodies"]), 16)
    g.cast_batch(rays)
    g.query_point_batch(np.array([[0.0, 1.0], [2.0, 3.0]], np.float32))
    q = np.zeros(1, abi.shape_query_dtype)
    q["rot_s"], q["kind"], q["p0"], q["p1"], q["layer_mask"] = 1.0, abi.POLY_RECT, 0.5, 0.5, 0xFFFFFFFFFFFFFFFF
    g.query_shape_batch(q)
    poses = g.get_poses()
    g.set_poses_indexed(np.array([1, 5, 7], np.uint32), poses[[1, 5, 7]])
    g.tick_poses(poses)
    g.close()
# ropes: cut, extend (with a move), dead particles
sc = scenes.c4_ropes(n_ropes=4, particles_per_rope=12, n_free_blocks=6)
g = GpuWorld(tuning=abi.default_tuning(substeps=2))
g.load_scene(sc)
g.tick()
g.rope_cut_after(1, 4)
g.rope_cut_after(2, 11)
b = g.get_bodies(); b["flags"] &= 15
kill = sc["rope_particles"][[2, 3, 30]]
b["flags"][kill] = 0
g.set_bodies(b)
g.rope_remove_dead_particles()
colls = sc["colliders"].copy(); colls["flags"][np.isin(colls["body"], kill)] = 0
g.set_colliders(colls)
g.set_constraints(sc["constraints"][~np.isin(sc["constraints"]["owner"], kill)])
ropes, pb = g.get_ropes()
if (ropes["particle_count"] >= 2).all():
    g.tick()
g.close()
# batched worlds + many colours
sc = scenes.c5_worlds(3, 8)
g = GpuWorld(tuning=abi.default_tuning(substeps=2))
g.load_scene(sc)
for _ in range(3):
    g.tick()
g.close()
g = GpuWorld(tuning=abi.default_tuning(substeps=1))
g.load_scene(scenes.star(60))
g.tick()
g.close()
# slabs (synchronous and stream-ordered exchange) and island shards, all in one process
import torch  # noqa: E402
sc = scenes.c3_mixed(40, 10, seed=3)
x0 = float(sc["bodies"]["x"].min()), float(sc["bodies"]["x"].max())
for asyn in (False, True):
    tr = slab.LocalTransport()
    ws, ds = [], []
    for r in range(2):
        w = GpuWorld(tuning=abi.default_tuning(substeps=2, flags=abi.TUNE_NO_SLEEPING))
        w.load_scene(sc)
        ws.append(w)
        ds.append(slab.SlabDriver(w, r, 2, x0[0], x0[1], halo=4.0, substeps=2, capacity=512, device=torch.device("cuda", 0), transport=tr))
    for _ in range(2):
        (slab.tick_local_async if asyn else slab.tick_local)(ds, tr)
    torch.cuda.synchronize()
    slab.cast_local(ds, scenes.c5_rays(sc["bodies"], 1, len(sc["bodies"]), 8))
    for w in ws:
        w.close()
ws, ds = [], []
for r in range(2):
    w = GpuWorld(tuning=abi.default_tuning(substeps=2))
    w.load_scene(sc)
    ws.append(w)
    ds.append(slab.IslandShardDriver(w, r, 2, capacity=len(sc["bodies"]), device=torch.device("cuda", 0)))
slab.tick_island_shards_local(ds)
for w in ws:
    w.close()
print("sanitize_small: done")

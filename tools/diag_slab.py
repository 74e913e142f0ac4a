"""Diagnostic: per-tick deviation of a 2-slab run from the single-world run (one process, cuda:0)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from moleengine_b200 import GpuWorld, abi, scenes, slab

n_slabs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sc = scenes.c3_mixed(120, 30, seed=0xC3)
nb = len(sc["bodies"])
x0 = float(sc["bodies"]["x"].min()), float(sc["bodies"]["x"].max())
tun = abi.default_tuning(substeps=4, flags=abi.TUNE_NO_SLEEPING)
ref = GpuWorld(tuning=tun); ref.load_scene(sc)
tr = slab.LocalTransport()
worlds, drivers = [], []
for r in range(n_slabs):
    w = GpuWorld(tuning=tun); w.load_scene(sc); worlds.append(w)
    drivers.append(slab.SlabDriver(w, r, n_slabs, x0[0], x0[1], halo=4.0, substeps=4, capacity=4096, device=torch.device("cuda", 0), transport=tr))
cuts = [slab.slab_bounds(r, n_slabs, x0[0], x0[1])[1] for r in range(n_slabs - 1)]
for t in range(40):
    ref.tick(); slab.tick_local(drivers, tr)
    want = ref.get_bodies()
    got = np.zeros(nb, abi.body_dtype); owners = np.zeros(nb, int)
    for w in worlds:
        b = w.get_bodies()
        owned = ((b["flags"] & 1) != 0) & ((b["flags"] & 96) == 0)
        got[owned] = b[owned]; owners += owned
    d = np.hypot(got["x"] - want["x"], got["y"] - want["y"])
    dc = np.min(np.abs(want["x"][:, None] - np.array(cuts + [1e9])[None, :]), axis=1)
    bad = d > 1e-4
    print(t, "owners ok", bool((owners == 1).all()), "max", float(d.max()), "n>1e-4", int(bad.sum()), "min dist-to-cut of bad", float(dc[bad].min()) if bad.any() else None,
          "max dist-to-cut of bad", float(dc[bad].max()) if bad.any() else None, "pairs", int(ref.stats()["n_pairs"]), [int(w.stats()["n_pairs"]) for w in worlds])

"""Replay the tick in which a long rope run goes non-finite through the oracle (f64 and f32, same coloured order)."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import oracle_py as O
from moleengine_b200 import GpuWorld, abi, scenes
n_ropes = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
sc = scenes.c4_ropes(n_ropes=n_ropes, particles_per_rope=64, n_free_blocks=3 * n_ropes)
tun = abi.default_tuning(substeps=4, flags=abi.TUNE_NO_SLEEPING)
g = GpuWorld(tuning=tun); g.load_scene(sc)
prev = sc["bodies"].copy()
for t in range(600):
    g.tick()
    b = g.get_bodies()
    if not (np.isfinite(b["x"]).all() and np.isfinite(b["vx"]).all()):
        print("GPU non-finite at tick", t, int((~np.isfinite(b["x"])).sum()), "bodies")
        break
    prev = b
else:
    print("finite"); sys.exit(0)
prev["flags"] &= 15
sc2 = dict(sc); sc2["bodies"] = prev
# fresh GPU world from the last finite state: is it reproducible, and in which substep count does it appear?
for sub in (1, 2, 3, 4):
    g2 = GpuWorld(tuning=abi.default_tuning(substeps=sub, flags=abi.TUNE_NO_SLEEPING)); g2.load_scene(sc2)
    g2.tick(1.0 / 60.0 * sub / 4.0)          # same dt per substep
    bb = g2.get_bodies()
    bad = np.flatnonzero(~np.isfinite(bb["x"]))
    print("GPU replay with", sub, "substeps:", len(bad), "non-finite", bad[:8].tolist())
    if len(bad) and sub == 1 or sub == 4:
        pairs = g2.debug_pairs()
        for f32 in (False, True):
            o = O.OracleWorld(use_f32=f32, tuning=abi.default_tuning(substeps=sub, flags=abi.TUNE_NO_SLEEPING)); o.load_scene(sc2)
            o.set_external_pairs(pairs)
            o.tick(1.0 / 60.0 * sub / 4.0, mode=O.MODE_COLOURED)
            ob = o.get_bodies()
            obad = np.flatnonzero(~np.isfinite(ob[:, 0]))
            print("   oracle", "f32" if f32 else "f64", "same order:", len(obad), "non-finite", obad[:8].tolist())
    g2.close()
    if len(bad):
        import os
        os.environ["SFG_DEBUG_NAN"] = "1"
        g4 = GpuWorld(tuning=abi.default_tuning(substeps=sub, flags=abi.TUNE_NO_SLEEPING)); g4.load_scene(sc2)
        g4.tick(1.0 / 60.0 * sub / 4.0)       # persistent kernel with the per-stage finite check
        g4.close()
        g3 = GpuWorld(tuning=abi.default_tuning(substeps=sub, flags=abi.TUNE_NO_SLEEPING | abi.TUNE_SEPARATE_LAUNCHES)); g3.load_scene(sc2)
        g3.tick(1.0 / 60.0 * sub / 4.0)
        bb3 = g3.get_bodies()
        print("   separate launches:", int((~np.isfinite(bb3["x"])).sum()), "non-finite")
        hi, lo, mult, col = g3.debug_contact_units()
        cbod = sc["colliders"]["body"]
        first_bad = int(bad[0])
        sel = (cbod[hi] == first_bad) | (cbod[lo] == first_bad)
        print("   units on body", first_bad, ":", [(int(h), int(l), int(c), int(sc["colliders"]["kind"][h]), int(sc["colliders"]["kind"][l])) for h, l, c in zip(hi[sel][:12], lo[sel][:12], col[sel][:12])])
        g3.close()
        # describe the first bad bodies: rope particle? which colliders, which partners
        parts = set(sc["rope_particles"].tolist())
        cb = sc["colliders"]["body"]
        for i in bad[:3]:
            mine = np.flatnonzero(cb == i)
            pr = pairs[np.isin(pairs[:, 0], mine) | np.isin(pairs[:, 1], mine)] if len(bad) and sub in (1, 4) else []
            print("   body", int(i), "rope particle" if int(i) in parts else "block", "state", {k: float(prev[k][i]) for k in ("x", "y", "vx", "vy", "w")}, "candidate pairs", len(pr))
        break

"""Where does the first non-finite body appear in a long rope run? (diagnostic)"""
import sys
import numpy as np
sys.path.insert(0, ".")
from moleengine_b200 import GpuWorld, abi, scenes
n_ropes = int(sys.argv[1]) if len(sys.argv) > 1 else 400
sc = scenes.c4_ropes(n_ropes=n_ropes, particles_per_rope=64, n_free_blocks=3 * n_ropes)
g = GpuWorld(tuning=abi.default_tuning(substeps=4)); g.load_scene(sc)
parts = set(sc["rope_particles"].tolist())
prev = sc["bodies"].copy()
for t in range(600):
    g.tick()
    b = g.get_bodies()
    bad = ~(np.isfinite(b["x"]) & np.isfinite(b["y"]) & np.isfinite(b["vx"]) & np.isfinite(b["rot_s"]))
    if bad.any():
        idx = np.flatnonzero(bad)
        print("tick", t, "first non-finite bodies", idx[:10].tolist(), "of", int(bad.sum()), "| rope particles among them:", sum(int(i) in parts for i in idx[:50]), "/", min(50, len(idx)))
        for i in idx[:4]:
            print("   body", int(i), "before:", {k: float(prev[k][i]) for k in ("x", "y", "vx", "vy", "w", "rot_s", "rot_b")}, "inv_mass", float(prev["inv_mass"][i]))
        st = g.stats(); print("   pairs", int(st["n_pairs"]), "colours", int(st["n_contact_colours"]))
        break
    prev = b
    if t % 100 == 99:
        v = np.hypot(b["vx"], b["vy"]); print("tick", t, "ok; max speed", float(v.max()), "pairs", int(g.stats()["n_pairs"]), "colours", int(g.stats()["n_contact_colours"]))
else:
    print("finite for 600 ticks")

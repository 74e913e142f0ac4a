"""Units per colour and how many of them carry the `near` hint (C3 at tick 15 by default)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moleengine_b200 import GpuWorld, abi, scenes  # noqa: E402
ticks = int(sys.argv[1]) if len(sys.argv) > 1 else 15
g = GpuWorld(tuning=abi.default_tuning(substeps=4))
g.load_scene(scenes.c3_mixed(1250, 800))
for _ in range(ticks):
    g.tick()
p = g.debug_pairs().astype(np.uint64)
h = g.debug_pair_hints()
hi, lo, mult, col = g.debug_contact_units()
key_p = p[:, 0] << np.uint64(32) | p[:, 1]
order = np.argsort(key_p)
key_u = hi.astype(np.uint64) << np.uint64(32) | lo.astype(np.uint64)
near_u = h[order][np.searchsorted(key_p[order], key_u)]
m = g.debug_manifolds()
n = len(hi)
touch = (m[:n, 2] > 0) | (m[n:2 * n, 2] > 0) if len(m) >= 2 * n else (m[:n, 2] > 0)
print("colour units near touching(last substep)")
for c in range(int(col.max()) + 1):
    s = col == c
    print(f"{c:3d} {int(s.sum()):8d} {int(near_u[s].sum()):8d} {int(touch[s].sum()):8d}")
print("total", n, int(near_u.sum()), int(touch.sum()))

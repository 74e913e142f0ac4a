"""Distribution of the in-tick manifold error (device f32 vs f64 oracle replaying the same coloured order, one substep) split by
WHEN in the Gauss-Seidel sweep the visit happened: the first four colours see inputs that carry no propagated rounding yet, later
visits see poses that already went through up to 2 x colours f32 corrections. This is the evidence behind oracle/parity.py's
TOL_MANIFOLD (1e-5, north_star) for early visits and TOL_MANIFOLD_LATE (3e-5) for late ones.
Usage: python tools/diag_manifold_dist.py > gpurun_out/manifold_dist.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
import parity  # noqa: E402
from moleengine_b200 import GpuWorld, abi, scenes  # noqa: E402

out = []
cases = [("parity_mix_400_s7", scenes.parity_mix(400, 7)), ("parity_mix_400_s11", scenes.parity_mix(400, 11)), ("parity_mix_1200_s3", scenes.parity_mix(1200, 3)),
         ("c5_worlds_12", scenes.c5_worlds(12, 16)), ("c5_worlds_6_dense", scenes.c5_worlds(6, 16, dense=True)), ("c3_mixed_60x40", scenes.c3_mixed(60, 40))]
for name, sc in cases:
    tun = abi.default_tuning(substeps=1)
    g = GpuWorld(tuning=tun)
    g.load_scene(sc)
    # let the scene develop contacts first (the device runs it), then compare ONE substep from that state
    for _ in range(30):
        g.tick()
    sc2 = dict(sc)
    sc2["bodies"] = g.get_bodies()
    sc2["bodies"]["flags"] &= 15
    g.close()
    g = GpuWorld(tuning=tun)
    g.load_scene(sc2)
    g.tick()
    o = O.OracleWorld(tuning=tun)
    o.load_scene(sc2)
    o.set_external_pairs(g.debug_pairs())
    o.tick(mode=O.MODE_COLOURED)
    gmr, omr, uh, ul, visit = parity.match_manifolds(g, o)
    cnt_g, cnt_o = gmr[:, 2].astype(int), omr[:, 2].astype(int)
    same = (cnt_g == cnt_o) & (cnt_o > 0)
    d = parity._manifold_diff(gmr, omr, cnt_o)
    row = {"scene": name, "entries_compared": int(same.sum()), "colours": int(visit.max() // 2 + 1)}
    for label, sel in (("early_visits(first 4 colours)", same & (visit < 8)), ("late_visits", same & (visit >= 8))):
        v = d[sel]
        v = v[np.isfinite(v) & (v < 1e-3)]            # equal-feature flips (O(1) differences, oracle/parity.py) are counted apart
        row[label] = {"n": int(len(v)), "p50": float(np.percentile(v, 50)) if len(v) else None, "p90": float(np.percentile(v, 90)) if len(v) else None,
                      "p99": float(np.percentile(v, 99)) if len(v) else None, "p99.9": float(np.percentile(v, 99.9)) if len(v) else None,
                      "max": float(v.max()) if len(v) else None, "over_1e-5": int((v > 1e-5).sum()), "over_3e-5": int((v > 3e-5).sum())}
    row["flips_or_worse(>1e-3)"] = int((d[same] >= 1e-3).sum())
    out.append(row)
    g.close()
print(json.dumps(out, indent=1))

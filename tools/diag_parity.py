import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import oracle_py as O
from moleengine_b200 import GpuWorld, abi, scenes

def state(gb):
    return np.stack([gb["x"], gb["y"], gb["rot_s"], gb["rot_b"], gb["vx"], gb["vy"], gb["w"]], 1).astype(np.float64)

def run(name, sc, substeps=1, ff=None):
    tun = abi.default_tuning(substeps=substeps, flags=abi.TUNE_SEPARATE_LAUNCHES)
    g = GpuWorld(tuning=tun); g.load_scene(sc); g.tick(forcefield=ff)
    got = state(g.get_bodies())
    o = O.OracleWorld(tuning=tun); o.load_scene(sc); o.set_external_pairs(g.debug_pairs()); o.tick(mode=O.MODE_COLOURED, forcefield=ff)
    ref = o.get_bodies()
    ep = np.abs(got[:, :4] - ref[:, :4]).max(1)
    ev = (np.abs(got[:, 4:] - ref[:, 4:]) / np.maximum(1.0, np.abs(ref[:, 4:]))).max(1)
    print(f"== {name}: pos {ep.max():.3e} vel {ev.max():.3e}")
    worst = np.argsort(-ev)[:6]
    b = sc["bodies"]
    rp = set(sc["rope_particles"].tolist())
    for i in worst:
        kind = "rope" if i in rp else ("kin" if not (b["flags"][i] & 12) else "dyn")
        colls = np.nonzero(sc["colliders"]["body"] == i)[0]
        print(f"  body {i} {kind} colls {colls.tolist()} kinds {sc['colliders']['kind'][colls].tolist()} ev {ev[i]:.3e} got {got[i,4:]} ref {ref[i,4:]}")
    # manifolds compare
    gm = g.debug_manifolds(); om = o.manifolds()
    hi, lo, mult, col = g.debug_contact_units()
    n = len(hi)
    # oracle entries are unit-major (mult copies adjacent) in ITS order; build dict
    od = {}
    cnt = {}
    for row in om:
        k = (int(row[0]), int(row[1])); c = cnt.get(k, 0); cnt[k] = c + 1
        od[(k, c)] = row
    bad = 0
    for u in range(n):
        for m in range(mult[u]):
            ge = gm[u + m * n]; oe = od[((int(hi[u]), int(lo[u])), m)]
            if int(ge[2]) != int(oe[2]) or np.abs(ge[3:] - oe[3:]).max() > 1e-3 * max(1, np.abs(oe[3:]).max()):
                bad += 1
                if bad < 6: print("  manifold diff", (hi[u], lo[u], m), "gpu", ge[2:], "ora", oe[2:])
    print("  manifold mismatches:", bad, "of", int(mult.sum()))
    g.close()

run("mix no ropes/constraints/kin", scenes.parity_mix(400, 7, n_ropes=0, with_constraints=False, with_kinematic=False))
run("mix kin only", scenes.parity_mix(400, 7, n_ropes=0, with_constraints=False, with_kinematic=True))
run("mix ropes no constraints", scenes.parity_mix(400, 7, n_ropes=3, with_constraints=False, with_kinematic=False))
run("mix all", scenes.parity_mix(400, 7))
run("c1", scenes.c1_sandbox_stack(), substeps=1)
run("c2 small", scenes.c2_circles(60, 30), substeps=1)

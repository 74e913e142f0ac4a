"""The reference sandbox's own scenes (scenes/*.ron, loaded by moleengine_b200/ron.py into the committed fixtures) on the
GPU against the oracle: one substep replayed in the GPU's coloured order (the strict gate), and 60 ticks against the f64
reference-order run stored in the fixture (different Gauss-Seidel order, so only gross agreement is asked there)."""
import glob
import os

import numpy as np
import pytest

from moleengine_b200 import GpuWorld, abi

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = sorted(glob.glob(os.path.join(GOLD, "ron_*.npz")))


def _load(path):
    f = np.load(path)
    return {k: f[k] for k in ("bodies", "colliders", "constraints", "ropes", "rope_particles")}, f


@pytest.mark.parametrize("path", FIXTURES)
def test_reference_scene_one_substep_parity(oracle, path):
    import parity
    sc, f = _load(path)
    ff = abi.gravity_field(*f["gravity"].tolist())
    tun = abi.default_tuning(substeps=1)
    g = GpuWorld(tuning=tun)
    g.load_scene(sc)
    o = oracle.OracleWorld(tuning=tun)
    o.load_scene(sc)
    for t in range(3):                                   # three ticks so that things are in contact
        if t:
            gb = g.get_bodies(); gb["flags"] &= 15
            o.set_bodies(gb)
        g.tick(forcefield=ff)
        o.set_external_pairs(g.debug_pairs())
        o.tick(mode=oracle.MODE_COLOURED, forcefield=ff)
    r = parity.compare_tick(g, o, sc, hops=3)
    print(os.path.basename(path), r)
    assert r["flips_ok"] and r["non_tie_mismatches"] == 0 and r["bad_outside_influence"] == 0
    assert r["manifolds_ok"] and r["pos_err_clean"] < 1e-4 and r["vel_err_clean"] < 1e-4
    g.close()


@pytest.mark.parametrize("path", FIXTURES)
def test_reference_scene_60_ticks(path):
    sc, f = _load(path)
    ff = abi.gravity_field(*f["gravity"].tolist())
    g = GpuWorld(tuning=abi.default_tuning(substeps=10))
    g.load_scene(sc)
    for _ in range(60):
        g.tick(forcefield=ff)
    gb = g.get_bodies()
    want = f["oracle_bodies_60"]
    got = np.stack([gb["x"], gb["y"]], 1).astype(np.float64)
    d = np.hypot(got[:, 0] - want[:, 0], got[:, 1] - want[:, 1])
    print(os.path.basename(path), {"median": float(np.median(d)), "p90": float(np.percentile(d, 90)), "max": float(d.max())})
    assert np.isfinite(got).all()
    # one second of simulation in a different (coloured, f32) Gauss-Seidel order: the bulk of the scene is where the reference puts it
    assert np.median(d) < 0.05 and np.percentile(d, 90) < 0.5
    g.close()

"""The long-run statistical gate (oracle/longrun.py) on the CPU: the oracle's two orders — the reference's DFS order and the
coloured order the device uses, both f64 — run the same scenes for hundreds of ticks and must agree on energy, momentum and the
penetration-depth distribution within the SAME stated tolerances the device is held to in tests/test_gpu_longrun.py. This keeps
the tolerances honest: they are what separates two correct Gauss-Seidel orders of a chaotic pile, not room for a device bug."""
import longrun
import pytest

from moleengine_b200 import abi


@pytest.mark.parametrize("name", ["c2", "c4"])
def test_two_correct_orders_pass_the_gate(oracle, name):
    sc, substeps, ticks = getattr(longrun, "scene_" + name)()
    tun = abi.default_tuning(substeps=substeps)
    ref = longrun.run_oracle(oracle, sc, tun, ticks, 50, oracle.MODE_REFERENCE)
    col = longrun.run_oracle(oracle, sc, tun, ticks, 50, oracle.MODE_COLOURED)
    dev = longrun.compare(col, ref, sc)
    print(name, dev)
    assert ref[-1][1]["pen_n"] > 50, "the scene never formed contacts"
    assert not longrun.check(dev, longrun.GATES[name]), longrun.check(dev, longrun.GATES[name])


def test_gate_catches_a_wrong_run(oracle):
    """The same gate must FAIL for a run that is physically different: gravity 3 % stronger."""
    sc, substeps, ticks = longrun.scene_c2()
    tun = abi.default_tuning(substeps=substeps)
    ref = longrun.run_oracle(oracle, sc, tun, 150, 50, oracle.MODE_REFERENCE)
    ff = abi.gravity_field()
    ff["terms"][0][0]["b"] *= 1.03
    w = oracle.OracleWorld(tuning=tun)
    w.load_scene(sc)
    out = []
    for t in range(1, 151):
        w.tick(mode=oracle.MODE_REFERENCE, forcefield=ff)
        if t % 50 == 0:
            out.append((t, longrun.sample(sc, w.get_bodies(), oracle)))
    bad = longrun.check(longrun.compare(out, ref, sc), longrun.GATES["c2"])
    assert bad, "a run under 3 % stronger gravity passed the long-run gate"


def test_c1_reference_run_comes_to_rest(oracle):
    """Config 1 at full spec in the reference order: the stack is at rest and asleep long before tick 600 (the device is compared
    with this run in tests/test_gpu_longrun.py)."""
    sc, substeps, ticks = longrun.scene_c1()
    ref = longrun.run_oracle(oracle, sc, abi.default_tuning(substeps=substeps), ticks, 100, oracle.MODE_REFERENCE)
    assert ref[-1][1]["ke"] < 1e-3 and abs(ref[-1][1]["mean_y"] - ref[1][1]["mean_y"]) < 1e-4
    assert ref[-1][1]["pen_n"] >= 100

"""Long-run statistical parity (BASELINE.json north_star: "over long runs, where the dynamics are chaotic, energy, momentum and
penetration statistics must agree"): the device runs BASELINE configs for hundreds of ticks next to the f64 oracle in
MODE_REFERENCE (the reference's own order: BVH pairs, DFS islands, sequential Gauss-Seidel, sleeping) and every 50 ticks the two
states are reduced to kinetic / potential energy, linear / angular momentum and the penetration-depth distribution of all touching
solid pairs (count, p50, p99, max). Tolerances are stated in oracle/longrun.py GATES."""
import json
import os

import longrun
import pytest

from moleengine_b200 import GpuWorld, abi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["c1", "c2", "c4"])
def test_long_run_statistics_follow_the_reference(oracle, name):
    sc, substeps, ticks = getattr(longrun, "scene_" + name)()
    tun = abi.default_tuning(substeps=substeps)
    every = 100 if name == "c1" else 50
    ref = longrun.run_oracle(oracle, sc, tun, ticks, every, oracle.MODE_REFERENCE)
    g = GpuWorld(tuning=tun)
    g.load_scene(sc)
    dev_run = longrun.run_gpu(g, sc, ticks, every, oracle)
    g.close()
    dev = longrun.compare(dev_run, ref, sc)
    report = {"scene": name, "ticks": ticks, "substeps": substeps, "deviation": {k: float(v) for k, v in dev.items()}, "gates": longrun.GATES[name],
              "device_last": dev_run[-1][1], "reference_last": ref[-1][1]}
    print(json.dumps(report))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"longrun_{name}.json"), "w") as f:
            json.dump(report, f, indent=1)
    assert ref[-1][1]["pen_n"] > 50, "the scene never formed contacts"
    bad = longrun.check(dev, longrun.GATES[name])
    assert not bad, bad

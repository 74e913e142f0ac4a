"""Where does a lone contact unit's latency go? Needs a -DSFG_PROBE build of the library (SFG_LIB=...): stage_contacts stamps
%clock64 after the unit header, the reach test, the remaining loads, each visit and the worklist append.
Usage: SFG_LIB=.../libsfg_probe.so python tools/diag_probe.py [settle_ticks]"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moleengine_b200 import GpuWorld, abi, scenes  # noqa: E402


def report(path, label):
    raw = np.fromfile(path, dtype=np.uint64)
    n = int(min(raw[0], 65536))
    t = raw[1:1 + 8 * n].reshape(n, 8).astype(np.int64)
    names = ["header + body loads + reach test", "remaining loads / setup", "visit 0 narrowphase", "visit 0 solve", "visit 1", "stores + WL append"]
    d = np.stack([t[:, 2] - t[:, 0], t[:, 3] - t[:, 2], t[:, 1] - t[:, 3], t[:, 4] - t[:, 1], t[:, 5] - t[:, 4], t[:, 6] - t[:, 5]], 1)
    ok = (t[:, :7] > 0).all(1)
    d = d[ok]
    print(f"{label}: {n} probed units ({ok.sum()} complete); cycles median / mean / p90")
    for i, nm in enumerate(names):
        print(f"  {nm:34s} {np.median(d[:, i]):8.0f} {d[:, i].mean():8.0f} {np.percentile(d[:, i], 90):8.0f}")
    tot = d.sum(1)
    print(f"  {'total':34s} {np.median(tot):8.0f} {tot.mean():8.0f} {np.percentile(tot, 90):8.0f}")


def main():
    settle = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    sc = scenes.c3_mixed(1250, 800)
    g = GpuWorld(tuning=abi.default_tuning(substeps=4))
    g.load_scene(sc)
    for _ in range(14):
        g.tick()
    os.environ["SFG_PROBE_OUT"] = "gpurun_out/probe_a.bin"
    os.environ["SFG_TRACE"] = "gpurun_out/trace_a.bin"
    g.tick()
    del os.environ["SFG_PROBE_OUT"], os.environ["SFG_TRACE"]
    print(g.stats())
    report("gpurun_out/probe_a.bin", "tick 15")
    if settle:
        for _ in range(settle):
            g.tick()
        os.environ["SFG_PROBE_OUT"] = "gpurun_out/probe_b.bin"
        os.environ["SFG_TRACE"] = "gpurun_out/trace_b.bin"
        g.tick()
        del os.environ["SFG_PROBE_OUT"], os.environ["SFG_TRACE"]
        print(g.stats())
        report("gpurun_out/probe_b.bin", f"tick {15 + settle}")


main()

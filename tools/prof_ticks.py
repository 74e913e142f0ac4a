"""Run a config to a given tick, then profile the next ticks only (ncu --profile-from-start off picks up cudaProfilerStart).
Usage: ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
           python tools/prof_ticks.py <c1|c2|c3|c4|c5> <warm_ticks> <profiled_ticks>"""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moleengine_b200 import GpuWorld, abi  # noqa: E402
from tools.bench_configs import build  # noqa: E402

name, warm, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
sc, substeps, _rays = build(name)
g = GpuWorld(tuning=abi.default_tuning(substeps=substeps, flags=int(os.environ.get("SFG_PROF_FLAGS", "0"))))   # e.g. 2 = one launch per stage
g.load_scene(sc)
for _ in range(warm):
    g.tick()
rt = ctypes.CDLL("libcuda.so.1")
rt.cuProfilerStart()
for _ in range(n):
    g.tick()
    if _rays is not None:
        g.cast_batch(_rays)
rt.cuProfilerStop()
st = g.stats()
print(name, "tick", warm + n, "pairs", int(st["n_pairs"]), "colours", int(st["n_contact_colours"]), "ms", float(st["ms_total"]), file=sys.stderr)

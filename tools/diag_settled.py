"""Stage trace of one tick of a config after `warm` ticks (SFG_TRACE dump read by tools/read_trace.py) + the phase times around it.
Usage: python tools/diag_settled.py <c2|c3|c4|c5> <warm_ticks> <trace_out.bin>"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moleengine_b200 import GpuWorld, abi  # noqa: E402
from tools.bench_configs import build  # noqa: E402

name, warm, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
sc, substeps, _ = build(name)
g = GpuWorld(tuning=abi.default_tuning(substeps=substeps, flags=int(os.environ.get("SFG_PROF_FLAGS", "0"))))   # e.g. 1 = no island pass
g.load_scene(sc)
for _ in range(warm):
    g.tick()
ph = {"broadphase": 0.0, "graph": 0.0, "solver": 0.0, "device_total": 0.0}
for _ in range(5):
    g.tick()
    st = g.stats()
    for k in ph:
        ph[k] += float(st["ms_" + ("total" if k == "device_total" else k)]) / 5
os.environ["SFG_TRACE"] = out
g.tick()
del os.environ["SFG_TRACE"]
st = g.stats()
print(json.dumps({"config": name, "tick": warm + 6, "pairs": int(st["n_pairs"]), "colours": int(st["n_contact_colours"]), "launches": int(st["kernel_launches"]),
                  "phases_ms": ph}))

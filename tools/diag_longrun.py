"""Long-run sanity on one GPU: every BASELINE scene for many ticks; finite state, bounded speeds, bodies inside the scene, sleeping."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from moleengine_b200 import GpuWorld, abi, scenes

def run(name, sc, substeps, ticks, ybound):
    g = GpuWorld(tuning=abi.default_tuning(substeps=substeps)); g.load_scene(sc)
    t0 = time.perf_counter()
    worst = 0.0
    for t in range(ticks):
        g.tick()
        worst = max(worst, float(g.stats()["ms_total"]) if t > 2 else 0.0)
    dt = time.perf_counter() - t0
    b = g.get_bodies()
    alive = (b["flags"] & abi.BODY_ALIVE) != 0
    v = np.hypot(b["vx"][alive], b["vy"][alive])
    st = g.stats()
    print(name, {"ticks": ticks, "ms_per_tick": round(dt / ticks * 1e3, 3), "worst_tick_ms": round(worst, 2), "finite": bool(np.isfinite(b["x"][alive]).all() and np.isfinite(v).all()),
                 "max_speed": float(v.max()), "median_speed": float(np.median(v)), "y_min": float(b["y"][alive].min()), "below_floor": int((b["y"][alive] < ybound).sum()),
                 "asleep": int(((b["flags"] & abi.BODY_ASLEEP) != 0).sum()), "islands": int(st["n_islands"]), "pairs": int(st["n_pairs"]), "colours": int(st["n_contact_colours"])})
    g.close()

run("c1", scenes.c1_sandbox_stack(), 10, 1200, -6.0)
run("c2", scenes.c2_circles(500, 200), 4, 900, -1.0)
run("c4", scenes.c4_ropes(10000, 64, 30000), 4, 400, -8.0)
run("c3_tenth", scenes.c3_mixed(400, 250), 4, 900, -1e9)
if len(sys.argv) > 1 and sys.argv[1] == "big":
    run("c5", scenes.c5_worlds(8192, 16), 4, 200, -6.0)
    run("c3", scenes.c3_mixed(1250, 800), 4, 400, -1e9)

import sys, numpy as np
sys.path.insert(0, ".")
from moleengine_b200 import GpuWorld, abi, scenes
sc = scenes.c3_mixed(1250, 800)
g = GpuWorld(tuning=abi.default_tuning(substeps=4)); g.load_scene(sc)
for t in range(40):
    g.tick()
    if t % 3 == 0 or t < 6:
        b = g.get_bodies()
        v2 = b["vx"].astype(np.float64)**2 + b["vy"].astype(np.float64)**2 + b["w"].astype(np.float64)**2
        st = g.stats()
        print(t, "slow bodies", int((v2 < 0.001).sum()), "islands", int(st["n_islands"]), "pairs", int(st["n_pairs"]), "ms", round(float(st["ms_total"]),3), "graph", round(float(st["ms_graph"]),3))

import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from moleengine_b200 import GpuWorld, abi, scenes
sc = scenes.c3_mixed(1250, 800)
nb = len(sc["bodies"])
g = GpuWorld(tuning=abi.default_tuning(substeps=4)); g.load_scene(sc)
host = torch.empty(nb * 16, dtype=torch.uint8).pin_memory()
h = host.numpy().view(np.float32).reshape(nb, 4)
lib = g.lib
lib.sfg_get_poses(g.h, 0, nb, abi.ptr(h))
for name, fn in (("set_poses", lambda: lib.sfg_set_poses(g.h, 0, nb, abi.ptr(h))), ("get_poses", lambda: lib.sfg_get_poses(g.h, 0, nb, abi.ptr(h))), ("tick", lambda: g.tick())):
    for _ in range(3): fn()
    ts = []
    for _ in range(15):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print(name, "median ms", np.median(ts) * 1e3, "min", min(ts) * 1e3, "max", max(ts) * 1e3)
# raw pinned copy bandwidth through torch
d = torch.empty(nb * 16, dtype=torch.uint8, device="cuda")
for _ in range(3): d.copy_(host, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): d.copy_(host, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print("torch H2D 16MB ms", dt * 1e3, "GB/s", nb * 16 / dt / 1e9)
t0 = time.perf_counter()
for _ in range(20): host.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print("torch D2H 16MB ms", dt * 1e3, "GB/s", nb * 16 / dt / 1e9)

"""Generates tests/golden/c3_tenth_settled300.npz: the 1/10-scale C3 world (100 000 bodies, seed 0xC3) after 300 ticks of the
ORACLE in the reference order (f64, MODE_REFERENCE, 4 substeps) — the bodies have landed on the shelves and lie in piles.
bench.py's CPU arms (--impl reference and cpu_baseline) load it so that their bounded sample is the same steady-state workload
the device arm's headline is quoted on, without spending minutes of every bench run on settling the scene on the CPU.
Nothing of the device path is involved. Takes ~10 minutes on 8 cores:  python tools/make_settled_fixture.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
from moleengine_b200 import abi, scenes  # noqa: E402

TICKS = 300


def main():
    sc = scenes.c3_mixed(400, 250, seed=0xC3)
    w = O.OracleWorld(tuning=abi.default_tuning(substeps=4))
    w.load_scene(sc)
    t0 = time.time()
    for t in range(TICKS):
        w.tick(mode=O.MODE_REFERENCE, threads=os.cpu_count() or 1)
        if t % 25 == 24:
            print(t + 1, "ticks", round(time.time() - t0, 1), "s", w.stats(), flush=True)
    st = w.get_bodies().astype(np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c3_tenth_settled300.npz"), state=st, ticks=np.int32(TICKS), nx=np.int32(400), ny=np.int32(250),
                        seed=np.int32(0xC3), stats=np.array(list(w.stats().values())))


if __name__ == "__main__":
    main()

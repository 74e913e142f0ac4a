"""Generates tests/golden/*.npz: regression fixtures of the ORACLE (not outputs of the reference — the reference is
Rust and cannot run here). They freeze the oracle's behaviour so later edits cannot silently change
what the GPU path is compared with. Re-run only when the oracle is deliberately changed:  python tools/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
from moleengine_b200 import abi, scenes  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def run(scene, ticks, mode, substeps):
    w = O.OracleWorld(tuning=abi.default_tuning(substeps=substeps))
    w.load_scene(scene)
    for _ in range(ticks):
        w.tick(mode=mode)
    return w


def main():
    os.makedirs(OUT, exist_ok=True)
    # C1, the reference's own CPU-runnable config: reference order, default 10 substeps, 60 ticks
    w = run(scenes.c1_sandbox_stack(), 60, O.MODE_REFERENCE, 10)
    np.savez_compressed(os.path.join(OUT, "c1_reference_60ticks.npz"), bodies=w.get_bodies(), stats=np.array(list(w.stats().values())))
    # a small mixed scene, one tick, coloured order: bodies + pair units + colours + manifolds
    sc = scenes.parity_mix(120, 3)
    w = run(sc, 1, O.MODE_COLOURED, 2)
    hi, lo, mult, col = w.pair_units()
    np.savez_compressed(os.path.join(OUT, "mix120_coloured_1tick.npz"), bodies=w.get_bodies(), hi=hi, lo=lo, mult=mult, colour=col,
                        manifolds=w.manifolds(), aabbs=w.aabbs())
    # casts on that scene (closed form and BVH walk)
    w = run(sc, 1, O.MODE_REFERENCE, 2)
    rays = np.zeros(64, abi.ray_dtype)
    ang = np.arange(64) * (2 * np.pi / 64) + 0.01
    rays["sx"], rays["sy"], rays["dx"], rays["dy"] = 0.3, 3.0, np.cos(ang), np.sin(ang)
    rays["max_distance"], rays["radius"] = 25.0, np.where(np.arange(64) % 4 == 0, 0.2, 0.0)
    np.savez_compressed(os.path.join(OUT, "mix120_casts.npz"), rays=rays, closed=w.cast_batch(rays, False), bvh=w.cast_batch(rays, True))


if __name__ == "__main__":
    main()

"""Generates tests/golden/ron_*.npz from the reference sandbox's scene files (run in the build container only: the GPU
box has no /root/reference). Each fixture holds the slot-indexed arrays moleengine_b200/ron.py produces for one scene
(derived data: bodies, colliders, constraints, ropes) plus the f64 oracle's body states after 60 ticks in the reference's
order, so that (a) the GPU tests can run the reference's own scenes without the reference tree and (b) a later edit of the
loader or of the oracle cannot silently change them.   python tools/make_ron_fixtures.py"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
from moleengine_b200 import abi, ron  # noqa: E402

SCENES = "/root/reference/examples/sandbox/scenes"
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    for path in sorted(glob.glob(os.path.join(SCENES, "*.ron"))):
        sc = ron.load_scene(path)
        if len(sc["bodies"]) == 0:
            continue                                    # minimal.ron / forest.ron: walls and scenery only
        w = O.OracleWorld(tuning=abi.default_tuning(substeps=10))
        w.load_scene(sc)
        ff = abi.gravity_field(*sc["gravity"])
        for _ in range(60):
            w.tick(mode=O.MODE_REFERENCE, forcefield=ff)
        np.savez_compressed(os.path.join(OUT, f"ron_{sc['name']}.npz"), bodies=sc["bodies"], colliders=sc["colliders"], constraints=sc["constraints"],
                            ropes=sc["ropes"], rope_particles=sc["rope_particles"], gravity=np.asarray(sc["gravity"], np.float64),
                            oracle_bodies_60=w.get_bodies(), oracle_stats_60=np.array(list(w.stats().values())))
        print(sc["name"], len(sc["bodies"]), "bodies")


if __name__ == "__main__":
    main()

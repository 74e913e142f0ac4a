"""Host-side constructors (moleengine_b200/shapes.py) against the oracle's restatement of collider.rs:222-337."""
import numpy as np

from moleengine_b200 import abi, scenes, shapes


def test_mass_properties_match_oracle(oracle):
    rng = np.random.default_rng(1)
    rows = []
    for kind in range(5):
        for _ in range(20):
            rows.append((kind, rng.uniform(0.1, 2.0), rng.uniform(0.1, 2.0), rng.choice([0.0, rng.uniform(0.01, 0.5)])))
    rows = np.array(rows)
    rows[rows[:, 0] == 0, 3] = np.maximum(rows[rows[:, 0] == 0, 3], 0.2)
    out = oracle.geometry(6, np.zeros((len(rows), 2)), rows)
    a = shapes.area(rows[:, 0].astype(np.uint32), rows[:, 1], rows[:, 2], rows[:, 3])
    m = shapes.second_moment_of_area(rows[:, 0].astype(np.uint32), rows[:, 1], rows[:, 2], rows[:, 3])
    assert np.allclose(a, out[:, 0], rtol=1e-13, atol=0) and np.allclose(m, out[:, 1], rtol=1e-13, atol=0)


def test_constructors():
    k, p0, p1, r = shapes.rounded_rect(6.0, 4.0, 1.0)           # collider.rs:90-102
    assert (int(k), float(p0), float(p1), float(r)) == (abi.POLY_RECT, 2.0, 1.0, 1.0)
    k, p0, p1, r = shapes.rounded_rect(1.0, 0.2, 5.0)           # radius reduced until it fits
    assert (float(p0), float(p1), float(r)) == (0.05, 0.05, 0.1)
    k, p0, p1, r = shapes.capsule(4.0, 1.0)
    assert (int(k), float(p0), float(r)) == (abi.POLY_SEGMENT, 2.0, 1.0)
    s, b = shapes.rotor(np.pi / 2)
    assert abs(s - np.cos(np.pi / 4)) < 1e-15 and abs(b + np.sin(np.pi / 4)) < 1e-15
    com, area, mom = shapes.compound_setup([shapes.circle(1.0), shapes.circle(1.0)], [(1.0, 0.0), (-1.0, 0.0)])
    assert np.allclose(com, 0) and abs(area - 2 * np.pi) < 1e-12 and abs(mom - 2 * (np.pi / 2 + np.pi)) < 1e-12
    # sic: centre of mass is divided by the collider count, not the total area (compound_collider.rs:17-23)
    com, _, _ = shapes.compound_setup([shapes.circle(1.0), shapes.circle(2.0)], [(1.0, 0.0), (3.0, 0.0)])
    assert abs(com[0] - (np.pi * 1.0 + 4 * np.pi * 3.0) / 2.0) < 1e-12


def test_scenes_are_deterministic_and_well_formed():
    for fn in (scenes.c1_sandbox_stack, lambda: scenes.c2_circles(20, 10), lambda: scenes.c3_mixed(30, 16), lambda: scenes.c4_ropes(5, 16, 10),
               lambda: scenes.c5_worlds(3, 4), lambda: scenes.parity_mix(60, 2)):
        a, b = fn(), fn()
        for key in ("bodies", "colliders", "constraints", "ropes", "rope_particles"):
            assert a[key].tobytes() == b[key].tobytes()
        nb = len(a["bodies"])
        cb = a["colliders"]["body"]
        assert cb.max() < nb and cb.min() >= -1
        assert np.all(a["bodies"]["flags"] & abi.BODY_ALIVE) and np.all(a["colliders"]["flags"] & abi.COLL_ALIVE)
        assert np.all(np.isfinite(a["bodies"]["inv_mass"])) and np.all(a["colliders"]["layer"] < 64)
    c1 = scenes.c1_sandbox_stack()
    assert len(c1["bodies"]) == 120 and len(c1["colliders"]) == 121 and c1["substeps"] == 10

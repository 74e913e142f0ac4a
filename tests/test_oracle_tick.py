"""CPU tests of the oracle's tick: frozen regression fixtures, internal consistency between its two Gauss-Seidel
orders, conservation laws, colouring validity and sleeping. (The reference has no tests above its collision
primitives — SURVEY.md §4 — so these and the closed forms of test_oracle_physics_kat.py are the only pins of the tick-level restatement.)"""
import os

import numpy as np

from moleengine_b200 import abi, scenes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _world(O, sc, substeps, f32=False, **kw):
    w = O.OracleWorld(use_f32=f32, tuning=abi.default_tuning(substeps=substeps, **kw))
    w.load_scene(sc)
    return w


def test_golden_c1_reference_order(oracle):
    g = np.load(os.path.join(GOLD, "c1_reference_60ticks.npz"))
    w = _world(oracle, scenes.c1_sandbox_stack(), 10)
    for _ in range(60):
        w.tick(mode=oracle.MODE_REFERENCE)
    assert np.allclose(w.get_bodies(), g["bodies"], rtol=0, atol=1e-9)
    # the stack is standing: boxes stayed in their columns, circles landed on top
    b = w.get_bodies()
    assert np.abs(b[:100, 0] - scenes.c1_sandbox_stack()["bodies"]["x"][:100]).max() < 0.05
    assert b[:100, 1].min() > -4.45 and b[100:, 1].min() > 5.4


def test_golden_mix_coloured(oracle):
    g = np.load(os.path.join(GOLD, "mix120_coloured_1tick.npz"))
    w = _world(oracle, scenes.parity_mix(120, 3), 2)
    w.tick(mode=oracle.MODE_COLOURED)
    hi, lo, mult, col = w.pair_units()
    assert np.array_equal(hi, g["hi"]) and np.array_equal(lo, g["lo"]) and np.array_equal(mult, g["mult"]) and np.array_equal(col, g["colour"])
    assert np.allclose(w.get_bodies(), g["bodies"], rtol=0, atol=1e-10)
    assert np.allclose(w.manifolds(), g["manifolds"], rtol=0, atol=1e-10)
    assert np.allclose(w.aabbs(), g["aabbs"], rtol=0, atol=0)


def test_golden_casts_and_bvh_equals_closed_form(oracle):
    g = np.load(os.path.join(GOLD, "mix120_casts.npz"))
    w = _world(oracle, scenes.parity_mix(120, 3), 2)
    w.tick(mode=oracle.MODE_REFERENCE)
    closed, bvh = w.cast_batch(g["rays"], False), w.cast_batch(g["rays"], True)
    assert np.allclose(closed, g["closed"], atol=1e-10) and np.allclose(bvh, g["bvh"], atol=1e-10)
    # the best-first BVH walk of the reference and the closed form agree
    assert np.array_equal(closed[:, 0], bvh[:, 0]) and np.allclose(closed[:, 1:], bvh[:, 1:], atol=1e-12)
    assert (closed[:, 0] >= 0).sum() > 40


def test_bvh_pairs_equal_closed_form_and_f32_equals_f64(oracle):
    sc = scenes.parity_mix(300, 9)
    sets = []
    for f32 in (False, True):
        for mode in (oracle.MODE_REFERENCE, oracle.MODE_COLOURED):       # BVH emission vs sort-and-sweep closed form
            w = _world(oracle, sc, 1, f32=f32, flags=abi.TUNE_NO_SLEEPING)
            w.tick(mode=mode)
            p = w.pairs().astype(np.uint64)
            sets.append(np.sort(p[:, 0] << np.uint64(32) | p[:, 1]))
    assert len(sets[0]) > 300
    assert all(np.array_equal(sets[0], s) for s in sets[1:])


def test_colouring_is_valid(oracle):
    sc = scenes.parity_mix(300, 9)
    w = _world(oracle, sc, 1)
    w.tick(mode=oracle.MODE_COLOURED)
    hi, lo, mult, col = w.pair_units()
    cb = sc["colliders"]["body"]
    sees = (sc["bodies"]["flags"] & (abi.BODY_MASS_FINITE | abi.BODY_INERTIA_FINITE)) != 0
    both = (cb[hi] >= 0) & (cb[lo] >= 0)
    assert np.array_equal(mult, np.where(both, 2, 1))                    # duplication quirk, physics.rs:641,651
    for c in np.unique(col):
        sel = col == c
        bs = np.concatenate([cb[hi[sel]], cb[lo[sel]]])
        bs = bs[bs >= 0]
        bs = bs[sees[bs]]
        assert len(bs) == len(np.unique(bs)), f"colour {c} touches a dynamic body twice"
    idx, kmult, kcol = w.constraint_units()
    k = sc["constraints"]
    for c in np.unique(kcol):
        sel = idx[kcol == c]
        bs = np.concatenate([k["owner"][sel], k["target"][sel]])
        bs = bs[bs >= 0]
        bs = bs[sees[bs]]
        assert len(bs) == len(np.unique(bs))


def test_momentum_is_conserved_without_external_forces(oracle):
    sc = scenes.parity_mix(150, 4, n_ropes=0, with_constraints=False, with_kinematic=False)
    sc["colliders"] = sc["colliders"][4:].copy()       # drop the statics: a closed system of dynamic bodies
    ff = np.zeros(1, abi.forcefield_dtype)
    m = 1.0 / sc["bodies"]["inv_mass"].astype(np.float64)
    for mode in (oracle.MODE_REFERENCE, oracle.MODE_COLOURED):
        w = _world(oracle, sc, 4, flags=abi.TUNE_NO_SLEEPING)
        b0 = sc["bodies"]
        p0 = np.array([(m * b0["vx"]).sum(), (m * b0["vy"]).sum()])
        for _ in range(20):
            w.tick(forcefield=ff, mode=mode)
        b = w.get_bodies()
        p1 = np.array([(m * b[:, 4]).sum(), (m * b[:, 5]).sum()])
        assert np.abs(p1 - p0).max() < 1e-9 * max(1.0, np.abs(m).sum())


def test_two_orders_agree_statistically(oracle):
    """DFS order (reference) and colour-major order (GPU) are different Gauss-Seidel sweeps of the same system:
    trajectories differ, aggregate behaviour must not."""
    sc = scenes.c2_circles(24, 12)
    res = []
    for mode in (oracle.MODE_REFERENCE, oracle.MODE_COLOURED):
        w = _world(oracle, sc, 4, flags=abi.TUNE_NO_SLEEPING)
        for _ in range(300):
            w.tick(mode=mode)
        b = w.get_bodies()
        res.append((b[:, 1].mean(), np.abs(b[:, 4:6]).mean(), b[:, 1].min()))
    (h0, v0, m0), (h1, v1, m1) = res
    assert abs(h0 - h1) < 0.05 * abs(h0) + 0.05      # pile height
    assert v0 < 0.1 and v1 < 0.1                      # both came to rest
    assert m0 > -0.2 and m1 > -0.2                    # nothing fell through the floor


def test_sleeping_matches_reference_rules(oracle):
    sc = scenes.c1_sandbox_stack()
    w = _world(oracle, sc, 10)
    slept_at = None
    for t in range(400):
        w.tick(mode=oracle.MODE_REFERENCE)
        st = w.stats()
        if st["islands"] == 0 and slept_at is None:
            slept_at = t
    assert slept_at is not None and slept_at > 10          # the whole stack is one island and eventually sleeps (physics.rs:711-741)
    frozen = w.get_bodies().copy()
    w.tick(mode=oracle.MODE_REFERENCE)
    assert np.array_equal(frozen, w.get_bodies())          # sleeping bodies are not touched (physics.rs:1150-1157)
    # disabling sleeping keeps every island awake
    w2 = _world(oracle, sc, 10, flags=abi.TUNE_NO_SLEEPING)
    for _ in range(60):
        w2.tick(mode=oracle.MODE_REFERENCE)
    assert w2.stats()["islands"] >= 1


def test_near_hint_is_a_priority_not_a_filter(oracle):
    """include/sfg_hash.h: the `near` bit goes in front of the colouring priority. So (a) the colouring stays a proper colouring
    of ALL pairs whatever the hints are, (b) near pairs outrank the rest: their colours are exactly the ones they would get if the
    other pairs did not exist, and (c) the hints the oracle computes itself (sfg_near_hint) are a mix of both values here and
    feeding them back as overrides changes nothing."""
    sc = scenes.parity_mix(300, 9)
    w = _world(oracle, sc, 1)
    w.tick(mode=oracle.MODE_COLOURED)
    pairs = w.pairs()
    hi0, lo0, _, col0 = w.pair_units()
    assert len(pairs) > 300
    rng = np.random.default_rng(5)
    hints = (rng.random(len(pairs)) < 0.4).astype(np.uint8)
    # (c): the oracle's own hints, fed back as overrides, change nothing
    own = w.pair_unit_hints()
    assert 0 < sum(own.values()) < len(own)
    w1 = _world(oracle, sc, 1)
    w1.set_pair_hints(pairs, np.array([own[(int(a), int(b))] for a, b in pairs], np.uint8))
    w1.tick(mode=oracle.MODE_COLOURED)
    hi1, lo1, _, col1 = w1.pair_units()
    key = lambda h, l: h.astype(np.uint64) << np.uint64(32) | l.astype(np.uint64)   # noqa: E731
    assert dict(zip(key(hi0, lo0).tolist(), col0.tolist())) == dict(zip(key(hi1, lo1).tolist(), col1.tolist()))
    # mixed hints
    w2 = _world(oracle, sc, 1)
    w2.set_pair_hints(pairs, hints)
    w2.tick(mode=oracle.MODE_COLOURED)
    hi2, lo2, _, col2 = w2.pair_units()
    assert sorted(key(hi2, lo2).tolist()) == sorted(key(hi0, lo0).tolist())          # (a) every pair is still there ...
    cb = sc["colliders"]["body"]
    sees = (sc["bodies"]["flags"] & (abi.BODY_MASS_FINITE | abi.BODY_INERTIA_FINITE)) != 0
    for c in np.unique(col2):                                                         # ... and no colour touches a dynamic body twice
        sel = col2 == c
        bs = np.concatenate([cb[hi2[sel]], cb[lo2[sel]]])
        bs = bs[bs >= 0]
        bs = bs[sees[bs]]
        assert len(bs) == len(np.unique(bs))
    # (b): the near pairs alone, as an external pair list
    near = dict(zip(key(pairs[:, 0], pairs[:, 1]).tolist(), hints.tolist()))
    near_pairs = pairs[hints == 1]
    w3 = _world(oracle, sc, 1)
    w3.set_external_pairs(near_pairs, np.ones(len(near_pairs), np.uint8))
    w3.tick(mode=oracle.MODE_COLOURED)
    hi3, lo3, _, col3 = w3.pair_units()
    alone = dict(zip(key(hi3, lo3).tolist(), col3.tolist()))
    mixed = dict(zip(key(hi2, lo2).tolist(), col2.tolist()))
    assert len(alone) > 50
    for k, c in alone.items():
        assert near[k] == 1 and mixed[k] == c

"""First GPU parity checks: broadphase pair set + colouring bit-exact, one-substep and one-tick parity with the oracle."""
import numpy as np
import pytest

from moleengine_b200 import GpuWorld, abi, scenes

pytestmark = pytest.mark.gpu


def _state(gb):
    return np.stack([gb["x"], gb["y"], gb["rot_s"], gb["rot_b"], gb["vx"], gb["vy"], gb["w"]], 1).astype(np.float64)


def _pairset(p):
    p = np.asarray(p, np.uint64).reshape(-1, 2)
    return np.sort(p[:, 0] << np.uint64(32) | p[:, 1])


@pytest.mark.parametrize("scene_fn", [scenes.c1_sandbox_stack, lambda: scenes.parity_mix(400, 7), lambda: scenes.c2_circles(60, 30),
                                      lambda: scenes.star(150), lambda: scenes.star(700, big_r=40.0, small_r=0.2), lambda: scenes.c5_worlds(6, 16)])
def test_pairs_and_colours_bit_exact(oracle, scene_fn):
    sc = scene_fn()
    tun = abi.default_tuning(substeps=1)
    g = GpuWorld(tuning=tun)
    g.load_scene(sc)
    g.tick()
    o32 = oracle.OracleWorld(use_f32=True, tuning=tun)
    o32.load_scene(sc)
    o32.tick(mode=oracle.MODE_COLOURED)     # pairs from the oracle's own BVH, `near` bits from its own evaluation of include/sfg_hash.h
    # AABBs bit-exact against the f32 oracle
    ga = g.debug_aabbs(len(sc["colliders"]))
    oa = o32.aabbs().astype(np.float32)
    alive = (sc["colliders"]["flags"] & abi.COLL_ALIVE) != 0
    assert np.array_equal(ga[alive].view(np.uint32), oa[alive].view(np.uint32))
    # candidate pair set bit-exact
    gp, op = _pairset(g.debug_pairs()), _pairset(o32.pairs())
    assert len(gp) > 0 and np.array_equal(gp, op)
    # the `near` bit of every pair's priority: computed independently on both sides from the same f32 tick-start state
    gh = dict(zip(map(tuple, g.debug_pairs().tolist()), g.debug_pair_hints().tolist()))
    oh = o32.pair_unit_hints()
    assert gh == oh
    print("pairs", len(gh), "near", sum(gh.values()))
    # colouring bit-exact
    hi, lo, mult, col = g.debug_contact_units()
    ohi, olo, omult, ocol = o32.pair_units()
    gk = dict(zip(zip(hi.tolist(), lo.tolist()), zip(mult.tolist(), col.tolist())))
    ok = dict(zip(zip(ohi.tolist(), olo.tolist()), zip(omult.tolist(), ocol.tolist())))
    assert gk == ok
    g.close()


@pytest.mark.parametrize("substeps", [1, 4])
@pytest.mark.parametrize("separate", [True, False])
def test_tick_matches_oracle_replay(oracle, substeps, separate):
    """north_star: the CPU reference is fed the same coloured order; manifolds 1e-5, positions / velocities 1e-4 relative.
    One substep is the strict gate; 4 substeps let f32 rounding compound, so the tolerance there is 1e-3."""
    import parity
    sc = scenes.parity_mix(400, 7)
    tun = abi.default_tuning(substeps=substeps, flags=(abi.TUNE_SEPARATE_LAUNCHES if separate else 0))
    g = GpuWorld(tuning=tun)
    g.load_scene(sc)
    g.tick()
    o = oracle.OracleWorld(tuning=tun)
    o.load_scene(sc)
    o.set_external_pairs(g.debug_pairs())
    o.tick(mode=oracle.MODE_COLOURED)
    tol = 1e-4 if substeps == 1 else 1e-3
    r = parity.compare_tick(g, o, sc, tol_state=tol, hops=3 if substeps == 1 else 8)
    print(r)
    assert r["flips_ok"] and r["non_tie_mismatches"] == 0
    assert r["ties"] <= 0.01 * r["entries"]
    if substeps == 1:
        assert r["manifolds_ok"]
        assert r["bad_outside_influence"] == 0
        assert r["influenced"] <= 0.3 * r["bodies"]
    else:
        # 4 substeps: f32 rounding is amplified by every contact it passes through (the dynamics are chaotic);
        # the strict gate is the single substep above, here almost every body must still agree to 1e-3
        assert r["vel_err_median"] < 1e-4 and r["pos_err_median"] < 1e-5
        assert r["bad_bodies"] <= 0.25 * r["bodies"]
        assert r["momentum_rel_err"] < 1e-3 and r["kinetic_rel_err"] < 1e-2
    g.close()


@pytest.mark.parametrize("name", ["c3_tenth", "c5_64"])
def test_pairs_hints_colours_islands_bit_exact_at_scale(oracle, name):
    """1/10-scale C3 (100 000 bodies) and 64-world C5, at the first tick and again after the device has run the scene for a
    while (bodies landing / piled up): tick-time boxes, candidate pairs, `near` bits, colours (f32 oracle, coloured order) and
    island ids (f32 oracle, reference order, its own BVH and DFS) all bit-exact from the SAME f32 state."""
    sc = scenes.c3_mixed(400, 250) if name == "c3_tenth" else scenes.c5_worlds(64, 16)
    tun = abi.default_tuning(substeps=4)
    tun["fall_asleep_frames"] = 1000000           # the island pass runs, nothing falls asleep
    g = GpuWorld(tuning=tun)
    g.load_scene(sc)
    alive = (sc["colliders"]["flags"] & abi.COLL_ALIVE) != 0
    for warm in (0, 30):
        for _ in range(warm):
            g.tick()
        state = g.get_bodies()
        state["flags"] &= 15
        g.tick()
        # coloured order, f32: boxes, pairs, hints, colours
        o = oracle.OracleWorld(use_f32=True, tuning=tun)
        o.load_scene(sc)
        o.set_bodies(state)
        o.tick(mode=oracle.MODE_COLOURED)
        ga, oa = g.debug_aabbs(len(sc["colliders"])), o.aabbs().astype(np.float32)
        assert np.array_equal(ga[alive].view(np.uint32), oa[alive].view(np.uint32))
        gp = g.debug_pairs()
        assert len(gp) > 50000 and np.array_equal(_pairset(gp), _pairset(o.pairs()))
        gh = dict(zip(map(tuple, gp.tolist()), g.debug_pair_hints().tolist()))
        assert gh == o.pair_unit_hints()
        hi, lo, mult, col = g.debug_contact_units()
        ohi, olo, omult, ocol = o.pair_units()
        key = lambda h, l: (h.astype(np.uint64) << np.uint64(32)) | l.astype(np.uint64)   # noqa: E731
        go, oo = np.argsort(key(hi, lo)), np.argsort(key(ohi, olo))
        assert np.array_equal(key(hi, lo)[go], key(ohi, olo)[oo]) and np.array_equal(mult[go], omult[oo]) and np.array_equal(col[go], ocol[oo])
        # reference order, f32: islands from the oracle's own BVH pairs and DFS
        o = oracle.OracleWorld(use_f32=True, tuning=tun)
        o.load_scene(sc)
        o.set_bodies(state)
        o.tick(mode=oracle.MODE_REFERENCE)
        table = lambda fb, es, bc, fl: sorted(zip(np.asarray(fb, np.uint64).tolist(), np.asarray(es, np.uint64).tolist(), np.asarray(bc, np.uint64).tolist()))  # noqa: E731
        got, want = table(*g.debug_islands()), table(*o.islands())
        assert len(got) > 0 and got == want
        print(name, "after", warm, "ticks: pairs", len(gp), "near", sum(gh.values()), "colours", int(col.max()) + 1, "islands", len(got))
    g.close()


def test_custom_layer_mask_filters_the_pairs(oracle):
    """CollisionMaskMatrix (collision.rs:113-183): ignore(1, 2), ignore_within(3), ignore_all(5) on a crowded scene whose colliders sit
    on layers 0..5; the candidate pairs, colours and the state after the tick follow the oracle given the same matrix, and differ
    from the default matrix's."""
    import parity
    sc = scenes.parity_mix(400, 11)
    rng = np.random.default_rng(5)
    keep = sc["colliders"]["layer"] == abi.ROPE_LAYER
    sc["colliders"]["layer"] = np.where(keep, sc["colliders"]["layer"], rng.integers(0, 6, len(sc["colliders"])).astype(np.uint32))
    m = abi.default_mask_matrix()

    def ignore(a, b):
        m[a] &= ~(np.uint64(1) << np.uint64(b)); m[b] &= ~(np.uint64(1) << np.uint64(a))
    ignore(1, 2)
    ignore(3, 3)
    for other in range(64):
        ignore(5, other)
    tun = abi.default_tuning(substeps=1)
    g = GpuWorld(tuning=tun, mask_matrix=m)
    g.load_scene(sc)
    g.tick()
    o32 = oracle.OracleWorld(use_f32=True, tuning=tun, mask_matrix=m)
    o32.load_scene(sc)
    o32.tick(mode=oracle.MODE_COLOURED)
    gp, op = _pairset(g.debug_pairs()), _pairset(o32.pairs())
    assert len(gp) > 0 and np.array_equal(gp, op)
    lay = sc["colliders"]["layer"]
    la, lb = lay[g.debug_pairs()[:, 0]], lay[g.debug_pairs()[:, 1]]
    assert not np.any(((la == 1) & (lb == 2)) | ((la == 2) & (lb == 1)) | ((la == 3) & (lb == 3)) | (la == 5) | (lb == 5))
    hi, lo, mult, col = g.debug_contact_units()
    ohi, olo, omult, ocol = o32.pair_units()
    assert dict(zip(zip(hi.tolist(), lo.tolist()), col.tolist())) == dict(zip(zip(ohi.tolist(), olo.tolist()), ocol.tolist()))
    # the default matrix finds more pairs on the same scene, and changing the matrix on a live world takes effect on the next tick
    g.set_mask_matrix(abi.default_mask_matrix())
    g.tick()
    assert len(g.debug_pairs()) > len(gp)
    g.close()
    # state parity of the tick under the custom matrix (f64 oracle replaying the device's order)
    g = GpuWorld(tuning=tun, mask_matrix=m)
    g.load_scene(sc)
    o = oracle.OracleWorld(tuning=tun, mask_matrix=m)
    o.load_scene(sc)
    g.tick()
    o.set_external_pairs(g.debug_pairs())
    o.tick(mode=oracle.MODE_COLOURED)
    r = parity.compare_tick(g, o, sc, tol_state=1e-4, hops=3)
    print(r)
    assert r["flips_ok"] and r["non_tie_mismatches"] == 0 and r["manifolds_ok"] and r["bad_outside_influence"] == 0
    g.close()

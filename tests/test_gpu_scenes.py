"""GPU parity on the other BASELINE configs (scaled down so the oracle finishes in seconds): ropes + attachment
constraints (C4), batched independent worlds with casts (C5), the circle pile (C2), plus edge cases."""
import numpy as np
import pytest

import parity
from moleengine_b200 import GpuWorld, SfgError, abi, scenes

pytestmark = pytest.mark.gpu


def _replay(oracle, sc, substeps=1, ff=None, f32=False):
    tun = abi.default_tuning(substeps=substeps)
    g = GpuWorld(tuning=tun)
    g.load_scene(sc)
    g.tick(forcefield=ff)
    o = oracle.OracleWorld(tuning=tun, use_f32=f32)
    o.load_scene(sc)
    o.set_external_pairs(g.debug_pairs())
    o.tick(mode=oracle.MODE_COLOURED, forcefield=ff)
    return g, o


def _check(r, ties_frac=0.02):
    print(r)
    assert r["flips_ok"] and r["non_tie_mismatches"] == 0
    assert r["ties"] <= ties_frac * max(r["entries"], 50)
    assert r["manifolds_ok"] and r["manifolds_compared"] > 0
    assert r["bad_outside_influence"] == 0
    assert r["influenced"] < r["bodies"]


def test_c4_ropes_and_attachments(oracle):
    sc = scenes.c4_ropes(n_ropes=24, particles_per_rope=32, n_free_blocks=60)
    # let the free blocks start right on the ropes so rope-body contacts (normal fix-up) are exercised
    g, o = _replay(oracle, sc)
    r = parity.compare_tick(g, o, sc)
    _check(r)
    # constraint colouring bit-exact
    idx, mult, col = g.debug_constraint_units()
    oi, om, oc = o.constraint_units()
    assert dict(zip(idx.tolist(), zip(mult.tolist(), col.tolist()))) == dict(zip(oi.tolist(), zip(om.tolist(), oc.tolist())))
    g.close()


def test_c4_ropes_settled_contacts(oracle):
    """Run the GPU for a while so blocks are lying on the ropes, then compare ONE tick from that state."""
    sc = scenes.c4_ropes(n_ropes=12, particles_per_rope=32, n_free_blocks=48)
    tun = abi.default_tuning(substeps=4)
    g = GpuWorld(tuning=tun)
    g.load_scene(sc)
    for _ in range(45):
        g.tick()
    sc2 = dict(sc)
    sc2["bodies"] = g.get_bodies()
    g.close()
    g, o = _replay(oracle, sc2)
    r = parity.compare_tick(g, o, sc2)
    _check(r, ties_frac=0.05)
    m = g.debug_manifolds()
    assert (m[:, 2] > 0).sum() > 10      # there really are contacts
    g.close()


def test_c5_worlds_and_casts(oracle):
    nw, bps = 6, 6
    sc = scenes.c5_worlds(n_worlds=nw, bodies_per_side=bps)
    g, o = _replay(oracle, sc)
    # no pair may cross worlds
    p = g.debug_pairs()
    dom = sc["colliders"]["domain"]
    assert len(p) > 0 and np.all(dom[p[:, 0]] == dom[p[:, 1]])
    _check(parity.compare_tick(g, o, sc))
    # batched raycasts + spherecasts over the tick-time grid vs the oracle's closed form (f32 oracle: same arithmetic)
    rays = scenes.c5_rays(sc["bodies"], nw, bps * bps, rays_per_world=64)
    hits = g.cast_batch(rays)
    o32 = oracle.OracleWorld(use_f32=True, tuning=abi.default_tuning(substeps=1))
    sc_after = dict(sc)
    sc_after["bodies"] = g.get_bodies()
    # oracle needs the same tick-time boxes: tick it from the same start, then give it the GPU's post-tick bodies
    o32.load_scene(sc)
    o32.set_external_pairs(g.debug_pairs())
    o32.tick(mode=oracle.MODE_COLOURED)
    o32.set_bodies(sc_after["bodies"])
    want = o32.cast_batch(rays, use_bvh=False)
    assert (hits["collider"] >= 0).sum() > len(rays) // 2
    same = hits["collider"] == want[:, 0].astype(np.int32)
    # a different collider is acceptable only on an exact-distance tie
    assert np.all(same | (np.abs(hits["t"] - want[:, 1]) <= 1e-4 * np.maximum(1.0, want[:, 1])))
    ok = same & (hits["collider"] >= 0)
    assert np.allclose(hits["t"][ok], want[ok, 1], rtol=1e-4, atol=1e-5)
    assert np.allclose(hits["px"][ok], want[ok, 2], rtol=1e-4, atol=1e-4) and np.allclose(hits["py"][ok], want[ok, 3], rtol=1e-4, atol=1e-4)
    assert np.allclose(hits["nx"][ok], want[ok, 4], atol=1e-3) and np.allclose(hits["ny"][ok], want[ok, 5], atol=1e-3)
    # point queries
    pts = np.stack([sc_after["bodies"]["x"][:200], sc_after["bodies"]["y"][:200]], 1).astype(np.float32)
    doms = (np.arange(200) // (bps * bps)).astype(np.uint32)
    out, cnt = g.query_point_batch(pts, doms, max_hits=8)
    for i in range(0, 200, 7):
        want_c = sorted(o32.query_point(float(pts[i, 0]), float(pts[i, 1]), int(doms[i])).tolist())
        got_c = sorted(out[i, : cnt[i]].tolist())
        assert got_c == want_c
    # shape queries (physics.rs:1215-1245): every polygon kind, with and without rounding, a layer mask that hides layer 0
    nq = 60
    qs = np.zeros(nq, abi.shape_query_dtype)
    rng = np.random.default_rng(5)
    pick = rng.integers(0, len(sc_after["bodies"]), nq)
    ang = rng.uniform(0, 6.28, nq)
    qs["x"] = sc_after["bodies"]["x"][pick] + rng.uniform(-0.4, 0.4, nq); qs["y"] = sc_after["bodies"]["y"][pick] + rng.uniform(-0.4, 0.4, nq)
    qs["rot_s"] = np.cos(ang / 2); qs["rot_b"] = -np.sin(ang / 2)
    qs["kind"] = np.arange(nq) % 5
    qs["p0"] = rng.uniform(0.2, 0.9, nq); qs["p1"] = rng.uniform(0.2, 0.9, nq)
    qs["circle_r"] = np.where(np.arange(nq) % 2 == 0, 0.15, 0.0) + np.where(qs["kind"] == 0, 0.3, 0.0)
    qs["layer_mask"] = np.where(np.arange(nq) % 7 == 3, np.uint64(0xFFFFFFFFFFFFFFFE), np.uint64(0xFFFFFFFFFFFFFFFF))
    qs["domain"] = (pick // (bps * bps)).astype(np.uint32)
    out, cnt = g.query_shape_batch(qs, max_hits=32)
    total = 0
    for i in range(nq):
        want_c = sorted(o32.query_shape(qs[i]).tolist())
        got_c = sorted(out[i, : cnt[i]].tolist())
        assert got_c == want_c, (i, got_c, want_c)
        total += len(want_c)
    assert total > nq // 2
    g.close()


def test_c2_pile_after_settling(oracle):
    sc = scenes.c2_circles(40, 25)
    tun = abi.default_tuning(substeps=4)
    g = GpuWorld(tuning=tun)
    g.load_scene(sc)
    for _ in range(90):
        g.tick()
    sc2 = dict(sc)
    sc2["bodies"] = g.get_bodies()
    st = g.stats()
    assert st["n_pairs"] > len(sc["bodies"])        # a real pile: more candidate pairs than bodies
    g.close()
    g, o = _replay(oracle, sc2)
    # A resting pile of circles is the worst case for zero-depth ties (circle corrections are exact, so every second
    # visit of a pair is decided by the last rounding bit) and Gauss-Seidel carries each tie through the pile within
    # the substep; so besides the per-entry checks this scene is judged on aggregate statistics.
    r = parity.compare_tick(g, o, sc2, hops=16)
    print(r)
    assert r["ties"] <= 0.02 * r["entries"]
    assert r["flips_ok"]
    assert r["non_tie_mismatches"] + r["mismatches_inside_influence"] <= r["ties"] + r["flips"]
    # Gauss-Seidel carries a tie further than any fixed number of hops in a resting pile (an error of 1e-4 m/s is all it
    # takes to be counted), and which entries tie depends on the last bit of 90 ticks of f32 history (it changes with every
    # rebuild of the kernel): some bodies beyond the 16-hop influence set exceed the tolerance (4 to 16 of 1000 in the builds
    # of this round), never a visible fraction of the pile. The aggregate statistics below are the gate for this scene.
    assert r["bad_outside_influence"] <= 0.05 * r["bodies"]
    assert r["pos_err_median"] < 1e-6 and r["vel_err_median"] < 1e-3
    assert r["momentum_rel_err"] < 2e-2 and r["kinetic_rel_err"] < 5e-2
    g.close()


def test_point_gravity_and_time_scale(oracle):
    sc = scenes.parity_mix(150, 3, n_ropes=1)
    ff = np.zeros(1, abi.forcefield_dtype)
    ff["n_terms"] = 2
    ff["terms"][0, 0] = (abi.FF_POINT_GRAVITY, 0.0, 3.0, 40.0, 0.5)
    ff["terms"][0, 1] = (abi.FF_GRAVITY, 0.5, -1.0, 0.0, 0.0)
    tun = abi.default_tuning(substeps=3)
    g = GpuWorld(tuning=tun)
    g.load_scene(sc)
    g.tick(1.0 / 60.0, 0.5, ff)           # physics.rs:411-415: ceil(0.5 * 3) = 2 substeps of dt = 0.5 / 60 / 2
    assert g.stats()["substeps"] == 2
    o = oracle.OracleWorld(tuning=tun)
    o.load_scene(sc)
    o.set_external_pairs(g.debug_pairs())
    o.tick(1.0 / 60.0, 0.5, ff, mode=oracle.MODE_COLOURED)
    r = parity.compare_tick(g, o, sc, tol_state=5e-4, hops=6)
    print(r)
    assert r["flips_ok"] and r["non_tie_mismatches"] == 0 and r["bad_outside_influence"] == 0
    g.close()


def test_edge_cases():
    g = GpuWorld()
    # empty world ticks and answers queries
    g.set_bodies(np.zeros(0, abi.body_dtype))
    g.set_colliders(np.zeros(0, abi.collider_dtype))
    g.tick()
    rays = np.zeros(3, abi.ray_dtype)
    rays["dx"], rays["max_distance"] = 1.0, 10.0
    assert np.all(g.cast_batch(rays)["collider"] == -1)
    # a single static collider, a dead slot and a NaN pose: no pairs, no crash
    sc = scenes.parity_mix(20, 1, n_ropes=0, with_constraints=False)
    sc["bodies"]["x"][3] = np.nan
    sc["colliders"]["flags"][6] = 0
    g.load_scene(sc)
    g.tick()
    b = g.get_bodies()
    finite = np.isfinite(b["x"])
    assert finite.sum() == len(b) - 1 and not finite[3]
    # invalid input is rejected with an error code, never a crash
    bad = sc["colliders"].copy()
    bad["kind"][0] = 9
    with pytest.raises(SfgError):
        g.set_colliders(bad)
    bad = sc["colliders"].copy()
    bad["body"][5] = 10_000
    with pytest.raises(SfgError):
        g.set_colliders(bad)
    one = np.zeros(1, abi.rope_dtype)
    one["particle_count"] = 1
    with pytest.raises(SfgError):       # reference: expect("Rope only had one particle") (solver.rs:120)
        g.set_ropes(one, np.zeros(1, np.int32))
    g.close()


def test_contacts_published(oracle):
    sc = scenes.parity_mix(200, 5)
    g, o = _replay(oracle, sc, substeps=2)
    gc = g.get_contacts()
    oc, on = o.contacts()
    gk = np.sort(gc["collider0"].astype(np.uint64) << np.uint64(32) | gc["collider1"].astype(np.uint64))
    ok = np.sort(oc[:, 0].astype(np.uint64) << np.uint64(32) | oc[:, 1].astype(np.uint64))
    # ContactInfo multiset (pairs appear once per entry, physics.rs:1097-1122); allow the zero-depth ties
    common = len(np.intersect1d(gk, ok))
    assert common >= 0.97 * max(len(np.unique(gk)), len(np.unique(ok)))
    g.close()


def test_pose_sync_roundtrip():
    """sfg_set_poses / sfg_get_poses (the HecsSyncManager path) leave velocities alone and round-trip exactly."""
    sc = scenes.parity_mix(80, 2, n_ropes=0, with_constraints=False)
    g = GpuWorld(tuning=abi.default_tuning(substeps=2))
    g.load_scene(sc)
    g.tick()
    b0 = g.get_bodies()
    p = g.get_poses()
    assert np.array_equal(p[:, 0], b0["x"]) and np.array_equal(p[:, 3], b0["rot_b"])
    p2 = p.copy()
    p2[:10, 0] += 0.25
    g.set_poses(p2)
    b1 = g.get_bodies()
    assert np.array_equal(b1["x"], p2[:, 0]) and np.array_equal(b1["vx"], b0["vx"]) and np.array_equal(b1["w"], b0["w"])
    g.tick()
    g.close()


def test_pair_buffer_overflow_reruns_the_search(oracle):
    """The candidate-pair buffer is sized from the previous tick; a world that suddenly contracts (poses written from the host,
    as user code may do between ticks) must overflow it once, re-run the search and still emit the exact pair set."""
    sc = scenes.c3_mixed(100, 30, seed=9)
    tun = abi.default_tuning(substeps=1)
    g = GpuWorld(tuning=tun)
    b0 = sc["bodies"].copy()
    b0["x"] *= 4.0; b0["y"] = 1.0 + (b0["y"] - b0["y"].min()) * 4.0       # sparse start: hardly any candidate pair
    sc0 = dict(sc); sc0["bodies"] = b0
    g.load_scene(sc0)
    g.tick()
    n_before = int(g.stats()["n_pairs"])
    b = sc["bodies"].copy()
    b["x"] *= 0.15; b["y"] = 1.0 + (b["y"] - b["y"].min()) * 0.15          # 44x the design density: boxes overlap everywhere
    g.set_bodies(b)
    g.set_colliders(sc["colliders"])
    g.tick()
    n_after = int(g.stats()["n_pairs"])
    # the buffer held max(1.25 * previous + 4096, 6 * colliders) pairs (+50 % allocation headroom)
    assert n_after > 1.5 * max(1.25 * n_before + 4096, 6 * len(sc["colliders"])) + 64, (n_before, n_after)
    sc2 = dict(sc); sc2["bodies"] = b
    o32 = oracle.OracleWorld(use_f32=True, tuning=tun)
    o32.load_scene(sc2)
    o32.tick(mode=oracle.MODE_COLOURED)
    key = lambda p: np.sort(np.asarray(p, np.uint64).reshape(-1, 2)[:, 0] << np.uint64(32) | np.asarray(p, np.uint64).reshape(-1, 2)[:, 1])  # noqa: E731
    assert np.array_equal(key(g.debug_pairs()), key(o32.pairs()))
    g.close()


@pytest.mark.parametrize("offset", [(1000.0, -640.0), (-8000.0, 3000.0)])
def test_far_from_the_origin(oracle, offset):
    """SURVEY.md "hard parts" 1: f32 positions lose 1/16 mm per ulp at |x| = 1000 m and v = (x - x_old) / dt would amplify that
    240x; the device keeps (x_old, dx) and pair-local frames, so a scene translated to the far corner of the 1 M-body world
    (and well beyond it) still meets the one-substep gates against the f64 oracle replaying the same order."""
    import parity
    sc = scenes.parity_mix(300, 5, offset=offset)
    tun = abi.default_tuning(substeps=1)
    g = GpuWorld(tuning=tun)
    g.load_scene(sc)
    g.tick()
    o = oracle.OracleWorld(tuning=tun)
    o.load_scene(sc)
    o.set_external_pairs(g.debug_pairs())
    o.tick(mode=oracle.MODE_COLOURED)
    r = parity.compare_tick(g, o, sc, tol_state=1e-4, hops=3)
    print(offset, r)
    assert r["flips_ok"] and r["non_tie_mismatches"] == 0 and r["bad_outside_influence"] == 0
    assert r["manifolds_ok"] and r["manifolds_compared"] > 50
    assert r["vel_err_clean"] < 1e-4
    g.close()


def test_restitution_added_by_a_partial_update(oracle):
    """Old velocities / external accelerations are only kept for worlds that have a collider with restitution (their one reader,
    solver.rs:555-576). The flag must come up when restitution arrives through sfg_update_colliders, not only through a full upload."""
    sc = scenes.parity_mix(300, 5)
    plain = sc["colliders"].copy()
    plain["restitution"] = 0.0
    assert (sc["colliders"]["restitution"] != 0).any()
    tun = abi.default_tuning(substeps=1)
    g = GpuWorld(tuning=tun)
    g.set_bodies(sc["bodies"])
    g.set_colliders(plain)                       # nothing bounces: the world starts without old velocities
    g.set_constraints(sc["constraints"])
    g.set_ropes(sc["ropes"], sc["rope_particles"])
    g.update_colliders(0, sc["colliders"])       # now some colliders do
    g.tick()
    o = oracle.OracleWorld(tuning=tun)
    o.load_scene(sc)
    o.set_external_pairs(g.debug_pairs())
    o.tick(mode=oracle.MODE_COLOURED)
    r = parity.compare_tick(g, o, sc)
    _check(r)
    g.close()


def test_near_hint_orders_the_colours():
    """The `near` bit in front of the colouring priority (include/sfg_hash.h): one hint per pair, every touching pair carries it in
    a falling scene, and near pairs sit in lower colours than the rest (that is what makes the high colours cheap)."""
    sc = scenes.c3_mixed(60, 40)
    g = GpuWorld(tuning=abi.default_tuning(substeps=4))
    g.load_scene(sc)
    for _ in range(12):
        g.tick()
    p = g.debug_pairs().astype(np.uint64)
    h = g.debug_pair_hints()
    assert h.dtype == np.uint8 and len(h) == len(p) and set(np.unique(h).tolist()) <= {0, 1}
    assert 0 < h.sum() < len(h)
    hi, lo, mult, col = g.debug_contact_units()
    key_p = p[:, 0] << np.uint64(32) | p[:, 1]
    order = np.argsort(key_p)
    key_u = hi.astype(np.uint64) << np.uint64(32) | lo.astype(np.uint64)
    near = h[order][np.searchsorted(key_p[order], key_u)].astype(bool)
    m = g.debug_manifolds()
    n = len(hi)
    touching = (m[:n, 2] > 0) | (m[n:2 * n, 2] > 0)
    assert touching.sum() > 20
    assert near[touching].all()                                  # the hint missed no contact
    assert col[near].mean() < col[~near].mean()
    assert col[near].max() < col.max()                           # the top colours hold far pairs only
    g.close()


def test_tick_poses_equals_set_tick_get():
    """sfg_tick_poses (Game::physics_tick in one call, game.rs:297-303) = sfg_set_poses + sfg_tick + sfg_get_poses, bit for bit,
    including the pose read-back that runs on the second stream next to the end of the tick."""
    sc = scenes.parity_mix(300, 3)
    tun = abi.default_tuning(substeps=4)
    a, b = GpuWorld(tuning=tun), GpuWorld(tuning=tun)
    a.load_scene(sc); b.load_scene(sc)
    poses = a.get_poses()
    for t in range(6):
        nudged = poses.copy()
        nudged[::17, 0] += 0.01 * (t + 1)                 # the game moves some entities between ticks
        a.set_poses(nudged); a.tick(); want = a.get_poses()
        got = b.tick_poses(nudged)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), t
        assert np.array_equal(b.get_bodies().view(np.uint8), a.get_bodies().view(np.uint8))
        poses = want
    keep = b.tick_poses(None)                              # no upload: the device's own poses are stepped
    a.tick()
    assert np.array_equal(keep.view(np.uint32), a.get_poses().view(np.uint32))
    a.close(); b.close()


def test_time_scale_and_slot_space_edge_cases():
    """ABI validation (round-1 advisor findings): a negative / NaN time scale gives zero substeps like Rust's saturating float-to-int
    cast (physics.rs:411-415) instead of undefined behaviour, an absurd one is refused, zero substeps in the tuning are refused at
    creation too, and shrinking the body slot space invalidates the colliders that may point beyond it."""
    sc = scenes.parity_mix(60, 2)
    g = GpuWorld(tuning=abi.default_tuning(substeps=4))
    g.load_scene(sc)
    before = g.get_bodies()
    for ts in (-1.0, float("nan"), 0.0):
        g.tick(1.0 / 60.0, ts)
        assert int(g.stats()["substeps"]) == 0
        after = g.get_bodies()
        for f in ("x", "y", "vx", "vy", "w"):
            assert np.array_equal(before[f], after[f]), (ts, f)          # no substep ran: nothing moved
    g.tick(1.0 / 60.0, 0.3)
    assert int(g.stats()["substeps"]) == 2                               # ceil(0.3 * 4)
    with pytest.raises(SfgError):
        g.tick(1.0 / 60.0, 1e9)
    with pytest.raises(SfgError):
        GpuWorld(tuning=abi.default_tuning(substeps=0))
    # a smaller slot space: the colliders uploaded before have to be uploaded again
    g.set_bodies(sc["bodies"][:10].copy())
    g.tick()
    assert int(g.stats()["n_colliders"]) == 0 and int(g.stats()["n_pairs"]) == 0
    g.close()

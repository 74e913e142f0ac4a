"""Rope topology edits on the device (rope.rs:147-165 cut_after, :88-110 extend_line, :226-286
remove_dead_particles): sfg_rope_cut_after / sfg_rope_extend / sfg_rope_remove_dead_particles change the library's rope table and
a few device words; the particle array is not re-uploaded and the solver's rope units are rebuilt on the device. Checked against
(a) the same edit done on the host followed by a full sfg_set_ropes — the two worlds must stay BIT-IDENTICAL, (b) the f64 oracle
replaying one tick after the edit, (c) a plain restatement of the reference's splitting rule for random sets of dead particles."""
import numpy as np
import pytest

import parity
from moleengine_b200 import GpuWorld, abi, scenes

pytestmark = pytest.mark.gpu
FIELDS = ("x", "y", "rot_s", "rot_b", "vx", "vy", "w")


def _scene():
    return scenes.c4_ropes(n_ropes=6, particles_per_rope=24, n_free_blocks=18)


def _same(a, b):
    ga, gb = a.get_bodies(), b.get_bodies()
    return all(np.array_equal(ga[f], gb[f]) for f in FIELDS)


def _host_ropes(sc):
    return [list(sc["rope_particles"][r["first_particle"]: r["first_particle"] + r["particle_count"]]) for r in sc["ropes"]], [r.copy() for r in sc["ropes"]]


def _upload(g, lists, params):
    ropes = np.zeros(len(lists), abi.rope_dtype)
    flat = []
    for i, (l, p) in enumerate(zip(lists, params)):
        ropes[i] = p
        ropes[i]["first_particle"], ropes[i]["particle_count"] = len(flat), len(l)
        flat.extend(l)
    g.set_ropes(ropes, np.array(flat, np.int32))


def test_cut_mid_run_equals_host_reupload_and_the_oracle(oracle):
    sc = _scene()
    tun = abi.default_tuning(substeps=4)
    a, b = GpuWorld(tuning=tun), GpuWorld(tuning=tun)
    a.load_scene(sc); b.load_scene(sc)
    for _ in range(20):
        a.tick(); b.tick()
    assert _same(a, b)
    lists, params = _host_ropes(sc)
    # cut rope 2 after its 10th particle and rope 4 after its LAST particle (which pops that particle: rope.rs:154-156)
    new = a.rope_cut_after(2, 9)
    assert new == len(lists)
    assert a.rope_cut_after(4, len(lists[4]) - 1) is None
    lists.append(lists[2][10:]); params.append(params[2].copy()); lists[2] = lists[2][:10]
    lists[4] = lists[4][:-1]
    _upload(b, lists, params)
    ra, pa = a.get_ropes()
    assert [list(pa[r["first_particle"]: r["first_particle"] + r["particle_count"]]) for r in ra] == [list(map(int, l)) for l in lists]
    # one tick after the edit against the f64 oracle (coloured order), then a longer run against the re-uploaded world
    state = a.get_bodies()
    a.tick(); b.tick()
    assert _same(a, b), "device-side cut and host re-upload diverged"
    sc2 = dict(sc)
    sc2["bodies"] = state.copy(); sc2["bodies"]["flags"] &= 15
    sc2["ropes"], sc2["rope_particles"] = ra, pa
    g1 = GpuWorld(tuning=abi.default_tuning(substeps=1)); g1.load_scene(sc2); g1.tick()
    o = oracle.OracleWorld(tuning=abi.default_tuning(substeps=1)); o.load_scene(sc2)
    o.set_external_pairs(g1.debug_pairs()); o.tick(mode=oracle.MODE_COLOURED)
    r = parity.compare_tick(g1, o, sc2)
    print(r)
    assert r["flips_ok"] and r["non_tie_mismatches"] == 0 and r["bad_outside_influence"] == 0
    g1.close()
    for _ in range(40):
        a.tick(); b.tick()
    assert _same(a, b)
    assert np.isfinite(parity.body_state(a.get_bodies())).all()
    assert int(a.stats()["n_ropes"]) == len(lists)
    a.close(); b.close()


def test_extend_appends_particles_without_reuploading_the_rope():
    sc = _scene()
    tun = abi.default_tuning(substeps=4)
    a, b = GpuWorld(tuning=tun), GpuWorld(tuning=tun)
    a.load_scene(sc); b.load_scene(sc)
    for _ in range(10):
        a.tick(); b.tick()
    # Rope::extend_line, six times on ropes that are NOT the last range of the particle array (each has to move to its end, and the
    # array outgrows its allocation on the way): two new particles after the rope's last one every time
    lists, params = _host_ropes(sc)
    gb = a.get_bodies()
    nb0 = len(sc["bodies"])
    from moleengine_b200 import shapes
    targets = [1, 2, 3, 0, 4, 1]
    new_b = np.zeros(2 * len(targets), abi.body_dtype)
    px, py = [], []
    tails = {}
    for k, r in enumerate(targets):
        last = tails.get(r, lists[r][-1])
        lx, ly = (float(gb["x"][last]), float(gb["y"][last])) if last < nb0 else (px[last - nb0], py[last - nb0])
        px += [lx + 0.1, lx + 0.2]; py += [ly, ly]
        tails[r] = nb0 + 2 * k + 1
    shapes.fill_bodies_particle(new_b, 0.02, np.array(px), np.array(py))
    new_c = np.zeros(len(new_b), abi.collider_dtype)
    shapes.fill_colliders(new_c, shapes.circle(np.full(len(new_b), 0.06)), body=nb0 + np.arange(len(new_b)), layer=abi.ROPE_LAYER, static_friction=None, dynamic_friction=1.5)
    for w in (a, b):
        cur = w.get_bodies()
        cur["flags"] &= 15
        w.set_bodies(np.concatenate([cur, new_b]))            # the slot space GROWS: ropes and constraints stay
        w.set_colliders(np.concatenate([sc["colliders"], new_c]))
    for k, r in enumerate(targets):
        slots = [nb0 + 2 * k, nb0 + 2 * k + 1]
        a.rope_extend(r, slots)
        lists[r] = lists[r] + slots
    b.set_constraints(sc["constraints"])
    _upload(b, lists, params)
    ra, pa = a.get_ropes()
    assert [list(pa[r["first_particle"]: r["first_particle"] + r["particle_count"]]) for r in ra] == [list(map(int, l)) for l in lists]
    for _ in range(30):
        a.tick(); b.tick()
    assert _same(a, b), "device-side extend and host re-upload diverged"
    gb = a.get_bodies()
    for chain in lists:
        d = np.hypot(np.diff(gb["x"][chain].astype(np.float64)), np.diff(gb["y"][chain].astype(np.float64)))
        assert np.abs(d - 0.1).max() < 0.03, "an extended rope is not a chain"
    a.close(); b.close()


def _reference_split(lists, alive):
    """RopeSet::remove_dead_particles (rope.rs:226-286), restated plainly: every rope is cut at its dead particles; the first
    surviving piece (possibly empty) stays in the rope, later pieces become new ropes appended in order; empty ropes are deleted."""
    out, fresh = [], []
    for l in lists:
        pieces, cur, seen_dead = [], [], False
        first = None
        for p in l:
            if alive[p]:
                cur.append(p)
            else:
                if first is None:
                    first = cur
                elif cur:
                    pieces.append(cur)
                cur = []
        if first is None:
            first = cur
        elif cur:
            pieces.append(cur)
        out.append(first)
        fresh.extend(pieces)
    return [l for l in out + fresh if l]


def test_dead_particle_split_follows_the_reference_rule():
    sc = _scene()
    rng = np.random.default_rng(3)
    for trial in range(4):
        g = GpuWorld(tuning=abi.default_tuning(substeps=2))
        g.load_scene(sc)
        g.tick()
        lists, params = _host_ropes(sc)
        bodies = g.get_bodies()
        bodies["flags"] &= 15
        alive = np.ones(len(bodies), bool)
        part = sc["rope_particles"]
        kill = rng.choice(part, size=[1, 3, 12, 40][trial], replace=False)
        # always include a run of consecutive dead particles, a rope's first particle and another rope's last one
        kill = np.unique(np.concatenate([kill, lists[0][5:8], lists[1][:1], lists[3][-1:]]))
        alive[kill] = False
        bodies["flags"][kill] = 0
        g.set_bodies(bodies)
        want = _reference_split(lists, alive)
        n = g.rope_remove_dead_particles()
        ra, pa = g.get_ropes()
        got = [list(map(int, pa[r["first_particle"]: r["first_particle"] + r["particle_count"]])) for r in ra]
        assert n == len(want) and got == [list(map(int, l)) for l in want], trial
        if all(len(l) >= 2 for l in want):
            # dead particles' colliders die with their bodies (entity_set.rs:168-176 on the host side)
            colls = sc["colliders"].copy()
            dead_c = np.isin(colls["body"], kill)
            colls["flags"][dead_c] = 0
            g.set_colliders(colls)
            g.set_constraints(sc["constraints"][~(np.isin(sc["constraints"]["owner"], kill))])
            for _ in range(5):
                g.tick()
            assert np.isfinite(parity.body_state(g.get_bodies())[alive]).all()
        else:
            with pytest.raises(Exception, match="Rope only had one particle"):      # solver.rs:120 panics
                g.tick()
        assert g.rope_remove_dead_particles() == len(want)                           # nothing more to do: a four-byte read-back
        g.close()


def test_indexed_pose_sync_equals_the_full_one():
    sc = scenes.parity_mix(200, 5)
    g = GpuWorld(tuning=abi.default_tuning(substeps=2))
    g.load_scene(sc)
    g.tick()
    full = g.get_poses()
    rng = np.random.default_rng(1)
    slots = rng.choice(len(full), size=57, replace=False).astype(np.uint32)
    assert np.array_equal(g.get_poses_indexed(slots), full[slots])
    new = full[slots].copy()
    new[:, 0] += 0.25
    g.set_poses_indexed(slots, new)
    after = g.get_poses()
    want = full.copy()
    want[slots] = new
    assert np.array_equal(after, want)
    with pytest.raises(Exception):
        g.set_poses_indexed(np.array([len(full)], np.uint32), new[:1])
    g.close()

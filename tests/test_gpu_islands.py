"""Islands and sleeping on the device against the oracle's restatement of
physics.rs:598-741 / 1097-1144.

IslandId {first_body, edge_sum} is integer work on the candidate-pair set, so it is compared bit-exactly (the oracle is
fed the GPU's pair set, which test_gpu_first.py already pins against the oracle's own broadphase). Falling asleep is a
threshold on velocities, so the long run compares WHEN bodies fall asleep (within a few ticks) and checks the exact
properties the reference guarantees: sleeping bodies do not move at all, their contacts stay published, and a poke
wakes exactly the poked island."""
import numpy as np
import pytest

from moleengine_b200 import GpuWorld, abi, scenes

pytestmark = pytest.mark.gpu


def _island_table(fb, es, bc, fl, mask):
    rows = sorted(zip(np.asarray(fb, np.uint64).tolist(), np.asarray(es, np.uint64).tolist(), np.asarray(bc, np.uint64).tolist(),
                      (np.asarray(fl, np.uint64) & np.uint64(mask)).tolist()))
    return rows


@pytest.mark.parametrize("scene_fn", [scenes.c1_sandbox_stack, lambda: scenes.parity_mix(400, 7), lambda: scenes.c2_circles(60, 30),
                                      lambda: scenes.c4_ropes(n_ropes=12, particles_per_rope=16, n_free_blocks=30)])
def test_island_ids_bit_exact(oracle, scene_fn):
    sc = scene_fn()
    tun = abi.default_tuning(substeps=2)
    g = GpuWorld(tuning=tun)
    g.load_scene(sc)
    o = oracle.OracleWorld(tuning=tun)
    o.load_scene(sc)
    for _ in range(3):
        g.tick()
        gb = g.get_bodies()
        gb["flags"] &= 15
        o.set_bodies(gb)                       # same state in, so the graph is the only thing compared
        o.set_external_pairs(g.debug_pairs())
        o.tick(mode=oracle.MODE_REFERENCE)
        got = _island_table(*g.debug_islands(), mask=1)
        want = _island_table(*o.islands(), mask=1)
        assert len(got) > 0 and got == want
        assert int(g.stats()["n_islands"]) == len(want)
    g.close()


def _stack_scene():
    sc = scenes.c1_sandbox_stack()
    return sc


def test_stack_falls_asleep_like_the_reference(oracle):
    sc = _stack_scene()
    tun = abi.default_tuning(substeps=10)
    g = GpuWorld(tuning=tun)
    g.load_scene(sc)
    o = oracle.OracleWorld(tuning=tun)
    o.load_scene(sc)
    nb = len(sc["bodies"])
    first_sleep_g = first_sleep_o = None
    g_asleep_ticks, o_asleep_ticks = [], []
    prev = None
    for t in range(160):
        g.tick()
        o.tick(mode=oracle.MODE_REFERENCE)
        gb = g.get_bodies()
        asleep = (gb["flags"] & abi.BODY_ASLEEP) != 0
        n_g = int(asleep.sum())
        ofb, oes, obc, ofl = o.islands()
        n_o = int(obc[(ofl & np.uint64(8)) != 0].sum())
        g_asleep_ticks.append(n_g); o_asleep_ticks.append(n_o)
        assert int(g.stats()["n_sleeping_bodies"]) == n_g
        if prev is not None and prev_asleep.any():
            both = asleep & prev_asleep
            # a sleeping body is not touched at all: not even gravity (physics.rs:1150-1157)
            for f in ("x", "y", "rot_s", "rot_b", "vx", "vy", "w"):
                assert np.array_equal(gb[f][both], prev[f][both])
        prev, prev_asleep = gb, asleep
        if first_sleep_g is None and n_g:
            first_sleep_g = t
        if first_sleep_o is None and n_o:
            first_sleep_o = t
    print("first asleep tick: gpu", first_sleep_g, "oracle", first_sleep_o, "final asleep bodies", g_asleep_ticks[-1], o_asleep_ticks[-1])
    assert first_sleep_g is not None and first_sleep_o is not None, "the stack never came to rest"
    # the stack is a chaotic system near rest; the f32 colour-ordered path and the f64 DFS-ordered reference settle within
    # a second of each other
    assert abs(first_sleep_g - first_sleep_o) <= 60
    assert g_asleep_ticks[-1] > 0.5 * nb and o_asleep_ticks[-1] > 0.5 * nb
    # contacts of sleeping islands stay published (physics.rs:1097-1105)
    assert len(g.get_contacts()) > 0
    g.close()


def test_poke_wakes_only_the_poked_island(oracle):
    """Two separate piles; after both sleep, give one body a velocity: its island is solved again, the other stays asleep."""
    sc = scenes.c1_sandbox_stack()
    b = sc["bodies"].copy()
    c = sc["colliders"].copy()
    nb, nc = len(b), len(c)
    b2, c2 = b.copy(), c.copy()
    b2["x"] += 200.0
    c2["body"] = np.where(c2["body"] >= 0, c2["body"] + nb, -1)
    static = c2["body"] < 0
    c2["x"][static] += 200.0
    sc2 = dict(sc)
    sc2["bodies"] = np.concatenate([b, b2]); sc2["colliders"] = np.concatenate([c, c2])
    tun = abi.default_tuning(substeps=10)
    g = GpuWorld(tuning=tun)
    g.load_scene(sc2)
    for t in range(600):
        g.tick()
        gb = g.get_bodies()
        asleep = (gb["flags"] & abi.BODY_ASLEEP) != 0
        if asleep[:nb].sum() > 0.5 * nb and asleep[nb:].sum() > 0.5 * nb:
            break
    assert asleep[:nb].sum() > 0.5 * nb and asleep[nb:].sum() > 0.5 * nb, "piles did not fall asleep"
    n_contacts_before = len(g.get_contacts())
    victim = int(np.flatnonzero(asleep[:nb])[0])
    poke = gb[victim:victim + 1].copy()
    poke["flags"] &= 15
    poke["vy"] = 3.0
    g.update_bodies(victim, poke)
    g.tick()
    gb2 = g.get_bodies()
    asleep2 = (gb2["flags"] & abi.BODY_ASLEEP) != 0
    assert not asleep2[victim]
    assert np.array_equal(asleep2[nb:], asleep[nb:])           # the far pile is untouched
    for f in ("x", "y", "rot_s", "rot_b"):
        assert np.array_equal(gb2[f][nb:][asleep[nb:]], gb[f][nb:][asleep[nb:]])
    assert len(g.get_contacts()) > 0 and n_contacts_before > 0
    g.close()


def test_no_sleeping_flag_skips_the_pass(oracle):
    sc = scenes.c1_sandbox_stack()
    g = GpuWorld(tuning=abi.default_tuning(substeps=4, flags=abi.TUNE_NO_SLEEPING))
    g.load_scene(sc)
    for _ in range(3):
        g.tick()
    assert int(g.stats()["n_islands"]) == 0
    fb, es, bc, fl = g.debug_islands()
    assert len(fb) == 0
    assert len(g.get_contacts()) > 0
    g.close()


def test_switching_sleeping_off_wakes_everything():
    """A stack that fell asleep must not stay frozen when the island pass is switched off (the pass is the only thing that
    rewrites the per-tick asleep flag): with sleeping off every body is integrated again, and a body pushed through the pose-only
    path moves its neighbours."""
    sc = scenes.c1_sandbox_stack()
    g = GpuWorld(tuning=abi.default_tuning(substeps=10))
    g.load_scene(sc)
    nb = len(sc["bodies"])
    for t in range(400):
        g.tick()
        gb = g.get_bodies()
        asleep = (gb["flags"] & abi.BODY_ASLEEP) != 0
        if asleep.sum() > 0.5 * nb:
            break
    assert asleep.sum() > 0.5 * nb, "the stack did not fall asleep"
    g.set_tuning(abi.default_tuning(substeps=10, flags=abi.TUNE_NO_SLEEPING))
    # lift one sleeping box through the pose-only path (no flags are re-uploaded there)
    cand = np.flatnonzero(asleep)
    victim = int(cand[np.argmax(gb["y"][cand])])          # the highest sleeping body: nothing lies on top of it
    poses = g.get_poses()
    poses[victim, 1] += 1.0
    g.set_poses(poses[victim:victim + 1], first=victim)
    y0 = float(poses[victim, 1])
    for _ in range(5):
        g.tick()
    gb2 = g.get_bodies()
    assert not ((gb2["flags"] & abi.BODY_ASLEEP) != 0).any(), "a body stayed asleep although the island pass is off"
    assert float(gb2["y"][victim]) < y0 - 1e-3, "the lifted box did not fall: it was never integrated"
    assert (gb2["vy"] != 0).sum() > 0.9 * nb, "bodies of the former sleeping island are not integrated"
    g.close()

"""Slab decomposition on the GPU: one world cut into x-slabs, each slab its own SfgWorld, halo bodies
exchanged through the C ABI's pack / unpack between substeps. Here all slabs live in one process on cuda:0 and the
"wire" is a device copy (slab.LocalTransport); bench.py runs the same driver over NCCL with one process per GPU.

Away from a cut a slab sees exactly the pairs the whole world sees, so those bodies must follow the single-world run
closely; next to a cut the Gauss-Seidel order differs, so only statistics are compared there (north_star)."""
import numpy as np
import pytest

from moleengine_b200 import GpuWorld, abi, scenes, slab

pytestmark = pytest.mark.gpu


def _state(b):
    return np.stack([b["x"], b["y"], b["rot_s"], b["rot_b"], b["vx"], b["vy"], b["w"]], 1).astype(np.float64)


def _assemble(worlds, n_bodies):
    """Every body from the rank that owned it in the last tick."""
    out = np.zeros(n_bodies, abi.body_dtype)
    owners = np.zeros(n_bodies, np.int32)
    for w in worlds:
        b = w.get_bodies()
        owned = ((b["flags"] & abi.BODY_ALIVE) != 0) & ((b["flags"] & (abi.BODY_GHOST | abi.BODY_INACTIVE)) == 0)
        out[owned] = b[owned]
        owners += owned
    return out, owners


@pytest.mark.parametrize("n_slabs", [2, 3])
def test_slabs_follow_the_single_world(n_slabs):
    import torch
    sc = scenes.c3_mixed(120, 30, seed=0xC3)
    nb = len(sc["bodies"])
    x0 = float(sc["bodies"]["x"].min()), float(sc["bodies"]["x"].max())
    tun = abi.default_tuning(substeps=4, flags=abi.TUNE_NO_SLEEPING)
    ref = GpuWorld(tuning=tun)
    ref.load_scene(sc)
    transport = slab.LocalTransport()
    worlds, drivers = [], []
    for r in range(n_slabs):
        w = GpuWorld(tuning=tun)
        w.load_scene(sc)
        worlds.append(w)
        drivers.append(slab.SlabDriver(w, r, n_slabs, x0[0], x0[1], halo=4.0, substeps=4, capacity=4096, device=torch.device("cuda", 0),
                                       transport=transport))
    ticks = 40
    for _ in range(ticks):
        ref.tick()
        slab.tick_local(drivers, transport)
    got, owners = _assemble(worlds, nb)
    alive = (sc["bodies"]["flags"] & abi.BODY_ALIVE) != 0
    assert np.array_equal(owners[alive], np.ones(int(alive.sum()), np.int32)), "every body has exactly one owner"
    want = ref.get_bodies()
    sg, sw = _state(got)[alive], _state(want)[alive]
    dpos = np.hypot(sg[:, 0] - sw[:, 0], sg[:, 1] - sw[:, 1])
    cuts = [slab.slab_bounds(r, n_slabs, x0[0], x0[1])[1] for r in range(n_slabs - 1)]
    dist_to_cut = np.min(np.abs(sw[:, 0:1] - np.array(cuts)[None, :]), axis=1)
    far = dist_to_cut > 12.0
    sent = sum(d.sent_records for d in drivers)
    print({"bodies": int(alive.sum()), "far": int(far.sum()), "dpos_far_max": float(dpos[far].max()), "dpos_far_median": float(np.median(dpos[far])),
           "dpos_near_max": float(dpos[~far].max()), "dpos_near_median": float(np.median(dpos[~far])), "records_last_tick": sent,
           "active_per_slab": [int(w.stats()["n_colliders"]) for w in worlds]})
    assert sent > 0, "nothing crossed the cuts"
    # far from a cut: the same pairs, the same colours in the neighbourhood, the same arithmetic
    assert np.median(dpos[far]) < 1e-6 and (dpos[far] < 1e-4).mean() > 0.98
    # near a cut: same physics, different Gauss-Seidel order at the seam
    assert np.median(dpos[~far]) < 1e-3 and dpos.max() < 0.5
    m = 1.0 / np.maximum(want["inv_mass"][alive].astype(np.float64), 1e-9)
    m[want["inv_mass"][alive] == 0] = 0.0
    pg, pw = (m[:, None] * sg[:, 4:6]).sum(0), (m[:, None] * sw[:, 4:6]).sum(0)
    kg, kw = 0.5 * (m * (sg[:, 4] ** 2 + sg[:, 5] ** 2)).sum(), 0.5 * (m * (sw[:, 4] ** 2 + sw[:, 5] ** 2)).sum()
    assert np.abs(pg - pw).max() <= 2e-3 * max(np.abs(pw).max(), 1.0)
    assert abs(kg - kw) <= 2e-3 * kw
    # each slab only carried its share of the world through the broadphase
    assert max(int(w.stats()["n_colliders"]) for w in worlds) < 0.75 * len(sc["colliders"])
    for w in worlds + [ref]:
        w.close()


def test_slab_mode_can_be_switched_off():
    sc = scenes.c3_mixed(40, 10, seed=5)
    tun = abi.default_tuning(substeps=2, flags=abi.TUNE_NO_SLEEPING)
    a, b = GpuWorld(tuning=tun), GpuWorld(tuning=tun)
    a.load_scene(sc); b.load_scene(sc)
    x0 = float(sc["bodies"]["x"].min()), float(sc["bodies"]["x"].max())
    b.slab_configure(x0[0] - 1.0, 0.5 * (x0[0] + x0[1]), 4.0)
    b.slab_begin()
    b.tick()                                              # half the world sits this tick out
    assert int(b.stats()["n_colliders"]) < int(0.8 * len(sc["colliders"]))
    gb = b.get_bodies()
    assert ((gb["flags"] & abi.BODY_INACTIVE) != 0).any()
    b.load_scene(sc)                                      # fresh state, slab off again
    b.slab_configure(0.0, 0.0, 0.0)
    a.tick(); b.tick()
    ga, gb = a.get_bodies(), b.get_bodies()
    for f in ("x", "y", "rot_s", "rot_b", "vx", "vy", "w", "flags"):
        assert np.array_equal(ga[f], gb[f])
    a.close(); b.close()


@pytest.mark.parametrize("n_shards", [2, 3])
def test_island_shards_are_bit_identical_to_one_gpu(n_shards):
    """independent islands of one world, one subset per rank, exact parity retained."""
    import torch
    sc = scenes.c3_mixed(120, 30, seed=0xC3)
    nb = len(sc["bodies"])
    tun = abi.default_tuning(substeps=4)                      # sleeping on: records live on the island's owner
    ref = GpuWorld(tuning=tun)
    ref.load_scene(sc)
    worlds, drivers = [], []
    for r in range(n_shards):
        w = GpuWorld(tuning=tun)
        w.load_scene(sc)
        worlds.append(w)
        drivers.append(slab.IslandShardDriver(w, r, n_shards, capacity=nb, device=torch.device("cuda", 0)))
    for t in range(25):
        ref.tick()
        slab.tick_island_shards_local(drivers)
        want = ref.get_bodies()
        sent = sum(d.sent_records for d in drivers)
        assert sent == int((want["flags"] & abi.BODY_ALIVE != 0).sum()), "every body is solved by exactly one rank"
        for w in worlds:
            got = w.get_bodies()
            for f in ("x", "y", "rot_s", "rot_b", "vx", "vy", "w"):
                assert np.array_equal(got[f], want[f]), (t, f)
    assert int(ref.stats()["n_islands"]) == int(worlds[0].stats()["n_islands"]) > n_shards
    assert min(d.sent_records for d in drivers) > 0
    # the ranks' contact lists partition the single-GPU list
    key = lambda c: sorted(zip(c["collider0"].tolist(), c["collider1"].tolist()))  # noqa: E731
    merged = sorted(sum((key(w.get_contacts()) for w in worlds), []))
    assert merged == key(ref.get_contacts())
    for w in worlds + [ref]:
        w.close()


def test_ropes_and_attachments_across_a_cut():
    """ropes / constraints crossing slabs (rope.rs:61-85 chains, constraint.rs:40-58 attachments). Four of the
    twelve rope-connected block pairs lie ACROSS the cut (particles and both end blocks on different ranks), free blocks are
    dropped on them. Ownership stays unique, the chains stay chains (link lengths, attachment distances) and the straddling ropes
    follow the single-world run statistically."""
    import torch
    sc = scenes.c4_ropes(n_ropes=12, particles_per_rope=32, n_free_blocks=36)
    nb = len(sc["bodies"])
    tun = abi.default_tuning(substeps=4, flags=abi.TUNE_NO_SLEEPING)
    ref = GpuWorld(tuning=tun)
    ref.load_scene(sc)
    transport = slab.LocalTransport()
    worlds, drivers = [], []
    for r in range(2):
        w = GpuWorld(tuning=tun)
        w.load_scene(sc)
        worlds.append(w)
        drivers.append(slab.SlabDriver(w, r, 2, -30.0, 30.0, halo=4.0, substeps=4, capacity=4096, device=torch.device("cuda", 0), transport=transport))
    rp = sc["rope_particles"]
    ropes = sc["ropes"]
    x0 = sc["bodies"]["x"]
    straddle = [i for i, r in enumerate(ropes) if x0[rp[r["first_particle"]]] < 0.0 < x0[rp[r["first_particle"] + r["particle_count"] - 1]]]
    assert len(straddle) == 4

    def chain_metrics(b):
        stretch, attach = 0.0, 0.0
        for i in straddle:
            r = ropes[i]
            p = rp[r["first_particle"]: r["first_particle"] + r["particle_count"]]
            d = np.hypot(np.diff(b["x"][p].astype(np.float64)), np.diff(b["y"][p].astype(np.float64)))
            stretch = max(stretch, float(np.abs(d - 0.1).max()))
        for k in sc["constraints"]:
            o, t = int(k["owner"]), int(k["target"])
            c, s = 1.0 - 2.0 * float(b["rot_b"][t]) ** 2, -2.0 * float(b["rot_s"][t]) * float(b["rot_b"][t])   # rotor -> rotation by theta: cos, sin
            ax = float(b["x"][t]) + c * float(k["off1x"]) - s * float(k["off1y"])
            ay = float(b["y"][t]) + s * float(k["off1x"]) + c * float(k["off1y"])
            attach = max(attach, float(np.hypot(float(b["x"][o]) - ax, float(b["y"][o]) - ay)))
        return stretch, attach

    for t in range(90):
        ref.tick()
        slab.tick_local(drivers, transport)
        if t % 15 == 14:
            got, owners = _assemble(worlds, nb)
            assert np.array_equal(owners, np.ones(nb, np.int32)), "every body has exactly one owner"
            want = ref.get_bodies()
            sg, ag = chain_metrics(got)
            sw, aw = chain_metrics(want)
            part = np.concatenate([rp[ropes[i]["first_particle"]: ropes[i]["first_particle"] + ropes[i]["particle_count"]] for i in straddle])
            dy = abs(float(got["y"][part].mean()) - float(want["y"][part].mean()))
            dmed = float(np.median(np.hypot(got["x"][part] - want["x"][part], got["y"][part] - want["y"][part])))
            print({"tick": t + 1, "stretch": (sg, sw), "attach": (ag, aw), "mean_y_diff": dy, "median_dev": dmed, "records": sum(d.sent_records for d in drivers)})
            assert np.isfinite(_state(got)).all()
            assert sg < max(2.0 * sw, 0.02) and ag < max(2.0 * aw, 0.05), "a chain across the cut came apart"
            assert dy < 0.1
    assert sum(d.sent_records for d in drivers) > 100, "no rope crossed the cut"
    for w in worlds + [ref]:
        w.close()


def test_ray_batches_over_slabs_match_the_single_world():
    """SURVEY.md §8e row 5 (physics.rs:1255-1311): every ray goes to the slabs its segment crosses, the reports are min-t reduced.
    In this scene the slab run is bit-identical to the single world away from the cuts, so the merged first hits must be the
    single-world first hits; long horizontal rays cross all three slabs."""
    import torch
    sc = scenes.c3_mixed(120, 30, seed=0xC3)
    x0 = float(sc["bodies"]["x"].min()), float(sc["bodies"]["x"].max())
    tun = abi.default_tuning(substeps=4, flags=abi.TUNE_NO_SLEEPING)
    ref = GpuWorld(tuning=tun)
    ref.load_scene(sc)
    transport = slab.LocalTransport()
    worlds, drivers = [], []
    for r in range(3):
        w = GpuWorld(tuning=tun)
        w.load_scene(sc)
        worlds.append(w)
        drivers.append(slab.SlabDriver(w, r, 3, x0[0], x0[1], halo=4.0, substeps=4, capacity=4096, device=torch.device("cuda", 0), transport=transport))
    for _ in range(30):
        ref.tick()
        slab.tick_local(drivers, transport)
    rng = np.random.default_rng(11)
    n = 4000
    rays = np.zeros(n, abi.ray_dtype)
    ys = sc["bodies"]["y"]
    rays["sx"] = rng.uniform(x0[0] - 5.0, x0[1] + 5.0, n)
    rays["sy"] = rng.uniform(float(ys.min()) - 2.0, float(ys.max()) + 6.0, n)
    ang = rng.uniform(-np.pi, np.pi, n)
    ang[: n // 4] = rng.choice([0.0, np.pi], n // 4) + rng.uniform(-0.02, 0.02, n // 4)       # long, nearly horizontal: cross every cut
    rays["dx"], rays["dy"] = np.cos(ang), np.sin(ang)
    rays["max_distance"] = np.where(np.arange(n) < n // 4, 400.0, rng.uniform(2.0, 60.0, n))
    rays["radius"] = np.where(rng.random(n) < 0.3, 0.25, 0.0)
    want = ref.cast_batch(rays)
    got, routed = slab.cast_local(drivers, rays)
    crossing = sum(routed) - n
    same_coll = got["collider"] == want["collider"]
    both_hit = same_coll & (want["collider"] >= 0)
    dt = np.abs(got["t"][both_hit] - want["t"][both_hit])
    print({"rays": n, "routed": routed, "extra_routes": crossing, "hits": int((want["collider"] >= 0).sum()), "same_collider": float(same_coll.mean()),
           "t_err_max": float(dt.max(initial=0.0))})
    assert crossing > n // 4, "the long rays did not cross the cuts"
    assert (want["collider"] >= 0).mean() > 0.3
    assert same_coll.mean() > 0.995 and (dt < 1e-4).mean() > 0.995
    # the merge rule on its own: min t, ties to the lowest slot, misses last
    a = np.zeros(3, abi.hit_dtype); b = np.zeros(3, abi.hit_dtype)
    a["collider"], a["t"] = [5, -1, 9], [2.0, 0.0, 1.0]
    b["collider"], b["t"] = [7, 3, 4], [1.5, 9.0, 1.0]
    m = slab.merge_hits([a, b])
    assert m["collider"].tolist() == [7, 3, 4] and m["t"].tolist() == [1.5, 9.0, 1.0]
    for w in worlds + [ref]:
        w.close()


def test_async_exchange_gives_the_same_bits_as_the_synchronous_one():
    """sfg_slab_pack_async / sfg_stream_fence / sfg_slab_unpack_async (no host synchronisation inside the tick) against the
    synchronous pack / unpack: every slab world ends every tick with the same bits."""
    import torch
    sc = scenes.c3_mixed(120, 30, seed=0xC3)
    x0 = float(sc["bodies"]["x"].min()), float(sc["bodies"]["x"].max())
    tun = abi.default_tuning(substeps=4, flags=abi.TUNE_NO_SLEEPING)
    sets = []
    for _ in range(2):
        transport = slab.LocalTransport()
        worlds, drivers = [], []
        for r in range(3):
            w = GpuWorld(tuning=tun)
            w.load_scene(sc)
            worlds.append(w)
            drivers.append(slab.SlabDriver(w, r, 3, x0[0], x0[1], halo=4.0, substeps=4, capacity=4096, device=torch.device("cuda", 0), transport=transport))
        sets.append((worlds, drivers, transport))
    for t in range(25):
        slab.tick_local(sets[0][1], sets[0][2])
        slab.tick_local_async(sets[1][1], sets[1][2])
        torch.cuda.synchronize()
        for wa, wb in zip(sets[0][0], sets[1][0]):
            ga, gb = wa.get_bodies(), wb.get_bodies()
            for f in ("x", "y", "rot_s", "rot_b", "vx", "vy", "w", "flags"):
                assert np.array_equal(ga[f], gb[f]), (t, f)
    for worlds, _, _ in sets:
        for w in worlds:
            w.close()

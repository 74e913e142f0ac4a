"""ORACLE — TEST INFRASTRUCTURE ONLY. Comparison helpers shared by tests/, smoke() and bench.py.

One-tick parity contract (BASELINE.json north_star): the f64 oracle replays the GPU's coloured order on the same
inputs and the same candidate pairs; contact manifolds agree to 1e-5 relative and positions / velocities to 1e-4
relative. One caveat is inherent to the reference algorithm, not to the GPU: every body-body pair is visited twice
per substep (physics.rs:641,651), and the second visit re-detects the pair exactly at the touching state left by
the first, where `depth <= 0` / `dist_sq >= r_sum^2` are decided by the last rounding bit. Such ties may flip between
f32 and f64. compare_tick() therefore (a) requires every manifold-count mismatch to be a zero-depth tie (oracle
lambda_n == 0 up to 1e-6 or GPU lambda_n == 0), (b) requires all bodies farther than `hops` contact-graph steps from a
tied entry to meet the tolerance, and (c) reports how many ties there were.
"""
import numpy as np


def body_state(gb):
    return np.stack([gb["x"], gb["y"], gb["rot_s"], gb["rot_b"], gb["vx"], gb["vy"], gb["w"]], 1).astype(np.float64)


def pair_keys(p):
    p = np.asarray(p, np.uint64).reshape(-1, 2)
    return np.sort(p[:, 0] << np.uint64(32) | p[:, 1])


def match_manifolds(g, o):
    """Align GPU manifolds (entry = copy * n_units + unit) with oracle manifolds (unit-major, copies adjacent).
    Returns (gm, om, unit_hi, unit_lo, copy) aligned row by row."""
    gm = g.debug_manifolds().astype(np.float64)
    hi, lo, mult, _ = g.debug_contact_units()
    n = len(hi)
    om = o.manifolds()
    ohi, olo, omult, _ = o.pair_units()
    start = np.concatenate([[0], np.cumsum(omult.astype(np.int64))])[:-1].astype(np.int64)
    okey = (ohi.astype(np.uint64) << np.uint64(32)) | olo.astype(np.uint64)
    order = np.argsort(okey)
    gkey = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    pos = order[np.searchsorted(okey[order], gkey)].astype(np.int64)
    assert np.array_equal(okey[pos], gkey), "pair units differ between GPU and oracle"
    rows_g, rows_o, uh, ul, cp = [], [], [], [], []
    for m in (0, 1):
        sel = np.nonzero(mult > m)[0].astype(np.int64)
        rows_g.append(gm[sel + m * n]); rows_o.append(om[start[pos[sel]] + m])
        uh.append(hi[sel]); ul.append(lo[sel]); cp.append(np.full(len(sel), m))
    return np.concatenate(rows_g), np.concatenate(rows_o), np.concatenate(uh), np.concatenate(ul), np.concatenate(cp)


def compare_tick(g, o, scene, tol_manifold=1e-5, tol_state=1e-4, hops=3):
    """g: GpuWorld after a tick, o: OracleWorld after the replayed tick. Returns a dict of measured errors."""
    gmr, omr, uh, ul, _ = match_manifolds(g, o)
    cnt_g, cnt_o = gmr[:, 2].astype(int), omr[:, 2].astype(int)
    mism = cnt_g != cnt_o
    # ties: the side that has the contact carries (numerically) zero normal correction
    lam = np.where(cnt_o > 0, np.abs(omr[:, 3]), np.abs(gmr[:, 3]))
    scale = np.maximum(np.abs(omr[:, 3]).max(initial=0.0), 1e-3)
    not_tie = mism & (lam > 1e-4 * scale)
    same = ~mism & (cnt_o > 0)
    # normals + offsets of matching manifolds (count 2 compares both points)
    d = np.abs(gmr[same][:, 4:] - omr[same][:, 4:])
    ref = np.maximum(1.0, np.abs(omr[same][:, 4:]))
    two = (cnt_o[same] == 2)[:, None]
    colmask = np.concatenate([np.ones((same.sum(), 6), bool), np.repeat(two, 4, 1)], 1)
    man_err = float((d / ref * colmask).max(initial=0.0))
    # body states
    got, want = body_state(g.get_bodies()), o.get_bodies()
    ep = np.abs(got[:, :4] - want[:, :4]).max(1) / np.maximum(1.0, np.abs(want[:, :2]).max(1))
    ev = (np.abs(got[:, 4:] - want[:, 4:]) / np.maximum(1.0, np.abs(want[:, 4:]))).max(1)
    bad = (ep > tol_state) | (ev > tol_state)
    # influence set of the tied entries: `hops` steps through entries that touched (either side)
    cb = scene["colliders"]["body"]
    b0, b1 = cb[uh], cb[ul]
    infl = np.zeros(len(got), bool)
    for arr in (b0[mism], b1[mism]):
        infl[arr[arr >= 0]] = True
    touching = (cnt_g > 0) | (cnt_o > 0)
    tb0, tb1 = b0[touching], b1[touching]
    ok = (tb0 >= 0) & (tb1 >= 0)
    tb0, tb1 = tb0[ok], tb1[ok]
    extra = []
    k = scene["constraints"]
    if len(k):
        kk = k[k["target"] >= 0]
        extra.append((kk["owner"], kk["target"]))
    rp = scene["rope_particles"]
    for r in scene["ropes"]:
        p = rp[r["first_particle"]: r["first_particle"] + r["particle_count"]]
        extra.append((p[:-1], p[1:]))
    for _ in range(hops):
        nxt = infl.copy()
        nxt[tb1[infl[tb0]]] = True
        nxt[tb0[infl[tb1]]] = True
        for a, b in extra:
            nxt[b[infl[a]]] = True
            nxt[a[infl[b]]] = True
        infl = nxt
    return {
        "entries": int(len(cnt_o)), "ties": int(mism.sum()), "non_tie_mismatches": int(not_tie.sum()),
        "manifold_err": man_err, "pos_err_clean": float(ep[~infl].max(initial=0.0)), "vel_err_clean": float(ev[~infl].max(initial=0.0)),
        "pos_err_all": float(ep.max(initial=0.0)), "vel_err_all": float(ev.max(initial=0.0)),
        "bad_bodies": int(bad.sum()), "bad_outside_influence": int((bad & ~infl).sum()), "influenced": int(infl.sum()),
        "bodies": int(len(got)),
    }

"""ORACLE — TEST INFRASTRUCTURE ONLY. Comparison helpers shared by tests/, smoke() and bench.py.

One-tick parity contract (BASELINE.json north_star): the f64 oracle replays the GPU's coloured order on the same
inputs and the same candidate pairs; contact manifolds agree to 1e-5 relative and positions / velocities to 1e-4
relative. One caveat is inherent to the reference algorithm, not to the GPU: every body-body pair is visited twice
per substep (physics.rs:641,651), and the second visit re-detects the pair exactly at the touching state left by
the first, where `depth <= 0` / `dist_sq >= r_sum^2` are decided by the last rounding bit. Such ties may flip between
f32 and f64. compare_tick() therefore (a) requires every manifold-count mismatch to be a zero-depth tie (oracle
lambda_n == 0 up to 1e-6 or GPU lambda_n == 0), (b) requires all bodies farther than `hops` contact-graph steps from a
tied entry to meet the tolerance, and (c) reports how many ties there were.

The narrowphase has a second kind of discontinuity: the choice between equally good features — two SAT axes of equal depth
(shape_shape.rs:160-215, strict `<`, first wins), two boundary features equally close to a circle centre
(collider.rs:796-892). A difference in the last bit of the inputs picks the other feature and the manifold changes by O(1)
although both answers are right. Such an entry is a "flip": same contact count on both sides, manifold off by more than
`flip_threshold`, and NOTHING upstream of it wrong — it is the first suspicious entry (in the solver's colour / copy order)
on both of its bodies. Flips seed the influence set like ties do and are counted; callers cap the count (a bug shows up as
many flips, or as mismatches that have no flip upstream of them).
"""
import numpy as np

# Contact manifolds: 1e-5 relative on the same inputs (north_star). On the same inputs means tests/test_gpu_narrowphase.py
# (15 000 random pose pairs through intersection_check alone) and, inside a tick, the visits of the first colours. A visit
# late in the sweep sees poses that already went through up to 2 x colours f32 corrections (absolute error a few 1e-7 m),
# and a rounded-corner contact normalises a vector as short as the overlap (centimetres): 3e-7 / 0.03 = 1e-5 of direction
# error that no narrowphase can avoid. Those visits get three times the tolerance.
TOL_MANIFOLD = 1e-5
TOL_MANIFOLD_LATE = 3e-5


def body_state(gb):
    return np.stack([gb["x"], gb["y"], gb["rot_s"], gb["rot_b"], gb["vx"], gb["vy"], gb["w"]], 1).astype(np.float64)


def pair_keys(p):
    p = np.asarray(p, np.uint64).reshape(-1, 2)
    return np.sort(p[:, 0] << np.uint64(32) | p[:, 1])


def match_manifolds(g, o):
    """Align GPU manifolds (entry = copy * n_units + unit) with oracle manifolds (unit-major, copies adjacent).
    Returns (gm, om, unit_hi, unit_lo, copy) aligned row by row."""
    gm = g.debug_manifolds().astype(np.float64)
    hi, lo, mult, colour = g.debug_contact_units()
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
        uh.append(hi[sel]); ul.append(lo[sel]); cp.append(2 * colour[sel].astype(np.int64) + m)   # order of the visit inside the sweep
    return np.concatenate(rows_g), np.concatenate(rows_o), np.concatenate(uh), np.concatenate(ul), np.concatenate(cp)


def _manifold_diff(gmr, omr, cnt_o):
    """relative difference of normals + offsets per entry (second point only where there is one); NaN on both sides = equal"""
    with np.errstate(invalid="ignore"):
        d = np.abs(gmr[:, 4:] - omr[:, 4:])
    d[np.isnan(gmr[:, 4:]) & np.isnan(omr[:, 4:])] = 0.0
    ref = np.maximum(1.0, np.abs(omr[:, 4:]))
    two = (cnt_o == 2)[:, None]
    colmask = np.concatenate([np.ones((len(cnt_o), 6), bool), np.repeat(two, 4, 1)], 1)
    return np.where(colmask, d / ref, 0.0).max(1, initial=0.0) if len(cnt_o) else np.zeros(0)


def compare_tick(g, o, scene, tol_manifold=1e-5, tol_state=1e-4, hops=3, flip_threshold=1e-3):
    """g: GpuWorld after a tick, o: OracleWorld after the replayed tick. Returns a dict of measured errors."""
    # the order the oracle replayed is the device's only if both computed the same `near` bits (include/sfg_hash.h): checked, not fed
    gh = dict(zip(map(tuple, g.debug_pairs().tolist()), g.debug_pair_hints().tolist()))
    oh = o.pair_unit_hints()
    hint_mismatches = sum(1 for k, v in oh.items() if gh.get(k) != v) + abs(len(gh) - len(oh))
    assert hint_mismatches == 0, f"{hint_mismatches} `near` hints differ between the device and the oracle"
    gmr, omr, uh, ul, visit = match_manifolds(g, o)
    cnt_g, cnt_o = gmr[:, 2].astype(int), omr[:, 2].astype(int)
    mism = cnt_g != cnt_o
    # ties: the side that has the contact carries (numerically) zero normal correction
    lam = np.where(cnt_o > 0, np.abs(omr[:, 3]), np.abs(gmr[:, 3]))
    scale = np.maximum(np.nanmax(np.abs(omr[:, 3]), initial=0.0), 1e-3)
    tie = mism & ~(lam > 1e-4 * scale)
    # flips: same count, manifold grossly different, first suspicious entry on both of its bodies (module docstring)
    cb_ = scene["colliders"]["body"]
    eb0, eb1 = cb_[uh], cb_[ul]
    mdiff = _manifold_diff(gmr, omr, cnt_o)
    with np.errstate(invalid="ignore"):
        gross = ~mism & (cnt_o > 0) & ~(mdiff <= flip_threshold)
    suspicious = gross | mism
    first_bad = np.full(len(scene["bodies"]) + 1, np.iinfo(np.int64).max, np.int64)
    for arr in (eb0, eb1):
        sel = suspicious & (arr >= 0)
        np.minimum.at(first_bad, arr[sel], visit[sel])
    upstream0 = np.where(eb0 >= 0, first_bad[np.maximum(eb0, 0)] < visit, False)
    upstream1 = np.where(eb1 >= 0, first_bad[np.maximum(eb1, 0)] < visit, False)
    flip = gross & ~upstream0 & ~upstream1
    tie = tie | flip          # from here on a flip is treated like a tie: it seeds the influence set
    # body states
    got, want = body_state(g.get_bodies()), o.get_bodies()
    with np.errstate(invalid="ignore"):
        ep = np.abs(got[:, :4] - want[:, :4]).max(1) / np.maximum(1.0, np.abs(want[:, :2]).max(1))
        ev = (np.abs(got[:, 4:] - want[:, 4:]) / np.maximum(1.0, np.abs(want[:, 4:]))).max(1)
    both_nan = np.isnan(got).any(1) & np.isnan(want).any(1)      # NaN poses stay NaN on both sides (shape_shape.rs:209-218)
    ep[both_nan] = 0.0
    ev[both_nan] = 0.0
    bad = ~((ep <= tol_state) & (ev <= tol_state))
    # influence set of the tied entries: `hops` steps through entries that touched (either side), constraints and rope links
    cb = scene["colliders"]["body"]
    b0, b1 = cb[uh], cb[ul]
    infl = np.zeros(len(got), bool)
    for arr in (b0[tie], b1[tie]):
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
    for _ in range(hops if tie.any() else 0):
        nxt = infl.copy()
        nxt[tb1[infl[tb0]]] = True
        nxt[tb0[infl[tb1]]] = True
        for a, b in extra:
            nxt[b[infl[a]]] = True
            nxt[a[infl[b]]] = True
        infl = nxt
    e_infl = np.where(b0 >= 0, infl[np.maximum(b0, 0)], False) | np.where(b1 >= 0, infl[np.maximum(b1, 0)], False)
    # manifolds of entries outside the influence set: counts must match, normals + offsets to tol (count 2 compares both points)
    same = ~mism & (cnt_o > 0) & ~e_infl
    man = mdiff[same]
    man_err = float(man.max(initial=0.0)) if not np.isnan(man).any() else float("nan")
    # the same over the visits of the first four colours only: their inputs carry no propagated rounding yet (see manifolds_ok)
    early = mdiff[same & (visit < 8)]
    man_err_early = float(early.max(initial=0.0)) if not np.isnan(early).any() else float("nan")
    return {
        "entries": int(len(cnt_o)), "ties": int((tie & ~flip).sum()), "flips": int(flip.sum()),
        "flips_ok": bool(flip.sum() <= max(1, len(cnt_o) // 500)),       # equal-feature choices are rare: one, or 0.2 % of the entries
        "non_tie_mismatches": int((mism & ~tie & ~e_infl).sum()),
        "mismatches_inside_influence": int((mism & ~tie & e_infl).sum()),
        "manifold_err": man_err, "manifold_err_early": man_err_early, "manifolds_compared": int(same.sum()),
        "manifolds_ok": bool(man_err_early < TOL_MANIFOLD and man_err < TOL_MANIFOLD_LATE),
        "pos_err_clean": float(ep[~infl].max(initial=0.0)), "vel_err_clean": float(ev[~infl].max(initial=0.0)),
        "pos_err_all": float(np.nanmax(ep, initial=0.0)), "vel_err_all": float(np.nanmax(ev, initial=0.0)),
        "bad_bodies": int(bad.sum()), "bad_outside_influence": int((bad & ~infl).sum()), "influenced": int(infl.sum()),
        "bodies": int(len(got)), "pos_err_median": float(np.nanmedian(ep)), "vel_err_median": float(np.nanmedian(ev)),
        "momentum_rel_err": _momentum_err(scene, got, want), "kinetic_rel_err": _kinetic_err(scene, got, want),
    }


def _mass(scene):
    b = scene["bodies"]
    fin = (b["flags"] & 4) != 0
    with np.errstate(divide="ignore"):
        m = np.where(fin & (b["inv_mass"] > 0), 1.0 / b["inv_mass"].astype(np.float64), 0.0)
        fi = (b["flags"] & 8) != 0
        i = np.where(fi & (b["inv_inertia"] > 0), 1.0 / b["inv_inertia"].astype(np.float64), 0.0)
    return m, i


def _momentum_err(scene, got, want):
    m, _ = _mass(scene)
    ok = ~(np.isnan(got).any(1) | np.isnan(want).any(1))
    pg = (m[ok, None] * got[ok, 4:6]).sum(0)
    pw = (m[ok, None] * want[ok, 4:6]).sum(0)
    scale = (m[ok] * np.linalg.norm(want[ok, 4:6], axis=1)).sum() + 1e-12
    return float(np.linalg.norm(pg - pw) / scale)


def _kinetic_err(scene, got, want):
    m, i = _mass(scene)
    ok = ~(np.isnan(got).any(1) | np.isnan(want).any(1))
    kg = 0.5 * (m[ok] * (got[ok, 4] ** 2 + got[ok, 5] ** 2) + i[ok] * got[ok, 6] ** 2).sum()
    kw = 0.5 * (m[ok] * (want[ok, 4] ** 2 + want[ok, 5] ** 2) + i[ok] * want[ok, 6] ** 2).sum()
    return float(abs(kg - kw) / (abs(kw) + 1e-12))

"""ORACLE — TEST INFRASTRUCTURE ONLY. Long-run statistical parity (BASELINE.json north_star: "over long runs, where the
dynamics are chaotic, energy, momentum and penetration statistics must agree").

A run is sampled every few ticks; each sample is reduced to scalars that do not depend on which body is which:
kinetic and potential energy, linear and angular momentum, and the distribution of penetration depths over all touching
solid pairs (how many contact points are deeper than 0.1 mm; p50, p99, max of all depths — evaluated by ONE routine, OracleWorld.penetrations, for device states and oracle states
alike, depth as solver.rs:425-431 measures it). Two runs of the same scene agree if every sampled scalar agrees within a
tolerance that is stated per quantity in the test; the tests also run the oracle itself in its two orders (reference DFS order
in f64, coloured order in f32) through the same gate, which shows how far apart two CORRECT runs of a chaotic pile end up.
"""
import numpy as np

G = 9.81


def masses(scene):
    b = scene["bodies"]
    with np.errstate(divide="ignore"):
        m = np.where(((b["flags"] & 4) != 0) & (b["inv_mass"] > 0), 1.0 / b["inv_mass"].astype(np.float64), 0.0)
        i = np.where(((b["flags"] & 8) != 0) & (b["inv_inertia"] > 0), 1.0 / b["inv_inertia"].astype(np.float64), 0.0)
    grav = ((b["flags"] & 2) == 0) & (m > 0)
    return m, i, grav


def sample(scene, state7, oracle_mod):
    """state7: (n, 7) x, y, rot_s, rot_b, vx, vy, w of every body slot -> dict of scalars."""
    s = np.asarray(state7, np.float64)
    m, inertia, grav = masses(scene)
    ok = np.isfinite(s).all(1)
    x, y, vx, vy, w = s[:, 0], s[:, 1], s[:, 4], s[:, 5], s[:, 6]
    ke = 0.5 * (m * (vx ** 2 + vy ** 2) + inertia * w ** 2)[ok].sum()
    pe = (m * G * y)[ok & grav].sum()
    px, py = (m * vx)[ok].sum(), (m * vy)[ok].sum()
    ang = (m * (x * vy - y * vx) + inertia * w)[ok].sum()
    b = scene["bodies"].copy()
    for k, f in enumerate(("x", "y", "rot_s", "rot_b", "vx", "vy", "w")):
        b[f] = s[:, k]
    ow = oracle_mod.OracleWorld()
    sc = dict(scene)
    sc["bodies"] = b
    ow.load_scene(sc)
    d = ow.penetrations()
    return {"ke": float(ke), "pe": float(pe), "px": float(px), "py": float(py), "L": float(ang), "nonfinite": int((~ok).sum()),
            "mean_y": float((m * y)[ok].sum() / max(m[ok].sum(), 1e-30)), "mass": float(m.sum()),
            "pen_n": int((d > 1e-4).sum()), "pen_p50": float(np.percentile(d, 50)) if len(d) else 0.0,
            "pen_p99": float(np.percentile(d, 99)) if len(d) else 0.0, "pen_max": float(d.max()) if len(d) else 0.0}


def run_oracle(oracle_mod, scene, tuning, ticks, every, mode, f32=False, threads=1):
    w = oracle_mod.OracleWorld(use_f32=f32, tuning=tuning)
    w.load_scene(scene)
    out = []
    for t in range(1, ticks + 1):
        w.tick(mode=mode, threads=threads)
        if t % every == 0 or t == ticks:
            out.append((t, sample(scene, w.get_bodies(), oracle_mod)))
    return out


def run_gpu(gpu_world, scene, ticks, every, oracle_mod):
    from parity import body_state
    out = []
    for t in range(1, ticks + 1):
        gpu_world.tick()
        if t % every == 0 or t == ticks:
            out.append((t, sample(scene, body_state(gpu_world.get_bodies()), oracle_mod)))
    return out


def compare(a, b, scene):
    """Largest deviation of every scalar over the samples of two runs, in the units the gates are stated in:
    energies relative to the scene's energy scale (mass * g * 1 m — the energy of lifting everything by a metre), momenta relative
    to mass * 1 m/s (angular: * 1 m), mean height in metres, penetration depths in metres, contact-point counts relative."""
    m, _, _ = masses(scene)
    M = max(float(m.sum()), 1e-30)
    r = {"ke": 0.0, "pe": 0.0, "px": 0.0, "py": 0.0, "L": 0.0, "mean_y": 0.0, "pen_p50": 0.0, "pen_p99": 0.0, "pen_max": 0.0, "pen_n": 0.0, "nonfinite": 0}
    xs = np.abs(scene["bodies"]["x"]).max() + 1.0
    for (ta, sa), (tb, sb) in zip(a, b):
        assert ta == tb
        r["ke"] = max(r["ke"], abs(sa["ke"] - sb["ke"]) / (M * G))
        r["pe"] = max(r["pe"], abs(sa["pe"] - sb["pe"]) / (M * G))
        r["px"] = max(r["px"], abs(sa["px"] - sb["px"]) / M)
        r["py"] = max(r["py"], abs(sa["py"] - sb["py"]) / M)
        r["L"] = max(r["L"], abs(sa["L"] - sb["L"]) / (M * xs))
        r["mean_y"] = max(r["mean_y"], abs(sa["mean_y"] - sb["mean_y"]))
        for k in ("pen_p50", "pen_p99", "pen_max"):
            r[k] = max(r[k], abs(sa[k] - sb[k]))
        r["pen_n"] = max(r["pen_n"], abs(sa["pen_n"] - sb["pen_n"]) / max(sb["pen_n"], 50))
        r["nonfinite"] = max(r["nonfinite"], sa["nonfinite"], sb["nonfinite"])
    return r


# ---- scenes and stated tolerances of the long-run gates (tests/test_gpu_longrun.py, tests/test_longrun_cpu.py) -------------
# Tolerances are in compare()'s units. They are about twice the spread measured between two CORRECT runs of the same scene —
# the f64 oracle in the reference's DFS order and in the coloured order — because that spread, not rounding, is what separates
# the device from the reference after hundreds of ticks of a chaotic pile (tests/test_longrun_cpu.py keeps that claim honest).
# Measured on B200 (device vs reference order, round 2): C1 mean height 5e-4 m, momenta < 1e-3, penetration p50 / p99 / max within
# 1.3e-4 / 3.4e-4 / 3.5e-4 m (both runs at rest and asleep); C2 mean height 0.028 m, KE 0.03, penetration p50 / p99 / max within
# 2.0e-3 / 8.0e-3 / 4.0e-2 m, deep-contact count within 28 %; C4 mean height 0.10 m, KE 0.08, penetration within 4e-4 / 8e-3 / 2e-2 m.
def scene_c1():
    """BASELINE config 1 at full spec: 600 ticks of 10 substeps, default tuning (sleeping on)."""
    from moleengine_b200 import scenes
    return scenes.c1_sandbox_stack(), 10, 600


def scene_c2():
    """BASELINE config 2 at 1/55 scale: 1800 circles falling into the container and coming to rest, 300 ticks of 4 substeps."""
    from moleengine_b200 import scenes
    return scenes.c2_circles(60, 30), 4, 300


def scene_c4():
    """BASELINE config 4 at 1/400 scale: 24 ropes of 32 particles between blocks, 60 free blocks dropped on them, 300 ticks of 4
    substeps; a wide static floor is appended (last collider slot) so that nothing falls out of the scene for ever."""
    import numpy as np
    from moleengine_b200 import abi, scenes, shapes
    sc = scenes.c4_ropes(n_ropes=24, particles_per_rope=32, n_free_blocks=60)
    floor = np.zeros(1, abi.collider_dtype)
    shapes.fill_colliders(floor, shapes.rounded_rect(400.0, 1.0, 0.0), x=0.0, y=-7.0)
    sc = dict(sc)
    sc["colliders"] = np.concatenate([sc["colliders"], floor])
    return sc, 4, 300


GATES = {
    "c1": {"mean_y": 0.02, "ke": 0.01, "px": 0.05, "py": 0.05, "L": 0.05, "pen_p50": 3e-4, "pen_p99": 1e-3, "pen_max": 2e-3, "pen_n": 0.3},
    "c2": {"mean_y": 0.08, "ke": 0.08, "px": 0.4, "py": 0.4, "L": 0.3, "pen_p50": 4e-3, "pen_p99": 2e-2, "pen_max": 8e-2, "pen_n": 0.5},
    "c4": {"mean_y": 0.25, "ke": 0.5, "px": 0.6, "py": 1.0, "L": 0.6, "pen_p50": 1.5e-3, "pen_p99": 3e-2, "pen_max": 8e-2, "pen_n": 0.9},
}


def check(dev, gates):
    """-> list of violated gates (empty = pass)."""
    bad = [f"{k}: {dev[k]:.4g} > {tol:.4g}" for k, tol in gates.items() if not (dev[k] <= tol)]
    if dev["nonfinite"]:
        bad.append(f"{dev['nonfinite']} non-finite bodies")
    return bad

"""Closed-form pins of the oracle's tick ABOVE the collision primitives (the reference has no tests there): bodies that touch
nothing follow the reference's own update rule exactly, so their state after n substeps has a closed form.
Rule (solver.rs:55-75, physics.rs:78-83): per substep  v += a dt  (only finite mass without ignores_gravity),  x += v dt,
rotation = from_angle(w dt) * rotation;  afterwards the velocity is re-derived from the pose difference (solver.rs:91-97), which
gives the same v back.  Hence after n substeps:  v_n = v_0 + n a dt,  x_n = x_0 + n v_0 dt + a dt^2 n (n + 1) / 2,  angle_n = angle_0 + n w dt.
Substep count and dt under time scaling: physics.rs:411-415."""
import math

import numpy as np
import pytest

from moleengine_b200 import abi, scenes, shapes

G = float(np.float32(-9.81))      # the force field crosses the ABI as f32


def _free_bodies():
    b = np.zeros(4, abi.body_dtype)
    shapes.fill_bodies_dynamic(b[0:1], 1.0, 0.2, 0.5, -100.0, 50.0, angle=0.3, vx=1.5, vy=2.0, w=0.7)
    shapes.fill_bodies_dynamic(b[1:2], 2.0, 0.4, 0.5, 0.0, 50.0, angle=-1.1, vx=-0.5, vy=0.25, w=-0.4)
    b["flags"][1] |= abi.BODY_IGNORES_GRAVITY
    shapes.fill_bodies_kinematic(b[2:3], 100.0, 50.0, angle=0.0, vx=0.75, vy=-0.5, w=1.3)
    shapes.fill_bodies_particle(b[3:4], 0.3, 200.0, 50.0)
    b["vx"][3], b["vy"][3] = 0.1, 3.0
    c = np.zeros(5, abi.collider_dtype)
    shapes.fill_colliders(c[0:1], shapes.rounded_rect(10.0, 0.2, 0.0), x=0.0, y=-1000.0)          # a floor nobody reaches
    shapes.fill_colliders(c[1:5], shapes.circle(np.full(4, 0.25)), body=np.arange(4))
    return scenes._scene("free_bodies", b, c, 10)


def _expect(sc, n, dt):
    b = sc["bodies"]
    x0, y0, vx0, vy0, w0 = (b[k].astype(np.float64) for k in ("x", "y", "vx", "vy", "w"))
    falls = ((b["flags"] & abi.BODY_MASS_FINITE) != 0) & ((b["flags"] & abi.BODY_IGNORES_GRAVITY) == 0)
    a = np.where(falls, G, 0.0)
    ang0 = -2.0 * np.arctan2(b["rot_b"].astype(np.float64), b["rot_s"].astype(np.float64))
    return (x0 + n * vx0 * dt, y0 + n * vy0 * dt + a * dt * dt * n * (n + 1) / 2.0, vx0, vy0 + n * a * dt, ang0 + n * w0 * dt, w0)


def _check(state, want, tol):
    x, y, vx, vy, ang, w = want
    assert np.abs(state[:, 0] - x).max() < tol and np.abs(state[:, 1] - y).max() < tol
    assert np.abs(state[:, 4] - vx).max() < tol and np.abs(state[:, 5] - vy).max() < tol and np.abs(state[:, 6] - w).max() < tol
    got = -2.0 * np.arctan2(state[:, 3], state[:, 2])            # DRotor2::from_angle: bv.xy = -sin(a / 2) (shapes.rotor)
    d = (got - ang + math.pi) % (2.0 * math.pi) - math.pi          # the rotor covers the angle twice: compare modulo a turn
    assert np.abs(d).max() < tol
    assert np.abs(state[:, 2] ** 2 + state[:, 3] ** 2 - 1.0).max() < max(tol, 1e-7)      # rotors stay unit (as unit as their f32 inputs were)


@pytest.mark.parametrize("mode_name", ["MODE_REFERENCE", "MODE_COLOURED"])
@pytest.mark.parametrize("substeps", [1, 4, 10])
def test_free_bodies_follow_the_closed_form(oracle, mode_name, substeps):
    sc = _free_bodies()
    # f32 inputs are exact in f64; the closed form and the tick then differ only by f64 rounding
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=substeps, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    ticks, frame_dt = 30, 1.0 / 60.0
    for _ in range(ticks):
        w.tick(frame_dt, mode=getattr(oracle, mode_name))
    _check(w.get_bodies(), _expect(sc, ticks * substeps, frame_dt / substeps), 1e-9)


@pytest.mark.parametrize("time_scale,substeps", [(0.5, 10), (0.25, 4), (1.0, 3), (2.0, 4), (0.33, 10)])
def test_time_scale_changes_substep_count_and_dt(oracle, time_scale, substeps):
    """physics.rs:411-415: substeps = ceil(time_scale * tuning.substeps), dt = time_scale * frame_dt / substeps."""
    sc = _free_bodies()
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=substeps, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    ticks, frame_dt = 12, 1.0 / 60.0
    for _ in range(ticks):
        w.tick(frame_dt, time_scale=time_scale)
    n = int(math.ceil(time_scale * substeps))
    _check(w.get_bodies(), _expect(sc, ticks * n, time_scale * frame_dt / n), 1e-9)


def _no_field():
    return np.zeros(1, abi.forcefield_dtype)


@pytest.mark.parametrize("mode_name", ["MODE_REFERENCE", "MODE_COLOURED"])
def test_two_overlapping_circles_are_pushed_apart_by_inverse_mass(oracle, mode_name):
    """solver.rs:428-452 + 496-618 in the one case with a closed form: two circles at rest overlapping by `depth` along x, no
    external field. The normal correction is lambda = -depth / (w0 + w1) along the contact normal, i.e. body i moves by
    depth * w_i / (w0 + w1) away from the other (no lever arm: the contact lies on the line of centres); the velocity the
    correction leaves behind is purely normal and separating, and the velocity step (restitution 0 by default) removes the
    relative normal velocity again — with masses m0, m1 that leaves both bodies at rest (equal and opposite momenta)."""
    r, depth = 0.5, 0.1
    b = np.zeros(2, abi.body_dtype)
    shapes.fill_bodies_dynamic(b[0:1], 1.0, 0.1, 1.0, -(r - depth / 2), 0.0)       # mass 1
    shapes.fill_bodies_dynamic(b[1:2], 1.0, 0.1, 3.0, +(r - depth / 2), 0.0)       # mass 3
    c = np.zeros(2, abi.collider_dtype)
    shapes.fill_colliders(c, shapes.circle(np.full(2, r)), body=np.arange(2))
    sc = scenes._scene("two_circles", b, c, 1)
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=1, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    w.tick(1.0 / 60.0, forcefield=_no_field(), mode=getattr(oracle, mode_name))
    st = w.get_bodies()
    w0, w1 = 1.0, 1.0 / 3.0
    x0 = float(b["x"][0]) - depth * w0 / (w0 + w1)
    x1 = float(b["x"][1]) + depth * w1 / (w0 + w1)
    assert abs(st[0, 0] - x0) < 1e-6 and abs(st[1, 0] - x1) < 1e-6                 # f32 inputs: the radii and positions carry 1e-8
    assert abs((st[1, 0] - st[0, 0]) - 2 * r) < 1e-6                                # exactly touching afterwards
    assert np.abs(st[:, 1]).max() < 1e-12 and np.abs(st[:, 4:7]).max() < 1e-4       # nothing sideways, nothing left moving


def test_circle_sunk_into_static_ground_is_lifted_onto_it(oracle):
    """The same correction against a collider without a body (inverse mass 0): the circle takes all of it."""
    r, depth = 0.5, 0.05
    b = np.zeros(1, abi.body_dtype)
    top = 0.1                                                                      # new_rounded_rect(w 10, h 0.2, r 0): half-height 0.1
    shapes.fill_bodies_dynamic(b, 1.0, 0.1, 1.0, 0.0, top + r - depth)
    c = np.zeros(2, abi.collider_dtype)
    shapes.fill_colliders(c[0:1], shapes.rounded_rect(10.0, 0.2, 0.0), x=0.0, y=0.0)
    shapes.fill_colliders(c[1:2], shapes.circle(np.full(1, r)), body=np.zeros(1, np.int32))
    sc = scenes._scene("circle_on_ground", b, c, 1)
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=1, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    w.tick(1.0 / 60.0, forcefield=_no_field())
    st = w.get_bodies()
    assert abs(st[0, 1] - (top + r)) < 1e-6 and abs(st[0, 0]) < 1e-12
    assert np.abs(st[0, 4:7]).max() < 1e-4


@pytest.mark.parametrize("compliance", [0.0, 2e-4])
@pytest.mark.parametrize("mode_name", ["MODE_REFERENCE", "MODE_COLOURED"])
def test_distance_constraint_closed_form_including_the_double_visit(oracle, mode_name, compliance):
    """solver.rs:187-269 with both offsets at the body origins (no lever arm): one visit moves the separation from L towards the
    target d by the fraction k = (w0 + w1) / (w0 + w1 + compliance / dt^2), split between the bodies by inverse mass. A two-body
    constraint is stored on BOTH bodies (physics.rs:519-532) and therefore visited twice per substep: separation after one
    substep = d - (d - L) (1 - k)^2. No field, no damping, bodies at rest."""
    L, d, dt = 2.0, 1.5, 1.0 / 60.0
    b = np.zeros(2, abi.body_dtype)
    shapes.fill_bodies_dynamic(b[0:1], 1.0, 0.1, 1.0, -L / 2, 0.25)        # mass 1
    shapes.fill_bodies_dynamic(b[1:2], 1.0, 0.1, 4.0, +L / 2, 0.25)        # mass 4
    c = np.zeros(2, abi.collider_dtype)
    shapes.fill_colliders(c, shapes.circle(np.full(2, 0.1)), body=np.arange(2))
    k = np.zeros(1, abi.constraint_dtype)
    k["owner"], k["target"], k["flags"], k["distance"], k["compliance"] = 0, 1, abi.LIMIT_EQ, d, compliance
    sc = scenes._scene("two_bodies_one_constraint", b, c, 1, constraints=k)
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=1, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    w.tick(dt, forcefield=_no_field(), mode=getattr(oracle, mode_name))
    st = w.get_bodies()
    w0, w1 = 1.0, 0.25
    kk = (w0 + w1) / (w0 + w1 + float(np.float32(compliance)) / dt ** 2)
    sep = d - (d - L) * (1.0 - kk) ** 2
    moved = L - sep                                                         # shared out by inverse mass, towards each other
    assert abs((st[1, 0] - st[0, 0]) - sep) < 1e-6
    assert abs(st[0, 0] - (-L / 2 + moved * w0 / (w0 + w1))) < 1e-6 and abs(st[1, 0] - (L / 2 - moved * w1 / (w0 + w1))) < 1e-6
    assert np.abs(st[:, 1] - 0.25).max() < 1e-9
    # the velocity is what the correction implies (solver.rs:91-97), and momentum is untouched
    assert abs(st[0, 4] - moved * w0 / (w0 + w1) / dt) < 1e-4 and abs(1.0 * st[0, 4] + 4.0 * st[1, 4]) < 1e-6


@pytest.mark.parametrize("mode_name", ["MODE_REFERENCE", "MODE_COLOURED"])
def test_two_particle_rope_closed_form(oracle, mode_name):
    """solver.rs:118-145 and 722-740 for a rope of two particles (no bending term): ONE distance step per substep (rope links are not
    duplicated, physics.rs:621-628) moves the separation from L to spacing + (L - spacing) (1 - k), k = (w0 + w1) / (w0 + w1 +
    compliance / dt^2), split by inverse mass; the velocity step then removes the fraction min(damping dt, 1) of the relative velocity."""
    L, spacing, dt, damping, compliance = 0.5, 0.3, 1.0 / 60.0, 20.0, 1e-4
    b = np.zeros(2, abi.body_dtype)
    shapes.fill_bodies_particle(b[0:1], 2.0, -L / 2, 1.0)
    shapes.fill_bodies_particle(b[1:2], 0.5, +L / 2, 1.0)
    c = np.zeros(2, abi.collider_dtype)
    shapes.fill_colliders(c, shapes.circle(np.full(2, 0.05)), body=np.arange(2), layer=abi.ROPE_LAYER)
    r = scenes.rope_params(1, spacing)
    r["compliance"], r["damping"], r["first_particle"], r["particle_count"] = compliance, damping, 0, 2
    sc = scenes._scene("two_particle_rope", b, c, 1, ropes=r, rope_particles=np.arange(2))
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=1, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    w.tick(dt, forcefield=_no_field(), mode=getattr(oracle, mode_name))
    st = w.get_bodies()
    w0, w1 = 0.5, 2.0
    k = (w0 + w1) / (w0 + w1 + float(np.float32(compliance)) / dt ** 2)
    moved = (L - float(np.float32(spacing))) * k                            # how much closer the particles end up
    assert abs((st[1, 0] - st[0, 0]) - (L - moved)) < 1e-6
    assert abs(st[0, 0] - (-L / 2 + moved * w0 / (w0 + w1))) < 1e-6 and abs(st[1, 0] - (L / 2 - moved * w1 / (w0 + w1))) < 1e-6
    # velocities: the correction's, less the damped share of the relative velocity; momentum (2 v0 + 0.5 v1) stays zero
    keep = 1.0 - min(damping * dt, 1.0)
    assert abs(st[0, 4] - keep * moved * w0 / (w0 + w1) / dt) < 1e-4 and abs(st[1, 4] + keep * moved * w1 / (w0 + w1) / dt) < 1e-4
    assert abs(2.0 * st[0, 4] + 0.5 * st[1, 4]) < 1e-6 and np.abs(st[:, 1] - 1.0).max() < 1e-9


@pytest.mark.parametrize("e,v_in,field,bounces", [(0.8, -2.0, False, True), (0.5, -0.5, False, True), (0.0, -2.0, False, False),
                                                   (0.8, -0.30, True, True), (0.8, -0.10, True, False)])
def test_restitution_closed_form_and_its_resting_threshold(oracle, e, v_in, field, bounces):
    """solver.rs:555-576: a circle arriving at static ground with normal speed |v| (its velocity AFTER the external acceleration of
    the substep, `old_velocities`) leaves with e |v| — unless v^2 < dt^2 |a_ext|^2 (solver.rs:569: slower than what one substep of
    the external field adds counts as resting), then with 0. The restitution of a pair is the larger of the two colliders'."""
    r, depth, top, dt = 0.5, 0.02, 0.1, 1.0 / 60.0
    g = G if field else 0.0
    b = np.zeros(1, abi.body_dtype)
    shapes.fill_bodies_dynamic(b, 1.0, 0.1, 1.0, 0.0, top + r - depth, vy=v_in - g * dt)     # so that the speed after the field's kick is v_in
    c = np.zeros(2, abi.collider_dtype)
    shapes.fill_colliders(c[0:1], shapes.rounded_rect(10.0, 0.2, 0.0), x=0.0, y=0.0)
    shapes.fill_colliders(c[1:2], shapes.circle(np.full(1, r)), body=np.zeros(1, np.int32), restitution=e)
    sc = scenes._scene("bounce", b, c, 1)
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=1, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    w.tick(dt, forcefield=abi.gravity_field() if field else _no_field())
    st = w.get_bodies()
    assert abs(st[0, 1] - (top + r)) < 1e-6                                  # lifted onto the ground
    want = float(np.float32(e)) * abs(v_in) if bounces else 0.0
    assert abs(st[0, 5] - want) < 2e-6, (st[0, 5], want)
    assert abs(st[0, 4]) < 1e-12 and abs(st[0, 6]) < 1e-9

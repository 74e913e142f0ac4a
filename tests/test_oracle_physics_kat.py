"""Closed-form pins of the oracle's tick ABOVE the collision primitives (the reference has no tests there): bodies that touch
nothing follow the reference's own update rule exactly, so their state after n substeps has a closed form.
Rule (solver.rs:55-75, physics.rs:78-83): per substep  v += a dt  (only finite mass without ignores_gravity),  x += v dt,
rotation = from_angle(w dt) * rotation;  afterwards the velocity is re-derived from the pose difference (solver.rs:91-97), which
gives the same v back.  Hence after n substeps:  v_n = v_0 + n a dt,  x_n = x_0 + n v_0 dt + a dt^2 n (n + 1) / 2,  angle_n = angle_0 + n w dt.
Substep count and dt under time scaling: physics.rs:411-415. The scenes live in kat_scenes.py (the device is held to the same
closed forms by test_gpu_physics_kat.py)."""
import math

import numpy as np
import pytest

import kat_scenes as K
from moleengine_b200 import abi


@pytest.mark.parametrize("mode_name", ["MODE_REFERENCE", "MODE_COLOURED"])
@pytest.mark.parametrize("substeps", [1, 4, 10])
def test_free_bodies_follow_the_closed_form(oracle, mode_name, substeps):
    sc = K.free_bodies()
    # f32 inputs are exact in f64; the closed form and the tick then differ only by f64 rounding
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=substeps, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    ticks, frame_dt = 30, 1.0 / 60.0
    for _ in range(ticks):
        w.tick(frame_dt, mode=getattr(oracle, mode_name))
    K.free_check(w.get_bodies(), K.free_expect(sc, ticks * substeps, frame_dt / substeps), 1e-9)


@pytest.mark.parametrize("time_scale,substeps", [(0.5, 10), (0.25, 4), (1.0, 3), (2.0, 4), (0.33, 10)])
def test_time_scale_changes_substep_count_and_dt(oracle, time_scale, substeps):
    """physics.rs:411-415: substeps = ceil(time_scale * tuning.substeps), dt = time_scale * frame_dt / substeps."""
    sc = K.free_bodies()
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=substeps, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    ticks, frame_dt = 12, 1.0 / 60.0
    for _ in range(ticks):
        w.tick(frame_dt, time_scale=time_scale)
    n = int(math.ceil(time_scale * substeps))
    K.free_check(w.get_bodies(), K.free_expect(sc, ticks * n, time_scale * frame_dt / n), 1e-9)


@pytest.mark.parametrize("mode_name", ["MODE_REFERENCE", "MODE_COLOURED"])
def test_two_overlapping_circles_are_pushed_apart_by_inverse_mass(oracle, mode_name):
    """solver.rs:428-452 + 496-618 in the one case with a closed form: two circles at rest overlapping by `depth` along x, no
    external field. The normal correction is lambda = -depth / (w0 + w1) along the contact normal, i.e. body i moves by
    depth * w_i / (w0 + w1) away from the other (no lever arm: the contact lies on the line of centres); the velocity the
    correction leaves behind is purely normal and separating, and the velocity step (restitution 0 by default) removes the
    relative normal velocity again — with masses m0, m1 that leaves both bodies at rest (equal and opposite momenta)."""
    r, depth = 0.5, 0.1
    sc = K.two_circles(r, depth)
    b = sc["bodies"]
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=1, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    w.tick(1.0 / 60.0, forcefield=K.no_field(), mode=getattr(oracle, mode_name))
    st = w.get_bodies()
    w0, w1 = 1.0, 1.0 / 3.0
    x0 = float(b["x"][0]) - depth * w0 / (w0 + w1)
    x1 = float(b["x"][1]) + depth * w1 / (w0 + w1)
    assert abs(st[0, 0] - x0) < 1e-6 and abs(st[1, 0] - x1) < 1e-6                 # f32 inputs: the radii and positions carry 1e-8
    assert abs((st[1, 0] - st[0, 0]) - 2 * r) < 1e-6                                # exactly touching afterwards
    assert np.abs(st[:, 1]).max() < 1e-12 and np.abs(st[:, 4:7]).max() < 1e-4       # nothing sideways, nothing left moving


def test_circle_sunk_into_static_ground_is_lifted_onto_it(oracle):
    """The same correction against a collider without a body (inverse mass 0): the circle takes all of it."""
    r = 0.5
    sc = K.circle_on_ground(r, 0.05)
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=1, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    w.tick(1.0 / 60.0, forcefield=K.no_field())
    st = w.get_bodies()
    assert abs(st[0, 1] - (K.GROUND_TOP + r)) < 1e-6 and abs(st[0, 0]) < 1e-12
    assert np.abs(st[0, 4:7]).max() < 1e-4


@pytest.mark.parametrize("compliance", [0.0, 2e-4])
@pytest.mark.parametrize("mode_name", ["MODE_REFERENCE", "MODE_COLOURED"])
def test_distance_constraint_closed_form_including_the_double_visit(oracle, mode_name, compliance):
    """solver.rs:187-269 with both offsets at the body origins (no lever arm): one visit moves the separation from L towards the
    target d by the fraction k = (w0 + w1) / (w0 + w1 + compliance / dt^2), split between the bodies by inverse mass. A two-body
    constraint is stored on BOTH bodies (physics.rs:519-532) and therefore visited twice per substep: separation after one
    substep = d - (d - L) (1 - k)^2. No field, no damping, bodies at rest."""
    L, d, dt = 2.0, 1.5, 1.0 / 60.0
    sc = K.two_bodies_one_constraint(L, d, compliance)
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=1, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    w.tick(dt, forcefield=K.no_field(), mode=getattr(oracle, mode_name))
    st = w.get_bodies()
    sep, m0, m1 = K.constraint_expect(L, d, compliance, dt)
    assert abs((st[1, 0] - st[0, 0]) - sep) < 1e-6
    assert abs(st[0, 0] - (-L / 2 + m0)) < 1e-6 and abs(st[1, 0] - (L / 2 - m1)) < 1e-6
    assert np.abs(st[:, 1] - 0.25).max() < 1e-9
    # the velocity is what the correction implies (solver.rs:91-97), and momentum is untouched
    assert abs(st[0, 4] - m0 / dt) < 1e-4 and abs(1.0 * st[0, 4] + 4.0 * st[1, 4]) < 1e-6


@pytest.mark.parametrize("mode_name", ["MODE_REFERENCE", "MODE_COLOURED"])
def test_two_particle_rope_closed_form(oracle, mode_name):
    """solver.rs:118-145 and 722-740 for a rope of two particles (no bending term): ONE distance step per substep (rope links are not
    duplicated, physics.rs:621-628) moves the separation from L to spacing + (L - spacing) (1 - k), k = (w0 + w1) / (w0 + w1 +
    compliance / dt^2), split by inverse mass; the velocity step then removes the fraction min(damping dt, 1) of the relative velocity."""
    L, spacing, dt, damping, compliance = 0.5, 0.3, 1.0 / 60.0, 20.0, 1e-4
    sc = K.two_particle_rope(L, spacing, damping, compliance)
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=1, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    w.tick(dt, forcefield=K.no_field(), mode=getattr(oracle, mode_name))
    st = w.get_bodies()
    m0, m1, keep = K.rope_expect(L, spacing, damping, compliance, dt)
    assert abs((st[1, 0] - st[0, 0]) - (L - m0 - m1)) < 1e-6
    assert abs(st[0, 0] - (-L / 2 + m0)) < 1e-6 and abs(st[1, 0] - (L / 2 - m1)) < 1e-6
    # velocities: the correction's, less the damped share of the relative velocity; momentum (2 v0 + 0.5 v1) stays zero
    assert abs(st[0, 4] - keep * m0 / dt) < 1e-4 and abs(st[1, 4] + keep * m1 / dt) < 1e-4
    assert abs(2.0 * st[0, 4] + 0.5 * st[1, 4]) < 1e-6 and np.abs(st[:, 1] - 1.0).max() < 1e-9


RESTITUTION_CASES = [(0.8, -2.0, False, True), (0.5, -0.5, False, True), (0.0, -2.0, False, False), (0.8, -0.30, True, True), (0.8, -0.10, True, False)]


@pytest.mark.parametrize("e,v_in,field,bounces", RESTITUTION_CASES)
def test_restitution_closed_form_and_its_resting_threshold(oracle, e, v_in, field, bounces):
    """solver.rs:555-576: a circle arriving at static ground with normal speed |v| (its velocity AFTER the external acceleration of
    the substep, `old_velocities`) leaves with e |v| — unless v^2 < dt^2 |a_ext|^2 (solver.rs:569: slower than what one substep of
    the external field adds counts as resting), then with 0. The restitution of a pair is the larger of the two colliders'."""
    r, dt = 0.5, 1.0 / 60.0
    g = K.G if field else 0.0
    sc = K.circle_on_ground(r, 0.02, vy=v_in - g * dt, restitution=e)       # so that the speed after the field's kick is v_in
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=1, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    w.tick(dt, forcefield=abi.gravity_field() if field else K.no_field())
    st = w.get_bodies()
    assert abs(st[0, 1] - (K.GROUND_TOP + r)) < 1e-6                         # lifted onto the ground
    want = float(np.float32(e)) * abs(v_in) if bounces else 0.0
    assert abs(st[0, 5] - want) < 2e-6, (st[0, 5], want)
    assert abs(st[0, 4]) < 1e-12 and abs(st[0, 6]) < 1e-9


SLIDING_CASES = [3.0, -3.0, 1.0, -1.0, 1.95, 1.90]


@pytest.mark.parametrize("mode_name", ["MODE_REFERENCE", "MODE_COLOURED"])
@pytest.mark.parametrize("v_t", SLIDING_CASES)
def test_friction_closed_form_including_the_signed_static_rule(oracle, mode_name, v_t):
    """kat_scenes.sliding_expect: static friction only for fast motion in ONE direction (the reference compares signed lambdas,
    solver.rs:462-467), dynamic friction bounded by mu_d |lambda_n| / dt (solver.rs:578-592); threshold at v_t dt = mu_s depth = 1.92 m/s."""
    depth, dt = 0.02, 1.0 / 60.0
    sc = K.sliding_particle(v_t, depth=depth)
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=1, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    w.tick(dt, forcefield=K.no_field(), mode=getattr(oracle, mode_name))
    st = w.get_bodies()
    x, vx = K.sliding_expect(v_t, depth, dt)
    assert abs(st[0, 0] - x) < 1e-6 and abs(st[0, 4] - vx) < 1e-4, (st[0], x, vx)
    assert abs(st[0, 1] - (K.GROUND_TOP + 0.5)) < 1e-6 and abs(st[0, 5]) < 1e-4


def test_kinematic_body_pushes_and_is_never_corrected(oracle):
    """body.rs:92-97, solver.rs:63,323-331: a body with infinite mass and inertia moves by its velocity alone (no gravity either);
    in a contact its inverse mass is 0, so the dynamic partner takes the whole correction."""
    r, depth, dt = 0.5, 0.1, 1.0 / 60.0
    b = np.zeros(2, abi.body_dtype)
    K.shapes.fill_bodies_kinematic(b[0:1], -(r - depth / 2), 0.0, vx=0.6)
    K.shapes.fill_bodies_dynamic(b[1:2], 1.0, 0.1, 1.0, +(r - depth / 2), 0.0)
    c = np.zeros(2, abi.collider_dtype)
    K.shapes.fill_colliders(c, K.shapes.circle(np.full(2, r)), body=np.arange(2))
    sc = K.scenes._scene("pusher", b, c, 1)
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=1, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    w.tick(dt)                                                               # default field: gravity acts on the dynamic body only
    st = w.get_bodies()
    vk = float(b["vx"][0])                                                   # 0.6 as f32
    xk = float(b["x"][0]) + vk * dt
    assert abs(st[0, 0] - xk) < 1e-12 and abs(st[0, 1]) < 1e-12 and abs(st[0, 4] - vk) < 1e-9 and abs(st[0, 5]) < 1e-12     # on rails
    # the dynamic circle ends exactly touching the kinematic one (the contact normal is the line of centres after the gravity step)
    d = math.hypot(st[1, 0] - st[0, 0], st[1, 1] - st[0, 1])
    assert abs(d - 2 * r) < 1e-6
    assert st[1, 0] > float(b["x"][1]) + depth * 0.9                          # it took (nearly) all of the overlap, plus the pusher's advance


def test_sensor_reports_the_contact_but_moves_nothing(oracle):
    """collider.rs `is_sensor`: a sensor pair is latched as a contact (physics.rs:1097-1122 publishes it) but is skipped by the position
    and velocity solves (solver.rs:323-331, 511-516)."""
    r, depth, dt = 0.5, 0.1, 1.0 / 60.0
    sc = K.two_circles(r, depth)
    sc["colliders"]["flags"][1] |= abi.COLL_SENSOR
    b = sc["bodies"]
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=1, flags=abi.TUNE_NO_SLEEPING))
    w.load_scene(sc)
    w.tick(dt, forcefield=K.no_field())
    st = w.get_bodies()
    assert abs(st[0, 0] - float(b["x"][0])) < 1e-12 and abs(st[1, 0] - float(b["x"][1])) < 1e-12 and np.abs(st[:, 4:7]).max() < 1e-12
    cs = w.contacts()
    assert len(cs) >= 1

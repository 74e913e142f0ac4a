"""The device held directly to the closed forms of the reference's update rules (tests/kat_scenes.py; derivations and reference
lines in test_oracle_physics_kat.py) — no oracle in between. Tolerances are f32: positions of order 1..10 carry 1e-6 per
operation, velocities are displacements divided by dt. The angular velocity is re-derived every substep from the difference of two
f32 rotors (solver.rs:91-97), a quantity of order w dt / 2: at 600 substeps per second that is 1e-7 against 5e-4, and the noise
random-walks through the substeps (measured after 300 substeps of 1/600 s: 6e-3 in w, 1.6e-3 in the angle) — hence `spin_tol` and `angle_tol`."""
import math

import numpy as np
import pytest

import kat_scenes as K
from moleengine_b200 import GpuWorld, abi

pytestmark = pytest.mark.gpu


def _run(sc, substeps, ticks, frame_dt=1.0 / 60.0, time_scale=None, field=None):
    g = GpuWorld(tuning=abi.default_tuning(substeps=substeps, flags=abi.TUNE_NO_SLEEPING))
    g.load_scene(sc)
    for _ in range(ticks):
        g.tick(frame_dt, time_scale, field)
    st = K.rows(g.get_bodies())
    g.close()
    return st


@pytest.mark.parametrize("substeps", [1, 4, 10])
def test_free_bodies_follow_the_closed_form(substeps):
    sc = K.free_bodies(spread=8.0, height=6.0)
    st = _run(sc, substeps, 30)
    K.free_check(st, K.free_expect(sc, 30 * substeps, 1.0 / 60.0 / substeps), 5e-4, vel_tol=5e-3, spin_tol=3e-2, angle_tol=5e-3)


@pytest.mark.parametrize("time_scale,substeps", [(0.5, 10), (0.25, 4), (2.0, 4), (0.33, 10)])
def test_time_scale_changes_substep_count_and_dt(time_scale, substeps):
    sc = K.free_bodies(spread=8.0, height=6.0)
    st = _run(sc, substeps, 12, time_scale=time_scale)
    n = int(math.ceil(time_scale * substeps))
    K.free_check(st, K.free_expect(sc, 12 * n, time_scale / 60.0 / n), 5e-4, vel_tol=5e-3, spin_tol=3e-2, angle_tol=5e-3)


def test_two_overlapping_circles_are_pushed_apart_by_inverse_mass():
    r, depth = 0.5, 0.1
    sc = K.two_circles(r, depth)
    b = sc["bodies"]
    st = _run(sc, 1, 1, field=K.no_field())
    w0, w1 = 1.0, 1.0 / 3.0
    assert abs(st[0, 0] - (float(b["x"][0]) - depth * w0 / (w0 + w1))) < 1e-5 and abs(st[1, 0] - (float(b["x"][1]) + depth * w1 / (w0 + w1))) < 1e-5, st
    assert abs((st[1, 0] - st[0, 0]) - 2 * r) < 1e-5
    assert np.abs(st[:, 1]).max() < 1e-6 and np.abs(st[:, 4:7]).max() < 1e-3, st


def test_circle_sunk_into_static_ground_is_lifted_onto_it():
    r = 0.5
    st = _run(K.circle_on_ground(r, 0.05), 1, 1, field=K.no_field())
    assert abs(st[0, 1] - (K.GROUND_TOP + r)) < 1e-5 and abs(st[0, 0]) < 1e-6, st
    assert np.abs(st[0, 4:7]).max() < 1e-3, st


@pytest.mark.parametrize("compliance", [0.0, 2e-4])
def test_distance_constraint_closed_form_including_the_double_visit(compliance):
    L, d, dt = 2.0, 1.5, 1.0 / 60.0
    st = _run(K.two_bodies_one_constraint(L, d, compliance), 1, 1, field=K.no_field())
    sep, m0, m1 = K.constraint_expect(L, d, compliance, dt)
    assert abs((st[1, 0] - st[0, 0]) - sep) < 1e-5, (st, sep)
    assert abs(st[0, 0] - (-L / 2 + m0)) < 1e-5 and abs(st[1, 0] - (L / 2 - m1)) < 1e-5, st
    assert abs(st[0, 4] - m0 / dt) < 2e-3 and abs(1.0 * st[0, 4] + 4.0 * st[1, 4]) < 2e-3, st


def test_two_particle_rope_closed_form():
    L, spacing, dt, damping, compliance = 0.5, 0.3, 1.0 / 60.0, 20.0, 1e-4
    st = _run(K.two_particle_rope(L, spacing, damping, compliance), 1, 1, field=K.no_field())
    m0, m1, keep = K.rope_expect(L, spacing, damping, compliance, dt)
    assert abs(st[0, 0] - (-L / 2 + m0)) < 1e-5 and abs(st[1, 0] - (L / 2 - m1)) < 1e-5, (st, m0, m1)
    assert abs(st[0, 4] - keep * m0 / dt) < 2e-3 and abs(st[1, 4] + keep * m1 / dt) < 2e-3, (st, keep)


@pytest.mark.parametrize("e,v_in,field,bounces", [(0.8, -2.0, False, True), (0.5, -0.5, False, True), (0.0, -2.0, False, False), (0.8, -0.30, True, True), (0.8, -0.10, True, False)])
def test_restitution_closed_form_and_its_resting_threshold(e, v_in, field, bounces):
    r, dt = 0.5, 1.0 / 60.0
    g = K.G if field else 0.0
    sc = K.circle_on_ground(r, 0.02, vy=v_in - g * dt, restitution=e)
    st = _run(sc, 1, 1, field=abi.gravity_field() if field else K.no_field())
    assert abs(st[0, 1] - (K.GROUND_TOP + r)) < 1e-5, st
    want = float(np.float32(e)) * abs(v_in) if bounces else 0.0
    assert abs(st[0, 5] - want) < 1e-3, (st[0, 5], want)


@pytest.mark.parametrize("v_t", [3.0, -3.0, 1.0, -1.0])
def test_friction_closed_form_including_the_signed_static_rule(v_t):
    depth, dt = 0.02, 1.0 / 60.0
    st = _run(K.sliding_particle(v_t, depth=depth), 1, 1, field=K.no_field())
    x, vx = K.sliding_expect(v_t, depth, dt)
    assert abs(st[0, 0] - x) < 1e-5 and abs(st[0, 4] - vx) < 1e-3, (st[0], x, vx)
    assert abs(st[0, 1] - (K.GROUND_TOP + 0.5)) < 1e-5 and abs(st[0, 5]) < 1e-3, st[0]

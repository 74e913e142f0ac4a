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

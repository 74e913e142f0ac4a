"""Scenes with closed-form answers under the reference's update rules (used by test_oracle_physics_kat.py on the CPU oracle and by
test_gpu_physics_kat.py on the device). Derivations and reference lines: test_oracle_physics_kat.py."""
import math

import numpy as np

from moleengine_b200 import abi, scenes, shapes  # noqa: F401 (re-exported for the tests)

G = float(np.float32(-9.81))      # the force field crosses the ABI as f32


def no_field():
    return np.zeros(1, abi.forcefield_dtype)


def free_bodies(spread=100.0, height=50.0):
    b = np.zeros(4, abi.body_dtype)
    shapes.fill_bodies_dynamic(b[0:1], 1.0, 0.2, 0.5, -spread, height, angle=0.3, vx=1.5, vy=2.0, w=0.7)
    shapes.fill_bodies_dynamic(b[1:2], 2.0, 0.4, 0.5, 0.0, height, angle=-1.1, vx=-0.5, vy=0.25, w=-0.4)
    b["flags"][1] |= abi.BODY_IGNORES_GRAVITY
    shapes.fill_bodies_kinematic(b[2:3], spread, height, angle=0.0, vx=0.75, vy=-0.5, w=1.3)
    shapes.fill_bodies_particle(b[3:4], 0.3, 2.0 * spread, height)
    b["vx"][3], b["vy"][3] = 0.1, 3.0
    c = np.zeros(5, abi.collider_dtype)
    shapes.fill_colliders(c[0:1], shapes.rounded_rect(10.0, 0.2, 0.0), x=0.0, y=-1000.0)          # a floor nobody reaches
    shapes.fill_colliders(c[1:5], shapes.circle(np.full(4, 0.25)), body=np.arange(4))
    return scenes._scene("free_bodies", b, c, 10)


def free_expect(sc, n, dt):
    """(x, y, vx, vy, angle, w) of every body after n substeps of dt."""
    b = sc["bodies"]
    x0, y0, vx0, vy0, w0 = (b[k].astype(np.float64) for k in ("x", "y", "vx", "vy", "w"))
    falls = ((b["flags"] & abi.BODY_MASS_FINITE) != 0) & ((b["flags"] & abi.BODY_IGNORES_GRAVITY) == 0)
    a = np.where(falls, G, 0.0)
    ang0 = -2.0 * np.arctan2(b["rot_b"].astype(np.float64), b["rot_s"].astype(np.float64))
    return (x0 + n * vx0 * dt, y0 + n * vy0 * dt + a * dt * dt * n * (n + 1) / 2.0, vx0, vy0 + n * a * dt, ang0 + n * w0 * dt, w0)


def free_check(state, want, tol, vel_tol=None, spin_tol=None, angle_tol=None):
    """state: (n, 7) rows of x, y, rot_s, rot_b, vx, vy, w."""
    vel_tol = tol if vel_tol is None else vel_tol
    spin_tol = vel_tol if spin_tol is None else spin_tol
    angle_tol = tol if angle_tol is None else angle_tol
    x, y, vx, vy, ang, w = want
    assert np.abs(state[:, 0] - x).max() < tol and np.abs(state[:, 1] - y).max() < tol, (state[:, :2], x, y)
    assert np.abs(state[:, 4] - vx).max() < vel_tol and np.abs(state[:, 5] - vy).max() < vel_tol, (state[:, 4:6], vx, vy)
    assert np.abs(state[:, 6] - w).max() < spin_tol, (state[:, 6], w)
    got = -2.0 * np.arctan2(state[:, 3], state[:, 2])            # DRotor2::from_angle: bv.xy = -sin(a / 2) (shapes.rotor)
    d = (got - ang + math.pi) % (2.0 * math.pi) - math.pi          # the rotor covers the angle twice: compare modulo a turn
    assert np.abs(d).max() < angle_tol, d
    assert np.abs(state[:, 2] ** 2 + state[:, 3] ** 2 - 1.0).max() < max(tol, 1e-6)      # rotors stay unit (as unit as their f32 inputs were)


def rows(bodies):
    """abi.body_dtype records -> (n, 7) f64 rows like OracleWorld.get_bodies."""
    return np.stack([bodies[k].astype(np.float64) for k in ("x", "y", "rot_s", "rot_b", "vx", "vy", "w")], axis=1)


def two_circles(r=0.5, depth=0.1):
    b = np.zeros(2, abi.body_dtype)
    shapes.fill_bodies_dynamic(b[0:1], 1.0, 0.1, 1.0, -(r - depth / 2), 0.0)       # mass 1
    shapes.fill_bodies_dynamic(b[1:2], 1.0, 0.1, 3.0, +(r - depth / 2), 0.0)       # mass 3
    c = np.zeros(2, abi.collider_dtype)
    shapes.fill_colliders(c, shapes.circle(np.full(2, r)), body=np.arange(2))
    return scenes._scene("two_circles", b, c, 1)


GROUND_TOP = 0.1                                                                   # new_rounded_rect(w 10, h 0.2, r 0): half-height 0.1


def circle_on_ground(r=0.5, depth=0.05, vy=0.0, restitution=0.0):
    b = np.zeros(1, abi.body_dtype)
    shapes.fill_bodies_dynamic(b, 1.0, 0.1, 1.0, 0.0, GROUND_TOP + r - depth, vy=vy)
    c = np.zeros(2, abi.collider_dtype)
    shapes.fill_colliders(c[0:1], shapes.rounded_rect(10.0, 0.2, 0.0), x=0.0, y=0.0)
    shapes.fill_colliders(c[1:2], shapes.circle(np.full(1, r)), body=np.zeros(1, np.int32), restitution=restitution)
    return scenes._scene("circle_on_ground", b, c, 1)


def two_bodies_one_constraint(L=2.0, d=1.5, compliance=0.0):
    b = np.zeros(2, abi.body_dtype)
    shapes.fill_bodies_dynamic(b[0:1], 1.0, 0.1, 1.0, -L / 2, 0.25)        # mass 1
    shapes.fill_bodies_dynamic(b[1:2], 1.0, 0.1, 4.0, +L / 2, 0.25)        # mass 4
    c = np.zeros(2, abi.collider_dtype)
    shapes.fill_colliders(c, shapes.circle(np.full(2, 0.1)), body=np.arange(2))
    k = np.zeros(1, abi.constraint_dtype)
    k["owner"], k["target"], k["flags"], k["distance"], k["compliance"] = 0, 1, abi.LIMIT_EQ, d, compliance
    return scenes._scene("two_bodies_one_constraint", b, c, 1, constraints=k)


def constraint_expect(L, d, compliance, dt):
    """(separation after one substep, how far each body moved towards the other) for masses 1 and 4."""
    w0, w1 = 1.0, 0.25
    kk = (w0 + w1) / (w0 + w1 + float(np.float32(compliance)) / dt ** 2)
    sep = d - (d - L) * (1.0 - kk) ** 2
    moved = L - sep
    return sep, moved * w0 / (w0 + w1), moved * w1 / (w0 + w1)


def two_particle_rope(L=0.5, spacing=0.3, damping=20.0, compliance=1e-4):
    b = np.zeros(2, abi.body_dtype)
    shapes.fill_bodies_particle(b[0:1], 2.0, -L / 2, 1.0)
    shapes.fill_bodies_particle(b[1:2], 0.5, +L / 2, 1.0)
    c = np.zeros(2, abi.collider_dtype)
    shapes.fill_colliders(c, shapes.circle(np.full(2, 0.05)), body=np.arange(2), layer=abi.ROPE_LAYER)
    r = scenes.rope_params(1, spacing)
    r["compliance"], r["damping"], r["first_particle"], r["particle_count"] = compliance, damping, 0, 2
    return scenes._scene("two_particle_rope", b, c, 1, ropes=r, rope_particles=np.arange(2))


def rope_expect(L, spacing, damping, compliance, dt):
    """(how far particle 0 (mass 2) and particle 1 (mass 0.5) moved towards each other, share of the relative velocity kept)."""
    w0, w1 = 0.5, 2.0
    k = (w0 + w1) / (w0 + w1 + float(np.float32(compliance)) / dt ** 2)
    moved = (L - float(np.float32(spacing))) * k
    return moved * w0 / (w0 + w1), moved * w1 / (w0 + w1), 1.0 - min(damping * dt, 1.0)


def sliding_particle(v_t, r=0.5, depth=0.02):
    """A particle body (mass 1, infinite inertia: no lever arms) with a circle collider, sunk `depth` into static ground, sliding at v_t."""
    b = np.zeros(1, abi.body_dtype)
    shapes.fill_bodies_particle(b, 1.0, 0.0, GROUND_TOP + r - depth)
    b["vx"] = v_t
    c = np.zeros(2, abi.collider_dtype)
    shapes.fill_colliders(c[0:1], shapes.rounded_rect(10.0, 0.2, 0.0), x=0.0, y=0.0)
    shapes.fill_colliders(c[1:2], shapes.circle(np.full(1, r)), body=np.zeros(1, np.int32))     # default material: mu_s 1.6, mu_d 1.5
    return scenes._scene("sliding_particle", b, c, 1)


def sliding_expect(v_t, depth, dt, mu_s=1.6, mu_d=1.5):
    """(x, vx) after one substep without external field, mass 1. Static friction (solver.rs:456-487) is SIGNED in the reference: it
    engages when lambda_t < mu_s lambda_n with lambda_n = -depth m < 0 and lambda_t = -(tangential motion) m, i.e. only for motion
    along +left_normal(n) larger than mu_s depth — for a body on the ground that is the +x direction — and then cancels the substep's
    tangential motion altogether. Otherwise the body slides, and the velocity step (solver.rs:578-592) takes
    min(|v_t|, mu_d |lambda_n| / dt) off the tangential velocity."""
    if v_t * dt > mu_s * depth:
        return 0.0, 0.0
    cut = min(abs(v_t), mu_d * depth / dt)
    return v_t * dt, v_t - math.copysign(cut, v_t)

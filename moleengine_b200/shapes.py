"""Host-side collider / body constructors, vectorised over numpy arrays.

Mirrors the reference's scene-building API so synthetic scenes are assembled exactly the way a
Starframe user would: `Collider::new_*` (collider.rs:63-125), `ColliderShape::area` /
`second_moment_of_area` (collider.rs:222-337), `Body::new_dynamic` / `new_particle` /
`new_kinematic` (body.rs:23-75), `CompoundColliderSetup` (compound_collider.rs:17-43, including its
centre-of-mass divided by the collider COUNT) and `PhysicsMaterial::default` (collider.rs:913-921).
All arithmetic is f64; results are rounded to f32 once when stored in the ABI records.
"""
import numpy as np

from . import abi

# collider.rs:419-424 — truncated exactly as the reference has them
SQRT_3 = 1.73205080757
PI6_COS = 0.866025403784
PI6_SIN = 0.5
PI6_TAN = 0.57735026919
PI3_SIN = PI6_COS


def rotor(angle):
    """uv::DRotor2::from_angle: {s: cos(a/2), bv.xy: -sin(a/2)} (positive angle = counter-clockwise)."""
    a = np.asarray(angle, dtype=np.float64)
    return np.cos(a / 2.0), -np.sin(a / 2.0)


# ---- Collider::new_* : return (kind, p0, p1, circle_r) arrays ------------------------------------
def circle(r):
    r = np.asarray(r, dtype=np.float64)
    z = np.zeros_like(r)
    return np.full(r.shape, abi.POLY_POINT, np.uint32), z, z.copy(), r


def rounded_rect(w, h, r):
    w, h, r = np.broadcast_arrays(np.asarray(w, np.float64), np.asarray(h, np.float64), np.asarray(r, np.float64))
    hw = np.maximum(w / 2.0 - r, 0.05)
    hh = np.maximum(h / 2.0 - r, 0.05)
    cr = np.minimum(np.minimum(r, w / 2.0), h / 2.0)
    return np.full(w.shape, abi.POLY_RECT, np.uint32), hw, hh, cr


def capsule(length, r):
    length, r = np.broadcast_arrays(np.asarray(length, np.float64), np.asarray(r, np.float64))
    return np.full(length.shape, abi.POLY_SEGMENT, np.uint32), length / 2.0, np.zeros_like(length), r.copy()


def hexagon(outer_r, r=0.0):
    outer_r, r = np.broadcast_arrays(np.asarray(outer_r, np.float64), np.asarray(r, np.float64))
    return np.full(outer_r.shape, abi.POLY_HEXAGON, np.uint32), outer_r.copy(), np.zeros_like(outer_r), r.copy()


def triangle(outer_r, r=0.0):
    outer_r, r = np.broadcast_arrays(np.asarray(outer_r, np.float64), np.asarray(r, np.float64))
    return np.full(outer_r.shape, abi.POLY_TRIANGLE, np.uint32), outer_r.copy(), np.zeros_like(outer_r), r.copy()


# ---- mass properties (collider.rs:222-337, 440-458) ----------------------------------------------
def _m_circle(r):
    return (np.pi / 2.0) * r ** 4


def _m_rect(hw, hh):
    return (4.0 / 3.0) * (hw ** 3 * hh + hw * hh ** 3)


def area(kind, p0, p1, cr):
    kind = np.asarray(kind)
    p0, p1, cr = (np.asarray(a, np.float64) for a in (p0, p1, cr))
    poly = np.select([kind == abi.POLY_RECT, kind == abi.POLY_TRIANGLE, kind == abi.POLY_HEXAGON],
                     [4.0 * p0 * p1, 3.0 * 0.25 * p0 * p0 / PI6_TAN, 3.0 * p0 * PI6_COS * p0], 0.0)
    side = np.select([kind == abi.POLY_SEGMENT, kind == abi.POLY_RECT, kind == abi.POLY_TRIANGLE, kind == abi.POLY_HEXAGON],
                     [4.0 * p0, 4.0 * (p0 + p1), 3.0 * p0 / PI6_TAN, 6.0 * p0], 0.0)
    return np.where(cr == 0.0, poly, poly + np.pi * cr * cr + side * cr)


def second_moment_of_area(kind, p0, p1, cr):
    kind = np.asarray(kind)
    p0, p1, cr = (np.asarray(a, np.float64) for a in (p0, p1, cr))
    point = _m_circle(cr)
    segment = _m_rect(p0, cr) + (_m_circle(cr) + (np.pi * cr ** 2 * p0 ** 2))
    caps_pt = lambda d2: _m_circle(cr) + (np.pi * cr ** 2) * d2  # noqa: E731
    # rect
    rect_poly = _m_rect(p0, p1)
    horiz = _m_rect(p0, cr / 2.0) + (2.0 * p0 * cr) * p1 ** 2
    vert = _m_rect(p1, cr / 2.0) + (2.0 * p1 * cr) * p0 ** 2
    rect_exp = 2.0 * (horiz + vert) + caps_pt(p0 ** 2 + p1 ** 2)
    # triangle (collider.rs "magic" 0.03608)
    tlen = p0 / PI6_TAN
    tri_poly = 0.03608 * tlen ** 4
    tri_exp = 3.0 * (_m_rect(tlen / 2.0, cr / 2.0) + (tlen * cr) * (p0 / 2.0) ** 2) + caps_pt(p0 ** 2)
    # hexagon
    hex_poly = (5.0 * SQRT_3 / 8.0) * p0 ** 4
    hex_exp = 6.0 * (_m_rect(p0 / 2.0, cr / 2.0) + (p0 * cr) * (PI6_COS * p0) ** 2) + caps_pt(p0 ** 2)
    rounded = cr != 0.0
    return np.select(
        [kind == abi.POLY_POINT, kind == abi.POLY_SEGMENT, kind == abi.POLY_RECT, kind == abi.POLY_TRIANGLE],
        [point, segment, np.where(rounded, rect_poly + rect_exp, rect_poly), np.where(rounded, tri_poly + tri_exp, tri_poly)],
        np.where(rounded, hex_poly + hex_exp, hex_poly))


# ---- record fillers ------------------------------------------------------------------------------
def fill_colliders(out, shape, x=0.0, y=0.0, angle=0.0, body=-1, layer=0, domain=0, sensor=False,
                   static_friction=1.6, dynamic_friction=1.5, restitution=0.0):
    """Fill abi.collider_dtype rows. static/dynamic friction None = Option::None (collider.rs:904-911)."""
    kind, p0, p1, cr = shape
    out["kind"], out["p0"], out["p1"], out["circle_r"] = kind, p0, p1, cr
    s, b = rotor(angle)
    out["x"], out["y"], out["rot_s"], out["rot_b"] = x, y, s, b
    out["body"], out["layer"], out["domain"] = body, layer, domain
    flags = abi.COLL_ALIVE | (abi.COLL_SENSOR if sensor else 0)
    if static_friction is not None:
        flags |= abi.COLL_HAS_STATIC_FRICTION
        out["static_friction"] = static_friction
    if dynamic_friction is not None:
        flags |= abi.COLL_HAS_DYNAMIC_FRICTION
        out["dynamic_friction"] = dynamic_friction
    out["restitution"] = restitution
    out["flags"] = flags
    return out


def fill_bodies_dynamic(out, area_, moment, density, x, y, angle=0.0, vx=0.0, vy=0.0, w=0.0):
    """Body::new_dynamic(ColliderInfo{area, second_moment_of_area}, density) (body.rs:36-45)."""
    mass = np.asarray(area_, np.float64) * density
    inertia = np.asarray(moment, np.float64) * density
    s, b = rotor(angle)
    out["x"], out["y"], out["rot_s"], out["rot_b"] = x, y, s, b
    out["vx"], out["vy"], out["w"] = vx, vy, w
    out["inv_mass"], out["inv_inertia"] = 1.0 / mass, 1.0 / inertia
    out["flags"] = abi.BODY_ALIVE | abi.BODY_MASS_FINITE | abi.BODY_INERTIA_FINITE
    return out


def fill_bodies_particle(out, mass, x, y):
    """Body::new_particle (body.rs:23-31): finite mass, infinite moment of inertia."""
    out["x"], out["y"], out["rot_s"], out["rot_b"] = x, y, 1.0, 0.0
    out["inv_mass"] = 1.0 / np.asarray(mass, np.float64)
    out["flags"] = abi.BODY_ALIVE | abi.BODY_MASS_FINITE
    return out


def fill_bodies_kinematic(out, x, y, angle=0.0, vx=0.0, vy=0.0, w=0.0):
    """Body::new_kinematic (body.rs:62-70)."""
    s, b = rotor(angle)
    out["x"], out["y"], out["rot_s"], out["rot_b"] = x, y, s, b
    out["vx"], out["vy"], out["w"] = vx, vy, w
    out["flags"] = abi.BODY_ALIVE
    return out


def compound_setup(shapes, offsets):
    """CompoundColliderSetup for ONE body: shapes = list of (kind, p0, p1, cr) scalars, offsets = list of (x, y).
    Returns (center_of_mass, area, second_moment) — compound_collider.rs:17-43 (sic: centre / count)."""
    areas = np.array([float(area(*s)) for s in shapes])
    moms = np.array([float(second_moment_of_area(*s)) for s in shapes])
    offs = np.asarray(offsets, np.float64).reshape(-1, 2)
    com = (areas[:, None] * offs).sum(0) / len(shapes)
    d2 = ((offs - com) ** 2).sum(1)
    return com, areas.sum(), (moms + areas * d2).sum()

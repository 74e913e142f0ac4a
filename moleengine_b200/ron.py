"""Scene ingestion: the reference sandbox's RON scenes -> the slot-indexed arrays of the C ABI.

Follows /root/reference/examples/sandbox/recipes.rs (`Recipe::spawn`, :204-520) and scenes/*.ron: `Block`, `Ball`,
`Capsule`, `GenericBody` (compound colliders through `CompoundColliderSetup`), `Blockchain` (capsule links tied by
attachment constraints), `Oscillator`, `RopeConnectedBlocks` (`Rope::spawn_line` + attachments) and `Player` (a particle
body with a frictionless capsule, player.rs:42-65). `Background*` recipes carry no physics and are skipped.
Bodies and colliders get slots in spawn order, exactly as a fresh `EntitySet` hands them out.

Only the physics side of a recipe is restated (no meshes, materials or ECS); poses go through f32 like the reference's
`Pose` component does (math.rs:124-129) wherever the recipe builds an `sf::Pose`.
"""
import math
import re

import numpy as np

from . import abi, shapes


# ---------------------------------------------------------------- a small RON reader
class Variant:
    """`Name(payload...)` or a bare `Name`."""

    def __init__(self, name, args):
        self.name, self.args = name, args

    def __repr__(self):
        return f"{self.name}{self.args!r}"


_TOKEN = re.compile(r"\s*(?:(//[^\n]*|/\*.*?\*/)|([A-Za-z_][A-Za-z0-9_]*)|(-?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?)|(\"(?:[^\"\\]|\\.)*\")|(.))", re.S)


def _tokens(text):
    pos, out = 0, []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            break
        pos = m.end()
        if m.group(1) is not None:
            continue
        if m.group(2) is not None:
            out.append(("id", m.group(2)))
        elif m.group(3) is not None:
            out.append(("num", float(m.group(3))))
        elif m.group(4) is not None:
            out.append(("str", m.group(4)[1:-1]))
        elif m.group(5) is not None and not m.group(5).isspace():
            out.append(("p", m.group(5)))
    return out


class _Parser:
    def __init__(self, text):
        self.t, self.i = _tokens(text), 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", None)

    def take(self, kind=None, val=None):
        tok = self.peek()
        if (kind and tok[0] != kind) or (val is not None and tok[1] != val):
            raise ValueError(f"RON: expected {val or kind}, got {tok} at token {self.i}")
        self.i += 1
        return tok

    def value(self):
        kind, val = self.peek()
        if kind == "num" or kind == "str":
            self.i += 1
            return val
        if kind == "id":
            self.i += 1
            if val in ("true", "false"):
                return val == "true"
            if self.peek() == ("p", "("):
                return Variant(val, self.paren())
            return Variant(val, None)
        if (kind, val) == ("p", "("):
            return self.paren()
        if (kind, val) == ("p", "["):
            self.i += 1
            items = []
            while self.peek() != ("p", "]"):
                items.append(self.value())
                if self.peek() == ("p", ","):
                    self.i += 1
            self.take("p", "]")
            return items
        raise ValueError(f"RON: unexpected token {kind} {val!r}")

    def paren(self):
        """( a: 1, b: 2 ) -> dict ; ( 1, 2 ) -> list ; ( x ) -> [x]"""
        self.take("p", "(")
        if self.peek()[0] == "id" and self.peek(1) == ("p", ":"):
            out = {}
            while self.peek() != ("p", ")"):
                name = self.take("id")[1]
                self.take("p", ":")
                out[name] = self.value()
                if self.peek() == ("p", ","):
                    self.i += 1
            self.take("p", ")")
            return out
        items = []
        while self.peek() != ("p", ")"):
            items.append(self.value())
            if self.peek() == ("p", ","):
                self.i += 1
        self.take("p", ")")
        return items


def parse(text):
    return _Parser(text).value()


# ---------------------------------------------------------------- recipe helpers
def _f32(x):
    return float(np.float32(x))


def _xy(v, default=(0.0, 0.0)):
    if v is None:
        return default
    if isinstance(v, dict):
        return (float(v.get("x", 0.0)), float(v.get("y", 0.0)))
    return (float(v[0]), float(v[1]))


def _angle(v):
    """sf::Angle: Deg(x) | Rad(x) | bare number (radians)."""
    if v is None:
        return 0.0
    if isinstance(v, Variant):
        a = float(v.args[0])
        return math.radians(_f32(a)) if v.name == "Deg" else a
    return float(v)


def _pose(p):
    """PoseBuilder / serde_pose: ( position: (x, y), rotation: Deg(a) ) -> (x, y, angle), through f32 like sf::Pose."""
    p = p or {}
    x, y = _xy(p.get("position"))
    return _f32(x), _f32(y), _f32(_angle(p.get("rotation")))


def _unwrap(v):
    """Newtype variants are written `Block (( ... ))`: one-element list around the struct."""
    while isinstance(v, list) and len(v) == 1:
        v = v[0]
    return v if v is not None else {}


def _polygon(v):
    """ColliderPolygon -> (kind, p0, p1)."""
    if v is None:
        return abi.POLY_POINT, 0.0, 0.0
    a = _unwrap(v.args) if v.args is not None else {}
    get = (lambda k: float(a[k])) if isinstance(a, dict) else None
    if v.name == "Point":
        return abi.POLY_POINT, 0.0, 0.0
    if v.name == "LineSegment":
        return abi.POLY_SEGMENT, get("hl"), 0.0
    if v.name == "Rect":
        return abi.POLY_RECT, get("hw"), get("hh")
    if v.name == "Triangle":
        return abi.POLY_TRIANGLE, get("outer_r"), 0.0
    if v.name == "Hexagon":
        return abi.POLY_HEXAGON, get("outer_r"), 0.0
    raise ValueError(f"unknown ColliderPolygon {v.name}")


class SceneBuilder:
    """Accumulates bodies / colliders / constraints / ropes in spawn order (fresh arenas: slot = insertion index)."""

    def __init__(self):
        self.bodies, self.colliders, self.constraints, self.ropes, self.rope_particles = [], [], [], [], []

    # -- EntitySet ---------------------------------------------------------------------------------
    def insert_body(self, kind, x, y, angle, area=0.0, moment=0.0, density=0.5, mass=0.0, vx=0.0, vy=0.0):
        b = np.zeros(1, abi.body_dtype)
        if kind == "dynamic":
            shapes.fill_bodies_dynamic(b, area, moment, density, x, y, angle, vx, vy)
        elif kind == "particle":
            shapes.fill_bodies_particle(b, mass, x, y)
            b["rot_s"], b["rot_b"] = shapes.rotor(angle)
        self.bodies.append(b)
        return len(self.bodies) - 1

    def insert_collider(self, shape, x=0.0, y=0.0, angle=0.0, body=-1, layer=0, sensor=False, static_friction=1.6, dynamic_friction=1.5,
                        restitution=0.0):
        c = np.zeros(1, abi.collider_dtype)
        sh = tuple(np.asarray([v]) for v in shape)
        shapes.fill_colliders(c, (sh[0].astype(np.uint32), sh[1], sh[2], sh[3]), x, y, angle, body, layer, 0, sensor, static_friction,
                              dynamic_friction, restitution)
        self.colliders.append(c)
        return len(self.colliders) - 1

    def insert_constraint(self, owner, target=-1, origin=(0.0, 0.0), target_origin=(0.0, 0.0), compliance=0.0, linear_damping=0.1,
                          angular_damping=0.0, distance=0.0, limit=abi.LIMIT_EQ, can_sleep=True):
        """ConstraintBuilder::new(owner)...build_attachment() / build_distance(d) (constraint.rs:79-176)."""
        k = np.zeros(1, abi.constraint_dtype)
        k["owner"], k["target"] = owner, target
        k["off0x"], k["off0y"], k["off1x"], k["off1y"] = origin[0], origin[1], target_origin[0], target_origin[1]
        k["compliance"], k["linear_damping"], k["angular_damping"], k["distance"] = compliance, linear_damping, angular_damping, distance
        k["flags"] = limit | (abi.CONSTRAINT_CAN_SLEEP if can_sleep else 0)
        self.constraints.append(k)

    # -- recipes.rs helpers --------------------------------------------------------------------------
    def spawn_block(self, blk):
        """recipes.rs:131-152. Returns the body slot or None for a static block."""
        blk = _unwrap(blk)
        w, h, r = float(blk.get("width", 1.0)), float(blk.get("height", 1.0)), float(blk.get("radius", 0.0))
        x, y, ang = _pose(blk.get("pose"))
        sh = tuple(float(np.asarray(v)) for v in shapes.rounded_rect(w, h, r))
        if blk.get("is_static", False):
            self.insert_collider(sh, x, y, ang)
            return None
        coll = self.insert_collider(sh)
        body = self.insert_body("dynamic", x, y, ang, float(shapes.area(*sh)), float(shapes.second_moment_of_area(*sh)), 0.5)
        self.colliders[coll]["body"] = body
        return body

    def spawn_static(self, pose, colls):
        """recipes.rs:161-180: coll.with_pose(solid.pose) — the collider's own pose is REPLACED by the solid's."""
        x, y, ang = pose
        for c in colls:
            self.insert_collider(c["shape"], x, y, ang, restitution=c.get("restitution", 0.0), sensor=c.get("sensor", False))

    def spawn_body(self, pose, colls, vel=(0.0, 0.0)):
        """recipes.rs:182-214 + compound_collider.rs:17-43 (centre of mass divided by the collider COUNT)."""
        x, y, ang = pose
        com, area, moment = shapes.compound_setup([c["shape"] for c in colls], [c.get("offset", (0.0, 0.0)) for c in colls])
        body = self.insert_body("dynamic", x, y, ang, area, moment, 0.5, vx=vel[0], vy=vel[1])
        for c in colls:
            ox, oy = c.get("offset", (0.0, 0.0))
            self.insert_collider(c["shape"], ox - com[0], oy - com[1], c.get("angle", 0.0), body, restitution=c.get("restitution", 0.0),
                                 sensor=c.get("sensor", False))
        return body

    def spawn_rope_line(self, start, end):
        """Rope::spawn_line with RopeParameters::default (rope.rs:27-43, 61-141). Returns (first particle body, last)."""
        spacing, thickness, mass = 0.1, 0.12, 0.02
        dx, dy = end[0] - start[0], end[1] - start[1]
        dist = math.hypot(dx, dy)
        segs = int(round(dist / spacing))
        count = segs + 1
        spacing = dist / segs
        sx, sy = spacing * dx / dist, spacing * dy / dist
        first_particle = len(self.rope_particles)
        px, py = start
        bodies = []
        for _ in range(count):
            b = self.insert_body("particle", px, py, 0.0, mass=mass)
            self.insert_collider(tuple(float(np.asarray(v)) for v in shapes.circle(thickness / 2.0)), body=b, layer=abi.ROPE_LAYER,
                                 static_friction=None, dynamic_friction=1.5)
            bodies.append(b)
            self.rope_particles.append(b)
            px += sx; py += sy
        r = np.zeros(1, abi.rope_dtype)
        r["spacing"], r["compliance"], r["bending_max_angle"] = spacing, 0.0000001, np.float32(np.deg2rad(np.float32(30.0)))
        r["bending_compliance"], r["damping"] = 0.2, 20.0
        r["first_particle"], r["particle_count"] = first_particle, count
        self.ropes.append(r)
        return bodies[0], bodies[-1]

    # -- Recipe::spawn --------------------------------------------------------------------------------
    def spawn(self, rec):
        name, a = rec.name, rec.args
        if name.startswith("Background"):
            return
        if name == "Player":                                  # player.rs:42-65: particle body, frictionless capsule, rotated 90 deg
            p = _unwrap(a)
            x, y = _xy(p.get("position"))
            b = self.insert_body("particle", _f32(x), _f32(y), _f32(math.radians(90.0)), mass=1.0)
            self.insert_collider(tuple(float(np.asarray(v)) for v in shapes.capsule(0.2, 0.1)), body=b, static_friction=None, dynamic_friction=None)
        elif name == "Block":
            self.spawn_block(a)
        elif name == "Ball":
            p = _unwrap(a)
            x, y = _xy(p.get("position"))
            coll = {"shape": tuple(float(np.asarray(v)) for v in shapes.circle(_f32(p.get("radius", 1.0)))), "restitution": float(p.get("restitution", 0.0))}
            pose = (_f32(x), _f32(y), 0.0)
            if p.get("is_static", False):
                self.spawn_static(pose, [coll])
            else:
                self.spawn_body(pose, [coll], _xy(p.get("start_velocity")))
        elif name == "Capsule":
            p = _unwrap(a)
            coll = {"shape": tuple(float(np.asarray(v)) for v in shapes.capsule(float(p.get("length", 1.0)), float(p.get("radius", 0.5))))}
            if p.get("is_static", False):
                self.spawn_static(_pose(p.get("pose")), [coll])
            else:
                self.spawn_body(_pose(p.get("pose")), [coll])
        elif name == "GenericBody":
            colls = []
            for c in a.get("colliders", []):
                c = _unwrap(c)
                shp = c.get("shape", {})
                kind, p0, p1 = _polygon(shp.get("polygon")) if "shape" in c else (abi.POLY_POINT, 0.0, 0.0)
                cr = float(shp.get("circle_r", 0.0)) if "shape" in c else 1.0       # Collider::default is a unit circle
                cp = c.get("pose") or {}
                ox, oy = _xy(cp.get("position"))
                ty = c.get("ty")
                colls.append({"shape": (kind, p0, p1, cr), "offset": (ox, oy), "angle": _angle(cp.get("rotation")),
                              "sensor": isinstance(ty, Variant) and ty.name == "Sensor"})
            self.spawn_body(_pose(a.get("pose")), colls)
        elif name == "Blockchain":                             # recipes.rs:283-365
            links = [(float(p[0]), float(p[1])) for p in a["links"]]
            if len(links) < 2:
                return
            width, spacing = float(a["width"]), float(a["spacing"])
            half_spacing, radius = spacing / 2.0, width / 2.0
            prev = None
            for (l1, l2) in zip(links[:-1], links[1:]):
                dx, dy = l2[0] - l1[0], l2[1] - l1[1]
                dist = math.hypot(dx, dy)
                cx, cy = (l1[0] + l2[0]) / 2.0, (l1[1] + l2[1]) / 2.0
                orientation = math.acos(dx / dist) * math.copysign(1.0, dy)
                full = dist - spacing
                coll = {"shape": tuple(float(np.asarray(v)) for v in shapes.capsule(full - width, radius))}
                cap = self.spawn_body((_f32(cx), _f32(cy), _f32(_f32(orientation))), [coll])
                half = full / 2.0
                if prev is not None:
                    self.insert_constraint(cap, prev[0], (-half - half_spacing, 0.0), (prev[1], 0.0), compliance=0.015)
                elif a.get("anchored_start", False):
                    self.insert_constraint(cap, -1, (-half - half_spacing, 0.0), l1)
                prev = (cap, half + half_spacing)
            if a.get("anchored_end", False):
                self.insert_constraint(prev[0], -1, (prev[1] + spacing / 2.0, 0.0), links[-1])
        elif name == "Oscillator":                             # recipes.rs:366-405
            x, y = _xy(a.get("position"))
            off = _f32(a["begin_length"]) / 2.0
            b1 = self.spawn_block({"pose": {"position": (_f32(x) + off, _f32(y))}})
            b2 = self.spawn_block({"pose": {"position": (_f32(x) - off, _f32(y))}})
            self.insert_constraint(b1, b2, compliance=float(a["compliance"]), linear_damping=0.0, distance=float(a["target_length"]))
        elif name == "RopeConnectedBlocks":                    # recipes.rs:406-481
            blk1, blk2 = _unwrap(a["block1"]), _unwrap(a["block2"])
            o1, o2 = _xy(a["offset1"]), _xy(a["offset2"])
            b1 = self.spawn_block(blk1)
            b2 = self.spawn_block(blk2)

            def world(blk, off):
                x, y, ang = _pose(blk.get("pose"))
                c, s = math.cos(ang), math.sin(ang)
                return (_f32(x + c * _f32(off[0]) - s * _f32(off[1])), _f32(y + s * _f32(off[0]) + c * _f32(off[1])))
            e1, e2 = world(blk1, o1), world(blk2, o2)
            first, last = self.spawn_rope_line(e1, e2)
            if b1 is not None:
                self.insert_constraint(first, b1, target_origin=o1)
            else:
                self.insert_constraint(first, -1, target_origin=e1)
            if b2 is not None:
                self.insert_constraint(last, b2, target_origin=o2)
            else:
                self.insert_constraint(last, -1, target_origin=e2)
        else:
            raise ValueError(f"unknown recipe {name}")

    def scene(self, name, gravity=(0.0, -9.81)):
        cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)  # noqa: E731
        return {"name": name, "bodies": cat(self.bodies, abi.body_dtype), "colliders": cat(self.colliders, abi.collider_dtype),
                "constraints": cat(self.constraints, abi.constraint_dtype), "ropes": cat(self.ropes, abi.rope_dtype),
                "rope_particles": np.asarray(self.rope_particles, np.int32), "substeps": 10, "gravity": gravity}


def load_scene_text(text, name="ron"):
    """Scene file (main.rs:210-230): ( gravity: (gx, gy), spawn_zone: .., recipes: [ .. ] ); gravity defaults to (0, -9.81)."""
    doc = parse(text)
    b = SceneBuilder()
    for rec in doc.get("recipes", []):
        b.spawn(rec)
    return b.scene(name, _xy(doc.get("gravity"), (0.0, -9.81)))


def load_scene(path):
    import os
    with open(path) as f:
        return load_scene_text(f.read(), os.path.splitext(os.path.basename(path))[0])

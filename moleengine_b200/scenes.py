"""Deterministic synthetic scenes for the BASELINE.json configs, built through the same
constructors a Starframe user calls (shapes.py). Each builder returns a dict of ABI record arrays:
bodies, colliders, constraints, ropes, rope_particles, plus `substeps` and a `name`.

PRNG: splitmix64 counters (seeded, vectorised), so scenes are identical on every machine.
"""
import numpy as np

from . import abi, shapes

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(seed, stream, n):
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03)
             + (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform(seed, stream, n, lo=0.0, hi=1.0):
    u = (_splitmix(seed, stream, n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return lo + (hi - lo) * u


def _scene(name, bodies, colliders, substeps, constraints=None, ropes=None, rope_particles=None):
    return {
        "name": name, "bodies": bodies, "colliders": colliders, "substeps": substeps,
        "constraints": np.zeros(0, abi.constraint_dtype) if constraints is None else constraints,
        "ropes": np.zeros(0, abi.rope_dtype) if ropes is None else ropes,
        "rope_particles": np.zeros(0, np.int32) if rope_particles is None else np.ascontiguousarray(rope_particles, np.int32),
    }


# ---- C1: headless sandbox-style scene (examples/sandbox/scenes/test.ron) -------------------------
def c1_sandbox_stack():
    nb = 120
    bodies = np.zeros(nb, abi.body_dtype)
    colliders = np.zeros(nb + 1, abi.collider_dtype)
    # static ground: Block(width 40, height 0.2, static) -> collider only (recipes.rs:131-153)
    shapes.fill_colliders(colliders[0:1], shapes.rounded_rect(40.0, 0.2, 0.0), x=0.0, y=-5.0)
    # 10x10 boxes, Body::new_dynamic(coll.info(), 0.5)
    i, j = np.meshgrid(np.arange(10), np.arange(10), indexing="xy")
    bx, by = ((i - 4.5) * 1.05).ravel(), (-4.4 + j).ravel()
    box = shapes.rounded_rect(np.full(100, 1.0), 1.0, 0.0)
    shapes.fill_bodies_dynamic(bodies[:100], shapes.area(*box), shapes.second_moment_of_area(*box), 0.5, bx, by)
    shapes.fill_colliders(colliders[1:101], box, body=np.arange(100))
    # 20 circles r = 0.5 through spawn_body (single-collider "compound": centre of mass = area * 0 / 1)
    k = np.arange(20)
    cx, cy = (k % 10 - 4.5) * 1.05, 6.1 + (k // 10) * 1.1
    circ = shapes.circle(np.full(20, 0.5))
    shapes.fill_bodies_dynamic(bodies[100:], shapes.area(*circ), shapes.second_moment_of_area(*circ), 0.5, cx, cy)
    shapes.fill_colliders(colliders[101:], circ, body=np.arange(100, 120))
    return _scene("c1_sandbox_stack", bodies, colliders, 10)


# ---- C2: circles falling into a static U-shaped container ----------------------------------------
def c2_circles(nx=500, ny=200, seed=0xC2):
    n = nx * ny
    bodies = np.zeros(n, abi.body_dtype)
    colliders = np.zeros(n + 3, abi.collider_dtype)
    half_w = nx * 0.65 / 2.0 + 7.5
    shapes.fill_colliders(colliders[0:1], shapes.rounded_rect(2.0 * half_w, 0.2, 0.0), x=0.0, y=-0.1)
    shapes.fill_colliders(colliders[1:2], shapes.rounded_rect(0.2, ny * 0.65 + 70.0, 0.0), x=-half_w, y=(ny * 0.65 + 70.0) / 2.0)
    shapes.fill_colliders(colliders[2:3], shapes.rounded_rect(0.2, ny * 0.65 + 70.0, 0.0), x=half_w, y=(ny * 0.65 + 70.0) / 2.0)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    x = (i.ravel() - (nx - 1) / 2.0) * 0.65 + uniform(seed, 1, n, -0.02, 0.02)
    y = 0.5 + j.ravel() * 0.65 + uniform(seed, 2, n, -0.02, 0.02)
    circ = shapes.circle(uniform(seed, 3, n, 0.2, 0.3))
    shapes.fill_bodies_dynamic(bodies, shapes.area(*circ), shapes.second_moment_of_area(*circ), 0.5, x, y)
    shapes.fill_colliders(colliders[3:], circ, body=np.arange(n))
    return _scene(f"c2_circles_{n}", bodies, colliders, 4)


def star(n_small=150, big_r=10.0, small_r=0.3):
    """One large dynamic disc ringed by touching small ones: every pair shares the big body, so the colouring needs
    n_small colours (more than the 64-colour window of the device kernel) and the big body's island holds everything."""
    n = n_small + 1
    bodies = np.zeros(n, abi.body_dtype)
    colliders = np.zeros(n, abi.collider_dtype)
    ang = np.arange(n_small) * (2.0 * np.pi / n_small)
    x = np.concatenate([[0.0], (big_r + small_r * 0.9) * np.cos(ang)])
    y = np.concatenate([[0.0], (big_r + small_r * 0.9) * np.sin(ang)])
    circ = shapes.circle(np.concatenate([[big_r], np.full(n_small, small_r)]))
    shapes.fill_bodies_dynamic(bodies, shapes.area(*circ), shapes.second_moment_of_area(*circ), 0.5, x, y)
    shapes.fill_colliders(colliders, circ, body=np.arange(n))
    return _scene(f"star_{n_small}", bodies, colliders, 4)


# ---- mixed bodies (C3 / C5 body mix) -------------------------------------------------------------
def _mixed_bodies(n, seed, x, y, domain=0, body_base=0):
    """25% Rect, 35% rounded Rect, 10% Hexagon, 10% Triangle (half of each rounded r = 0.1), 5% capsule,
    5% circle, 10% compounds of 2-3 colliders (a bar + round ends, after compound_colliders.ron:33-43).
    Returns (bodies[n], colliders[m]) with collider.body = body_base + local index."""
    u = uniform(seed, 10, n)
    a = uniform(seed, 11, n, 0.3, 0.6)
    b = uniform(seed, 12, n, 0.3, 0.6)
    r = uniform(seed, 13, n, 0.05, 0.15)
    hx = uniform(seed, 14, n, 0.4, 0.7)
    tr = uniform(seed, 15, n, 0.5, 0.8)
    hl = uniform(seed, 16, n, 0.3, 0.5)
    cr = uniform(seed, 17, n, 0.3, 0.5)
    half = uniform(seed, 18, n) < 0.5
    ang = uniform(seed, 19, n, -np.pi, np.pi)
    three = uniform(seed, 20, n) < 0.5

    kind = np.zeros(n, np.uint32)
    p0 = np.zeros(n)
    p1 = np.zeros(n)
    crr = np.zeros(n)
    is_rect = u < 0.25
    is_rrect = (u >= 0.25) & (u < 0.60)
    is_hex = (u >= 0.60) & (u < 0.70)
    is_tri = (u >= 0.70) & (u < 0.80)
    is_cap = (u >= 0.80) & (u < 0.85)
    is_circ = (u >= 0.85) & (u < 0.90)
    is_comp = u >= 0.90
    kind[is_rect | is_rrect | is_comp] = abi.POLY_RECT
    p0[is_rect | is_rrect] = a[is_rect | is_rrect]
    p1[is_rect | is_rrect] = b[is_rect | is_rrect]
    crr[is_rrect] = r[is_rrect]
    kind[is_hex] = abi.POLY_HEXAGON
    p0[is_hex] = hx[is_hex]
    crr[is_hex & half] = 0.1
    kind[is_tri] = abi.POLY_TRIANGLE
    p0[is_tri] = tr[is_tri]
    crr[is_tri & half] = 0.1
    kind[is_cap] = abi.POLY_SEGMENT
    p0[is_cap] = hl[is_cap]
    crr[is_cap] = 0.15
    kind[is_circ] = abi.POLY_POINT
    crr[is_circ] = cr[is_circ]
    # compounds: bar Rect{hw = a, hh = 0.12} at local origin + Hexagon{0.2, r 0.1} at (+-(a + 0.1), 0)
    p0[is_comp] = a[is_comp]
    p1[is_comp] = 0.12

    # single-collider mass properties
    ar = shapes.area(kind, p0, p1, crr)
    mo = shapes.second_moment_of_area(kind, p0, p1, crr)
    # compound mass properties, vectorised CompoundColliderSetup (sic: centre of mass / collider COUNT)
    hex_shape = (np.full(n, abi.POLY_HEXAGON, np.uint32), np.full(n, 0.2), np.zeros(n), np.full(n, 0.1))
    a_hex = shapes.area(*hex_shape)
    m_hex = shapes.second_moment_of_area(*hex_shape)
    off = a + 0.1
    n_sub = np.where(three, 3, 2)
    # colliders: bar at 0, hex at +off, (hex at -off if three)
    com_x = np.where(three, 0.0, a_hex * off) / n_sub
    comp_area = ar + a_hex * (n_sub - 1)
    comp_mom = (mo + ar * com_x ** 2) + (m_hex + a_hex * (off - com_x) ** 2) + np.where(three, m_hex + a_hex * (-off - com_x) ** 2, 0.0)
    ar = np.where(is_comp, comp_area, ar)
    mo = np.where(is_comp, comp_mom, mo)

    bodies = np.zeros(n, abi.body_dtype)
    shapes.fill_bodies_dynamic(bodies, ar, mo, 0.5, x, y, ang)

    n_extra = int((is_comp * (n_sub - 1)).sum())
    colliders = np.zeros(n + n_extra, abi.collider_dtype)
    body_idx = body_base + np.arange(n)
    # primary collider of every body (compound: the bar, shifted by -com)
    shapes.fill_colliders(colliders[:n], (kind, p0, p1, crr), x=np.where(is_comp, -com_x, 0.0), y=0.0, body=body_idx, domain=domain)
    # extra colliders of compounds
    ci = np.nonzero(is_comp)[0]
    e1 = colliders[n:n + len(ci)]
    shapes.fill_colliders(e1, tuple(s[ci] for s in hex_shape), x=(off - com_x)[ci], y=0.0, body=body_idx[ci], domain=domain)
    c3 = np.nonzero(is_comp & three)[0]
    e2 = colliders[n + len(ci):]
    shapes.fill_colliders(e2, tuple(s[c3] for s in hex_shape), x=(-off - com_x)[c3], y=0.0, body=body_idx[c3], domain=domain)
    return bodies, colliders


# ---- C3: mixed bodies in a large world -------------------------------------------------------------
def c3_mixed(nx=1250, ny=800, seed=0xC3, rows_per_shelf=8):
    """nx*ny bodies of the mix on a 1.6 m grid, centred on the origin, resting space broken up by static shelves:
    strips of 100 x 0.2 under every `rows_per_shelf` rows (so piles form and every island stays bounded)."""
    n = nx * ny
    pitch = 1.6
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    shelf = j.ravel() // rows_per_shelf
    x = (i.ravel() - (nx - 1) / 2.0) * pitch + uniform(seed, 1, n, -0.05, 0.05)
    y = (j.ravel() - (ny - 1) / 2.0) * pitch + shelf * 1.0 + uniform(seed, 2, n, -0.05, 0.05)
    bodies, dyn_colls = _mixed_bodies(n, seed, x, y)
    n_shelves = (ny + rows_per_shelf - 1) // rows_per_shelf
    width = nx * pitch + 4.0
    per_row = int(np.ceil(width / 100.0))
    sx = (np.arange(per_row) - (per_row - 1) / 2.0) * 100.0
    sy = (np.arange(n_shelves) * rows_per_shelf - (ny - 1) / 2.0) * pitch + np.arange(n_shelves) * 1.0 - 1.1
    gx, gy = np.meshgrid(sx, sy, indexing="xy")
    statics = np.zeros(gx.size, abi.collider_dtype)
    shapes.fill_colliders(statics, shapes.rounded_rect(np.full(gx.size, 100.0), 0.2, 0.0), x=gx.ravel(), y=gy.ravel())
    colliders = np.concatenate([statics, dyn_colls])
    return _scene(f"c3_mixed_{n}", bodies, colliders, 4)


# ---- C4: rope-heavy scene --------------------------------------------------------------------------
def rope_params(n=1, spacing=0.1):
    """RopeParameters::default (rope.rs:27-43)."""
    r = np.zeros(n, abi.rope_dtype)
    r["spacing"], r["compliance"], r["bending_max_angle"] = spacing, 0.0000001, np.float32(np.deg2rad(np.float32(30.0)))
    r["bending_compliance"], r["damping"] = 0.2, 20.0
    return r


def c4_ropes(n_ropes=10000, particles_per_rope=64, n_free_blocks=30000, seed=0xC4):
    """RopeConnectedBlocks (recipes.rs:406-481) x n_ropes: Rope::spawn_line between two 1x1 dynamic blocks with an
    attachment constraint at each end, laid out on a grid of cells 9 m x 4 m; free 0.5 x 0.5 blocks dropped from above."""
    ppr = particles_per_rope
    seg = ppr - 1
    spacing = 0.1
    cols = int(np.ceil(np.sqrt(n_ropes * 4.0 / 9.0))) or 1
    r = np.arange(n_ropes)
    ox = (r % cols - (cols - 1) / 2.0) * 9.0
    oy = (r // cols) * 4.0
    length = seg * spacing
    n_blocks = 2 * n_ropes
    n_part = n_ropes * ppr
    nb = n_blocks + n_part + n_free_blocks
    bodies = np.zeros(nb, abi.body_dtype)
    # blocks: rope ends sit 0.55 from the block centres like scenes/ropes.ron (offset1 (0.55, ..), offset2 (-0.55, ..))
    blk = shapes.rounded_rect(np.full(n_blocks, 1.0), 1.0, 0.0)
    bxl, bxr = ox - length / 2.0 - 0.55, ox + length / 2.0 + 0.55
    shapes.fill_bodies_dynamic(bodies[:n_blocks], shapes.area(*blk), shapes.second_moment_of_area(*blk), 0.5,
                               np.stack([bxl, bxr], 1).ravel(), np.repeat(oy, 2))
    # particles: Body::new_particle(0.02) along the line (rope.rs:113-141)
    k = np.arange(ppr)
    px = (ox[:, None] - length / 2.0 + k[None, :] * spacing).ravel()
    py = np.repeat(oy, ppr)
    shapes.fill_bodies_particle(bodies[n_blocks:n_blocks + n_part], 0.02, px, py)
    # free blocks dropped onto the ropes
    fb = shapes.rounded_rect(np.full(n_free_blocks, 0.5), 0.5, 0.0)
    fr = np.arange(n_free_blocks) % max(n_ropes, 1)
    fx = ox[fr] + uniform(seed, 1, n_free_blocks, -2.5, 2.5)
    fy = oy[fr] + 1.0 + (np.arange(n_free_blocks) // max(n_ropes, 1)) * 0.8 + uniform(seed, 2, n_free_blocks, 0.0, 0.2)
    shapes.fill_bodies_dynamic(bodies[n_blocks + n_part:], shapes.area(*fb), shapes.second_moment_of_area(*fb), 0.5, fx, fy)

    # static ground far below so blocks eventually land
    n_ground = cols
    colliders = np.zeros(n_ground + nb, abi.collider_dtype)
    gx = (np.arange(cols) - (cols - 1) / 2.0) * 9.0
    shapes.fill_colliders(colliders[:n_ground], shapes.rounded_rect(np.full(cols, 9.0), 0.2, 0.0), x=gx, y=-6.0)
    c = colliders[n_ground:]
    shapes.fill_colliders(c[:n_blocks], blk, body=np.arange(n_blocks))
    # rope colliders: circle r = thickness / 2 on ROPE_LAYER with the rope material (static friction None) (rope.rs:122-124)
    shapes.fill_colliders(c[n_blocks:n_blocks + n_part], shapes.circle(np.full(n_part, 0.06)), body=n_blocks + np.arange(n_part),
                          layer=abi.ROPE_LAYER, static_friction=None, dynamic_friction=1.5)
    shapes.fill_colliders(c[n_blocks + n_part:], fb, body=n_blocks + n_part + np.arange(n_free_blocks))

    ropes = rope_params(n_ropes, spacing)
    ropes["first_particle"] = r * ppr
    ropes["particle_count"] = ppr
    rope_particles = (n_blocks + np.arange(n_part)).astype(np.int32)
    # ConstraintBuilder::new(particle).with_target(block).with_target_origin(offset).build_attachment()
    cons = np.zeros(2 * n_ropes, abi.constraint_dtype)
    cons["owner"] = np.stack([n_blocks + r * ppr, n_blocks + r * ppr + ppr - 1], 1).ravel()
    cons["target"] = np.arange(n_blocks)
    cons["off1x"] = np.tile([0.55, -0.55], n_ropes)
    cons["linear_damping"] = 0.1
    cons["flags"] = abi.LIMIT_EQ | abi.CONSTRAINT_CAN_SLEEP
    return _scene(f"c4_ropes_{n_ropes}x{ppr}", bodies, colliders, 4, cons, ropes, rope_particles)


# ---- C5: batched independent worlds ----------------------------------------------------------------
def c5_worlds(n_worlds=8192, bodies_per_side=16, seed=0xC5, first_world=0, dense=False):
    """Each world: 4 static walls + bodies_per_side^2 bodies of the C3 mix on a grid. Worlds are independent `domain`s sharing
    coordinates; seed = 0xC5 + GLOBAL world index (first_world + i), so a rank that builds only its share of the worlds gets the
    same worlds a single GPU would.
    Default = SURVEY.md §8(d)'s C5: a 20 x 20 box (walls as scenes/minimal.ron:4-15, the box made square so that 256 bodies of
    the C3 mix — about 205 m^2 of them — have room to fall and pile up), bodies on a 1.15 m grid.
    dense=True is round 1's variant: minimal.ron's own 20 x 10 box with the rows squeezed to 0.56 m — the bodies fill the box
    completely and start out overlapping; kept as a stress case (tools/bench_configs.py c5d)."""
    bps = bodies_per_side
    per = bps * bps
    if dense:
        wall_shapes = [(20.0, 0.2, 0.0, 5.0), (20.0, 0.2, 0.0, -5.0), (0.2, 10.0, 10.0, 0.0), (0.2, 10.0, -10.0, 0.0)]
        y0, pitch_y = -4.2, 0.56
    else:
        wall_shapes = [(20.0, 0.2, 0.0, 10.0), (20.0, 0.2, 0.0, -10.0), (0.2, 20.0, 10.0, 0.0), (0.2, 20.0, -10.0, 0.0)]
        y0, pitch_y = -(bps - 1) / 2.0 * 1.15, 1.15
    all_b, all_c = [], []
    i, j = np.meshgrid(np.arange(bps), np.arange(bps), indexing="xy")
    for wld in range(n_worlds):
        s = seed + first_world + wld
        x = (i.ravel() - (bps - 1) / 2.0) * 1.15 + uniform(s, 1, per, -0.02, 0.02)
        y = y0 + j.ravel() * pitch_y + uniform(s, 2, per, -0.01, 0.01)
        b, c = _mixed_bodies(per, s, x, y, domain=wld, body_base=wld * per)
        walls = np.zeros(4, abi.collider_dtype)
        for k, (w_, h_, wx, wy) in enumerate(wall_shapes):
            shapes.fill_colliders(walls[k:k + 1], shapes.rounded_rect(w_, h_, 0.0), x=wx, y=wy, domain=wld)
        all_b.append(b)
        all_c.append(walls)
        all_c.append(c)
    return _scene(f"c5_worlds_{n_worlds}x{per}" + ("_dense" if dense else ""), np.concatenate(all_b), np.concatenate(all_c), 4)


def c5_rays(scene_bodies, n_worlds, bodies_per_world, rays_per_world=64):
    """64 rays per world fanned from body 0 of each world: max_distance 30, radius 0 for the first 48, 0.25 for the last 16."""
    rays = np.zeros(n_worlds * rays_per_world, abi.ray_dtype)
    w = np.repeat(np.arange(n_worlds), rays_per_world)
    k = np.tile(np.arange(rays_per_world), n_worlds)
    ang = (k + 0.5) * (2.0 * np.pi / rays_per_world)
    rays["sx"] = scene_bodies["x"][w * bodies_per_world]
    rays["sy"] = scene_bodies["y"][w * bodies_per_world]
    rays["dx"], rays["dy"] = np.cos(ang), np.sin(ang)
    rays["radius"] = np.where(k < rays_per_world * 3 // 4, 0.0, 0.25)
    rays["max_distance"] = 30.0
    rays["domain"] = w
    return rays


# ---- small everything-at-once scene for parity tests ---------------------------------------------
def parity_mix(n=400, seed=7, n_ropes=3, with_constraints=True, spread=1.15, offset=(0.0, 0.0), with_kinematic=True):
    """A dense blob of the body mix over a static floor + walls, a few ropes tied to bodies and to the world,
    a kinematic mover, a sensor and a restitution pair — every solver branch in a few hundred bodies."""
    side = int(np.ceil(np.sqrt(n)))
    idx = np.arange(n)
    x = (idx % side - (side - 1) / 2.0) * spread + uniform(seed, 1, n, -0.03, 0.03) + offset[0]
    y = 0.9 + (idx // side) * spread * 0.9 + uniform(seed, 2, n, -0.03, 0.03) + offset[1]
    bodies, dyn_colls = _mixed_bodies(n, seed, x, y)
    bodies["vx"] = uniform(seed, 3, n, -1.0, 1.0)
    bodies["vy"] = uniform(seed, 4, n, -2.0, 0.5)
    bodies["w"] = uniform(seed, 5, n, -2.0, 2.0)
    dyn_colls["restitution"][::7] = 0.6
    dyn_colls["flags"][5::11] &= ~np.uint32(abi.COLL_HAS_STATIC_FRICTION)
    half = side * spread / 2.0 + 1.5
    statics = np.zeros(4, abi.collider_dtype)
    shapes.fill_colliders(statics[0:1], shapes.rounded_rect(2 * half + 2.0, 0.4, 0.0), x=offset[0], y=offset[1] - 0.2)
    shapes.fill_colliders(statics[1:2], shapes.rounded_rect(0.4, 12.0, 0.1), x=offset[0] - half, y=offset[1] + 5.0)
    shapes.fill_colliders(statics[2:3], shapes.rounded_rect(0.4, 12.0, 0.1), x=offset[0] + half, y=offset[1] + 5.0, angle=0.05)
    shapes.fill_colliders(statics[3:4], shapes.circle(1.0), x=offset[0], y=offset[1] + 2.0, sensor=True)
    extra_b, extra_c = [], []
    nb = n
    if with_kinematic:
        kb = np.zeros(1, abi.body_dtype)
        shapes.fill_bodies_kinematic(kb, offset[0] - half + 1.0, offset[1] + 1.0, vx=1.5, vy=0.0, w=0.5)
        kc = np.zeros(1, abi.collider_dtype)
        shapes.fill_colliders(kc, shapes.rounded_rect(1.5, 0.4, 0.1), body=nb)
        extra_b.append(kb); extra_c.append(kc)
        nb += 1
    ropes = rope_params(n_ropes)
    rope_particles = []
    cons = []
    for r in range(n_ropes):
        ppr = 12 + 5 * r
        ropes["first_particle"][r] = len(rope_particles)
        ropes["particle_count"][r] = ppr
        pb = np.zeros(ppr, abi.body_dtype)
        px = offset[0] - half + 1.0 + r * 1.3 + np.arange(ppr) * 0.1
        py = np.full(ppr, offset[1] + 4.0 + 0.7 * r) + 0.02 * np.sin(np.arange(ppr) * 1.7)
        shapes.fill_bodies_particle(pb, 0.02, px, py)
        pc = np.zeros(ppr, abi.collider_dtype)
        shapes.fill_colliders(pc, shapes.circle(np.full(ppr, 0.06)), body=nb + np.arange(ppr), layer=abi.ROPE_LAYER,
                              static_friction=None)
        rope_particles.extend(range(nb, nb + ppr))
        extra_b.append(pb); extra_c.append(pc)
        if with_constraints:
            c = np.zeros(2, abi.constraint_dtype)
            c["owner"] = [nb, nb + ppr - 1]
            c["target"] = [-1, r]                       # first particle pinned to the world, last tied to body r
            c["off1x"] = [px[0], 0.2]
            c["off1y"] = [py[0], 0.1]
            c["linear_damping"] = 0.1
            c["flags"] = abi.LIMIT_EQ | abi.CONSTRAINT_CAN_SLEEP
            cons.append(c)
        nb += ppr
    if with_constraints and n >= 8:
        c = np.zeros(3, abi.constraint_dtype)
        c["owner"] = [4, 5, 6]
        c["target"] = [5, -1, 7]
        c["distance"] = [1.3, 0.5, 0.9]
        c["off0x"] = [0.1, 0.0, -0.1]
        c["off1x"] = [0.0, x[5], 0.1]
        c["off1y"] = [0.1, y[5] + 0.6, 0.0]
        c["compliance"] = [0.0, 0.001, 0.0]
        c["linear_damping"] = [0.1, 2.0, 0.1]
        c["angular_damping"] = [0.0, 1.0, 0.5]
        c["flags"] = [abi.LIMIT_EQ, abi.LIMIT_LT, abi.LIMIT_GT]
        cons.append(c)
    all_bodies = np.concatenate([bodies] + extra_b) if extra_b else bodies
    all_colls = np.concatenate([statics, dyn_colls] + extra_c)
    constraints = np.concatenate(cons) if cons else None
    return _scene(f"parity_mix_{n}", all_bodies, all_colls, 4, constraints, ropes if n_ropes else None,
                  np.array(rope_particles, np.int32))

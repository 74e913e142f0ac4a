"""Known-answer tests: the reference's own 11 collision unit tests, ported 1:1, run against the oracle
(collider.rs:952-1167, shape_shape.rs:490-555, query.rs:310-499). They pin the oracle's collision primitives;
nothing above them (tick, solver) is pinned by the reference."""
import math

import numpy as np
import pytest

RECT, TRI, HEX = (2, 0.5, 0.8, 0.0), (3, 1.0, 0.0, 0.0), (4, 1.0, 0.0, 0.0)
TEST_POLYGONS = [RECT, TRI, HEX]


def rot(sb, v):
    s, b = sb
    fx, fy = s * v[0] + b * v[1], s * v[1] - b * v[0]
    return np.array([s * fx + b * fy, s * fy - b * fx])


def edge_count(O, sh):
    return int(O.geometry(5, [0, 0], sh)[0, 0])


def get_edge(O, sh, i):
    r = O.geometry(4, [i, 0], sh)[0]
    return dict(n=r[0:2], start=r[2:4], dir=r[4:6], len=r[6])


def mirrored(e):
    return dict(n=-e["n"], start=-e["start"], dir=-e["dir"], len=e["len"])


def symmetrical(sh):
    return sh[0] != 3


def edges_with_mirrors(O, sh):
    es = [get_edge(O, sh, i) for i in range(edge_count(O, sh))]
    if symmetrical(sh):
        es += [mirrored(e) for e in es]
    return es


def cbp(O, sh, pt):
    r = O.geometry(0, pt, sh)[0]
    return r[0:2], bool(r[2])


def test_closest_boundary_points(oracle):
    O = oracle
    for sh in TEST_POLYGONS:
        for e in edges_with_mirrors(O, sh):
            for t in [0.3, 0.5, 0.7]:
                pt_on = e["start"] + t * e["len"] * e["dir"]
                pt_in = pt_on - 0.05 * e["n"]
                cp, interior = cbp(O, sh, pt_in)
                assert interior and abs(np.dot(cp - e["start"], e["n"])) < 1e-4 and abs(np.dot(cp - pt_in, e["dir"])) < 1e-4
            for t in [0.0, 0.01, 0.3, 0.45, 0.5, 0.55, 0.7, 0.99, 1.0]:
                pt_on = e["start"] + t * e["len"] * e["dir"]
                pt_out = pt_on + e["n"]
                cp, interior = cbp(O, sh, pt_out)
                assert (not interior) and abs(np.dot(cp - e["start"], e["n"])) < 1e-4 and abs(np.dot(cp - pt_out, e["dir"])) < 1e-4
            p = e["start"] - 0.1 * e["dir"] + e["n"]
            cp, interior = cbp(O, sh, p)
            assert (not interior) and np.sum((cp - e["start"]) ** 2) < 1e-4
            end = e["start"] + e["len"] * e["dir"]
            p = end + 0.1 * e["dir"] + e["n"]
            cp, interior = cbp(O, sh, p)
            assert (not interior) and np.sum((cp - end) ** 2) < 1e-4


def unit_circle(n):
    return [np.array([math.cos(i * math.tau / n), math.sin(i * math.tau / n)]) for i in range(n)]


def test_supporting_edges_match_edge_list(oracle):
    O = oracle
    for sh in TEST_POLYGONS:
        for d in unit_circle(20):
            r = O.geometry(2, d, sh)[0]
            supp = dict(start=r[2:4], dir=r[4:6], len=r[6])
            best, best_v = None, None
            for i in range(edge_count(O, sh)):
                e = get_edge(O, sh, i)
                if symmetrical(sh) and np.dot(e["n"], d) < 0:
                    e = mirrored(e)
                v = np.dot(e["n"], d)
                if best is None or v >= best_v:   # Iterator::max_by keeps the last maximum
                    best, best_v = e, v
            if not np.dot(best["dir"], d) < 0:
                best = dict(start=best["start"] + best["len"] * best["dir"], dir=-best["dir"], len=best["len"])
            assert np.sum((best["start"] - supp["start"]) ** 2) < 1e-4
            assert np.sum((best["dir"] - supp["dir"]) ** 2) < 1e-4
            assert abs(best["len"] - supp["len"]) < 1e-4


def test_projected_extent_matches_edge_list(oracle):
    O = oracle
    for sh in TEST_POLYGONS:
        for d in unit_circle(20):
            proj = O.geometry(1, d, sh)[0, 0]
            far = -1e300
            for i in range(edge_count(O, sh)):
                pp = np.dot(d, get_edge(O, sh, i)["start"])
                if symmetrical(sh) and pp < 0:
                    pp = -pp
                far = max(far, pp)
            assert abs(proj - far) < 1e-4


def test_round_inward_preserves_edge_distance(oracle):
    O = oracle
    for sh in TEST_POLYGONS:
        r = O.geometry(7, [0.2, 0], sh)[0]
        rounded = (int(r[0]), r[1], r[2], r[3])
        for i in range(edge_count(O, sh)):
            orig, new = get_edge(O, sh, i), get_edge(O, rounded, i)
            outer_start = new["start"] + rounded[3] * new["n"]
            assert abs(np.dot(outer_start, new["n"]) - np.dot(orig["start"], orig["n"])) < 1e-4


def test_second_moment_of_area_changes_smoothly_and_monotonically(oracle):
    O = oracle
    for sh in TEST_POLYGONS:
        prev = O.geometry(6, [0, 0], sh)[0, 1]
        for shrink in range(1, 20):
            r = O.geometry(7, [shrink * 0.02, 0], sh)[0]
            new = O.geometry(6, [0, 0], (int(r[0]), r[1], r[2], r[3]))[0, 1]
            assert -0.05 < new - prev < 0.0
            prev = new


def test_clip_various_edges(oracle):
    O = oracle
    s2 = 1 / math.sqrt(2)
    assert O.clip_edge([1, 1, 1, 0, 2], [1, 0, s2, s2, 2])[0] == 0          # Intersects
    d = rot(O.from_angle(math.pi / 6), [1, 0])
    k, enters, exits = O.clip_edge([1, 1, 1, 0, 2], [2, 0, d[0], d[1], 2])
    assert k == 1 and enters == 0.0 and abs(exits - 1 / math.cos(math.pi / 6)) < 1e-3
    d = rot(O.from_angle(7 * math.pi / 8), [1, 0])
    k, enters, exits = O.clip_edge([1, 1, 1, 0, 2], [4, 0, d[0], d[1], 2])
    assert k == 1 and abs(enters - 1 / math.cos(math.pi / 8)) < 1e-3 and exits == 2.0


# ---- query.rs tests ----
def ray_collider(O, ray, pose, sh):
    r = O.spherecast_collider([[ray[0], ray[1], ray[2], ray[3], 0.0]], [pose], [sh])[0]
    return (r[1] if r[0] else None)


def ray_circle(O, ray, c, r):
    return ray_collider(O, ray, (c[0], c[1], 1.0, 0.0), (0, 0, 0, r))


def xf(O, pose, ray):
    sb = pose[2:4]
    s = rot(sb, ray[0:2]) + np.array(pose[0:2])
    d = rot(sb, ray[2:4])
    return [s[0], s[1], d[0], d[1]]


def norm(x, y):
    m = math.hypot(x, y)
    return [x / m, y / m]


def test_ray_circle(oracle):
    O = oracle
    assert ray_circle(O, [0, 0, 0, 1], (0, 2), 1.0) is not None
    assert ray_circle(O, [0, 0] + norm(1, 1), (0, 2), 1.0) is None


def _hit(O, pose, sh, ray, expected):
    t = ray_collider(O, xf(O, pose, ray), pose, sh)
    assert t is not None and abs(t - expected) < 1e-4


def _hit_circle(O, pose, sh, ray, c):
    a = ray_collider(O, xf(O, pose, ray), pose, sh)
    b = ray_circle(O, ray, c, sh[3])
    assert (a is None) == (b is None)
    if a is not None:
        assert abs(a - b) < 1e-4


def _miss(O, pose, sh, ray):
    assert ray_collider(O, xf(O, pose, ray), pose, sh) is None


def test_ray_capsule(oracle):
    O = oracle
    sb = O.from_angle(math.radians(65.0))
    pose = (5.0, 3.5, sb[0], sb[1])
    cap = (1, 2.0, 0.0, 1.0)   # new_capsule(4, 1)
    _hit(O, pose, cap, [0, -2, 0, 1], 1.0)
    _hit(O, pose, cap, [0, -2] + norm(1, 1), math.sqrt(2))
    _hit_circle(O, pose, cap, [0, -2] + norm(2.1, 1), (2, 0))
    _miss(O, pose, cap, [0, -2, 1, 0])
    _hit_circle(O, pose, cap, [-3, -2] + norm(1, 1), (-2, 0))
    _hit(O, pose, cap, [-3, -2] + norm(2, 1), math.sqrt(5))
    _hit_circle(O, pose, cap, [-2.5, -2, 0, 1], (-2, 0))
    _hit_circle(O, pose, cap, [-500, 0, 1, 0], (-2, 0))
    _miss(O, pose, cap, [-500, 3, 1, 0])


def test_ray_rect(oracle):
    O = oracle
    sb = O.from_angle(math.radians(2.0))
    pose = (-5.0, 8.3, sb[0], sb[1])
    rect = (2, 2.0, 1.0, 0.0)  # new_rect(4, 2)
    _hit(O, pose, rect, [0, -2, 0, 1], 1.0)
    _hit(O, pose, rect, [0, -2] + norm(1, 1), math.sqrt(2))
    _miss(O, pose, rect, [0, -2] + norm(2.1, 1))
    _miss(O, pose, rect, [0, -2, 1, 0])
    _miss(O, pose, rect, [-3, -2, 1, 0])
    _hit(O, pose, rect, [-3, -2] + norm(1, 1), math.sqrt(2))
    _hit(O, pose, rect, [-3, -2] + norm(2, 1), math.sqrt(5))
    _hit(O, pose, rect, [-3, -2] + norm(1, 2), math.sqrt(5))


def test_ray_rounded_rect(oracle):
    O = oracle
    sb = O.from_angle(math.radians(23.0))
    pose = (500.0, 8.5, sb[0], sb[1])
    rr = (2, 2.0, 1.0, 1.0)    # new_rounded_rect(6, 4, 1)
    _hit(O, pose, rr, [0, -3, 0, 1], 1.0)
    _hit(O, pose, rr, [0, -3] + norm(1, 1), math.sqrt(2))
    _hit_circle(O, pose, rr, [0, -3] + norm(2.1, 1), (2, -1))
    _miss(O, pose, rr, [0, -3, 1, 0])
    _miss(O, pose, rr, [-4, -3, 1, 0])
    _hit_circle(O, pose, rr, [-4, -3] + norm(1, 1), (-2, -1))
    _hit(O, pose, rr, [-4, -3] + norm(2, 1), math.sqrt(5))
    _hit(O, pose, rr, [-4, -3] + norm(1, 2), math.sqrt(5))
    _hit_circle(O, pose, rr, [-2.5, -3, 0, 1], (-2, -1))


def test_inside_always_misses(oracle):
    O = oracle
    pose = (0.0, 0.0, 1.0, 0.0)
    for sh in [(0, 0, 0, 1.0), (1, 1.0, 0, 0.5), (2, 1.0, 0.5, 0.0), (2, 0.75, 0.25, 0.25)]:
        angle = 0.0
        while angle < 2.0 * math.tau:
            assert ray_collider(O, [0, 0, math.cos(angle), math.sin(angle)], pose, sh) is None
            angle += 0.05

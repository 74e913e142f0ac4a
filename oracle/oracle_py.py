"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes handle over oracle/_build/liboracle.so (the CPU restatement
of the Starframe tick, oracle/sf_world.hpp). Imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; the product package never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
MODE_REFERENCE, MODE_COLOURED = 0, 1
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.sfo_create.restype = C.c_void_p
        lib.sfo_create.argtypes = [C.c_int]
        lib.sfo_n_pairs.restype = lib.sfo_n_pair_units.restype = lib.sfo_n_constraint_units.restype = C.c_size_t
        lib.sfo_n_entries.restype = lib.sfo_n_contacts.restype = lib.sfo_n_islands.restype = lib.sfo_penetrations.restype = C.c_size_t
        lib.sfo_query_point.restype = C.c_uint32
        lib.sfo_query_shape.restype = C.c_uint32
        vp = C.c_void_p
        for name, args in {
            "sfo_destroy": [vp], "sfo_set_tuning": [vp, vp], "sfo_set_mask_matrix": [vp, vp],
            "sfo_set_bodies": [vp, C.c_uint32, vp], "sfo_set_colliders": [vp, C.c_uint32, vp],
            "sfo_set_constraints": [vp, C.c_uint32, vp], "sfo_set_ropes": [vp, C.c_uint32, vp, C.c_uint32, vp],
            "sfo_set_external_pairs": [vp, vp, C.c_size_t],
            "sfo_set_external_pair_hints": [vp, vp, vp, C.c_size_t],
            "sfo_tick": [vp, C.c_double, C.c_int, C.c_double, vp, C.c_int, C.c_int],
            "sfo_get_bodies": [vp, C.c_uint32, C.c_uint32, vp], "sfo_n_pairs": [vp], "sfo_get_pairs": [vp, vp],
            "sfo_n_pair_units": [vp], "sfo_get_pair_units": [vp, vp, vp, vp, vp], "sfo_get_pair_unit_hints": [vp, vp], "sfo_n_constraint_units": [vp],
            "sfo_get_constraint_units": [vp, vp, vp, vp], "sfo_n_entries": [vp], "sfo_get_manifolds": [vp, vp],
            "sfo_n_contacts": [vp], "sfo_get_contacts": [vp, vp, vp], "sfo_get_aabbs": [vp, vp],
            "sfo_cast_batch": [vp, C.c_uint32, vp, vp, C.c_int],
            "sfo_query_point": [vp, C.c_double, C.c_double, C.c_uint32, vp, C.c_uint32], "sfo_get_stats": [vp, vp], "sfo_query_shape": [vp, vp, vp, C.c_uint32], "sfo_n_islands": [vp], "sfo_penetrations": [vp, vp, C.c_size_t], "sfo_get_islands": [vp, vp, vp, vp, vp],
            "sfo_intersection_check": [C.c_int, C.c_uint32, vp, vp, vp],
            "sfo_spherecast_collider": [C.c_int, C.c_uint32, vp, vp, vp, vp],
            "sfo_geometry": [C.c_int, C.c_uint32, C.c_uint32, vp, vp, vp], "sfo_clip_edge": [vp, vp, vp],
            "sfo_ray_aabb": [vp, vp, vp], "sfo_from_angle": [C.c_double, vp],
        }.items():
            getattr(lib, name).argtypes = args
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleWorld:
    """Same record arrays in as the GPU library (abi.*_dtype); f64 (reference-faithful) or f32 arithmetic."""

    def __init__(self, use_f32=False, tuning=None, mask_matrix=None):
        self.lib = load()
        self.h = C.c_void_p(self.lib.sfo_create(1 if use_f32 else 0))
        self._keep = []
        if tuning is not None:
            self.lib.sfo_set_tuning(self.h, _p(tuning))
        if mask_matrix is not None:
            m = np.ascontiguousarray(mask_matrix, np.uint64)
            self.lib.sfo_set_mask_matrix(self.h, _p(m))
        self.n_bodies = self.n_colliders = 0

    def __del__(self):
        try:
            if self.h:
                self.lib.sfo_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def load_scene(self, scene):
        b, c = scene["bodies"], scene["colliders"]
        self.lib.sfo_set_bodies(self.h, len(b), _p(b))
        self.lib.sfo_set_colliders(self.h, len(c), _p(c))
        k = scene["constraints"]
        self.lib.sfo_set_constraints(self.h, len(k), _p(k) if len(k) else None)
        r, rp = scene["ropes"], np.ascontiguousarray(scene["rope_particles"], np.int32)
        self.lib.sfo_set_ropes(self.h, len(r), _p(r) if len(r) else None, len(rp), _p(rp) if len(rp) else None)
        self.n_bodies, self.n_colliders = len(b), len(c)

    def set_bodies(self, b):
        self.lib.sfo_set_bodies(self.h, len(b), _p(b))
        self.n_bodies = len(b)

    def set_external_pairs(self, pairs, hints=None):
        """Candidate pairs of the device's broadphase. The `near` bit in front of each pair's colouring priority (include/sfg_hash.h)
        is computed by the oracle itself from its tick-start state; `hints` overrides it (tests of the priority rule only)."""
        p = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
        self.lib.sfo_set_external_pairs(self.h, _p(p), len(p))
        if hints is not None:
            self.set_pair_hints(p, hints)

    def set_pair_hints(self, pairs, hints):
        """Override the `near` bits the oracle would compute, keyed by pair (tests of the priority rule)."""
        p = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
        h = np.ascontiguousarray(hints, np.uint8)
        assert len(h) == len(p)
        self.lib.sfo_set_external_pair_hints(self.h, _p(p), _p(h), len(p))

    def tick(self, frame_dt=1.0 / 60.0, time_scale=None, forcefield=None, mode=MODE_REFERENCE, threads=1):
        from moleengine_b200 import abi  # record dtypes only (no device code is touched)
        ff = abi.gravity_field() if forcefield is None else forcefield
        self.lib.sfo_tick(self.h, frame_dt, 0 if time_scale is None else 1, 0.0 if time_scale is None else time_scale,
                          _p(ff), mode, threads)

    def get_bodies(self):
        """(n, 7) f64: x, y, rot_s, rot_b, vx, vy, w"""
        out = np.zeros((self.n_bodies, 7))
        self.lib.sfo_get_bodies(self.h, 0, self.n_bodies, _p(out))
        return out

    def pairs(self):
        n = self.lib.sfo_n_pairs(self.h)
        out = np.zeros((n, 2), np.uint32)
        if n:
            self.lib.sfo_get_pairs(self.h, _p(out))
        return out

    def pair_units(self):
        n = self.lib.sfo_n_pair_units(self.h)
        a = [np.zeros(n, np.uint32) for _ in range(4)]
        if n:
            self.lib.sfo_get_pair_units(self.h, *[_p(x) for x in a])
        return a  # hi, lo, mult, colour

    def pair_unit_hints(self):
        """The `near` bit the oracle computed for each pair unit of the last MODE_COLOURED tick, as {(hi, lo): bit}."""
        n = self.lib.sfo_n_pair_units(self.h)
        h = np.zeros(n, np.uint8)
        if n:
            self.lib.sfo_get_pair_unit_hints(self.h, _p(h))
        hi, lo, _, _ = self.pair_units()
        return dict(zip(zip(hi.tolist(), lo.tolist()), h.tolist()))

    def constraint_units(self):
        n = self.lib.sfo_n_constraint_units(self.h)
        a = [np.zeros(n, np.uint32) for _ in range(3)]
        if n:
            self.lib.sfo_get_constraint_units(self.h, *[_p(x) for x in a])
        return a  # idx, mult, colour

    def manifolds(self):
        """(n_entries, 14): hi, lo, count, lambda, normal.xy, 2 x (off0.xy, off1.xy)"""
        n = self.lib.sfo_n_entries(self.h)
        out = np.zeros((n, 14))
        if n:
            self.lib.sfo_get_manifolds(self.h, _p(out))
        return out

    def contacts(self):
        n = self.lib.sfo_n_contacts(self.h)
        c = np.zeros((n, 2), np.uint32)
        nn = np.zeros((n, 2))
        if n:
            self.lib.sfo_get_contacts(self.h, _p(c), _p(nn))
        return c, nn

    def aabbs(self):
        out = np.zeros((self.n_colliders, 4))
        self.lib.sfo_get_aabbs(self.h, _p(out))
        return out

    def cast_batch(self, rays, use_bvh=False):
        """(n, 6): collider (-1 none), t, point.xy, normal.xy"""
        out = np.zeros((len(rays), 6))
        self.lib.sfo_cast_batch(self.h, len(rays), _p(rays), _p(out), 1 if use_bvh else 0)
        return out

    def query_point(self, x, y, domain=0, cap=64):
        out = np.zeros(cap, np.int32)
        n = self.lib.sfo_query_point(self.h, x, y, domain, _p(out), cap)
        return out[: min(n, cap)]

    def query_shape(self, query, cap=64):
        """query: one record of moleengine_b200.abi.shape_query_dtype."""
        out = np.zeros(cap, np.int32)
        q = np.ascontiguousarray(query).reshape(1)
        n = self.lib.sfo_query_shape(self.h, _p(q), _p(out), cap)
        return out[: min(n, cap)]

    def penetrations(self):
        """Depth of every contact point (> 0) between solid colliders in the CURRENT state (solver.rs:425-431's depth): a statistic
        for long-run parity gates, evaluated by the same code for device states (loaded with set_bodies) and oracle states."""
        cap = 1 << 16
        while True:
            out = np.zeros(cap)
            n = self.lib.sfo_penetrations(self.h, _p(out), cap)
            if n <= cap:
                return out[:n]
            cap = int(n)

    def islands(self):
        """Every island of the last MODE_REFERENCE tick: first_body, edge_sum, body_count, flags (1 cannot sleep, 8 asleep)."""
        n = self.lib.sfo_n_islands(self.h)
        a = [np.zeros(n, np.uint64) for _ in range(4)]
        if n:
            self.lib.sfo_get_islands(self.h, *[_p(x) for x in a])
        return a

    def stats(self):
        out = np.zeros(8, np.uint64)
        self.lib.sfo_get_stats(self.h, _p(out))
        return dict(zip(["pairs", "contact_entries", "constraint_entries", "pair_colours", "constraint_colours",
                         "islands", "sleeping", "substeps"], out.tolist()))


# ---- standalone primitives -------------------------------------------------------------------------
def intersection_check(poses, shapes, use_f32=False):
    """poses (n, 2, 4) x, y, s, b ; shapes (n, 2, 4) kind, p0, p1, r -> (n, 12): count, n.xy, 2 x (off0.xy, off1.xy), pad"""
    poses = np.ascontiguousarray(poses, np.float64)
    shapes = np.ascontiguousarray(shapes, np.float64)
    n = len(poses)
    out = np.zeros((n, 12))
    load().sfo_intersection_check(int(use_f32), n, _p(poses), _p(shapes), _p(out))
    return out


def spherecast_collider(rays5, poses, shapes, use_f32=False):
    """rays5 (n, 5) sx, sy, dx, dy, radius -> (n, 6): hit, t, normal.xy, point.xy"""
    rays5, poses, shapes = (np.ascontiguousarray(a, np.float64) for a in (rays5, poses, shapes))
    out = np.zeros((len(rays5), 6))
    load().sfo_spherecast_collider(int(use_f32), len(rays5), _p(rays5), _p(poses), _p(shapes), _p(out))
    return out


def geometry(op, args, shapes, use_f32=False):
    args, shapes = np.ascontiguousarray(args, np.float64).reshape(-1, 2), np.ascontiguousarray(shapes, np.float64).reshape(-1, 4)
    out = np.zeros((len(args), 7))
    load().sfo_geometry(int(use_f32), op, len(args), _p(args), _p(shapes), _p(out))
    return out


def clip_edge(target, edge):
    t, e, out = np.asarray(target, np.float64), np.asarray(edge, np.float64), np.zeros(3)
    load().sfo_clip_edge(_p(t), _p(e), _p(out))
    return out


def ray_aabb(ray, aabb):
    r, a, out = np.asarray(ray, np.float64), np.asarray(aabb, np.float64), np.zeros(2)
    load().sfo_ray_aabb(_p(r), _p(a), _p(out))
    return out


def from_angle(angle):
    out = np.zeros(2)
    load().sfo_from_angle(float(angle), _p(out))
    return out

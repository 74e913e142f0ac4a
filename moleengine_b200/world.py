"""Thin Python handle over the C ABI (one SfgWorld = one device + one stream).

This is the harness-level wrapper used by tests and bench.py; it moves structured numpy arrays
(abi.*_dtype) across the boundary unchanged. It never computes physics itself.
"""
import ctypes as C

import numpy as np

from . import abi


class SfgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sfg error {code}: {msg}")
        self.code = code


class GpuWorld:
    def __init__(self, tuning=None, mask_matrix=None, device=0):
        self.lib = abi.load()
        self.h = C.c_void_p()
        t = abi.default_tuning() if tuning is None else tuning
        m = abi.default_mask_matrix() if mask_matrix is None else np.ascontiguousarray(mask_matrix, dtype=np.uint64)
        rc = self.lib.sfg_world_create(abi.ptr(t), abi.ptr(m), device, C.byref(self.h))
        if rc != abi.SFG_OK:
            raise SfgError(rc, "sfg_world_create failed (no CUDA device? there is no CPU fallback)")
        self.n_bodies = 0

    def close(self):
        if self.h:
            self.lib.sfg_world_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != abi.SFG_OK:
            raise SfgError(rc, self.lib.sfg_last_error(self.h).decode())

    # ---- uploads -------------------------------------------------------------------------------
    def set_tuning(self, tuning):
        self._ck(self.lib.sfg_set_tuning(self.h, abi.ptr(tuning)))

    def set_mask_matrix(self, m):
        self._ck(self.lib.sfg_set_mask_matrix(self.h, abi.ptr(np.ascontiguousarray(m, dtype=np.uint64))))

    def set_bodies(self, bodies):
        assert bodies.dtype == abi.body_dtype
        self._ck(self.lib.sfg_set_bodies(self.h, len(bodies), abi.ptr(bodies)))
        self.n_bodies = len(bodies)

    def update_bodies(self, first, bodies):
        self._ck(self.lib.sfg_update_bodies(self.h, first, len(bodies), abi.ptr(bodies)))

    def set_colliders(self, colliders):
        assert colliders.dtype == abi.collider_dtype
        self._ck(self.lib.sfg_set_colliders(self.h, len(colliders), abi.ptr(colliders)))

    def update_colliders(self, first, colliders):
        self._ck(self.lib.sfg_update_colliders(self.h, first, len(colliders), abi.ptr(colliders)))

    def set_constraints(self, constraints):
        assert constraints.dtype == abi.constraint_dtype
        self._ck(self.lib.sfg_set_constraints(self.h, len(constraints), abi.ptr(constraints) if len(constraints) else None))

    def set_ropes(self, ropes, particle_bodies):
        assert ropes.dtype == abi.rope_dtype
        pb = np.ascontiguousarray(particle_bodies, dtype=np.int32)
        self._ck(self.lib.sfg_set_ropes(self.h, len(ropes), abi.ptr(ropes) if len(ropes) else None, len(pb),
                                        abi.ptr(pb) if len(pb) else None))

    # ---- rope topology edits on the device (rope.rs:147-165, 88-110, 226-286) ------------------------------
    def rope_cut_after(self, rope, particle_index):
        """-> index of the new rope holding the tail, or None if the last particle was popped."""
        out = C.c_uint32()
        self._ck(self.lib.sfg_rope_cut_after(self.h, rope, particle_index, C.byref(out)))
        return None if out.value == 0xFFFFFFFF else out.value

    def rope_extend(self, rope, particle_bodies):
        pb = np.ascontiguousarray(particle_bodies, dtype=np.int32)
        self._ck(self.lib.sfg_rope_extend(self.h, rope, len(pb), abi.ptr(pb)))

    def rope_remove_dead_particles(self):
        n = C.c_uint32()
        self._ck(self.lib.sfg_rope_remove_dead_particles(self.h, C.byref(n)))
        return n.value

    def get_ropes(self):
        """-> (ropes, particle_bodies) as the library holds them now (compacted)."""
        nr, npart = C.c_uint32(), C.c_uint32()
        self._ck(self.lib.sfg_get_ropes(self.h, None, 0, C.byref(nr), None, 0, C.byref(npart)))
        ropes, pb = np.zeros(nr.value, abi.rope_dtype), np.zeros(npart.value, np.int32)
        if nr.value:
            self._ck(self.lib.sfg_get_ropes(self.h, abi.ptr(ropes), nr.value, C.byref(nr), abi.ptr(pb), npart.value, C.byref(npart)))
        return ropes, pb

    def load_scene(self, scene):
        """scene: dict with bodies / colliders / constraints / ropes / rope_particles (see scenes.py)."""
        self.set_bodies(scene["bodies"])
        self.set_colliders(scene["colliders"])
        self.set_constraints(scene.get("constraints", np.zeros(0, abi.constraint_dtype)))
        self.set_ropes(scene.get("ropes", np.zeros(0, abi.rope_dtype)), scene.get("rope_particles", np.zeros(0, np.int32)))

    # ---- step + readback -----------------------------------------------------------------------
    def tick(self, frame_dt=1.0 / 60.0, time_scale=None, forcefield=None):
        ff = abi.gravity_field() if forcefield is None else forcefield
        self._ck(self.lib.sfg_tick(self.h, frame_dt, 0 if time_scale is None else 1,
                                   0.0 if time_scale is None else time_scale, abi.ptr(ff)))

    # ---- the tick in three parts + slab halo exchange (moleengine_b200/slab.py drives these) ----------------
    def tick_begin(self, frame_dt=1.0 / 60.0, time_scale=None, forcefield=None):
        ff = abi.gravity_field() if forcefield is None else forcefield
        self._ck(self.lib.sfg_tick_begin(self.h, frame_dt, 0 if time_scale is None else 1, 0.0 if time_scale is None else time_scale, abi.ptr(ff)))

    def tick_substeps(self, count):
        self._ck(self.lib.sfg_tick_substeps(self.h, count))

    def tick_end(self):
        self._ck(self.lib.sfg_tick_end(self.h))

    def slab_configure(self, x_lo, x_hi, halo):
        self._ck(self.lib.sfg_slab_configure(self.h, x_lo, x_hi, halo))

    def island_shard_configure(self, rank, n_ranks):
        self._ck(self.lib.sfg_island_shard_configure(self.h, rank, n_ranks))

    def slab_begin(self):
        self._ck(self.lib.sfg_slab_begin(self.h))

    def slab_pack(self, mode, side, device_ptr, capacity_records):
        n = C.c_uint32()
        self._ck(self.lib.sfg_slab_pack(self.h, mode, side, C.c_void_p(device_ptr), capacity_records, C.byref(n)))
        return n.value

    def slab_unpack(self, device_ptr, capacity_records):
        self._ck(self.lib.sfg_slab_unpack(self.h, C.c_void_p(device_ptr), capacity_records))

    def slab_pack_async(self, mode, side, device_ptr, capacity_records):
        self._ck(self.lib.sfg_slab_pack_async(self.h, mode, side, C.c_void_p(device_ptr), capacity_records))

    def slab_unpack_async(self, device_ptr, capacity_records):
        self._ck(self.lib.sfg_slab_unpack_async(self.h, C.c_void_p(device_ptr), capacity_records))

    def stream_fence(self, external_stream, direction):
        """direction 0: `external_stream` (a cudaStream_t as an int) waits for the world's stream; 1: the world's stream waits for it."""
        self._ck(self.lib.sfg_stream_fence(self.h, C.c_void_p(external_stream), direction))

    def get_bodies(self, first=0, count=None):
        count = self.n_bodies - first if count is None else count
        out = np.zeros(count, abi.body_dtype)
        self._ck(self.lib.sfg_get_bodies(self.h, first, count, abi.ptr(out)))
        return out

    def set_poses(self, poses, first=0):
        """poses: (n, 4) float32 rows of x, y, rot_s, rot_b (hecs_sync.rs sync_hecs_to_physics)."""
        p = np.ascontiguousarray(poses, np.float32).reshape(-1, 4)
        self._ck(self.lib.sfg_set_poses(self.h, first, len(p), abi.ptr(p)))

    def get_poses(self, first=0, count=None, out=None):
        count = self.n_bodies - first if count is None else count
        out = np.zeros((count, 4), np.float32) if out is None else out
        self._ck(self.lib.sfg_get_poses(self.h, first, count, abi.ptr(out)))
        return out

    def set_poses_indexed(self, slots, poses):
        """Poses of the listed body slots only (HecsSyncManager: entities whose options say hecs_to_physics)."""
        sl = np.ascontiguousarray(slots, np.uint32)
        p = np.ascontiguousarray(poses, np.float32).reshape(-1, 4)
        assert len(sl) == len(p)
        self._ck(self.lib.sfg_set_poses_indexed(self.h, len(sl), abi.ptr(sl), abi.ptr(p)))

    def get_poses_indexed(self, slots):
        sl = np.ascontiguousarray(slots, np.uint32)
        out = np.zeros((len(sl), 4), np.float32)
        self._ck(self.lib.sfg_get_poses_indexed(self.h, len(sl), abi.ptr(sl), abi.ptr(out)))
        return out

    def tick_poses(self, poses_in=None, frame_dt=1.0 / 60.0, time_scale=None, forcefield=None, out=None):
        """Game::physics_tick in one call: poses of all bodies in (or None), tick, poses of all bodies out (sfg_tick_poses)."""
        ff = abi.gravity_field() if forcefield is None else forcefield
        out = np.zeros((self.n_bodies, 4), np.float32) if out is None else out
        pin = None
        if poses_in is not None:
            pin = np.ascontiguousarray(poses_in, np.float32)
            assert pin.shape == (self.n_bodies, 4)
        self._ck(self.lib.sfg_tick_poses(self.h, None if pin is None else abi.ptr(pin), frame_dt, 0 if time_scale is None else 1,
                                         0.0 if time_scale is None else time_scale, abi.ptr(ff), abi.ptr(out)))
        return out

    def get_contacts(self, capacity=1 << 20):
        out = np.zeros(capacity, abi.contact_dtype)
        n = C.c_size_t()
        self._ck(self.lib.sfg_get_contacts(self.h, abi.ptr(out), capacity, C.byref(n)))
        if n.value > capacity:
            return self.get_contacts(n.value)
        return out[: n.value]

    def stats(self):
        s = np.zeros(1, abi.stats_dtype)
        self._ck(self.lib.sfg_get_stats(self.h, abi.ptr(s)))
        return s[0]

    def cast_batch(self, rays):
        assert rays.dtype == abi.ray_dtype
        out = np.zeros(len(rays), abi.hit_dtype)
        self._ck(self.lib.sfg_cast_batch(self.h, len(rays), abi.ptr(rays), abi.ptr(out)))
        return out

    def cast_kernel_ms(self):
        """Device time of the k_cast_batch launch of the last cast_batch (copies excluded)."""
        ms = C.c_float()
        self._ck(self.lib.sfg_get_cast_kernel_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def query_point_batch(self, points, domains=None, max_hits=8):
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 2)
        dom = None if domains is None else np.ascontiguousarray(domains, dtype=np.uint32)
        out = np.zeros((len(pts), max_hits), np.int32)
        cnt = np.zeros(len(pts), np.uint32)
        self._ck(self.lib.sfg_query_point_batch(self.h, len(pts), abi.ptr(pts), abi.ptr(dom), max_hits, abi.ptr(out), abi.ptr(cnt)))
        return out, cnt

    def query_shape_batch(self, queries, max_hits=16):
        assert queries.dtype == abi.shape_query_dtype
        out = np.zeros((len(queries), max_hits), np.int32)
        cnt = np.zeros(len(queries), np.uint32)
        self._ck(self.lib.sfg_query_shape_batch(self.h, len(queries), abi.ptr(queries), max_hits, abi.ptr(out), abi.ptr(cnt)))
        return out, cnt

    # ---- parity hooks ----------------------------------------------------------------------------
    def debug_pairs(self):
        n = C.c_size_t()
        self._ck(self.lib.sfg_debug_get_pairs(self.h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 2), np.uint32)
        if n.value:
            self._ck(self.lib.sfg_debug_get_pairs(self.h, abi.ptr(out), n.value, C.byref(n)))
        return out

    def debug_pair_hints(self):
        """Per pair of debug_pairs(): 1 = `near` (top bit of the colouring priority, include/sfg_hash.h)."""
        n = C.c_size_t()
        self._ck(self.lib.sfg_debug_get_pair_hints(self.h, None, 0, C.byref(n)))
        out = np.zeros(n.value, np.uint8)
        if n.value:
            self._ck(self.lib.sfg_debug_get_pair_hints(self.h, abi.ptr(out), n.value, C.byref(n)))
        return out

    def debug_contact_units(self):
        n = C.c_size_t()
        self._ck(self.lib.sfg_debug_get_contact_entries(self.h, None, None, None, 0, C.byref(n)))
        hi = np.zeros(n.value, np.uint32)
        lo = np.zeros(n.value, np.uint32)
        col = np.zeros(n.value, np.uint32)
        if n.value:
            self._ck(self.lib.sfg_debug_get_contact_entries(self.h, abi.ptr(hi), abi.ptr(lo), abi.ptr(col), n.value, C.byref(n)))
        return hi, lo & 0x7FFFFFFF, (lo >> 31) + 1, col

    def debug_constraint_units(self):
        n = C.c_size_t()
        self._ck(self.lib.sfg_debug_get_constraint_entries(self.h, None, None, 0, C.byref(n)))
        idx = np.zeros(n.value, np.uint32)
        col = np.zeros(n.value, np.uint32)
        if n.value:
            self._ck(self.lib.sfg_debug_get_constraint_entries(self.h, abi.ptr(idx), abi.ptr(col), n.value, C.byref(n)))
        return idx & 0x7FFFFFFF, (idx >> 31) + 1, col

    def debug_manifolds(self):
        n = C.c_size_t()
        self._ck(self.lib.sfg_debug_get_manifolds(self.h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 14), np.float32)
        if n.value:
            self._ck(self.lib.sfg_debug_get_manifolds(self.h, abi.ptr(out), n.value, C.byref(n)))
        return out

    def debug_islands(self):
        """Islands of the last tick: first_body, edge_sum, body_count, flags (1 cannot sleep, 2 moved, 4 fast, 8 asleep)."""
        n = C.c_size_t()
        self._ck(self.lib.sfg_debug_get_islands(self.h, None, None, None, None, 0, C.byref(n)))
        fb, es, bc, fl = np.zeros(n.value, np.uint32), np.zeros(n.value, np.uint64), np.zeros(n.value, np.uint32), np.zeros(n.value, np.uint32)
        if n.value:
            self._ck(self.lib.sfg_debug_get_islands(self.h, abi.ptr(fb), abi.ptr(es), abi.ptr(bc), abi.ptr(fl), n.value, C.byref(n)))
        return fb, es, bc, fl

    def debug_aabbs(self, n_colliders):
        out = np.zeros((n_colliders, 4), np.float32)
        self._ck(self.lib.sfg_debug_get_aabbs(self.h, abi.ptr(out), n_colliders))
        return out

"""ctypes / numpy view of the C ABI in include/starframe_gpu.h.

The numpy dtypes below are byte-for-byte the C records (checked against sizeof in the tests), so
scene builders fill structured arrays and hand their buffers straight to the library.
There is no CPU fallback: if the CUDA library has not been built, loading fails loudly.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFG_LIB") or os.path.join(HERE, "libstarframe_gpu.so")   # SFG_LIB: an experimental build of the same CUDA library (A/B timing)

SFG_OK, SFG_E_INVALID, SFG_E_CUDA, SFG_E_CAPACITY, SFG_E_NO_DEVICE, SFG_E_STATE = 0, 1, 2, 3, 4, 5

BODY_ALIVE, BODY_IGNORES_GRAVITY, BODY_MASS_FINITE, BODY_INERTIA_FINITE, BODY_ASLEEP = 1, 2, 4, 8, 16
BODY_GHOST, BODY_INACTIVE = 32, 64
POLY_POINT, POLY_SEGMENT, POLY_RECT, POLY_TRIANGLE, POLY_HEXAGON = 0, 1, 2, 3, 4
COLL_ALIVE, COLL_SENSOR, COLL_HAS_STATIC_FRICTION, COLL_HAS_DYNAMIC_FRICTION = 1, 2, 4, 8
LIMIT_EQ, LIMIT_LT, LIMIT_GT, CONSTRAINT_CAN_SLEEP = 0, 1, 2, 4
FF_NONE, FF_GRAVITY, FF_POINT_GRAVITY = 0, 1, 2
TUNE_NO_SLEEPING, TUNE_SEPARATE_LAUNCHES, TUNE_COUNT_ENTRIES = 1, 2, 4
ROPE_LAYER = 63  # collision.rs:130

tuning_dtype = np.dtype([("substeps", "<u4"), ("sleep_vel_threshold", "<f4"), ("fall_asleep_frames", "<u4"),
                         ("max_expected_acceleration", "<f4"), ("cell_size", "<f4"), ("flags", "<u4"),
                         ("reserved", "<u4", (2,))])
body_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("rot_s", "<f4"), ("rot_b", "<f4"), ("vx", "<f4"), ("vy", "<f4"),
                       ("w", "<f4"), ("flags", "<u4"), ("inv_mass", "<f4"), ("inv_inertia", "<f4"),
                       ("reserved", "<u4", (2,))])
collider_dtype = np.dtype([("kind", "<u4"), ("layer", "<u4"), ("flags", "<u4"), ("body", "<i4"), ("p0", "<f4"),
                           ("p1", "<f4"), ("circle_r", "<f4"), ("domain", "<u4"), ("x", "<f4"), ("y", "<f4"),
                           ("rot_s", "<f4"), ("rot_b", "<f4"), ("static_friction", "<f4"),
                           ("dynamic_friction", "<f4"), ("restitution", "<f4"), ("reserved", "<u4")])
constraint_dtype = np.dtype([("owner", "<i4"), ("target", "<i4"), ("flags", "<u4"), ("reserved", "<u4"),
                             ("compliance", "<f4"), ("linear_damping", "<f4"), ("angular_damping", "<f4"),
                             ("distance", "<f4"), ("off0x", "<f4"), ("off0y", "<f4"), ("off1x", "<f4"),
                             ("off1y", "<f4")])
rope_dtype = np.dtype([("spacing", "<f4"), ("compliance", "<f4"), ("bending_max_angle", "<f4"),
                       ("bending_compliance", "<f4"), ("damping", "<f4"), ("first_particle", "<u4"),
                       ("particle_count", "<u4"), ("reserved", "<u4")])
force_term_dtype = np.dtype([("kind", "<u4"), ("a", "<f4"), ("b", "<f4"), ("c", "<f4"), ("d", "<f4")])
forcefield_dtype = np.dtype([("n_terms", "<u4"), ("terms", force_term_dtype, (4,))])
contact_dtype = np.dtype([("collider0", "<u4"), ("collider1", "<u4"), ("nx", "<f4"), ("ny", "<f4")])
ray_dtype = np.dtype([("sx", "<f4"), ("sy", "<f4"), ("dx", "<f4"), ("dy", "<f4"), ("radius", "<f4"),
                      ("max_distance", "<f4"), ("domain", "<u4"), ("reserved", "<u4")])
shape_query_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("rot_s", "<f4"), ("rot_b", "<f4"), ("p0", "<f4"), ("p1", "<f4"),
                              ("circle_r", "<f4"), ("kind", "<u4"), ("layer_mask", "<u8"), ("domain", "<u4"), ("reserved", "<u4")])
hit_dtype = np.dtype([("collider", "<i4"), ("t", "<f4"), ("px", "<f4"), ("py", "<f4"), ("nx", "<f4"), ("ny", "<f4"),
                      ("reserved", "<u4", (2,))])
stats_dtype = np.dtype([("n_bodies", "<u4"), ("n_colliders", "<u4"), ("n_pairs", "<u4"), ("n_contact_entries", "<u4"),
                        ("n_constraint_entries", "<u4"), ("n_contact_colours", "<u4"), ("n_constraint_colours", "<u4"),
                        ("n_rope_particles", "<u4"), ("n_ropes", "<u4"), ("n_contacts", "<u4"), ("substeps", "<u4"),
                        ("grid_levels", "<u4"), ("grid_cells", "<u4"), ("n_islands", "<u4"), ("n_sleeping_bodies", "<u4"),
                        ("kernel_launches", "<u4"), ("ms_broadphase", "<f4"), ("ms_graph", "<f4"), ("ms_solver", "<f4"),
                        ("ms_total", "<f4"), ("cell_size", "<f4"), ("n_stages", "<u4"), ("n_touching_units", "<u4"), ("n_touching_entries", "<u4")])

# every symbol include/starframe_gpu.h declares (tests check the library exports each one)
EXPORTS = [
    "sfg_abi_version", "sfg_world_create", "sfg_world_destroy", "sfg_world_clear", "sfg_last_error", "sfg_set_tuning",
    "sfg_set_mask_matrix", "sfg_default_tuning", "sfg_default_mask_matrix", "sfg_set_bodies", "sfg_update_bodies",
    "sfg_set_colliders", "sfg_update_colliders", "sfg_set_constraints", "sfg_set_ropes", "sfg_rope_cut_after", "sfg_rope_extend", "sfg_rope_remove_dead_particles", "sfg_get_ropes", "sfg_tick", "sfg_tick_begin", "sfg_tick_substeps", "sfg_tick_end", "sfg_slab_configure", "sfg_island_shard_configure", "sfg_slab_begin", "sfg_slab_pack", "sfg_slab_unpack", "sfg_slab_pack_async", "sfg_slab_unpack_async", "sfg_stream_fence", "sfg_get_bodies",
    "sfg_get_contacts", "sfg_set_poses", "sfg_get_poses", "sfg_set_poses_indexed", "sfg_get_poses_indexed", "sfg_tick_poses", "sfg_get_stats", "sfg_cast_batch", "sfg_get_cast_kernel_ms", "sfg_query_point_batch", "sfg_query_shape_batch", "sfg_debug_get_pairs", "sfg_debug_get_pair_hints",
    "sfg_debug_get_contact_entries", "sfg_debug_get_constraint_entries", "sfg_debug_get_manifolds", "sfg_debug_get_aabbs", "sfg_debug_get_islands",
    "sfg_debug_intersection_check", "sfg_debug_spherecast_collider", "sfg_debug_geometry",
]

_lib = None


def load():
    """Load libstarframe_gpu.so (built in-tree by __graft_entry__.build()). Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "moleengine_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u32, i32, sz = C.c_void_p, C.c_uint32, C.c_int, C.c_size_t
    lib.sfg_abi_version.restype = i32
    lib.sfg_world_create.argtypes = [vp, vp, i32, C.POINTER(vp)]
    lib.sfg_world_destroy.argtypes = [vp]
    lib.sfg_world_destroy.restype = None
    lib.sfg_world_clear.argtypes = [vp]
    lib.sfg_last_error.argtypes = [vp]
    lib.sfg_last_error.restype = C.c_char_p
    lib.sfg_set_tuning.argtypes = [vp, vp]
    lib.sfg_set_mask_matrix.argtypes = [vp, vp]
    lib.sfg_default_tuning.argtypes = [vp]
    lib.sfg_default_tuning.restype = None
    lib.sfg_default_mask_matrix.argtypes = [vp]
    lib.sfg_default_mask_matrix.restype = None
    lib.sfg_set_bodies.argtypes = [vp, u32, vp]
    lib.sfg_update_bodies.argtypes = [vp, u32, u32, vp]
    lib.sfg_set_colliders.argtypes = [vp, u32, vp]
    lib.sfg_update_colliders.argtypes = [vp, u32, u32, vp]
    lib.sfg_set_constraints.argtypes = [vp, u32, vp]
    lib.sfg_set_ropes.argtypes = [vp, u32, vp, u32, vp]
    lib.sfg_rope_cut_after.argtypes = [vp, u32, u32, C.POINTER(u32)]
    lib.sfg_rope_extend.argtypes = [vp, u32, u32, vp]
    lib.sfg_rope_remove_dead_particles.argtypes = [vp, C.POINTER(u32)]
    lib.sfg_get_ropes.argtypes = [vp, vp, u32, C.POINTER(u32), vp, u32, C.POINTER(u32)]
    lib.sfg_set_poses_indexed.argtypes = [vp, u32, vp, vp]
    lib.sfg_get_poses_indexed.argtypes = [vp, u32, vp, vp]
    lib.sfg_tick.argtypes = [vp, C.c_double, i32, C.c_double, vp]
    lib.sfg_tick_begin.argtypes = [vp, C.c_double, i32, C.c_double, vp]
    lib.sfg_tick_substeps.argtypes = [vp, u32]
    lib.sfg_tick_end.argtypes = [vp]
    lib.sfg_slab_configure.argtypes = [vp, C.c_float, C.c_float, C.c_float]
    lib.sfg_slab_begin.argtypes = [vp]
    lib.sfg_island_shard_configure.argtypes = [vp, u32, u32]
    lib.sfg_slab_pack.argtypes = [vp, i32, i32, vp, u32, C.POINTER(u32)]
    lib.sfg_slab_unpack.argtypes = [vp, vp, u32]
    lib.sfg_slab_pack_async.argtypes = [vp, i32, i32, vp, u32]
    lib.sfg_slab_unpack_async.argtypes = [vp, vp, u32]
    lib.sfg_stream_fence.argtypes = [vp, vp, i32]
    lib.sfg_get_bodies.argtypes = [vp, u32, u32, vp]
    lib.sfg_get_contacts.argtypes = [vp, vp, sz, C.POINTER(sz)]
    lib.sfg_set_poses.argtypes = [vp, u32, u32, vp]
    lib.sfg_get_poses.argtypes = [vp, u32, u32, vp]
    lib.sfg_tick_poses.argtypes = [vp, vp, C.c_double, i32, C.c_double, vp, vp]
    lib.sfg_get_stats.argtypes = [vp, vp]
    lib.sfg_cast_batch.argtypes = [vp, u32, vp, vp]
    lib.sfg_get_cast_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.sfg_query_point_batch.argtypes = [vp, u32, vp, vp, u32, vp, vp]
    lib.sfg_query_shape_batch.argtypes = [vp, u32, vp, u32, vp, vp]
    lib.sfg_debug_get_pairs.argtypes = [vp, vp, sz, C.POINTER(sz)]
    lib.sfg_debug_get_pair_hints.argtypes = [vp, vp, sz, C.POINTER(sz)]
    lib.sfg_debug_get_contact_entries.argtypes = [vp, vp, vp, vp, sz, C.POINTER(sz)]
    lib.sfg_debug_get_constraint_entries.argtypes = [vp, vp, vp, sz, C.POINTER(sz)]
    lib.sfg_debug_get_manifolds.argtypes = [vp, vp, sz, C.POINTER(sz)]
    lib.sfg_debug_get_aabbs.argtypes = [vp, vp, sz]
    lib.sfg_debug_get_islands.argtypes = [vp, vp, vp, vp, vp, sz, C.POINTER(sz)]
    lib.sfg_debug_intersection_check.argtypes = [i32, u32, vp, vp, vp]
    lib.sfg_debug_spherecast_collider.argtypes = [i32, u32, vp, vp, vp, vp]
    lib.sfg_debug_geometry.argtypes = [i32, u32, u32, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("sfg_abi_version",):
            fn.restype = i32
    _lib = lib
    return lib


def ptr(a):
    """Pointer to a C-contiguous numpy array's buffer (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def default_tuning(**kw):
    t = np.zeros(1, tuning_dtype)
    t["substeps"], t["sleep_vel_threshold"], t["fall_asleep_frames"], t["max_expected_acceleration"] = 10, 0.001, 10, 10.0
    for k, v in kw.items():
        t[k] = v
    return t


def default_mask_matrix():
    m = np.full(64, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    m[63] &= np.uint64(~(1 << 63) & 0xFFFFFFFFFFFFFFFF)
    return m


def gravity_field(gx=0.0, gy=-9.81):
    ff = np.zeros(1, forcefield_dtype)
    ff["n_terms"] = 1
    ff["terms"][0, 0]["kind"] = FF_GRAVITY
    ff["terms"][0, 0]["a"], ff["terms"][0, 0]["b"] = gx, gy
    return ff

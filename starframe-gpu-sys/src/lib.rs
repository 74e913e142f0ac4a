//! Raw declarations of `include/starframe_gpu.h`, 1:1. UNCOMPILED in this repository's image (no Rust toolchain);
//! `tests/test_rust_binding.py` parses this file and checks every `#[repr(C)]` record (field order, types, size) against
//! the C header and against the numpy dtypes the tested ctypes binding uses, and every `extern "C"` function against the
//! header's declarations.
#![allow(non_camel_case_types)]
use core::ffi::c_void;
use std::os::raw::{c_char, c_int};

pub const SFG_ABI_VERSION: c_int = 1;
pub const SFG_OK: c_int = 0;
pub const SFG_E_INVALID: c_int = 1;
pub const SFG_E_CUDA: c_int = 2;
pub const SFG_E_CAPACITY: c_int = 3;
pub const SFG_E_NO_DEVICE: c_int = 4;
pub const SFG_E_STATE: c_int = 5;

pub const SFG_TUNE_NO_SLEEPING: u32 = 1;
pub const SFG_TUNE_SEPARATE_LAUNCHES: u32 = 2;
pub const SFG_TUNE_COUNT_ENTRIES: u32 = 4;
pub const SFG_MAX_SUBSTEPS: u32 = 4095;
pub const SFG_BODY_ALIVE: u32 = 1;
pub const SFG_BODY_IGNORES_GRAVITY: u32 = 2;
pub const SFG_BODY_MASS_FINITE: u32 = 4;
pub const SFG_BODY_INERTIA_FINITE: u32 = 8;
pub const SFG_BODY_ASLEEP: u32 = 16;
pub const SFG_BODY_GHOST: u32 = 32;
pub const SFG_BODY_INACTIVE: u32 = 64;
pub const SFG_POLY_POINT: u32 = 0;
pub const SFG_POLY_SEGMENT: u32 = 1;
pub const SFG_POLY_RECT: u32 = 2;
pub const SFG_POLY_TRIANGLE: u32 = 3;
pub const SFG_POLY_HEXAGON: u32 = 4;
pub const SFG_COLL_ALIVE: u32 = 1;
pub const SFG_COLL_SENSOR: u32 = 2;
pub const SFG_COLL_HAS_STATIC_FRICTION: u32 = 4;
pub const SFG_COLL_HAS_DYNAMIC_FRICTION: u32 = 8;
pub const SFG_LIMIT_EQ: u32 = 0;
pub const SFG_LIMIT_LT: u32 = 1;
pub const SFG_LIMIT_GT: u32 = 2;
pub const SFG_LIMIT_MASK: u32 = 3;
pub const SFG_CONSTRAINT_CAN_SLEEP: u32 = 4;
pub const SFG_FF_NONE: u32 = 0;
pub const SFG_FF_GRAVITY: u32 = 1;
pub const SFG_FF_POINT_GRAVITY: u32 = 2;

#[repr(C)]
pub struct SfgWorld {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SfgTuning {
    pub substeps: u32,
    pub sleep_vel_threshold: f32,
    pub fall_asleep_frames: u32,
    pub max_expected_acceleration: f32,
    pub cell_size: f32,
    pub flags: u32,
    pub reserved: [u32; 2],
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SfgBody {
    pub x: f32,
    pub y: f32,
    pub rot_s: f32,
    pub rot_b: f32,
    pub vx: f32,
    pub vy: f32,
    pub w: f32,
    pub flags: u32,
    pub inv_mass: f32,
    pub inv_inertia: f32,
    pub reserved: [u32; 2],
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SfgCollider {
    pub kind: u32,
    pub layer: u32,
    pub flags: u32,
    pub body: i32,
    pub p0: f32,
    pub p1: f32,
    pub circle_r: f32,
    pub domain: u32,
    pub x: f32,
    pub y: f32,
    pub rot_s: f32,
    pub rot_b: f32,
    pub static_friction: f32,
    pub dynamic_friction: f32,
    pub restitution: f32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SfgConstraint {
    pub owner: i32,
    pub target: i32,
    pub flags: u32,
    pub reserved: u32,
    pub compliance: f32,
    pub linear_damping: f32,
    pub angular_damping: f32,
    pub distance: f32,
    pub off0x: f32,
    pub off0y: f32,
    pub off1x: f32,
    pub off1y: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SfgRope {
    pub spacing: f32,
    pub compliance: f32,
    pub bending_max_angle: f32,
    pub bending_compliance: f32,
    pub damping: f32,
    pub first_particle: u32,
    pub particle_count: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SfgForceTerm {
    pub kind: u32,
    pub a: f32,
    pub b: f32,
    pub c: f32,
    pub d: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SfgForceField {
    pub n_terms: u32,
    pub terms: [SfgForceTerm; 4],
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SfgContactInfo {
    pub collider0: u32,
    pub collider1: u32,
    pub nx: f32,
    pub ny: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SfgRay {
    pub sx: f32,
    pub sy: f32,
    pub dx: f32,
    pub dy: f32,
    pub radius: f32,
    pub max_distance: f32,
    pub domain: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SfgShapeQuery {
    pub x: f32,
    pub y: f32,
    pub rot_s: f32,
    pub rot_b: f32,
    pub p0: f32,
    pub p1: f32,
    pub circle_r: f32,
    pub kind: u32,
    pub layer_mask: u64,
    pub domain: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SfgCastHit {
    pub collider: i32,
    pub t: f32,
    pub px: f32,
    pub py: f32,
    pub nx: f32,
    pub ny: f32,
    pub reserved: [u32; 2],
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SfgStats {
    pub n_bodies: u32,
    pub n_colliders: u32,
    pub n_pairs: u32,
    pub n_contact_entries: u32,
    pub n_constraint_entries: u32,
    pub n_contact_colours: u32,
    pub n_constraint_colours: u32,
    pub n_rope_particles: u32,
    pub n_ropes: u32,
    pub n_contacts: u32,
    pub substeps: u32,
    pub grid_levels: u32,
    pub grid_cells: u32,
    pub n_islands: u32,
    pub n_sleeping_bodies: u32,
    pub kernel_launches: u32,
    pub ms_broadphase: f32,
    pub ms_graph: f32,
    pub ms_solver: f32,
    pub ms_total: f32,
    pub cell_size: f32,
    pub n_stages: u32,
    pub n_touching_units: u32,
    pub n_touching_entries: u32,
}

extern "C" {
    // lifetime: PhysicsWorld::new (physics.rs:366), clear (:386)
    pub fn sfg_abi_version() -> c_int;
    pub fn sfg_world_create(tuning: *const SfgTuning, mask_matrix: *const u64, device: c_int, out: *mut *mut SfgWorld) -> c_int;
    pub fn sfg_world_destroy(w: *mut SfgWorld);
    pub fn sfg_world_clear(w: *mut SfgWorld) -> c_int;
    pub fn sfg_last_error(w: *const SfgWorld) -> *const c_char;
    // public fields `consts`, `mask_matrix` (physics.rs:353-354)
    pub fn sfg_set_tuning(w: *mut SfgWorld, tuning: *const SfgTuning) -> c_int;
    pub fn sfg_set_mask_matrix(w: *mut SfgWorld, mask_matrix: *const u64) -> c_int;
    pub fn sfg_default_tuning(out: *mut SfgTuning);
    pub fn sfg_default_mask_matrix(out: *mut u64);
    // entity_set / constraint_set / rope_set contents, uploaded by arena slot
    pub fn sfg_set_bodies(w: *mut SfgWorld, n_slots: u32, bodies: *const SfgBody) -> c_int;
    pub fn sfg_update_bodies(w: *mut SfgWorld, first_slot: u32, count: u32, bodies: *const SfgBody) -> c_int;
    pub fn sfg_set_colliders(w: *mut SfgWorld, n_slots: u32, colliders: *const SfgCollider) -> c_int;
    pub fn sfg_update_colliders(w: *mut SfgWorld, first_slot: u32, count: u32, colliders: *const SfgCollider) -> c_int;
    pub fn sfg_set_constraints(w: *mut SfgWorld, n: u32, constraints: *const SfgConstraint) -> c_int;
    pub fn sfg_set_ropes(w: *mut SfgWorld, n_ropes: u32, ropes: *const SfgRope, n_particles: u32, particle_bodies: *const i32) -> c_int;
    // rope topology edits on the device (rope.rs:147-165 cut_after, :61-85 extend_line, :226-286 remove_dead_particles)
    pub fn sfg_rope_cut_after(w: *mut SfgWorld, rope: u32, particle_index: u32, new_rope_out: *mut u32) -> c_int;
    pub fn sfg_rope_extend(w: *mut SfgWorld, rope: u32, n_new: u32, particle_bodies: *const i32) -> c_int;
    pub fn sfg_rope_remove_dead_particles(w: *mut SfgWorld, n_ropes_out: *mut u32) -> c_int;
    pub fn sfg_get_ropes(w: *mut SfgWorld, ropes_out: *mut SfgRope, rope_capacity: u32, n_ropes_out: *mut u32, particle_bodies_out: *mut i32,
                         particle_capacity: u32, n_particles_out: *mut u32) -> c_int;
    // PhysicsWorld::tick (physics.rs:396-1158), whole or in three parts
    pub fn sfg_tick(w: *mut SfgWorld, frame_dt: f64, has_time_scale: c_int, time_scale: f64, ff: *const SfgForceField) -> c_int;
    pub fn sfg_tick_begin(w: *mut SfgWorld, frame_dt: f64, has_time_scale: c_int, time_scale: f64, ff: *const SfgForceField) -> c_int;
    pub fn sfg_tick_substeps(w: *mut SfgWorld, count: u32) -> c_int;
    pub fn sfg_tick_end(w: *mut SfgWorld) -> c_int;
    // one world over several GPUs: x-slabs with halo exchange, or island shards
    pub fn sfg_slab_configure(w: *mut SfgWorld, x_lo: f32, x_hi: f32, halo: f32) -> c_int;
    pub fn sfg_island_shard_configure(w: *mut SfgWorld, rank: u32, n_ranks: u32) -> c_int;
    pub fn sfg_slab_begin(w: *mut SfgWorld) -> c_int;
    pub fn sfg_slab_pack(w: *mut SfgWorld, mode: c_int, side: c_int, device_out: *mut c_void, capacity_records: u32, count_out: *mut u32) -> c_int;
    pub fn sfg_slab_unpack(w: *mut SfgWorld, device_in: *const c_void, capacity_records: u32) -> c_int;
    pub fn sfg_slab_pack_async(w: *mut SfgWorld, mode: c_int, side: c_int, device_out: *mut c_void, capacity_records: u32) -> c_int;
    pub fn sfg_slab_unpack_async(w: *mut SfgWorld, device_in: *const c_void, capacity_records: u32) -> c_int;
    pub fn sfg_stream_fence(w: *mut SfgWorld, external_stream: *mut c_void, direction: c_int) -> c_int;
    // results
    pub fn sfg_get_bodies(w: *mut SfgWorld, first_slot: u32, count: u32, out: *mut SfgBody) -> c_int;
    pub fn sfg_get_contacts(w: *mut SfgWorld, out: *mut SfgContactInfo, capacity: usize, n_out: *mut usize) -> c_int;
    // pose-only sync: HecsSyncManager::sync_hecs_to_physics / sync_physics_to_hecs (hecs_sync.rs:121-227); rows of x, y, rot_s, rot_b
    pub fn sfg_set_poses(w: *mut SfgWorld, first_slot: u32, count: u32, poses_xysb: *const f32) -> c_int;
    pub fn sfg_get_poses(w: *mut SfgWorld, first_slot: u32, count: u32, poses_xysb: *mut f32) -> c_int;
    pub fn sfg_set_poses_indexed(w: *mut SfgWorld, n: u32, slots: *const u32, poses_xysb: *const f32) -> c_int;
    pub fn sfg_get_poses_indexed(w: *mut SfgWorld, n: u32, slots: *const u32, poses_xysb: *mut f32) -> c_int;
    // Game::physics_tick (game.rs:297-303) in one call
    pub fn sfg_tick_poses(w: *mut SfgWorld, poses_in_xysb: *const f32, frame_dt: f64, has_time_scale: c_int, time_scale: f64,
                          forcefield: *const SfgForceField, poses_out_xysb: *mut f32) -> c_int;
    pub fn sfg_get_stats(w: *mut SfgWorld, out: *mut SfgStats) -> c_int;
    // PhysicsWorld::raycast / spherecast (physics.rs:1255-1311), query_point (:1188-1210), query_shape (:1215-1245), batched
    pub fn sfg_cast_batch(w: *mut SfgWorld, n: u32, rays: *const SfgRay, out: *mut SfgCastHit) -> c_int;
    pub fn sfg_get_cast_kernel_ms(w: *mut SfgWorld, out_ms: *mut f32) -> c_int;
    pub fn sfg_query_point_batch(w: *mut SfgWorld, n: u32, points_xy: *const f32, domains: *const u32, max_hits: u32,
                                 out_colliders: *mut i32, out_counts: *mut u32) -> c_int;
    pub fn sfg_query_shape_batch(w: *mut SfgWorld, n: u32, queries: *const SfgShapeQuery, max_hits: u32, out_colliders: *mut i32,
                                 out_counts: *mut u32) -> c_int;
}

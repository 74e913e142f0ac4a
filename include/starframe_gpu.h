/* starframe_gpu.h — C ABI of the B200-native physics tick for Starframe (MoleTrooper/moleengine).
 *
 * This is the drop-in boundary: a Rust `-sys` crate binds exactly these
 * symbols and a thin wrapper re-creates the reference's `PhysicsWorld` API on top of them.
 * The reference has no FFI for this path (it is safe Rust end to end), so each entry point
 * cites the Rust item it replaces under /root/reference/src.
 *
 * Conventions
 *  - Every function returns an int status (SFG_OK = 0); nothing throws or unwinds.
 *  - All buffers are caller-allocated HOST memory unless the name says `_device`.
 *  - Arrays are indexed by ARENA SLOT (thunderdome `Index::slot()`); the host keeps generations
 *    and does the pre-tick housekeeping (physics.rs:399-431) before uploading.
 *  - Device state is f32 structure-of-arrays; the records below are the 16-byte-aligned rows the
 *    library scatters into float4 columns on upload.
 *  - One SfgWorld = one CUDA device + one stream; a world is single-threaded from the caller's
 *    point of view (the reference's tick/query methods take `&mut self`).
 */
#ifndef STARFRAME_GPU_H
#define STARFRAME_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFG_ABI_VERSION 1

/* status codes */
#define SFG_OK 0
#define SFG_E_INVALID 1   /* bad argument (null pointer, out-of-range slot, unknown enum) */
#define SFG_E_CUDA 2      /* CUDA runtime error; text in sfg_last_error() */
#define SFG_E_CAPACITY 3  /* an internal buffer hit its hard cap (pairs, cells) */
#define SFG_E_NO_DEVICE 4 /* no usable CUDA device — there is NO CPU fallback */
#define SFG_E_STATE 5     /* call order violated (e.g. cast before any tick) */

typedef struct SfgWorld SfgWorld;

/* physics.rs:311-349 TuningConstants (+ GPU-only knobs, all optional = 0) */
typedef struct SfgTuning {
    uint32_t substeps;               /* default 10 */
    float sleep_vel_threshold;       /* default 0.001 */
    uint32_t fall_asleep_frames;     /* default 10; UINT32_MAX disables sleeping */
    float max_expected_acceleration; /* default 10.0 */
    float cell_size;                 /* broadphase level-0 cell edge; 0 = auto */
    uint32_t flags;                  /* SFG_TUNE_* */
    uint32_t reserved[2];
} SfgTuning;
#define SFG_TUNE_NO_SLEEPING 1u      /* skip island / sleeping bookkeeping entirely */
#define SFG_TUNE_SEPARATE_LAUNCHES 2u /* debug: one launch per colour instead of the persistent kernel */
#define SFG_TUNE_COUNT_ENTRIES 4u     /* measurement: count the contact entries that found a contact (SfgStats.n_touching_entries); one more
                                      * warp-aggregated atomic per group of touching units in the solver, so off by default */
#define SFG_MAX_SUBSTEPS 4095u       /* substeps of one tick (after time scaling) are capped: per-substep tables are sized by it, manifold tags carry 12 bits of it */

/* body.rs:7-13 Body. flags carry the Mass enum discriminants (body.rs:105-129). */
typedef struct SfgBody {
    float x, y, rot_s, rot_b;        /* pose: translation + rotor {s, bv.xy} (math.rs:12) */
    float vx, vy, w;                 /* physics.rs:48-53 Velocity */
    uint32_t flags;                  /* SFG_BODY_* */
    float inv_mass, inv_inertia;     /* Mass::inv() (body.rs:121-128); 0 when infinite */
    uint32_t reserved[2];
} SfgBody;
#define SFG_BODY_ALIVE 1u
#define SFG_BODY_IGNORES_GRAVITY 2u
#define SFG_BODY_MASS_FINITE 4u
#define SFG_BODY_INERTIA_FINITE 8u
#define SFG_BODY_INACTIVE 64u        /* output only: in a slab-decomposed world, outside this rank's slab and halo (stale copy) */
#define SFG_BODY_GHOST 32u           /* output only: in a slab-decomposed world, a halo copy owned by a neighbouring slab */
#define SFG_BODY_ASLEEP 16u          /* output only (sfg_get_bodies): the body's island was skipped in the last tick (physics.rs:711-741) */

/* collider.rs:19-29 Collider, :197-200 ColliderShape, :378-410 ColliderPolygon, :903-911 material */
typedef struct SfgCollider {
    uint32_t kind;                   /* SFG_POLY_* */
    uint32_t layer;                  /* 0..63 (collision.rs:113) */
    uint32_t flags;                  /* SFG_COLL_* */
    int32_t body;                    /* body slot, or -1 = static (entity_set.rs:55 coll_bodies) */
    float p0, p1, circle_r;          /* hl | hw,hh | outer_r ; Minkowski circle radius */
    uint32_t domain;                 /* independent-world id (batched worlds); 0 for a single world */
    float x, y, rot_s, rot_b;        /* pose relative to the body, or to the world if static */
    float static_friction, dynamic_friction, restitution;
    uint32_t reserved;
} SfgCollider;
#define SFG_POLY_POINT 0u
#define SFG_POLY_SEGMENT 1u
#define SFG_POLY_RECT 2u
#define SFG_POLY_TRIANGLE 3u
#define SFG_POLY_HEXAGON 4u
#define SFG_COLL_ALIVE 1u
#define SFG_COLL_SENSOR 2u                 /* ColliderType::Sensor */
#define SFG_COLL_HAS_STATIC_FRICTION 4u    /* static_friction_coef: Some(_) */
#define SFG_COLL_HAS_DYNAMIC_FRICTION 8u   /* dynamic_friction_coef: Some(_) */

/* constraint.rs:12-58 Constraint (only ConstraintType::Distance exists) */
typedef struct SfgConstraint {
    int32_t owner;                   /* body slot */
    int32_t target;                  /* body slot or -1 = world */
    uint32_t flags;                  /* SFG_LIMIT_* | SFG_CONSTRAINT_CAN_SLEEP */
    uint32_t reserved;
    float compliance, linear_damping, angular_damping, distance;
    float off0x, off0y, off1x, off1y;
} SfgConstraint;
#define SFG_LIMIT_EQ 0u
#define SFG_LIMIT_LT 1u
#define SFG_LIMIT_GT 2u
#define SFG_LIMIT_MASK 3u
#define SFG_CONSTRAINT_CAN_SLEEP 4u

/* rope.rs:16-50 — the solver-visible part of RopeParameters + the particle chain */
typedef struct SfgRope {
    float spacing, compliance, bending_max_angle, bending_compliance, damping;
    uint32_t first_particle;         /* offset into the particle body-slot array */
    uint32_t particle_count;         /* >= 2 */
    uint32_t reserved;
} SfgRope;

/* forcefield.rs — the generic trait becomes data: a sum of up to 4 terms */
typedef struct SfgForceTerm {
    uint32_t kind;                   /* SFG_FF_* */
    float a, b, c, d;                /* GRAVITY: (gx, gy) ; POINT: (px, py, strength, falloff) */
} SfgForceTerm;
typedef struct SfgForceField {
    uint32_t n_terms;
    SfgForceTerm terms[4];
} SfgForceField;
#define SFG_FF_NONE 0u
#define SFG_FF_GRAVITY 1u
#define SFG_FF_POINT_GRAVITY 2u

/* physics.rs:112-118 ContactInfo (island id is internal) */
typedef struct SfgContactInfo {
    uint32_t collider0, collider1;   /* slots; pair order as the reference: [later, earlier] */
    float nx, ny;                    /* first contact point's normal, away from collider0 */
} SfgContactInfo;

/* query.rs:32-35 Ray + the spherecast arguments of physics.rs:1266 */
typedef struct SfgRay {
    float sx, sy, dx, dy;            /* start, unit direction */
    float radius, max_distance;
    uint32_t domain;
    uint32_t reserved;
} SfgRay;
/* physics.rs:1215-1219 query_shape arguments: PhysicsPose + ColliderShape + CollisionLayerMask (collision.rs:112-128) */
typedef struct SfgShapeQuery {
    float x, y, rot_s, rot_b;        /* pose */
    float p0, p1, circle_r;          /* ColliderShape, same encoding as SfgCollider */
    uint32_t kind;                   /* SFG_POLY_* */
    uint64_t layer_mask;             /* bit l set = colliders on layer l are reported */
    uint32_t domain;
    uint32_t reserved;
} SfgShapeQuery;
/* physics.rs:131-149 CastHit; collider = -1 means None */
typedef struct SfgCastHit {
    int32_t collider;
    float t, px, py, nx, ny;
    uint32_t reserved[2];
} SfgCastHit;

typedef struct SfgStats {
    uint32_t n_bodies, n_colliders;          /* live counts */
    uint32_t n_pairs;                        /* candidate pairs kept (>=1 body, distinct bodies) */
    uint32_t n_contact_entries;              /* after the x2 duplication quirk (physics.rs:641,651) */
    uint32_t n_constraint_entries;
    uint32_t n_contact_colours, n_constraint_colours;
    uint32_t n_rope_particles, n_ropes;
    uint32_t n_contacts;                     /* entries with a non-zero latched manifold */
    uint32_t substeps;                       /* of the last tick */
    uint32_t grid_levels, grid_cells;
    uint32_t n_islands, n_sleeping_bodies;
    uint32_t kernel_launches;                /* kernels launched by the last tick */
    float ms_broadphase, ms_graph, ms_solver, ms_total; /* CUDA-event times of the last tick */
    float cell_size;
    uint32_t n_stages;                       /* stages per substep of the persistent solver's table */
    uint32_t n_touching_units;               /* SFG_TUNE_COUNT_ENTRIES: pair units whose first visit found a contact, summed over the substeps */
    uint32_t n_touching_entries;             /* the same counted in contact entries (a body-body unit is two entries), summed over the substeps */
} SfgStats;

/* ---- lifetime: PhysicsWorld::new (physics.rs:366), clear (:386) ---- */
int sfg_abi_version(void);
int sfg_world_create(const SfgTuning* tuning, const uint64_t mask_matrix[64], int device, SfgWorld** out);
void sfg_world_destroy(SfgWorld* w);
int sfg_world_clear(SfgWorld* w);
const char* sfg_last_error(const SfgWorld* w);

/* ---- public mutable fields `consts`, `mask_matrix` (physics.rs:353-354) ---- */
int sfg_set_tuning(SfgWorld* w, const SfgTuning* tuning);
int sfg_set_mask_matrix(SfgWorld* w, const uint64_t mask_matrix[64]);
void sfg_default_tuning(SfgTuning* out);                 /* TuningConstants::default */
void sfg_default_mask_matrix(uint64_t out[64]);          /* collision.rs:134-140 */

/* ---- entity_set / constraint_set / rope_set contents (entity_set.rs:66-165,
 *      constraint_set.rs:24-48, rope.rs:185-210), uploaded by slot ---- */
int sfg_set_bodies(SfgWorld* w, uint32_t n_slots, const SfgBody* bodies);
int sfg_update_bodies(SfgWorld* w, uint32_t first_slot, uint32_t count, const SfgBody* bodies);
int sfg_set_colliders(SfgWorld* w, uint32_t n_slots, const SfgCollider* colliders);
int sfg_update_colliders(SfgWorld* w, uint32_t first_slot, uint32_t count, const SfgCollider* colliders);
int sfg_set_constraints(SfgWorld* w, uint32_t n, const SfgConstraint* constraints);
int sfg_set_ropes(SfgWorld* w, uint32_t n_ropes, const SfgRope* ropes, uint32_t n_particles,
                  const int32_t* particle_bodies);

/* ---- rope topology edits without re-uploading the ropes (rope.rs:147-165 Rope::cut_after, :88-110 Rope::extend_line,
 *      :226-286 RopeSet::remove_dead_particles). The library keeps one range per rope on the host and the particle -> body
 *      array on the device; the solver's rope units are rebuilt on the device before the next tick. Ropes are addressed by their
 *      index in the table (sfg_set_ropes order; new ropes are appended; sfg_rope_remove_dead_particles compacts the table and
 *      reports its new size; sfg_get_ropes returns it). ---- */
int sfg_rope_cut_after(SfgWorld* w, uint32_t rope, uint32_t particle_index, uint32_t* new_rope_out);   /* *new_rope_out = UINT32_MAX: the last particle was popped */
int sfg_rope_extend(SfgWorld* w, uint32_t rope, uint32_t n_new, const int32_t* particle_bodies);
int sfg_rope_remove_dead_particles(SfgWorld* w, uint32_t* n_ropes_out);
int sfg_get_ropes(SfgWorld* w, SfgRope* ropes_out, uint32_t rope_capacity, uint32_t* n_ropes_out, int32_t* particle_bodies_out,
                  uint32_t particle_capacity, uint32_t* n_particles_out);

/* ---- PhysicsWorld::tick (physics.rs:396-1158) ---- */
int sfg_tick(SfgWorld* w, double frame_dt, int has_time_scale, double time_scale, const SfgForceField* ff);
/* The same tick in three calls, for callers that must act between substeps (slab halo exchange):
 * begin = boxes, pairs, islands, colouring; substeps(count) = the next `count` XPBD substeps; end = contacts, sleeping, stats. */
int sfg_tick_begin(SfgWorld* w, double frame_dt, int has_time_scale, double time_scale, const SfgForceField* ff);
int sfg_tick_substeps(SfgWorld* w, uint32_t count);
int sfg_tick_end(SfgWorld* w);

/* ---- one large world cut into x-slabs, one per GPU (north_star: "spatial slabs, boundary-body halo exchange over NVLink").
 * Every rank uploads the WHOLE world (same slots everywhere) and then owns the bodies whose x lies in [x_lo, x_hi); bodies
 * outside the slab and its halo sit the tick out on this rank. Per tick the caller runs
 *   sfg_slab_begin; pack(mode 0, side) -> send to that neighbour / receive -> unpack      (refresh: who is near the cut)
 *   sfg_tick_begin; then per substep: sfg_tick_substeps(1); pack(mode 1) -> exchange -> unpack (after EVERY substep: ownership follows position, so both
 *   copies of a body near a cut must be bitwise equal when the next refresh decides who owns it); sfg_tick_end
 * The buffers are DEVICE memory owned by the caller ((capacity_records + 1) * 64 bytes; record 0 is a header carrying the
 * count), so the transport (NCCL send/recv, peer copies) stays outside the library. side 0 = towards lower x, 1 = higher x.
 * A slab world never sleeps (islands across a cut are not tracked). x_hi <= x_lo switches slab mode off. */
int sfg_slab_configure(SfgWorld* w, float x_lo, float x_hi, float halo);
/* Island sharding (the reference's own parallelism, physics.rs:895-947 / 1070-1082, one level up): every rank holds the whole
 * world, solves only the islands with hash(first_body) % n_ranks == rank and skips the others for the tick; after sfg_tick
 * the ranks exchange what they solved: sfg_slab_pack(mode 2, side 0) -> all-gather -> sfg_slab_unpack of every other rank's
 * buffer. Bit-identical to one GPU (islands share nothing). n_ranks <= 1 switches it off. */
int sfg_island_shard_configure(SfgWorld* w, uint32_t rank, uint32_t n_ranks);
int sfg_slab_begin(SfgWorld* w);
int sfg_slab_pack(SfgWorld* w, int mode, int side, void* device_out, uint32_t capacity_records, uint32_t* count_out);   /* mode 0 refresh, 1 substep sync, 2 island shard */
int sfg_slab_unpack(SfgWorld* w, const void* device_in, uint32_t capacity_records);
/* The same two without any host synchronisation: the kernels are only queued on the world's stream and the record count stays on
 * the device (record 0 of the buffer; read it back after sfg_tick_end to check it against the capacity). sfg_stream_fence orders
 * the world's stream against the transport's stream without the host waiting: direction 0 = `external_stream` (a cudaStream_t)
 * waits for everything queued on the world's stream so far, 1 = the world's stream waits for everything queued on
 * `external_stream` so far. With these a halo exchange is pack, fence(0), send / receive on the transport's stream, fence(1),
 * unpack — all asynchronous; the tick's only host synchronisation is the one at its end. */
int sfg_slab_pack_async(SfgWorld* w, int mode, int side, void* device_out, uint32_t capacity_records);
int sfg_slab_unpack_async(SfgWorld* w, const void* device_in, uint32_t capacity_records);
int sfg_stream_fence(SfgWorld* w, void* external_stream, int direction);

/* ---- results: bodies written back (physics.rs:1150-1157), contacts (:1097-1122, :1165) ---- */
int sfg_get_bodies(SfgWorld* w, uint32_t first_slot, uint32_t count, SfgBody* out);
int sfg_get_contacts(SfgWorld* w, SfgContactInfo* out, size_t capacity, size_t* n_out);

/* ---- pose-only sync, the GPU side of HecsSyncManager (physics/hecs_sync.rs:121-227): the game loop pushes f32
 *      `Pose`s of both-ways entities in before the tick (Game::physics_tick, game.rs:297-303) and pulls poses out after.
 *      poses: `count` rows of (x, y, rot_s, rot_b); velocities and masses are untouched. ---- */
int sfg_set_poses(SfgWorld* w, uint32_t first_slot, uint32_t count, const float* poses_xysb);
int sfg_get_poses(SfgWorld* w, uint32_t first_slot, uint32_t count, float* poses_xysb);
/* The same for a LIST of body slots: HecsSyncManager only touches the entities whose HecsSyncOptions say hecs_to_physics /
 * physics_to_hecs (hecs_sync.rs:149-158, 196-204), so only those poses cross the bus: n rows of (x, y, rot_s, rot_b) for slots[i]. */
int sfg_set_poses_indexed(SfgWorld* w, uint32_t n, const uint32_t* slots, const float* poses_xysb);
int sfg_get_poses_indexed(SfgWorld* w, uint32_t n, const uint32_t* slots, float* poses_xysb);
/* Game::physics_tick in one call (game.rs:297-303): poses of ALL body slots in (NULL = keep the device's), sfg_tick, poses of
 * all body slots out (NULL = none). Same results as sfg_set_poses + sfg_tick + sfg_get_poses; the upload is queued in front of
 * the tick instead of being waited for and the download overlaps the end of the tick. Pinned host buffers make both copies
 * run at PCIe rate. */
int sfg_tick_poses(SfgWorld* w, const float* poses_in_xysb, double frame_dt, int has_time_scale, double time_scale,
                   const SfgForceField* forcefield, float* poses_out_xysb);
int sfg_get_stats(SfgWorld* w, SfgStats* out);

/* ---- PhysicsWorld::raycast / spherecast (physics.rs:1255-1311), batched ---- */
int sfg_cast_batch(SfgWorld* w, uint32_t n, const SfgRay* rays, SfgCastHit* out);
/* measurement: device time of the k_cast_batch launch of the last sfg_cast_batch (CUDA events on the world's stream, copies excluded) */
int sfg_get_cast_kernel_ms(SfgWorld* w, float* out_ms);
/* ---- PhysicsWorld::query_point (physics.rs:1188-1210), batched: for each point, up to
 *      `max_hits` collider slots (-1 padded) ---- */
int sfg_query_point_batch(SfgWorld* w, uint32_t n, const float* points_xy, const uint32_t* domains,
                          uint32_t max_hits, int32_t* out_colliders, uint32_t* out_counts);
/* PhysicsWorld::query_shape (physics.rs:1215-1245), batched: up to max_hits collider slots per query (-1 padded) and the
 * full count. Like every query it uses the spatial index and boxes of the last tick and the current poses. */
int sfg_query_shape_batch(SfgWorld* w, uint32_t n, const SfgShapeQuery* queries, uint32_t max_hits, int32_t* out_colliders,
                          uint32_t* out_counts);

/* ---- parity hooks (tests only; not part of the drop-in surface) ---- */
/* candidate pairs of the last tick as [later, earlier] collider slots, unsorted */
int sfg_debug_get_pairs(SfgWorld* w, uint32_t* out_pairs, size_t capacity_pairs, size_t* n_out);
/* per pair, in the same order: 1 = the pair may touch during the tick. A scheduling hint computed at tick start that is the
 * top bit of the pair's colouring priority (include/sfg_hash.h); the CPU oracle is fed these bits to replay the same order */
int sfg_debug_get_pair_hints(SfgWorld* w, uint8_t* out_hints, size_t capacity_pairs, size_t* n_out);
/* contact entries of the last tick: (collider_hi, collider_lo | copy<<31) and colour */
int sfg_debug_get_contact_entries(SfgWorld* w, uint32_t* out_hi, uint32_t* out_lo_copy, uint32_t* out_colour,
                                  size_t capacity, size_t* n_out);
int sfg_debug_get_constraint_entries(SfgWorld* w, uint32_t* out_index_copy, uint32_t* out_colour,
                                     size_t capacity, size_t* n_out);
/* per contact entry: count, lambda_n and 2 x (normal.xy, offsets[0].xy, offsets[1].xy) = 14 floats */
int sfg_debug_get_manifolds(SfgWorld* w, float* out, size_t capacity_entries, size_t* n_out);
/* per collider slot: tick-time AABB (min.xy, max.xy) */
int sfg_debug_get_aabbs(SfgWorld* w, float* out, size_t capacity_colliders);
/* islands of the last tick (physics.rs:598-705): IslandId {first_body, edge_sum} (physics.rs:178-185), body count,
 * flags: 1 cannot sleep, 2 a body moved since the last tick, 4 a body is fast after the solve, 8 asleep (skipped) */
int sfg_debug_get_islands(SfgWorld* w, uint32_t* first_body, uint64_t* edge_sum, uint32_t* body_count, uint32_t* flags,
                          size_t capacity, size_t* n_out);
/* standalone device narrowphase: n pairs of (pose, shape) -> 14 floats each (as above, lambda=0).
 * poses: 2*n rows of (x, y, rot_s, rot_b); shapes: 2*n rows of (kind, p0, p1, circle_r) as floats. */
int sfg_debug_intersection_check(int device, uint32_t n, const float* poses, const float* shapes, float* out);
/* standalone device ray_collider / spherecast_collider: out rows (hit, t, nx, ny, px, py) */
int sfg_debug_spherecast_collider(int device, uint32_t n, const SfgRay* rays, const float* poses,
                                  const float* shapes, float* out);
/* standalone device geometry tables (collider.rs:692-892): op 0 = closest_boundary_point
 * (out: x, y, is_interior), 1 = projected_extent (out: extent, 0, 0), 2 = supporting_edge
 * (in: dir; out rows of 7: normal.xy, start.xy, dir.xy, length), 3 = point_collider_bool.
 * args: n rows of (x, y); shapes: n rows of (kind, p0, p1, circle_r); stride of out = 7 floats. */
int sfg_debug_geometry(int device, uint32_t op, uint32_t n, const float* args, const float* shapes, float* out);

#ifdef __cplusplus
}
#endif
#endif /* STARFRAME_GPU_H */

// Device-resident state of one physics world and the parameter blocks handed to kernels.
// Layout (all structure-of-arrays, 16-byte columns, indexed by ARENA SLOT unless noted):
//
//  bodies     P  float4 (x_old, y_old, dx, dy)  position at substep start + displacement so far
//             Q  float4 (s, b, s_old, b_old)    rotor now / at substep start
//             V  float4 (vx, vy, w, -)          OV float4 velocity after forces, before the solve
//             MI float4 (1/m, 1/I, flags, -)    EA float2 external acceleration (restitution gate)
//  colliders  CS float4 (p0, p1, circle_r, kind) CL float4 local pose (x, y, s, b)
//             CM float4 (mu_s, mu_d, e, flags)   CI int4 (body, layer, domain, flags)   AABB float4
//  pair units (sorted colour-major, rebuilt every tick)
//             H  (body0, body1, flags, mu_s)   S (hi, lo, mu_d, e)   G0 G1 shapes   L0 L1 local poses
//             RC float: body centres farther apart than this cannot touch (bounding radii + local offsets + margin)
//             M0 (count | tag << 8, lambda_n, n.x, n.y)  M1 / M2 point 0 / 1 (off0.xy, off1.xy); entry = copy * n_units + unit
//                tag = (tick mod 4096) << 12 | substep + 1 of the visit that wrote the entry: an entry with another tag counts as
//                "no contact", so visits that find nothing write nothing and the column is cleared only every 4096 ticks
//             WL u32: per colour (same ranges as the units) the units that found a contact in the current substep;
//                wl_count[substep * n_colours + colour] = how many. The velocity sweep walks these lists only.
//
// Keeping (x_old, dx) instead of x lets every kernel form differences of nearby points without the
// cancellation that absolute f32 coordinates suffer far from the origin, and makes the velocity
// update v = dx / dt exact to f32 rounding (SURVEY.md "hard parts" 1).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sfg {

constexpr int MAX_LEVELS = 8;
#ifndef SFG_SOLVER_THREADS
#define SFG_SOLVER_THREADS 512      // threads per CTA of the persistent solver, one CTA per SM: 117 registers per thread, no spills.
                                    // Measured on C3 / settled C3 / C5 (solver ms): 384: 0.89 / 4.76 / 16.9, 448: 0.86 / 4.65 / 16.5,
                                    // 512: 0.85 / 4.48 / 15.7, 640 (96 registers, 64 B of spills): 0.87 / 4.59 / 16.5, 768 (80, 176 B): 0.90 / 4.77 / 16.9
#endif

#ifndef SFG_STEAL_MIN_ITEMS
#define SFG_STEAL_MIN_ITEMS 32u    // items per CTA from which a stage's tail is stolen (smaller stages: the global atomics cost more than they return)
#endif
#ifndef SFG_STEAL_PERCENT
#define SFG_STEAL_PERCENT 15u       // share of a big stage's items that is not dealt to CTAs up front but taken from one global counter by
                                    // whichever warp runs out of work first (evens out the CTAs; 0 = off)
#endif

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct GridParams {
    float ox, oy;
    int n_levels;                 // grid levels; colliders larger than the top cell go to the "huge" list
    uint32_t n_domains;
    float inv_cell[MAX_LEVELS];
    float cell[MAX_LEVELS];
    int nx[MAX_LEVELS], ny[MAX_LEVELS];
    uint32_t base[MAX_LEVELS];    // first key of each level: level 0 = all small colliders, level l >= 1 = ALL colliders of that size class
    uint32_t n_cells;             // key n_cells = huge list; 0xFFFFFFFF = dead
    // Large colliders are found by small ones through one of two extra key ranges, never through base[l >= 1]:
    //   proxy grid (level-0 geometry): one entry per level-0 cell a large collider's box overlaps ("rasterised"), so a long
    //   thin platform costs its few neighbours a 3x3 look-up instead of costing every body under a coarse cell a box test;
    //   big-plain levels: large colliders that would need more than RASTER_MAX proxies, same cells as base[l].
    uint32_t bp_base[MAX_LEVELS]; // first key of big-plain level l (l >= 1)
    uint32_t proxy_base;          // first key of the proxy grid
    uint32_t n_keys;              // one past the last key
};
constexpr uint32_t RASTER_MAX = 2048;
enum : uint32_t { ET_HOME = 0u, ET_BIG_PLAIN = 1u, ET_PROXY = 2u };   // entry type, kept in SI.z bits 16..17

struct Bounds {                   // filled by k_aabb with ordered-int atomics
    int min_x, min_y, max_x, max_y;   // float bits mapped to ordered ints
    int max_extent;
    uint32_t max_domain;
    uint32_t n_live;
    uint32_t pad;
    double sum_extent;
};

// unit flags
enum : uint32_t {
    UF_MULT2 = 1u, UF_SENSOR = 2u, UF_HAS_SF = 4u, UF_HAS_DF = 8u, UF_SEES0 = 16u, UF_SEES1 = 32u,
    UF_ROPE0 = 64u, UF_ROPE1 = 128u, UF_ASLEEP = 256u,
};
// body flag bits kept in MI.z (same values as SFG_BODY_*)
// BF_ALIVE means "takes part in this tick"; BF_EXISTS is what the host uploaded. They differ only in a slab-decomposed world,
// where the per-tick classification (sfg_api.cu k_slab_classify) switches off bodies outside the slab and its halo.
enum : uint32_t { BF_ALIVE = 1u, BF_NOGRAV = 2u, BF_MASS = 4u, BF_INERTIA = 8u, BF_ASLEEP = 16u /* set per tick by the island pass */,
                  BF_EXISTS = 32u, BF_GHOST = 64u, BF_ZONE_L = 128u, BF_ZONE_R = 256u,
                  BF_FOREIGN = 512u /* island owned by another rank this tick (island sharding); always together with BF_ASLEEP */ };
// collider flag bits kept in CM.w / CI.w (same values as SFG_COLL_*)
enum : uint32_t { CF_ALIVE = 1u, CF_SENSOR = 2u, CF_HAS_SF = 4u, CF_HAS_DF = 8u };

struct ForceParams { uint32_t n; uint32_t kind[4]; float a[4], b[4], c[4], d[4]; };

// Everything a solver stage needs. Passed by value (kernel parameter space).
struct SolverCtx {
    // bodies
    float4 *P, *Q, *V, *OV, *MI;
    float2 *EA, *PCD;
    const int *rope_next, *rope_prev;
    uint32_t n_bodies;
    uint32_t tag_base;            // (tick number mod 4096) << 12: the tag of a manifold entry is tag_base + substep + 1, so an entry written by an
                                  // earlier tick never reads as this tick's and M0 needs no clearing between ticks (sfg_api.cu build_graph)
    uint32_t bouncy;              // some collider has restitution != 0: only then are old velocities / external accelerations kept (solver.rs:555-576)
    // pair units
    const float4 *H, *S, *G0, *G1, *L0, *L1;
    const float* RC;
    float4 *M0, *M1, *M2;
    float2* LN;
    uint32_t* WL;
    uint32_t* wl_count;
    uint32_t* steal;              // [substep * n_stages + stage]: next item of that stage's stolen tail (k_solve_persistent)
    uint32_t* far_min;            // [substep * n_stages + first stage of a merged far-only run]: the first stage of that run with a surviving unit (0xFFFFFFFF = none)
    uint32_t* stat_touch;         // SFG_TUNE_COUNT_ENTRIES: [0] += units, [1] += entries whose first visit found a contact (nullptr = not counted)
    const uint32_t* near_count;   // per contact colour: units sorted first inside the colour because they may touch this tick
    uint32_t n_units, n_colours;
    // constraint units (sorted colour-major)
    const float4 *K0, *K1, *K2;   // (owner, target, flags, -) (compliance, lin damp, ang damp, distance) (off0.xy, off1.xy)
    uint32_t n_kunits;
    // rope units
    const int4* RU;               // (curr, next, after, rope) body slots, sorted by colour i mod 3
    const int4* RD;               // (curr, next, rope, -) damping pairs sorted by colour i mod 2
    const float4 *RPA, *RPB;      // per rope (spacing, compliance, max_angle, bend_compliance) (damping, -, -, -)
    const int* rp_body;           // particle -> body slot
    const int* rp_rope;           // particle -> rope
    float2* LAT;                  // per body slot: lateral correction; x = NaN means unset
    uint32_t n_particles;
    // step
    float dt, inv_dt, inv_dt_sq;
    ForceParams ff;
};

// stage table of the persistent solver kernel
enum StageKind : uint32_t {
    ST_INTEGRATE = 0, ST_ROPE_STEP, ST_CONSTRAINTS, ST_PRECONTACT, ST_CONTACTS, ST_DERIVE, ST_CONTACT_VEL,
    ST_CONSTRAINT_DAMP, ST_ROPE_DAMP, ST_ROPE_LATERAL,
};
struct Stage { uint32_t kind, begin, end, colour; };   // colour: index of the contact colour (contact stages only)
constexpr uint32_t M0_COUNT_MASK = 0xFFu;

}  // namespace sfg

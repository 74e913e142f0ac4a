/* Deterministic 32-bit priorities for the Jones-Plassmann constraint colouring.
 *
 * The reference (MoleTrooper/moleengine) has no colouring: it runs sequential Gauss-Seidel in
 * DFS island order (src/physics.rs:598-705, src/physics/solver.rs:275-490). The GPU path
 * replaces that order with a colour-major order; the colouring is defined here, once, so that
 * the CUDA implementation and the CPU oracle agree bit-exactly:
 *
 *   priority(entry) = (sfg_hash3(a, b, c), a, b, c)   compared lexicographically, higher first
 *   colour(entry)   = smallest colour not used by any conflicting entry of HIGHER priority
 *
 * For a contact entry (a, b, c) = (later collider slot, earlier collider slot, copy 0|1);
 * for a user-constraint entry (a, b, c) = (constraint index, 0x5fc0ffee, copy 0|1).
 * Two entries conflict iff they share a body that sees forces (finite mass or inertia).
 *
 * Contact entries carry one more bit in front of the hash: `near` = the pair may touch during this tick (a hint taken
 * from the poses and velocities at tick start, sfg_broadphase.cuh k_pair_unit_setup). Near pairs outrank all others, so
 * they take the lowest colours and the high colours hold only pairs that a single distance test rejects — those stages
 * of the solver cost next to nothing. It is a scheduling hint, not a filter: every pair is still visited, and any
 * priority order gives a valid Gauss-Seidel order. The hint is DEFINED below (sfg_near_hint) in f32 with every operation
 * rounded exactly once (no fused multiply-add, correctly rounded square roots), so the CUDA kernel and the CPU oracle compute
 * the same bit from the same tick-start state independently; tests compare the two (sfg_debug_get_pair_hints) before they
 * compare colours.
 */
#ifndef SFG_HASH_H
#define SFG_HASH_H
#include <stdint.h>

#if defined(__CUDACC__)
#define SFG_HD __host__ __device__ static __forceinline__
#else
#define SFG_HD static inline
#endif

SFG_HD uint32_t sfg_fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
SFG_HD uint32_t sfg_hash3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t h = sfg_fmix32(a * 0x9e3779b1u + 0x7f4a7c15u);
    h = sfg_fmix32(h ^ (b * 0x85ebca77u + 0x165667b1u));
    h = sfg_fmix32(h ^ (c * 0x27d4eb2fu + 0x9e3779b9u));
    return h;
}

SFG_HD uint32_t sfg_pair_priority(uint32_t hi, uint32_t lo, uint32_t near_hint) {
    return (near_hint ? 0x80000000u : 0u) | (sfg_hash3(hi, lo, 0u) >> 1);
}

#define SFG_CONSTRAINT_TAG 0x5fc0ffeeu

/* ---- the `near` hint: exactly rounded f32 arithmetic on both sides ----
 * Device code uses the _rn intrinsics (never contracted into FMAs, IEEE square root whatever -prec-sqrt says); host code must be
 * compiled with -ffp-contract=off (oracle/Makefile does). Inputs are the f32 tick-start state: body position x = P.x + P.z,
 * rotor (s, b), velocity; collider local pose and shape. */
#if defined(__CUDA_ARCH__)
#define SFG_FADD(a, b) __fadd_rn((a), (b))
#define SFG_FSUB(a, b) __fsub_rn((a), (b))
#define SFG_FMUL(a, b) __fmul_rn((a), (b))
#define SFG_FSQRT(a) __fsqrt_rn((a))
#else
#include <math.h>
#define SFG_FADD(a, b) ((float)(a) + (float)(b))
#define SFG_FSUB(a, b) ((float)(a) - (float)(b))
#define SFG_FMUL(a, b) ((float)(a) * (float)(b))
#define SFG_FSQRT(a) sqrtf((a))
#endif

/* radius of the circle around a collider's own origin that contains its rounded shape; kind = SFG_POLY_* (0 point, 1 segment,
 * 2 rect: p0, p1 = half extents, 3 triangle, 4 hexagon: p0 = outer radius) */
SFG_HD float sfg_bound_radius(uint32_t kind, float p0, float p1, float circle_r) {
    float core = 0.0f;
    if (kind == 2u) core = SFG_FSQRT(SFG_FADD(SFG_FMUL(p0, p0), SFG_FMUL(p1, p1)));
    else if (kind != 0u) core = p0;
    return SFG_FADD(core, circle_r);
}
/* world position of a bodied collider's origin: rotor (s, b) applied to the local offset (ultraviolet's sandwich, same operation
 * order as the tick-time boxes), plus the body position */
SFG_HD void sfg_collider_centre(float bx, float by, float rs, float rb, float lx, float ly, float* cx, float* cy) {
    const float fx = SFG_FADD(SFG_FMUL(rs, lx), SFG_FMUL(rb, ly));
    const float fy = SFG_FSUB(SFG_FMUL(rs, ly), SFG_FMUL(rb, lx));
    *cx = SFG_FADD(SFG_FADD(SFG_FMUL(rs, fx), SFG_FMUL(rb, fy)), bx);
    *cy = SFG_FADD(SFG_FSUB(SFG_FMUL(rs, fy), SFG_FMUL(rb, fx)), by);
}
/* near = the two bounding circles are NOT farther apart than the bodies can close in one frame at their current relative
 * velocity (with a 2 cm + 50 % allowance): gap <= 0.02 + 1.5 * frame_dt * |v1 - v0|. NaN anywhere counts as near. */
SFG_HD uint32_t sfg_near_hint(float c0x, float c0y, float r0, float c1x, float c1y, float r1, float v0x, float v0y, float v1x, float v1y,
                              float frame_dt) {
    const float dx = SFG_FSUB(c1x, c0x), dy = SFG_FSUB(c1y, c0y);
    const float dist = SFG_FSQRT(SFG_FADD(SFG_FMUL(dx, dx), SFG_FMUL(dy, dy)));
    const float gap = SFG_FSUB(dist, SFG_FADD(r0, r1));
    const float wx = SFG_FSUB(v1x, v0x), wy = SFG_FSUB(v1y, v0y);
    const float speed = SFG_FSQRT(SFG_FADD(SFG_FMUL(wx, wx), SFG_FMUL(wy, wy)));
    const float reach = SFG_FADD(0.02f, SFG_FMUL(SFG_FMUL(1.5f, frame_dt), speed));
    return (gap > reach) ? 0u : 1u;
}

#endif /* SFG_HASH_H */

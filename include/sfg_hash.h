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
 * priority order gives a valid Gauss-Seidel order. The oracle is handed the hint bits together with the pairs
 * (sfg_debug_get_pair_hints) and applies sfg_pair_priority() to them, so both sides colour identically.
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

#endif /* SFG_HASH_H */

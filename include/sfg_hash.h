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

#define SFG_CONSTRAINT_TAG 0x5fc0ffeeu

#endif /* SFG_HASH_H */

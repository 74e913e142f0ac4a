// Spatial-hash broadphase and constraint-graph build (once per tick).
//   k_aabb            tick-time AABB per collider (bit-exact f32 restatement of physics.rs:447-459) + world bounds
//   k_cell_count / k_cell_emit   hierarchical uniform grid: level by AABB extent, key = cell of the AABB centre; large
//                     colliders also get one proxy entry per level-0 cell they overlap (sfg_state.cuh GridParams)
//   radix sort        sfg_sort.cuh
//   k_gather_sorted / k_cell_starts
//   k_pairs           one lane per collider: forward half of the 3x3 block on its own level, 3x3 of the proxy grid (large
//                     partners), coarser levels only where populated; candidate tests = strict AABB overlap + layer mask +
//                     same domain + ">=1 body, distinct bodies" (physics.rs:461-474, 551-555, 579); hits compacted with warp ballots
//   k_pair_unit_setup / k_adj_fill / k_colour_dataflow / k_build_pair_units   graph colouring (include/sfg_hash.h)
// The emitted SET equals the reference's BVH result; the order does not.
#pragma once
#include <cooperative_groups.h>
#include "../../include/sfg_hash.h"
#include "sfg_geom.cuh"
#include "sfg_state.cuh"

namespace sfg {
namespace cg = cooperative_groups;

SFG_DI int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
inline float ord2f_host(int i) { int j = i >= 0 ? i : i ^ 0x7FFFFFFF; float f; memcpy(&f, &j, 4); return f; }

__global__ void __launch_bounds__(256) k_aabb(uint32_t n_colls, const float4* __restrict__ CS, const float4* __restrict__ CL,
                                              const int4* __restrict__ CI, const float4* __restrict__ P, const float4* __restrict__ Q,
                                              const float4* __restrict__ V, const float4* __restrict__ MI, float frame_dt, float pad,
                                              float4* __restrict__ AABB, Bounds* __restrict__ bounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = false;
    float4 box = make_float4(0, 0, 0, 0);
    uint32_t domain = 0;
    if (i < n_colls) {
        const int4 ci = CI[i];
        // a collider whose body sits this tick out (outside the slab and its halo) is not in the broadphase
        if ((ci.w & CF_ALIVE) && (ci.x < 0 || (__float_as_uint(MI[ci.x].z) & BF_ALIVE))) {
            live = true;
            domain = uint32_t(ci.z);
            const Shape s = mkshape(CS[i]);
            const float4 l = CL[i];
            if (ci.x >= 0) {
                const float4 p = P[ci.x], q = Q[ci.x], v = V[ci.x];
                const V2 bt = mk(__fadd_rn(p.x, p.z), __fadd_rn(p.y, p.w));
                const Rot br = mkrot(q.x, q.y);
                // body.pose * coll.pose
                V2 rt = rot_apply_exact(br, mk(l.x, l.y));
                V2 t = mk(__fadd_rn(rt.x, bt.x), __fadd_rn(rt.y, bt.y));
                Rot r = rot_mul_exact(br, mkrot(l.z, l.w));
                box = shape_aabb_exact(s, t, r);
                const float ex = __fmul_rn(v.x, frame_dt), ey = __fmul_rn(v.y, frame_dt);   // collision.rs:49-63
                if (ex < 0.f) box.x = __fadd_rn(box.x, ex); else box.z = __fadd_rn(box.z, ex);
                if (ey < 0.f) box.y = __fadd_rn(box.y, ey); else box.w = __fadd_rn(box.w, ey);
                box.x = __fsub_rn(box.x, pad); box.y = __fsub_rn(box.y, pad);                // collision.rs:39-46
                box.z = __fadd_rn(box.z, pad); box.w = __fadd_rn(box.w, pad);
            } else {
                box = shape_aabb_exact(s, mk(l.x, l.y), mkrot(l.z, l.w));
            }
            AABB[i] = box;
        } else {
            // not in this tick's broadphase: k_cell_keys gives a non-finite box the dead key, the cast kernels never reach it
            const float qnan = __int_as_float(0x7FC00000);
            AABB[i] = make_float4(qnan, qnan, qnan, qnan);
        }
    }
    // bounds: warp reduce, one atomic per warp
    const bool fin = live && isfinite(box.x) && isfinite(box.y) && isfinite(box.z) && isfinite(box.w);
    float mnx = fin ? box.x : 3.0e38f, mny = fin ? box.y : 3.0e38f, mxx = fin ? box.z : -3.0e38f, mxy = fin ? box.w : -3.0e38f;
    float ext = fin ? fmaxf(box.z - box.x, box.w - box.y) : 0.f;
    float sum = ext;
    uint32_t cnt = fin ? 1u : 0u, dom = live ? domain : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xFFFFFFFFu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xFFFFFFFFu, mny, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xFFFFFFFFu, mxx, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xFFFFFFFFu, mxy, o));
        ext = fmaxf(ext, __shfl_xor_sync(0xFFFFFFFFu, ext, o)); sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o); dom = max(dom, __shfl_xor_sync(0xFFFFFFFFu, dom, o));
    }
    // one set of atomics per CTA (per-warp atomics on 8 hot addresses cost more than the whole kernel)
    __shared__ float s_f[5][8];
    __shared__ uint32_t s_u[2][8];
    __shared__ float s_sum[8];
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_f[0][warp] = mnx; s_f[1][warp] = mny; s_f[2][warp] = mxx; s_f[3][warp] = mxy; s_f[4][warp] = ext; s_u[0][warp] = cnt; s_u[1][warp] = dom; s_sum[warp] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) {
            mnx = fminf(mnx, s_f[0][k]); mny = fminf(mny, s_f[1][k]); mxx = fmaxf(mxx, s_f[2][k]); mxy = fmaxf(mxy, s_f[3][k]);
            ext = fmaxf(ext, s_f[4][k]); cnt += s_u[0][k]; dom = max(dom, s_u[1][k]); sum += s_sum[k];
        }
        if (cnt) {
            atomicMin(&bounds->min_x, f2ord(mnx)); atomicMin(&bounds->min_y, f2ord(mny));
            atomicMax(&bounds->max_x, f2ord(mxx)); atomicMax(&bounds->max_y, f2ord(mxy));
            atomicMax(&bounds->max_extent, f2ord(ext));
            atomicMax(&bounds->max_domain, dom);
            atomicAdd(&bounds->n_live, cnt);
            atomicAdd(&bounds->sum_extent, double(sum));
        }
    }
}

SFG_DI int level_of(const GridParams& g, float ext) {
    int l = 0;
    while (l < g.n_levels && !(ext <= g.cell[l])) ++l;
    return l;   // == n_levels -> huge list
}
SFG_DI void cell_of(const GridParams& g, int l, float cx, float cy, int* ix, int* iy) {
    int x = int(floorf((cx - g.ox) * g.inv_cell[l])), y = int(floorf((cy - g.oy) * g.inv_cell[l]));
    *ix = min(max(x, 0), g.nx[l] - 1);
    *iy = min(max(y, 0), g.ny[l] - 1);
}

SFG_DI uint32_t cell_key(const GridParams& g, uint32_t base, int l, uint32_t domain, int ix, int iy) {
    return base + (domain * uint32_t(g.ny[l]) + uint32_t(iy)) * uint32_t(g.nx[l]) + uint32_t(ix);
}
// entries a collider contributes to the sorted list: its home entry, plus proxies or a big-plain entry if it is large
SFG_DI uint32_t cell_entry_count(const GridParams& g, float4 b, int* level, int* ix0, int* iy0, int* ix1, int* iy1) {
    const float ext = fmaxf(b.z - b.x, b.w - b.y);
    const int l = level_of(g, ext);
    *level = l;
    if (l == 0 || l >= g.n_levels) return 1u;
    cell_of(g, 0, b.x, b.y, ix0, iy0);
    cell_of(g, 0, b.z, b.w, ix1, iy1);
    const uint32_t n = uint32_t(*ix1 - *ix0 + 1) * uint32_t(*iy1 - *iy0 + 1);
    return n <= RASTER_MAX ? 1u + n : 2u;
}
__global__ void __launch_bounds__(256) k_cell_count(const __grid_constant__ GridParams g, uint32_t n_colls, const float4* __restrict__ AABB,
                                                    uint32_t* __restrict__ counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_colls) return;
    const float4 b = AABB[i];          // k_aabb writes NaN boxes for colliders that sit the tick out
    uint32_t n = 0;
    if (isfinite(b.x) && isfinite(b.y) && isfinite(b.z) && isfinite(b.w)) { int l, a0, a1, a2, a3; n = cell_entry_count(g, b, &l, &a0, &a1, &a2, &a3); }
    counts[i] = n;
}
__global__ void __launch_bounds__(256) k_cell_emit(const __grid_constant__ GridParams g, uint32_t n_colls, const float4* __restrict__ AABB,
                                                   const int4* __restrict__ CI, const uint32_t* __restrict__ offsets,
                                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_colls) return;
    const float4 b = AABB[i];
    if (!(isfinite(b.x) && isfinite(b.y) && isfinite(b.z) && isfinite(b.w))) return;
    int l, ix0, iy0, ix1, iy1;
    const uint32_t n = cell_entry_count(g, b, &l, &ix0, &iy0, &ix1, &iy1);
    const uint32_t domain = uint32_t(CI[i].z);
    uint32_t o = offsets[i];
    if (l >= g.n_levels) { keys[o] = g.n_cells; vals[o] = i; return; }
    int ix, iy;
    cell_of(g, l, 0.5f * (b.x + b.z), 0.5f * (b.y + b.w), &ix, &iy);
    keys[o] = cell_key(g, g.base[l], l, domain, ix, iy); vals[o] = i; ++o;
    if (n == 1u) return;
    if (n == 2u) { keys[o] = cell_key(g, g.bp_base[l], l, domain, ix, iy); vals[o] = i; return; }
    for (int y = iy0; y <= iy1; ++y)
        for (int x = ix0; x <= ix1; ++x) { keys[o] = cell_key(g, g.proxy_base, 0, domain, x, y); vals[o] = i; ++o; }
}

// sorted copies so the pair search streams instead of gathering: SA = box, SI = (slot, body, layer | level << 8 | type << 16, domain)
__global__ void __launch_bounds__(256) k_gather_sorted(const __grid_constant__ GridParams g, uint32_t n, const uint32_t* __restrict__ keys,
                                                       const uint32_t* __restrict__ vals, const float4* __restrict__ AABB,
                                                       const int4* __restrict__ CI, float4* __restrict__ SA, int4* __restrict__ SI) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t slot = vals[j];
    const uint32_t key = keys[j];
    const int4 ci = CI[slot];
    int level = g.n_levels;
    uint32_t type = ET_HOME;
    if (key >= g.proxy_base) { type = ET_PROXY; level = 0; }
    else if (key > g.n_cells) { type = ET_BIG_PLAIN; for (int l = 1; l < g.n_levels; ++l) if (key >= g.bp_base[l]) level = l; }
    else if (key < g.n_cells) { for (int l = 0; l < g.n_levels; ++l) if (key >= g.base[l]) level = l; }
    SA[j] = AABB[slot];
    SI[j] = make_int4(int(slot), ci.x, ci.y | (level << 8) | int(type << 16), ci.z);
}

// start[c] = first sorted index whose key >= c, for c in [0, n_cells + 1]; start[n_cells + 1] = number of live entries
// (n_cells here = the number of keys, g.n_keys)
__global__ void __launch_bounds__(256) k_cell_starts(const uint32_t* __restrict__ keys, uint32_t n, uint32_t n_cells, uint32_t* __restrict__ start) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    // entry j owns the keys (key[j - 1], key[j]]; a virtual key n_cells + 1 follows the last live entry
    uint32_t k = 0, kp = 1;                     // empty range for threads past the end
    if (j <= n) {
        k = j < n ? min(keys[j], n_cells + 1) : n_cells + 1;
        kp = j == 0 ? 0u : min(keys[j - 1], n_cells + 1) + 1u;
    }
    const bool wide = k + 1u > kp + 32u;       // a run of empty cells (an unpopulated level): the whole warp fills it
    if (!wide) for (uint32_t cidx = kp; cidx <= k; ++cidx) start[cidx] = j;
    uint32_t todo = __ballot_sync(0xFFFFFFFFu, wide);
    while (todo) {
        const int src = __ffs(int(todo)) - 1;
        const uint32_t a = __shfl_sync(0xFFFFFFFFu, kp, src), b = __shfl_sync(0xFFFFFFFFu, k, src), v = __shfl_sync(0xFFFFFFFFu, j, src);
        for (uint32_t cidx = a + lane; cidx <= b; cidx += 32u) start[cidx] = v;
        todo &= todo - 1u;
    }
}
// One lane per HOME entry (in cell-sorted order; home entries sort before the big-plain and proxy ranges). Same-level
// partners are searched in the FORWARD half neighbourhood only (rest of the own cell + cell x+1, then cells x-1..x+1 of row
// y+1): sorted order is a total order, so every unordered pair inside a 3x3 block is met exactly once. A small (level-0)
// collider then looks for large partners in the proxy grid (3x3 around its own cell; a large collider has a proxy in every
// level-0 cell its box overlaps, and the pair is taken only from the proxy whose cell holds the lower-left corner of the two
// boxes' intersection, so it is met exactly once) and in the big-plain levels (3x3); a large collider searches the levels
// from its own upwards. Nobody searches downwards. Hits are compacted with a warp ballot into a per-warp shared-memory
// staging buffer and flushed 32+ at a time behind ONE atomic reservation.
constexpr int PAIR_STAGE = 64;
__global__ void __launch_bounds__(256) k_pairs(const __grid_constant__ GridParams g, uint32_t n_live, const float4* __restrict__ SA,
                                               const int4* __restrict__ SI, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ start,
                                               const unsigned long long* __restrict__ mask, uint32_t* __restrict__ cursor,
                                               uint2* __restrict__ pairs, uint32_t capacity) {
    __shared__ uint2 stage[8][PAIR_STAGE];
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t j = gw * 32 + lane;
    const bool active = j < n_live;
    float4 A = make_float4(0, 0, 0, 0);
    int4 I = make_int4(0, -1, 0, 0);
    if (active) { A = SA[j]; I = SI[j]; }
    const int my_level = (I.z >> 8) & 0xFF, my_layer = I.z & 0xFF;
    const int nl = g.n_levels;
    const float cx = 0.5f * (A.x + A.z), cy = 0.5f * (A.y + A.w);
    const unsigned long long my_mask = active ? mask[my_layer] : 0ull;
    // Segments of a lane, in order: 0, 1 = own level (rest of the own cell + cell x+1, then cells x-1..x+1 of row y+1);
    // 2, 3, 4 = proxy rows y-1, y, y+1 (small colliders only); then the RARE ranges: 3 rows per populated coarser level
    // (big-plain levels for a small collider, levels above its own for a large one) and the huge list. In most worlds
    // every large collider is rasterised and nothing is huge, so the rare mask is 0 and a lane has 5 cheap segments.
    const bool small = my_level == 0 && nl > 0, huge = my_level >= nl;
    uint32_t rare_mask = 0;
    if (active) {
        for (int l = small ? 1 : my_level + 1; l < nl; ++l) {
            const uint32_t kb = small ? g.bp_base[l] : g.base[l], ke = kb + uint32_t(g.nx[l]) * uint32_t(g.ny[l]) * g.n_domains;
            if (__ldg(start + ke) > __ldg(start + kb)) rare_mask |= 1u << l;
        }
        const uint32_t h0 = huge ? j + 1 : __ldg(start + g.n_cells);
        if (__ldg(start + g.n_cells + 1) > h0) rare_mask |= 1u << 31;
    }
    int ix0 = 0, iy0 = 0;
    if (active && !huge) cell_of(g, my_level, cx, cy, &ix0, &iy0);
    const int own_nx = huge ? 1 : g.nx[my_level], own_ny = huge ? 1 : g.ny[my_level];
    const uint32_t own_row = huge ? 0u : g.base[my_level] + (uint32_t(I.w) * uint32_t(own_ny) + uint32_t(iy0)) * uint32_t(own_nx);
    int seg = active ? (huge ? 5 : 0) : 1000, rare_sub = 0;
    bool seg_proxy = false, exhausted = !active;
    uint32_t cur = 0, end = 0;
    uint32_t n_buf = 0;
    const uint32_t lt = (1u << lane) - 1u;
    auto flush = [&]() {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(cursor, n_buf);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        for (uint32_t k = lane; k < n_buf; k += 32) if (base + k < capacity) pairs[base + k] = stage[warp][k];
        __syncwarp();
        n_buf = 0;
    };
    auto rows3 = [&](uint32_t base, int l, int dy) {       // cells x-1..x+1 of row y+dy at level l, key range starting at `base`
        int ix, iy;
        cell_of(g, l, cx, cy, &ix, &iy);
        const int y = iy + dy;
        if (y >= 0 && y < g.ny[l]) {
            const uint32_t row = base + (uint32_t(I.w) * uint32_t(g.ny[l]) + uint32_t(y)) * uint32_t(g.nx[l]);
            cur = start[row + max(ix - 1, 0)]; end = start[row + min(ix + 1, g.nx[l] - 1) + 1];
        } else { cur = 0; end = 0; }
    };
    for (;;) {
        while (cur >= end && !exhausted) {
            seg_proxy = false;
            if (seg == 0) { cur = j + 1; end = start[own_row + min(ix0 + 1, own_nx - 1) + 1]; }
            else if (seg == 1) {
                if (iy0 + 1 < own_ny) { cur = start[own_row + own_nx + max(ix0 - 1, 0)]; end = start[own_row + own_nx + min(ix0 + 1, own_nx - 1) + 1]; }
            } else if (seg < 5) {
                const int y = iy0 + seg - 3;                 // level 0: the proxy grid has the geometry of the own level
                if (small && y >= 0 && y < own_ny) {
                    const uint32_t row = g.proxy_base + (uint32_t(I.w) * uint32_t(own_ny) + uint32_t(y)) * uint32_t(own_nx);
                    cur = start[row + max(ix0 - 1, 0)]; end = start[row + min(ix0 + 1, own_nx - 1) + 1];
                    seg_proxy = true;
                }
            } else if (rare_mask == 0u) exhausted = true;
            else {
                const int l = __ffs(int(rare_mask)) - 1;
                if (l == 31) { cur = huge ? j + 1 : start[g.n_cells]; end = start[g.n_cells + 1]; rare_mask = 0u; }   // huge-vs-huge forward only
                else {
                    rows3(small ? g.bp_base[l] : g.base[l], l, rare_sub - 1);
                    if (++rare_sub == 3) { rare_sub = 0; rare_mask &= ~(1u << l); }
                }
            }
            ++seg;
        }
        const bool more = cur < end;
        if (!__any_sync(0xFFFFFFFFu, more)) break;
        bool hit = false;
        uint32_t hi = 0, lo = 0;
        if (more) {
            const uint32_t cj = cur++;
            const float4 B = __ldg(SA + cj);
            if (aabb_overlap(A, B)) {       // cells of one key share the domain; only the huge list mixes them
                const int4 J = __ldg(SI + cj);
                bool take = J.w == I.w;
                if (take && seg_proxy) {    // only from the proxy in the cell of the intersection's lower-left corner
                    int mx, my;
                    cell_of(g, 0, fmaxf(A.x, B.x), fmaxf(A.y, B.y), &mx, &my);
                    take = __ldg(keys + cj) == cell_key(g, g.proxy_base, 0, uint32_t(I.w), mx, my);
                }
                if (take) {
                    const bool i_later = uint32_t(I.x) > uint32_t(J.x);
                    hi = i_later ? uint32_t(I.x) : uint32_t(J.x); lo = i_later ? uint32_t(J.x) : uint32_t(I.x);
                    const int l_lo = i_later ? (J.z & 0xFF) : my_layer;
                    const unsigned long long mrow = i_later ? my_mask : mask[J.z & 0xFF];   // mask_matrix.get(later.layer, earlier.layer)
                    const bool bodies_ok = (I.y >= 0 || J.y >= 0) && !(I.y >= 0 && I.y == J.y);
                    hit = ((mrow >> l_lo) & 1ull) && bodies_ok;
                }
            }
        }
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, hit);
        if (hit) stage[warp][n_buf + __popc(ballot & lt)] = make_uint2(hi, lo);
        n_buf += __popc(ballot);
        __syncwarp();
        if (n_buf >= 32) flush();
    }
    if (n_buf) flush();
}

// ---------------------------------------------------------------- graph colouring
// unit = one candidate pair (or one user constraint); ub = its dynamic bodies (or -1); prio = sfg_hash3
// world position of a collider's origin at tick start, in the exactly rounded arithmetic of include/sfg_hash.h
SFG_DI V2 collider_centre(int4 ci, float4 l, const float4* __restrict__ P, const float4* __restrict__ Q) {
    if (ci.x < 0) return mk(l.x, l.y);
    const float4 p = P[ci.x], q = Q[ci.x];
    V2 c;
    sfg_collider_centre(__fadd_rn(p.x, p.z), __fadd_rn(p.y, p.w), q.x, q.y, l.x, l.y, &c.x, &c.y);
    return c;
}
// per candidate pair: the bodies it couples (only bodies that see forces count for conflicts), its colouring priority
// (include/sfg_hash.h) and the degree of its bodies. The priority's top bit is the `near` hint: the colliders' bounding
// circles are closer than the bodies can come in one frame at their current relative speed (a hint, not a proof). Its
// definition — sfg_near_hint, exactly rounded f32 — is shared with the CPU oracle, which computes it on its own.
SFG_DI int4 crec_ci(const float4* __restrict__ CREC, uint32_t slot) {
    const float4 v = __ldg(CREC + size_t(slot) * 4 + 3);
    return make_int4(__float_as_int(v.x), __float_as_int(v.y), __float_as_int(v.z), __float_as_int(v.w));
}
__global__ void __launch_bounds__(256) k_pair_unit_setup(uint32_t n_units, const uint2* __restrict__ ids, const float4* __restrict__ CREC,
                                                         const float4* __restrict__ MI, const float4* __restrict__ P,
                                                         const float4* __restrict__ Q, const float4* __restrict__ V, float frame_dt,
                                                         int2* __restrict__ ub, uint32_t* __restrict__ prio, uint32_t* __restrict__ deg,
                                                         uint8_t* __restrict__ ucls) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    const uint2 id = ids[u];
    // whole 64-byte collider records (sfg_api.cu k_unpack_colliders): (shape, local pose, material, ids)
    const float4* r0 = CREC + size_t(id.x) * 4; const float4* r1 = CREC + size_t(id.y) * 4;
    const float4 cs0 = __ldg(r0), cl0 = __ldg(r0 + 1), cs1 = __ldg(r1), cl1 = __ldg(r1 + 1);
    const int4 c0 = crec_ci(CREC, id.x), c1 = crec_ci(CREC, id.y);
    int b0 = c0.x, b1 = c1.x;
    V2 v0 = mk(0.f, 0.f), v1 = v0;
    if (b0 >= 0) { const float4 v = V[b0]; v0 = mk(v.x, v.y); }
    if (b1 >= 0) { const float4 v = V[b1]; v1 = mk(v.x, v.y); }
    const V2 t0 = collider_centre(c0, cl0, P, Q), t1 = collider_centre(c1, cl1, P, Q);
    const Shape s0 = mkshape(cs0), s1 = mkshape(cs1);
    const bool near_hint = sfg_near_hint(t0.x, t0.y, sfg_bound_radius(s0.kind, s0.p0, s0.p1, s0.r), t1.x, t1.y, sfg_bound_radius(s1.kind, s1.p0, s1.p1, s1.r),
                                         v0.x, v0.y, v1.x, v1.y, frame_dt) != 0u;
    if (b0 >= 0 && !(__float_as_uint(MI[b0].z) & (BF_MASS | BF_INERTIA))) b0 = -1;
    if (b1 >= 0 && !(__float_as_uint(MI[b1].z) & (BF_MASS | BF_INERTIA))) b1 = -1;
    ub[u] = make_int2(b0, b1);
    prio[u] = sfg_pair_priority(id.x, id.y, near_hint ? 1u : 0u);
    {   // order inside a colour (k_unit_sort_keys): far units last, the rest grouped by the EXACT pair of polygon kinds and the
        // visit count, so that the lanes of a warp take the same kind switches and loop counts all the way through the SAT
        // (grouping by edge count alone left rectangles next to segments and triangles next to hexagons: the SAT loops ran at
        // 8-10 active lanes in the settled pile)
        const uint32_t k0 = s0.kind, k1 = s1.kind;        // 0 point, 1 segment, 2 rect, 3 triangle, 4 hexagon
        uint32_t path;
        if (k0 == K_POINT && k1 == K_POINT) path = 0;
        else if (k0 == K_POINT || k1 == K_POINT) path = max(k0, k1);                      // 1..4: circle against that polygon kind
        else path = 5u + (k0 - 1u) * 4u + (k1 - 1u);                                      // 5..20: polygon against polygon, ordered
        const uint32_t mult2 = (c0.x >= 0 && c1.x >= 0) ? 1u : 0u;
        ucls[u] = uint8_t(path * 2 + mult2 + (near_hint ? 0u : 64u));
    }
    if (b0 >= 0) atomicAdd(deg + b0, 1u);
    if (b1 >= 0) atomicAdd(deg + b1, 1u);
}
__global__ void __launch_bounds__(256) k_constraint_unit_setup(uint32_t n_units, const int4* __restrict__ KI /* owner, target, flags, - */,
                                                               const float4* __restrict__ MI, uint2* __restrict__ ids, int2* __restrict__ ub,
                                                               uint32_t* __restrict__ prio, uint32_t* __restrict__ deg) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    int b0 = KI[u].x, b1 = KI[u].y;
    if (b0 >= 0 && !(__float_as_uint(MI[b0].z) & (BF_MASS | BF_INERTIA))) b0 = -1;
    if (b1 >= 0 && !(__float_as_uint(MI[b1].z) & (BF_MASS | BF_INERTIA))) b1 = -1;
    ids[u] = make_uint2(u, 0u);
    ub[u] = make_int2(b0, b1);
    prio[u] = sfg_hash3(u, SFG_CONSTRAINT_TAG, 0u);
    if (b0 >= 0) atomicAdd(deg + b0, 1u);
    if (b1 >= 0) atomicAdd(deg + b1, 1u);
}
// adjacency lists body -> (unit, priority of the unit); also the bucket key of the colouring pass (descending priority)
__global__ void __launch_bounds__(256) k_adj_fill(uint32_t n_units, const int2* __restrict__ ub, const uint32_t* __restrict__ prio,
                                                  const uint32_t* __restrict__ adj_start, uint32_t* __restrict__ cursor,
                                                  uint2* __restrict__ adj, uint32_t* __restrict__ adj_body, int4* __restrict__ urec,
                                                  int bucket_shift, uint32_t* __restrict__ bucket_key, uint32_t* __restrict__ order) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    const int2 b = ub[u];
    const uint32_t p = prio[u];
    // every list slot also remembers whose list it is and which side of the unit it stands for (top bit)
    if (b.x >= 0) { const uint32_t q = adj_start[b.x] + atomicAdd(cursor + b.x, 1u); adj[q] = make_uint2(u, p); adj_body[q] = uint32_t(b.x); }
    if (b.y >= 0) { const uint32_t q = adj_start[b.y] + atomicAdd(cursor + b.y, 1u); adj[q] = make_uint2(u, p); adj_body[q] = uint32_t(b.y) | 0x80000000u; }
    urec[u] = make_int4(b.x, b.y, 0, 0);       // k_adj_rank fills in the two ranks
    bucket_key[u] = (~p) >> bucket_shift;
    order[u] = u;
}

// Sort key of a coloured unit: colour-major; inside a colour, units whose bounding circles are clearly apart at tick
// start come last (the contact stage rejects them with one distance test, so whole warps take the short path), and the
// rest are grouped by narrowphase path (circle-circle / circle-polygon by polygon kind / polygon-polygon by edge count)
// and visit count, so the 32 lanes of a warp run the same branches. The order inside a colour cannot change results
// (no shared dynamic body), only divergence.
__global__ void __launch_bounds__(256) k_unit_sort_keys(uint32_t n_units, const uint32_t* __restrict__ colour, const uint8_t* __restrict__ ucls,
                                                        uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ near_count) {
    __shared__ uint32_t s_near[64];
    if (threadIdx.x < 64) s_near[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n_units) {
        const uint32_t cls = ucls ? ucls[u] : 0u;            // k_pair_unit_setup; constraints have no classes
        const uint32_t col = colour[u];
        keys[u] = col * 128u + cls;
        vals[u] = u;                                          // the sort's payload: the unit's index before the sort
        // how many units of each colour sort into its "may touch" head: the solver deals those out a few lanes per warp
        // (sfg_solver.cuh k_solve_persistent). Counted per block in shared memory, one global atomic per colour and block.
        if (near_count && !(cls & 64u)) {
            if (col < 64u) atomicAdd(&s_near[col], 1u);
            else atomicAdd(near_count + col, 1u);
        }
    }
    __syncthreads();
    if (near_count && threadIdx.x < 64 && s_near[threadIdx.x]) atomicAdd(near_count + threadIdx.x, s_near[threadIdx.x]);
}

constexpr uint32_t UNCOLOURED = 0xFFFFFFFFu;

// Greedy colouring in descending priority (include/sfg_hash.h): a unit takes the smallest colour unused by its
// higher-priority neighbours (units that share a body with it). The definition is sequential; here it runs in ONE pass
// as a dataflow with a ticket per body. Units are edges between bodies, so "the higher-priority neighbours of u" are the
// units ahead of u in the priority order of its (at most) two bodies. Every body carries one 64-bit word
//     ticket (24 bits: how many of its units are coloured) | mask (40 bits: which of colours 0..39 they took).
// A unit knows its rank in each body's list (k_adj_rank) and polls the two words until both tickets equal its ranks: at that moment exactly its higher-priority neighbours are coloured, the masks ARE
// their colours, and the unit takes the lowest free bit, ors it in and bumps both tickets — two loads and two stores per
// unit instead of a colour word per neighbour. Only one unit at a time holds a body's current ticket, so the words need
// no atomics, and the polled word is the data, so they need no fences. If colours 0..39 are all taken the unit falls
// back to reading its neighbours' colour words (polled until set) for the exact smallest free colour.
// The result is the fixed point Jones-Plassmann rounds converge to, bit-identical to the oracle's greedy loop.
// Progress: the highest-priority uncoloured unit is never kept waiting; units are bucket-sorted by the top bits of the
// priority (k_adj_fill + one radix pass) and dealt to warps in 32-unit chunks of that order, a bucket spans far fewer
// chunks than there are co-resident warps (checked on the host), so the warp owning that unit's chunk is running it.
// ctrl: [1] max colour + 1, [4] 1, [5] != 0: a body has 2^24 or more units (the host reports SFG_E_CAPACITY), [6] != 0: the watchdog
// fired. hist = units per colour.
constexpr int CM_BITS = 40;
constexpr unsigned long long CM_MASK = (1ull << CM_BITS) - 1ull;
SFG_DI unsigned long long ld_relaxed64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
SFG_DI void st_relaxed64(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
SFG_DI bool unit_higher(uint2 vp, uint32_t pu, uint2 iu, const uint2* __restrict__ ids) {   // does neighbour vp = (unit, priority) come before (pu, iu)?
    if (vp.y != pu) return vp.y > pu;
    const uint2 iv = ids[vp.x];
    return iv.x > iu.x || (iv.x == iu.x && iv.y > iu.y);
}
// rank of every adjacency entry inside its body's list = how many of that body's other units come before it. One thread per
// list slot: the lanes of a warp walk the same one to three lists, so the walk is broadcast loads (the same scan done by
// each unit for itself touches 2 x degree scattered sectors and was most of the colouring time in crowded worlds).
__global__ void __launch_bounds__(256) k_adj_rank(uint32_t n_slots, const uint2* __restrict__ adj, const uint32_t* __restrict__ adj_body,
                                                  const uint2* __restrict__ ids, const uint32_t* __restrict__ adj_start,
                                                  int4* __restrict__ urec, uint2* __restrict__ uslot, uint32_t* ctrl, uint32_t n_bodies) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_slots || q >= adj_start[n_bodies]) return;      // slots in use = sum of the degrees (static sides have no list)
    const uint2 me = adj[q];
    const uint32_t ab = adj_body[q], body = ab & 0x7FFFFFFFu;
    const uint32_t a0 = adj_start[body], a1 = adj_start[body + 1];
    uint32_t r = 0;
    for (uint32_t a = a0; a < a1; ++a) {
        const uint2 vp = adj[a];
        if (vp.x == me.x) continue;
        if (vp.y != me.y) r += vp.y > me.y ? 1u : 0u;
        else {                                                 // equal hashes (rare): the pair's collider slots decide
            const uint2 iv = ids[vp.x], iu = ids[me.x];
            r += (iv.x > iu.x || (iv.x == iu.x && iv.y > iu.y)) ? 1u : 0u;
        }
    }
    if (a1 - a0 >= (1u << (64 - CM_BITS))) ctrl[5] = 1u;
    // the colouring pass reads ONE 16-byte record per unit: (body 0, body 1, rank in body 0's list, rank in body 1's list).
    // A list longer than the 40 colours of the body word is "crowded" (top bit of the rank): its units also write their colour
    // next to their list entry (adj_colour, k_colour_dataflow), so they have to know where that entry is.
    const bool crowded = a1 - a0 > uint32_t(CM_BITS);
    reinterpret_cast<int*>(urec + me.x)[(ab >> 31) ? 3 : 2] = int(r | (crowded ? 0x80000000u : 0u));
    if (crowded) reinterpret_cast<uint32_t*>(uslot + me.x)[(ab >> 31) ? 1 : 0] = q;
}
// Exact smallest colour unused by the higher-priority neighbours of unit u, for a unit whose two body words show colours
// 0..39 all taken. The WHOLE WARP does it for one of its lanes (all arguments are that lane's values, broadcast): the 32
// lanes stride over the two lists and or their bit sets together. A lane doing this alone keeps its 31 neighbours from
// polling for the 50 us a 300-entry list takes, and with them every unit queued behind their tickets — measured as ticks
// of 0.2 to 1 s in a rope pile with 290 colours. Every higher-priority neighbour is coloured by the time this runs (its
// ticket came up); a colour word still unset only means its store is in flight, then the pass is repeated. A crowded body's
// colours sit next to its list (adj_colour): contiguous reads instead of a gather per neighbour. One pass collects colours
// base..base+511 in a bit set; beyond that, the next 512.
SFG_DI uint32_t colour_scan_warp(uint32_t lane, uint32_t u, int bx, int by, bool crowded0, bool crowded1, uint32_t pu, uint2 iu,
                                 const uint2* __restrict__ ids, const uint32_t* __restrict__ adj_start, const uint2* __restrict__ adj,
                                 const uint32_t* colour, const uint32_t* adj_colour, uint32_t* ctrl) {
    for (uint32_t base = 0;; base += 512u) {
        uint32_t m[16];
        bool beyond, retry;
        uint32_t spins = 0;
        do {
            retry = false; beyond = false;
#pragma unroll
            for (int i = 0; i < 16; ++i) m[i] = 0u;
            for (int sd = 0; sd < 2; ++sd) {
                const int body = sd ? by : bx;
                if (body < 0 || (sd && by == bx)) continue;
                const bool crowded = sd ? crowded1 : crowded0;
                const uint32_t q0 = adj_start[body], q1 = adj_start[body + 1];
                for (uint32_t q = q0 + lane; q < q1; q += 32u) {
                    const uint2 vp = adj[q];
                    const uint32_t cv = crowded ? ld_relaxed(adj_colour + q) : ld_relaxed(colour + vp.x);
                    if (vp.x == u || !unit_higher(vp, pu, iu, ids)) continue;
                    if (cv == 0xFFFFFFFFu) { retry = true; continue; }
                    if (cv < base) continue;
                    const uint32_t rel = cv - base;
                    if (rel < 512u) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) if (int(rel >> 5) == i) m[i] |= 1u << (rel & 31u);   // static indices: the set stays in registers
                    } else beyond = true;
                }
            }
            retry = __any_sync(0xFFFFFFFFu, retry);
            if (retry) __nanosleep(100);
        } while (retry && ++spins < (1u << 20));              // bounded: see the watchdog in k_colour_dataflow
        if (retry) ctrl[6] = 2u;                               // a neighbour's colour never arrived: reported like the watchdog
        beyond = __any_sync(0xFFFFFFFFu, beyond);
        uint32_t found = 0xFFFFFFFFu;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t all = __reduce_or_sync(0xFFFFFFFFu, m[i]);
            if (found == 0xFFFFFFFFu && all != 0xFFFFFFFFu) found = base + uint32_t(i) * 32u + uint32_t(__ffs(int(~all)) - 1);
        }
        if (found != 0xFFFFFFFFu) return found;
        if (!beyond) return base + 512u;
    }
}
#ifndef SFG_COLOUR_THREADS
#define SFG_COLOUR_THREADS 1024   // one 1024-thread CTA per SM (56 registers). Measured: 512 x 3 (40 registers) and 512 x 4 (32, spilling)
#define SFG_COLOUR_BLOCKS 1       // resident CTAs per SM are 2-5 % slower on the graph phase of C3, C2 and C5: the pass is not short of warps
#endif
__global__ void __launch_bounds__(SFG_COLOUR_THREADS, SFG_COLOUR_BLOCKS) k_colour_dataflow(uint32_t n_units, const uint32_t* __restrict__ order, const int2* __restrict__ ub,
                                                         const uint32_t* __restrict__ prio, const uint2* __restrict__ ids,
                                                         const uint32_t* __restrict__ adj_start, const uint2* __restrict__ adj,
                                                         const int4* __restrict__ urec, const uint2* __restrict__ uslot, uint32_t* adj_colour,
                                                         uint32_t* colour, unsigned long long* body_word, uint32_t* ctrl, uint32_t* hist, uint32_t hist_cap) {
    __shared__ uint32_t s_hist[64];
    __shared__ uint32_t s_max;
    if (threadIdx.x < 64) s_hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warps_per_cta = blockDim.x >> 5;
    const uint32_t n_warps = gridDim.x * warps_per_cta;
    // chunk c -> CTA c mod gridDim (spread over SMs), warp slot (c / gridDim) mod warps_per_cta
    for (uint32_t chunk = (threadIdx.x >> 5) * gridDim.x + blockIdx.x; chunk * 32u < n_units; chunk += n_warps) {
        const uint32_t idx = chunk * 32u + lane;
        bool done = idx >= n_units;
        uint32_t u = 0, pu = 0;
        int2 b = make_int2(-1, -1);
        uint2 iu = make_uint2(0, 0);
        unsigned long long want0 = 0, want1 = 0;        // ticket the body must show, already shifted
        uint32_t step0 = 1;                            // a unit whose two sides are the same body sits twice in that list
        bool crowded0 = false, crowded1 = false;       // that body's list is longer than the 40 colours of its word
        if (!done) {
            u = order[idx];
            const int4 rec = urec[u];
            b = make_int2(rec.x, rec.y);
            crowded0 = rec.z < 0; crowded1 = rec.w < 0;
            if (b.x >= 0) want0 = (unsigned long long)(uint32_t(rec.z) & 0x7FFFFFFFu) << CM_BITS;
            if (b.y == b.x && b.x >= 0) { b.y = -1; step0 = 2; }
            if (b.y >= 0) want1 = (unsigned long long)(uint32_t(rec.w) & 0x7FFFFFFFu) << CM_BITS;
        }
        uint32_t polls = 0;
        while (__any_sync(0xFFFFFFFFu, !done)) {
            bool ready = false;
            unsigned long long w0 = 0ull, w1 = 0ull;
            if (!done) {
                w0 = b.x >= 0 ? ld_relaxed64(body_word + b.x) : 0ull;
                w1 = b.y >= 0 ? ld_relaxed64(body_word + b.y) : 0ull;
                // watchdog: a ticket that never comes up (it cannot, by the progress argument above) must not hang the device
                if (++polls > (1u << 22)) { ctrl[6] = 1u; done = true; }
                else ready = (w0 & ~CM_MASK) == want0 && (w1 & ~CM_MASK) == want1;
            }
            const unsigned long long used = (w0 | w1) & CM_MASK;
            uint32_t result = used != CM_MASK ? uint32_t(__ffsll((long long)~used) - 1) : 0u;
            // lanes whose 40 mask bits are all taken: the warp scans their lists together, one lane's unit at a time
            uint32_t scan = __ballot_sync(0xFFFFFFFFu, ready && used == CM_MASK);
            while (scan) {
                const int L = __ffs(int(scan)) - 1;
                scan &= scan - 1u;
                const uint32_t Lu = __shfl_sync(0xFFFFFFFFu, u, L);
                const int Lbx = __shfl_sync(0xFFFFFFFFu, b.x, L);
                const int Lby = __shfl_sync(0xFFFFFFFFu, step0 == 2 ? b.x : b.y, L);
                const bool Lc0 = __shfl_sync(0xFFFFFFFFu, crowded0 ? 1 : 0, L) != 0, Lc1 = __shfl_sync(0xFFFFFFFFu, (step0 == 2 ? crowded0 : crowded1) ? 1 : 0, L) != 0;
                uint32_t Lpu = 0; uint2 Liu = make_uint2(0, 0);
                if (int(lane) == L) { Lpu = prio[u]; Liu = ids[u]; }
                Lpu = __shfl_sync(0xFFFFFFFFu, Lpu, L); Liu.x = __shfl_sync(0xFFFFFFFFu, Liu.x, L); Liu.y = __shfl_sync(0xFFFFFFFFu, Liu.y, L);
                const uint32_t r = colour_scan_warp(lane, Lu, Lbx, Lby, Lc0, Lc1, Lpu, Liu, ids, adj_start, adj, colour, adj_colour, ctrl);
                if (int(lane) == L) result = r;
            }
            if (ready) {
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(colour + u), "r"(result) : "memory");
                if (crowded0 || crowded1) {             // the colour also goes next to the unit's entries in crowded lists
                    const uint2 sl = uslot[u];
                    if (crowded0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(adj_colour + sl.x), "r"(result) : "memory");
                    if (crowded1) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(adj_colour + sl.y), "r"(result) : "memory");
                }
                const unsigned long long bit = result < uint32_t(CM_BITS) ? 1ull << result : 0ull;
                if (b.x >= 0) st_relaxed64(body_word + b.x, (w0 | bit) + ((unsigned long long)step0 << CM_BITS));
                if (b.y >= 0) st_relaxed64(body_word + b.y, (w1 | bit) + (1ull << CM_BITS));
                if (result < 64) atomicAdd(&s_hist[result], 1u);
                else if (result < hist_cap) atomicAdd(hist + result, 1u);
                atomicMax(&s_max, result + 1u);
                done = true;
            } else if (!done) __nanosleep(100);
        }
    }
    __syncthreads();
    if (threadIdx.x < 64 && s_hist[threadIdx.x]) atomicAdd(hist + threadIdx.x, s_hist[threadIdx.x]);
    if (threadIdx.x == 0) { if (s_max) atomicMax(ctrl + 1, s_max); ctrl[4] = 1u; }
}

// after the stable sort by colour: write the streaming columns of each pair unit
__global__ void __launch_bounds__(256) k_build_pair_units(uint32_t n_units, const uint32_t* __restrict__ perm, const uint2* __restrict__ ids,
                                                          const float4* __restrict__ CREC,
                                                          const float4* __restrict__ MI, const int* __restrict__ rope_next,
                                                          const int* __restrict__ rope_prev, float4* __restrict__ H, float4* __restrict__ S,
                                                          float4* __restrict__ G0, float4* __restrict__ G1, float4* __restrict__ L0,
                                                          float4* __restrict__ L1, float* __restrict__ RC, uint32_t* __restrict__ counts) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = s < n_units;
    uint2 id = make_uint2(0u, 0u);
    if (in) id = ids[__ldcs(perm + s)];
    {   // statistics: how many units are body-body pairs (visited twice per substep, physics.rs:641,651) -> counts[0]
        const bool two = in && __float_as_int(__ldg(CREC + size_t(id.x) * 4 + 3).x) >= 0 && __float_as_int(__ldg(CREC + size_t(id.y) * 4 + 3).x) >= 0;
        const uint32_t n2 = __syncthreads_count(two);
        if (threadIdx.x == 0 && n2) atomicAdd(counts, n2);
    }
    if (!in) return;
    const float4* r0 = CREC + size_t(id.x) * 4; const float4* r1 = CREC + size_t(id.y) * 4;      // whole 64-byte collider records
    const float4 g0 = __ldg(r0), l0 = __ldg(r0 + 1), m0 = __ldg(r0 + 2), g1 = __ldg(r1), l1 = __ldg(r1 + 1), m1 = __ldg(r1 + 2);
    const int4 c0 = crec_ci(CREC, id.x), c1 = crec_ci(CREC, id.y);
    const uint32_t f0 = uint32_t(c0.w), f1 = uint32_t(c1.w);
    uint32_t fl = 0;
    if (c0.x >= 0 && c1.x >= 0) fl |= UF_MULT2;
    if ((f0 | f1) & CF_SENSOR) fl |= UF_SENSOR;
    if ((f0 & f1) & CF_HAS_SF) fl |= UF_HAS_SF;
    if ((f0 & f1) & CF_HAS_DF) fl |= UF_HAS_DF;
    if (c0.x >= 0) {
        const uint32_t bf = __float_as_uint(MI[c0.x].z);
        if (bf & (BF_MASS | BF_INERTIA)) fl |= UF_SEES0;
        if (bf & BF_ASLEEP) fl |= UF_ASLEEP;
        if (rope_next && rope_next[c0.x] >= 0 && rope_prev[c0.x] >= 0) fl |= UF_ROPE0;
    }
    if (c1.x >= 0) {
        const uint32_t bf = __float_as_uint(MI[c1.x].z);
        if (bf & (BF_MASS | BF_INERTIA)) fl |= UF_SEES1;
        if (bf & BF_ASLEEP) fl |= UF_ASLEEP;
        if (rope_next && rope_next[c1.x] >= 0 && rope_prev[c1.x] >= 0) fl |= UF_ROPE1;
    }
    // 100 bytes per unit stream out here (0.4 GB in the settled C3 pile) while the 64-byte collider records are gathered at random:
    // the stores are marked streaming (evict-first) so that they do not push the records out of L2
    __stcs(H + s, make_float4(__int_as_float(c0.x), __int_as_float(c1.x), __uint_as_float(fl), (m0.x + m1.x) / 2.0f));
    __stcs(S + s, make_float4(__uint_as_float(id.x), __uint_as_float(id.y), (m0.y + m1.y) / 2.0f, fmaxf(m0.z, m1.z)));
    __stcs(G0 + s, g0); __stcs(G1 + s, g1);
    __stcs(L0 + s, l0); __stcs(L1 + s, l1);
    // bodies (or static collider origins) farther apart than this cannot touch: the colliders' bounding radii, never less
    // than circle_circle's "closer than sqrt(0.001) always collides" distance, plus how far each collider sits from its
    // body, plus an f32 safety margin (sfg_solver.cuh stage_contacts)
    float reach = fmaxf(bound_radius(mkshape(g0)) + bound_radius(mkshape(g1)), 0.0332f);
    if (c0.x >= 0) reach += sqrtf(l0.x * l0.x + l0.y * l0.y);
    if (c1.x >= 0) reach += sqrtf(l1.x * l1.x + l1.y * l1.y);
    __stcs(RC + s, reach * 1.001f + 2e-3f);
}
__global__ void __launch_bounds__(256) k_build_constraint_units(uint32_t n_units, const uint32_t* __restrict__ perm, const int4* __restrict__ KI,
                                                                const float4* __restrict__ KA, const float4* __restrict__ KB,
                                                                float4* __restrict__ K0, float4* __restrict__ K1, float4* __restrict__ K2) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_units) return;
    const uint32_t u = perm[s];
    const int4 ki = KI[u];
    K0[s] = make_float4(__int_as_float(ki.x), __int_as_float(ki.y), __int_as_float(ki.z), __uint_as_float(u));
    K1[s] = KA[u];
    K2[s] = KB[u];
}

}  // namespace sfg

// CUB-free device primitives for the broadphase / graph build: an LSD radix sort whose ranking runs
// in shared memory (8-bit digits, warp match-any multisplit, stable) with ONE kernel per digit (the digit histograms of all
// passes are taken in one sweep up front, a tile's place among its predecessors comes from a decoupled look-back inside the
// scatter kernel), and a one-pass exclusive scan (decoupled look-back).
// Written for sm_100a: 256-thread CTAs, 2048-key tiles, grids sized by the tile count.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sfg {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;   // keys per CTA
constexpr int SORT_WARPS = SORT_THREADS / 32;

// ---- digit histograms of every pass in one sweep over the keys: ghist[pass * 256 + digit] (zeroed by the caller) ----
constexpr int SORT_MAX_PASSES = 4;
__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist_all(const uint32_t* __restrict__ keys, uint32_t n, int passes, uint32_t* __restrict__ ghist) {
    __shared__ uint32_t sh[SORT_MAX_PASSES][256];
    for (int p = 0; p < SORT_MAX_PASSES; ++p) sh[p][threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * SORT_THREADS + threadIdx.x; i < n; i += gridDim.x * SORT_THREADS) {
        const uint32_t k = keys[i];
        for (int p = 0; p < passes; ++p) atomicAdd(&sh[p][(k >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for (int p = 0; p < passes; ++p) { const uint32_t c = sh[p][threadIdx.x]; if (c) atomicAdd(ghist + p * 256 + threadIdx.x, c); }
}

// ---- exclusive scan helpers (used by the scatter kernel and the scan kernels) ----
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nw ? wsum[lane] : 0, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, o); if (lane >= o) wi += t; }
        wsum[lane] = wi - w;
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    uint32_t res = wsum[warp] + inc - v;
    __syncthreads();
    return res;
}


// ---- scatter: stable rank inside the tile in shared memory, the tile re-ordered there, then a coalesced write ----
// Order inside a tile is (warp, round, lane): warp w owns keys [w*256, (w+1)*256) of the tile. After the ranking every key knows
// its place in the tile's digit-sorted order; keys and values are moved there in shared memory, and the tile is written out in
// that order, so that the threads of a warp store runs of consecutive addresses (one run per digit) instead of 32 scattered words.
// Where the tile's keys of digit d start in the output = (keys of smaller digits anywhere: exclusive scan of the pass's global
// histogram) + (keys of digit d in the tiles in front of this one: decoupled look-back, one 32-bit word per tile and digit,
// flag << 30 | count, exactly like k_scan_lookback below but 256 chains side by side, thread d walking chain d). Tiles take their
// numbers from a ticket counter, so a tile's predecessors are always running or done.
constexpr uint32_t SORT_FLAG_AGG = 1u << 30, SORT_FLAG_PREFIX = 2u << 30, SORT_VALUE_MASK = (1u << 30) - 1u;
__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                               uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n,
                                                               int shift, const uint32_t* __restrict__ ghist, uint32_t* __restrict__ state,
                                                               uint32_t* __restrict__ ticket) {
    __shared__ uint32_t warp_hist[SORT_WARPS][256];
    __shared__ uint32_t digit_base[256];          // global position of the tile's first key of each digit, minus its place in the tile
    __shared__ uint32_t s_key[SORT_TILE], s_val[SORT_TILE];
    __shared__ uint32_t s_tot, s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < SORT_WARPS * 256; i += SORT_THREADS) (&warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * SORT_TILE + warp * (32 * SORT_ITEMS);
    uint32_t key[SORT_ITEMS], val[SORT_ITEMS], rank[SORT_ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const uint32_t i = base + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? keys_in[i] : 0xFFFFFFFFu;
        val[r] = ok ? vals_in[i] : 0u;
        const uint32_t d = ok ? ((key[r] >> shift) & 255u) : 256u;   // 256 = out of range, ranked nowhere
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t b = 0;
        if (ok && lane == leader) { b = warp_hist[warp][d]; warp_hist[warp][d] = b + __popc(peers); }
        b = __shfl_sync(0xFFFFFFFFu, b, leader);
        rank[r] = b + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    // exclusive scan of each digit over the warps and of the digit totals over the digits; the tile's count of digit d is published
    // at once (AGGREGATE), the look-back for what lies in front of it waits until the tile has been re-ordered in shared memory:
    // by then the predecessors have usually published their prefixes, and the walk is a step or two
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) { uint32_t c = warp_hist[w][d]; warp_hist[w][d] = run; run += c; }
    uint32_t* my = state + size_t(tile) * 256 + d;
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(my), "r"((tile == 0 ? SORT_FLAG_PREFIX : SORT_FLAG_AGG) | run) : "memory");
    const uint32_t local = block_exclusive_scan(run, &s_tot);            // where digit d starts inside the sorted tile
    const uint32_t smaller = block_exclusive_scan(ghist[d], &s_tot);     // keys of smaller digits in the whole array
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) warp_hist[w][d] += local;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const uint32_t i = base + r * 32 + lane;
        if (i < n) {
            const uint32_t dg = (key[r] >> shift) & 255u;
            const uint32_t pos = warp_hist[warp][dg] + rank[r];
            s_key[pos] = key[r];
            s_val[pos] = val[r];
        }
    }
    {
        uint32_t before = 0;
        if (tile != 0) {
            // the words of SORT_LOOKBACK predecessors are fetched at once (independent loads), then consumed nearest first
            constexpr int SORT_LOOKBACK = 8;
            bool done = false;
            for (int p = int(tile) - 1; p >= 0 && !done; p -= SORT_LOOKBACK) {
                uint32_t w[SORT_LOOKBACK];
#pragma unroll
                for (int k = 0; k < SORT_LOOKBACK; ++k) {
                    w[k] = SORT_FLAG_PREFIX;                                   // in front of tile 0: an empty prefix
                    if (p - k >= 0) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(w[k]) : "l"(state + size_t(p - k) * 256 + d) : "memory");
                }
#pragma unroll
                for (int k = 0; k < SORT_LOOKBACK; ++k) {
                    if (done) continue;
                    uint32_t x = w[k];
                    while ((x >> 30) == 0u) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(x) : "l"(state + size_t(p - k) * 256 + d) : "memory");
                    before += x & SORT_VALUE_MASK;
                    if ((x >> 30) == 2u) done = true;
                }
            }
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(my), "r"(SORT_FLAG_PREFIX | (before + run)) : "memory");
        }
        digit_base[d] = smaller + before - local;
    }
    __syncthreads();
    const uint32_t tile_n = min(uint32_t(SORT_TILE), n - tile * SORT_TILE);
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const uint32_t j = r * SORT_THREADS + threadIdx.x;
        if (j < tile_n) {
            const uint32_t k = s_key[j];
            const uint32_t pos = digit_base[(k >> shift) & 255u] + j;
            keys_out[pos] = k;
            vals_out[pos] = s_val[j];
        }
    }
}

// ---- exclusive scan (u32) in ONE pass: decoupled look-back over 2048-element tiles ----
// Every tile publishes its own sum as soon as it has it (flag AGGREGATE) and, once it knows the sum of everything in front of
// it, its inclusive prefix (flag PREFIX); a tile finds what is in front of it by walking back over its predecessors' words,
// adding aggregates until it meets a prefix. Tiles take their numbers from a ticket counter in the order they start running,
// so a tile's predecessors are always running or done (no deadlock whatever the block scheduling order is). One word per tile:
// flag << 62 | value. Replaces reduce + scan-of-sums + apply (three launches and two dependent round trips through a tiny kernel).
constexpr unsigned long long SCAN_FLAG_AGG = 1ull << 62, SCAN_FLAG_PREFIX = 2ull << 62, SCAN_VALUE_MASK = 0xFFFFFFFFull;
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_lookback(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n,
                                                                unsigned long long* __restrict__ state, uint32_t* __restrict__ ticket,
                                                                uint32_t* __restrict__ total_out) {
    __shared__ uint32_t s_tile, s_tot, s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile, n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t v[SCAN_ITEMS], s = 0;
    const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = base + k < n ? in[base + k] : 0; s += v[k]; }
    uint32_t e = block_exclusive_scan(s, &s_tot);
    if (threadIdx.x == 0) {
        const uint32_t total = s_tot;
        uint32_t exclusive = 0;
        if (tile == 0) {
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(state), "l"(SCAN_FLAG_PREFIX | total) : "memory");
        } else {
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(state + tile), "l"(SCAN_FLAG_AGG | total) : "memory");
            for (uint32_t p = tile; p-- > 0;) {
                unsigned long long w;
                do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(state + p) : "memory"); } while ((w >> 62) == 0ull);
                exclusive += uint32_t(w & SCAN_VALUE_MASK);
                if ((w >> 62) == 2ull) break;
            }
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(state + tile), "l"(SCAN_FLAG_PREFIX | (unsigned long long)(exclusive + total)) : "memory");
        }
        s_prefix = exclusive;
        if (tile + 1 == n_tiles && total_out) *total_out = exclusive + total;
    }
    __syncthreads();
    e += s_prefix;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { if (base + k < n) out[base + k] = e; e += v[k]; }
}

// Host-side drivers. `tmp` must hold at least scan_tmp_words(n) uint32 (8-byte aligned: every caller passes a cudaMalloc'ed block or an even offset into one).
inline size_t scan_tmp_words(size_t n) { return 2 * ((n + SCAN_TILE - 1) / SCAN_TILE + 2); }
// exclusive scan of in[0..n) into out[0..n) (in == out allowed); the grand total goes to *total_dev if not null
inline int scan_exclusive(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* tmp, uint32_t* total_dev, cudaStream_t st) {
    if (n == 0) { if (total_dev) cudaMemsetAsync(total_dev, 0, 4, st); return 0; }
    const uint32_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    cudaMemsetAsync(tmp, 0, size_t(nb + 1) * 8, st);                       // ticket (first 8 bytes) + one state word per tile
    k_scan_lookback<<<nb, SCAN_THREADS, 0, st>>>(in, out, n, reinterpret_cast<unsigned long long*>(tmp) + 1, tmp, total_dev);
    return 1;
}

// tmp layout (uint32): [SORT_MAX_PASSES][256] global digit histograms, 8 ticket words, then [passes][tiles][256] look-back words
inline size_t sort_tmp_words(size_t n) {
    size_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
    return SORT_MAX_PASSES * 256 + 8 + size_t(SORT_MAX_PASSES) * 256 * tiles;
}
// Sorts (keys, vals) by the low `bits` bits of the key, stable. Ping-pongs between (k0, v0) and (k1, v1);
// returns 0 if the result is in (k0, v0), 1 if in (k1, v1); *launches is incremented by the kernels launched.
inline int radix_sort_pairs(uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, uint32_t n, int bits, uint32_t* tmp,
                            cudaStream_t st, int* launches) {
    if (n == 0 || bits <= 0) return 0;
    const uint32_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
    const int passes = bits >= 32 ? SORT_MAX_PASSES : (bits + 7) / 8;
    uint32_t* ghist = tmp;
    uint32_t* tickets = tmp + SORT_MAX_PASSES * 256;
    uint32_t* state = tickets + 8;
    cudaMemsetAsync(tmp, 0, (SORT_MAX_PASSES * 256 + 8 + size_t(passes) * 256 * tiles) * 4, st);
    k_sort_hist_all<<<tiles, SORT_THREADS, 0, st>>>(k0, n, passes, ghist);
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        uint32_t *ki = cur ? k1 : k0, *vi = cur ? v1 : v0, *ko = cur ? k0 : k1, *vo = cur ? v0 : v1;
        k_sort_scatter<<<tiles, SORT_THREADS, 0, st>>>(ki, vi, ko, vo, n, 8 * p, ghist + p * 256, state + size_t(p) * 256 * tiles, tickets + p);
        cur ^= 1;
    }
    if (launches) *launches += 1 + passes;
    return cur;
}

}  // namespace sfg

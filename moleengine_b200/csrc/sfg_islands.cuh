// Islands and sleeping on the device (once per tick, between the broadphase and the colouring).
//
// Reference: /root/reference/src/physics.rs:598-741 (recursive DFS islands + sleeping filter), :1097-1144 (contact
// retention, fall asleep), :178-202 (IslandId / SleepingIsland). The DFS is replaced by a lock-free union-find; the
// smallest body slot of each component (an atomicMin per body) is exactly IslandId::first_body (DFS roots are taken in
// ascending slot order, :674-677). Roots are hooked by a HASH of the slot, not by the slot: bodies are usually numbered
// along rows of the scene, and hooking by slot builds chains as long as a row before path halving catches up.
// IslandId::edge_sum is a wrapping sum over directed graph edges,
//      two-body edge (rope link, two-body constraint, body-body pair) seen from body a:  (a + 1) * (b + 1)
//      one-body edge (body-world constraint, body-static pair):                           (a + 1)
// (:634, :651, :662, :669) and does not depend on the visit order, so atomics reproduce it bit for bit; every two-body
// edge is stored on both endpoints (:519-532, :556-569) and therefore counted twice.
//
// Sleeping records live in a table indexed by first_body: the reference keeps at most one record per IslandId and drops
// every record that no island matched in the current tick (:739), so two records can never share a first_body.
#pragma once
#include "../../include/sfg_hash.h"
#include "sfg_state.cuh"

namespace sfg {

enum : uint32_t { IF_NO_SLEEP = 1u, IF_MOVING = 2u, IF_FAST = 4u, IF_ASLEEP = 8u, IF_ROOT = 16u, IF_FOREIGN = 32u /* solved by another rank */ };

struct IslandCtx {
    uint32_t n_bodies, n_units, n_constraints, n_links;
    const uint2* U_ids;           // candidate pairs (collider slots)
    const int4* CI;               // collider -> body
    const int4* KI;               // constraints (owner, target, flags, -)
    const int4* RU;               // rope links (curr, next, after, rope)
    float4* MI;
    const float4* V;
    uint32_t* parent;             // union-find forest, then the island's representative (root) for every body
    uint32_t* isl_first;          // per root: smallest slot of the island = IslandId::first_body
    unsigned long long* isl_sum;  // per root: edge_sum
    uint32_t* isl_flags;          // per root: IF_*
    uint32_t* isl_count;          // per root: bodies
    unsigned long long* sl_sum;   // sleeping record at first_body
    uint32_t* sl_ticks;
    uint32_t* sl_valid;
    uint32_t* counters;           // [0] islands, [1] sleeping bodies, [2] asleep islands
    float sleep_vel_threshold;
    uint32_t fall_asleep_frames;
    uint32_t shard_rank, shard_n;   // island sharding: this rank solves the islands with fmix32(first_body) % shard_n == shard_rank
    uint32_t sleeping;              // 0: islands are only built for sharding, nothing falls asleep
};

__device__ __forceinline__ uint32_t uf_find(const uint32_t* parent, uint32_t x) {
    uint32_t p;
    while ((p = __ldcg(parent + x)) != x) x = p;
    return x;
}
// find with path halving: every node on the way is re-pointed at its grandparent. Racing writers only ever store an
// ancestor of the node, so the forest stays a forest with the same roots.
__device__ __forceinline__ uint32_t uf_find_halve(uint32_t* parent, uint32_t x) {
    uint32_t p = __ldcg(parent + x);
    while (p != x) {
        const uint32_t gp = __ldcg(parent + p);
        if (gp != p) __stcg(parent + x, gp);
        x = p; p = gp;
    }
    return x;
}
// hook the root with the larger hash under the other one (sfg_fmix32 is a bijection, so this is a total order)
__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
    uint32_t ra = uf_find_halve(parent, a), rb = uf_find_halve(parent, b);
    while (ra != rb) {
        if (sfg_fmix32(ra) < sfg_fmix32(rb)) { const uint32_t t = ra; ra = rb; rb = t; }
        const uint32_t old = atomicCAS(parent + ra, ra, rb);
        if (old == ra) return;
        ra = uf_find_halve(parent, old);
        rb = uf_find_halve(parent, rb);
    }
}

__global__ void __launch_bounds__(256) k_isl_init(IslandCtx c) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= c.n_bodies) return;
    c.parent[b] = b; c.isl_first[b] = 0xFFFFFFFFu; c.isl_sum[b] = 0ull; c.isl_flags[b] = 0u; c.isl_count[b] = 0u;
    float4 m = c.MI[b];
    const uint32_t fl = __float_as_uint(m.z);
    if (fl & (BF_ASLEEP | BF_FOREIGN)) { m.z = __uint_as_float(fl & ~(BF_ASLEEP | BF_FOREIGN)); c.MI[b] = m; }
    if (b < 4) c.counters[b] = 0u;
}

// edge index space: [0, n_units) pairs, then constraints, then rope links
__device__ __forceinline__ bool isl_edge(const IslandCtx& c, uint32_t i, int* a, int* b, uint32_t* no_sleep) {
    *no_sleep = 0u;
    if (i < c.n_units) {
        const uint2 id = c.U_ids[i];
        *a = c.CI[id.x].x; *b = c.CI[id.y].x;
        return true;
    }
    i -= c.n_units;
    if (i < c.n_constraints) {
        const int4 k = c.KI[i];
        *a = k.x; *b = k.y;
        if (!(uint32_t(k.z) & 4u)) *no_sleep = 1u;      // SFG_CONSTRAINT_CAN_SLEEP
        return true;
    }
    i -= c.n_constraints;
    if (i < c.n_links) {
        const int4 r = c.RU[i];
        *a = r.x; *b = r.y;
        *no_sleep = 1u;                                   // physics.rs:620
        return true;
    }
    return false;
}
// The edges are united in phases of growing size (1/8, 1/8, 1/4, 1/2 of them, each phase a stride-8 sample of the whole list, so
// every phase touches every neighbourhood), with the forest flattened in between: after the first phases almost every edge finds
// both ends under the same parent in one look and is done, instead of chasing two chains to the root of a pile.
__global__ void __launch_bounds__(256) k_isl_union(IslandCtx c, uint32_t lo, uint32_t width, uint32_t n_edges, uint32_t fast) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = (t / width) * 8u + lo + (t % width);
    if (i >= n_edges) return;
    int a, b; uint32_t ns;
    if (!isl_edge(c, i, &a, &b, &ns)) return;
    if (a < 0 && b < 0) return;
    // edge_sum and "never sleeps" are gathered per BODY here (fire-and-forget reductions on the edge's first dynamic body) and
    // folded into the island's root by k_isl_bodies_pre: wrapping sums and ORs do not care about the order
    const uint32_t home = uint32_t(a >= 0 ? a : b);
    atomicAdd(c.isl_sum + home, a >= 0 && b >= 0 ? 2ull * ((unsigned long long)(a) + 1ull) * ((unsigned long long)(b) + 1ull)
                                                 : (unsigned long long)(home) + 1ull);
    if (ns) atomicOr(c.isl_flags + home, IF_NO_SLEEP);
    if (a < 0 || b < 0 || a == b) return;
    if (fast && __ldcg(c.parent + a) == __ldcg(c.parent + b)) return;     // same parent: same island already
    uf_union(c.parent, uint32_t(a), uint32_t(b));
}
__global__ void __launch_bounds__(256) k_isl_flatten(IslandCtx c) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= c.n_bodies) return;
    const uint32_t p = __ldcg(c.parent + b);
    if (p == b) return;
    const uint32_t root = uf_find(c.parent, p);
    if (root != p) __stcg(c.parent + b, root);
}
// per body: final label, island size, "moved between frames" test on the incoming velocities (physics.rs:722-731)
__global__ void __launch_bounds__(256) k_isl_bodies_pre(IslandCtx c) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= c.n_bodies) return;
    if (!(__float_as_uint(c.MI[b].z) & BF_ALIVE)) { c.sl_valid[b] = 0u; return; }
    const uint32_t root = uf_find(c.parent, b);
    c.parent[b] = root;
    const float4 v = c.V[b];
    const uint32_t moving = v.x * v.x + v.y * v.y + v.z * v.z >= c.sleep_vel_threshold ? 1u : 0u;
    // this body's share of the island's edge_sum / no-sleep flag (k_isl_union); the root's own share is already in place
    const unsigned long long part = b != root ? c.isl_sum[b] : 0ull;
    const uint32_t part_ns = b != root ? (c.isl_flags[b] & IF_NO_SLEEP) : 0u;
    const uint32_t act = __activemask();
    const uint32_t peers = __match_any_sync(act, root);
    const uint32_t peers_moving = __ballot_sync(act, moving != 0u) & peers;
    const uint32_t peers_ns = __ballot_sync(act, part_ns != 0u) & peers;
    unsigned long long total = 0ull;
    for (uint32_t m = peers; m; m &= m - 1u) total += __shfl_sync(peers, part, __ffs(int(m)) - 1);
    if (int(threadIdx.x & 31u) == __ffs(int(peers)) - 1) {   // lanes run in slot order: the group's leader holds its smallest slot
        atomicAdd(c.isl_count + root, uint32_t(__popc(peers)));
        atomicMin(c.isl_first + root, b);
        if (total) atomicAdd(c.isl_sum + root, total);
        const uint32_t want = (peers_moving ? IF_MOVING : 0u) | (peers_ns ? IF_NO_SLEEP : 0u);
        if (want && (__ldcg(c.isl_flags + root) & want) != want) atomicOr(c.isl_flags + root, want);
    }
}
// per island root: the retain() of physics.rs:714-737 and the record clean-up of :739
__global__ void __launch_bounds__(256) k_isl_decide(IslandCtx c) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= c.n_bodies) return;
    if (!(__float_as_uint(c.MI[b].z) & BF_ALIVE) || c.parent[b] != b) return;
    uint32_t fl = c.isl_flags[b] | IF_ROOT;
    const uint32_t first = c.isl_first[b];          // the record of this island, if any, sits at its first_body
    const bool match = c.sl_valid[first] && c.sl_sum[first] == c.isl_sum[b];
    bool continues = false, keep = true;
    if (c.sleeping && match && !(fl & IF_MOVING)) { continues = true; keep = c.sl_ticks[first] < c.fall_asleep_frames; }
    if (!continues) c.sl_valid[first] = 0u;
    if (!keep) { fl |= IF_ASLEEP; atomicAdd(c.counters + 2, 1u); }
    // islands are independent (physics.rs:155-172), so each rank can solve a subset of them and the results are exactly
    // those of one GPU; the owner is a pure function of the island's identity
    if (c.shard_n > 1u && sfg_fmix32(first) % c.shard_n != c.shard_rank) fl |= IF_FOREIGN;
    c.isl_flags[b] = fl;
    atomicAdd(c.counters + 0, 1u);
}
__global__ void __launch_bounds__(256) k_isl_mark(IslandCtx c) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= c.n_bodies) return;
    float4 m = c.MI[b];
    const uint32_t fl = __float_as_uint(m.z);
    if (!(fl & BF_ALIVE)) return;
    const uint32_t root = c.parent[b];
    if (c.isl_first[root] != b) c.sl_valid[b] = 0u;   // a record is only ever found through the island whose first_body it names
    const uint32_t ifl = c.isl_flags[root];
    if (ifl & (IF_ASLEEP | IF_FOREIGN)) {
        m.z = __uint_as_float(fl | BF_ASLEEP | ((ifl & IF_FOREIGN) ? BF_FOREIGN : 0u));
        c.MI[b] = m;
        if (ifl & IF_ASLEEP) atomicAdd(c.counters + 1, 1u);
    }
}
// after the solve: physics.rs:1128-1144
__global__ void __launch_bounds__(256) k_isl_bodies_post(IslandCtx c) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= c.n_bodies) return;
    const uint32_t fl = __float_as_uint(c.MI[b].z);
    if (!(fl & BF_ALIVE) || (fl & BF_ASLEEP)) return;
    const float4 v = c.V[b];
    if (!(v.x * v.x + v.y * v.y + v.z * v.z < c.sleep_vel_threshold)) {
        const uint32_t root = c.parent[b];
        if (!(__ldcg(c.isl_flags + root) & IF_FAST)) atomicOr(c.isl_flags + root, IF_FAST);
    }
}
__global__ void __launch_bounds__(256) k_isl_fall_asleep(IslandCtx c) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= c.n_bodies) return;
    if (!(__float_as_uint(c.MI[b].z) & BF_ALIVE) || c.parent[b] != b) return;
    const uint32_t fl = c.isl_flags[b];
    if (!c.sleeping || (fl & (IF_ASLEEP | IF_NO_SLEEP | IF_FAST | IF_FOREIGN))) return;
    const uint32_t first = c.isl_first[b];
    if (c.sl_valid[first] && c.sl_sum[first] == c.isl_sum[b]) c.sl_ticks[first] += 1u;
    else { c.sl_valid[first] = 1u; c.sl_sum[first] = c.isl_sum[b]; c.sl_ticks[first] = 0u; }
}

// ---- contact list with retention (physics.rs:1097-1122) ----
struct ContactRec { uint32_t c0, c1; float nx, ny; uint32_t root, pad; unsigned long long sum; };

// keep last tick's contacts of islands that are asleep now (record found and slept long enough)
__global__ void __launch_bounds__(256) k_contacts_retain(uint32_t n_old, const ContactRec* __restrict__ old_list, const unsigned long long* sl_sum,
                                                         const uint32_t* sl_ticks, const uint32_t* sl_valid, uint32_t n_bodies,
                                                         uint32_t fall_asleep_frames, ContactRec* __restrict__ out, uint32_t* cursor) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_old) return;
    const ContactRec r = old_list[i];
    if (r.root >= n_bodies) return;
    if (sl_valid[r.root] && sl_sum[r.root] == r.sum && sl_ticks[r.root] >= fall_asleep_frames) out[atomicAdd(cursor, 1u)] = r;
}
// this tick's latched contacts (solver.rs:318-320), tagged with the island they belong to. One thread per unit looks at both of
// its entries (the second visit of a body-body pair latches its own normal) and reads the unit's header and island once; the
// records of a warp go out behind one atomic.
__global__ void __launch_bounds__(256) k_contacts_append(uint32_t n_units, const float2* __restrict__ LN,
                                                         const float4* __restrict__ H, const float4* __restrict__ S,
                                                         const uint32_t* __restrict__ parent, const uint32_t* __restrict__ isl_first,
                                                         const unsigned long long* __restrict__ isl_sum, ContactRec* __restrict__ out,
                                                         uint32_t* cursor) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    float2 n0 = make_float2(0.f, 0.f), n1 = n0;
    bool t0 = false, t1 = false;
    if (u < n_units) {
        n0 = LN[u]; n1 = LN[u + n_units];
        t0 = !(isnan(n0.x) && isnan(n0.y)); t1 = !(isnan(n1.x) && isnan(n1.y));
    }
    const uint32_t cnt = (t0 ? 1u : 0u) + (t1 ? 1u : 0u);
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (int(lane) >= o) incl += v; }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (total == 0u) return;
    uint32_t base = 0;
    if (lane == 31u) base = atomicAdd(cursor, total);
    base = __shfl_sync(0xFFFFFFFFu, base, 31) + incl - cnt;
    if (cnt == 0u) return;
    const float4 h = H[u], s = S[u];
    const int b0 = __float_as_int(h.x), b1 = __float_as_int(h.y);
    ContactRec r;
    r.c0 = __float_as_uint(s.x); r.c1 = __float_as_uint(s.y); r.pad = 0u;
    const uint32_t rep = parent[b0 >= 0 ? b0 : b1];
    r.root = isl_first[rep];          // IslandId::first_body: where the sleeping record of this island would sit
    r.sum = isl_sum[rep];
    if (t0) { r.nx = n0.x; r.ny = n0.y; out[base++] = r; }
    if (t1) { r.nx = n1.x; r.ny = n1.y; out[base] = r; }
}
__global__ void __launch_bounds__(256) k_contacts_export(uint32_t n, const ContactRec* __restrict__ list, uint32_t* __restrict__ out4) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ContactRec r = list[i];
    out4[4 * i + 0] = r.c0; out4[4 * i + 1] = r.c1; out4[4 * i + 2] = __float_as_uint(r.nx); out4[4 * i + 3] = __float_as_uint(r.ny);
}
__global__ void __launch_bounds__(256) k_isl_export(IslandCtx c, uint32_t cap, uint32_t* cursor, uint32_t* first_body, unsigned long long* sums,
                                                    uint32_t* counts, uint32_t* flags) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= c.n_bodies) return;
    if (!(__float_as_uint(c.MI[b].z) & BF_ALIVE) || c.parent[b] != b) return;
    const uint32_t pos = atomicAdd(cursor, 1u);
    if (pos >= cap) return;
    first_body[pos] = c.isl_first[b]; sums[pos] = c.isl_sum[b]; counts[pos] = c.isl_count[b]; flags[pos] = c.isl_flags[b];
}

}  // namespace sfg

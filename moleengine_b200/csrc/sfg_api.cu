// C ABI of the B200-native Starframe physics tick (include/starframe_gpu.h): world lifetime, slot-indexed
// uploads, sfg_tick orchestration (broadphase -> constraint graph -> colouring -> XPBD substeps), readback,
// batched casts and the parity hooks. There is NO CPU fallback: every entry point that computes needs a
// CUDA device and reports SFG_E_NO_DEVICE / SFG_E_CUDA otherwise.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/starframe_gpu.h"
#include "sfg_broadphase.cuh"
#include "sfg_cast.cuh"
#include "sfg_islands.cuh"
#include "sfg_solver.cuh"
#include "sfg_sort.cuh"

using namespace sfg;

namespace {

// Device buffers grow with 50 % headroom and NEVER call cudaFree while a world is in use: on this platform a cudaFree in the
// middle of a run was measured at 90-270 ms (it synchronises the device and unmaps), i.e. a visible hitch every time a pile
// grows past a buffer. The outgrown allocation goes to a graveyard that is emptied when the world is cleared or destroyed;
// the dead memory is bounded by the geometric growth (less than twice the live size).
struct Graveyard {
    std::vector<void*> dead;
    void add(void* q) { dead.push_back(q); }
    bool empty() const { return dead.empty(); }
    void bury_all() { for (void* q : dead) cudaFree(q); dead.clear(); }   // cudaFree synchronises the device first
};
// The graveyard belongs to the world whose buffers it holds (SfgWorld::graveyard): clearing or destroying one world frees only
// what that world retired, on its own device — never another world's buffers (a cudaFree is a device-wide synchronisation of
// 90-270 ms here, exactly the hitch the graveyard exists to avoid; batched worlds and several GPUs in one process would all pay
// it). A DBuf picks up its world's graveyard when it is constructed as a member of that world (sfg_world_create sets the
// thread-local pointer around `new SfgWorld`).
thread_local Graveyard* tl_ctor_yard = nullptr;
template <class T>
struct DBuf {
    T* p = nullptr;
    size_t cap = 0;
    Graveyard* yard = tl_ctor_yard;
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    ~DBuf() { release(); }              // sfg_world_destroy deletes the world on its own device: nothing a release list forgot can leak
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        static const bool dbg = getenv("SFG_DEBUG_ALLOC") != nullptr;     // debug: report every (re)allocation and what it cost
        const auto t0 = std::chrono::steady_clock::now();
        if (p) { if (yard) yard->add(p); else cudaFree(p); }
        p = nullptr; cap = 0;
        size_t want = n + n / 2 + 64;   // 50 % headroom: pair counts creep up as piles form
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess && yard && !yard->empty()) {     // out of memory: now the graveyard has to go
            cudaGetLastError();
            yard->bury_all();
            e = cudaMalloc(&p, want * sizeof(T));
        }
        if (e == cudaSuccess) cap = want;
        if (dbg) fprintf(stderr, "SFG_DEBUG_ALLOC: %zu bytes, cudaMalloc %.2f ms\n", want * sizeof(T),
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Pair-sized buffers are reserved for 6 candidate pairs per live collider from the first tick on (2-8 pairs per
// body; C3 peaks at 3.5, C5 at 7): a cudaMalloc in the middle of a run costs 100-700 ms on this platform when it has to map
// fresh memory, which showed up as a hitch every time a growing pile outgrew its buffers. ~330 B per reserved pair.
constexpr size_t RESERVE_PAIRS_PER_COLLIDER = 6;
constexpr uint32_t SLAB_ZONE_CAP = 1u << 18;   // owned bodies within the halo of one cut (262 144 x 4 B per side)

// per rope, uploaded before k_rope_build: where the rope's particles sit and where its units go in the colour-major lists
struct RopeBuildRec { uint32_t first, count, ru_off[3], rd_off[2], pad; };

struct TickHeader {            // small device block read back once per tick phase
    Bounds bounds;
    uint32_t n_pairs;
    uint32_t barrier;
    uint32_t n_entries;
    uint32_t pad[1];
};

}  // namespace

struct SfgWorld {
    Graveyard* graveyard = nullptr;          // this world's outgrown device buffers (freed by sfg_world_clear / sfg_world_destroy)
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;          // the island pass runs here, next to the colouring (both only read the pair list)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool islands_in_flight = false;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_cast[2] = {nullptr, nullptr};
    cudaEvent_t ev_fence = nullptr;          // sfg_stream_fence
    float last_cast_ms = 0.f;
    uint32_t* pin = nullptr;                 // 1 KB of page-locked host memory: where the tick's small read-backs land (a pageable
                                             // target makes cudaMemcpyAsync stage the copy through the driver)
    uint32_t n_k_two = 0;                    // two-body constraints (visited twice per substep)
    std::string err;
    SfgTuning tuning{};
    uint64_t mask[64];
    int n_sms = 148;
    bool ticked = false;
    float frame_dt = 1.0f / 60.0f;   // of the tick in flight

    // bodies
    uint32_t n_bodies = 0;
    DBuf<float4> P, Q, V, OV, MI;
    DBuf<float2> EA, PCD, LAT;
    DBuf<int> rope_next, rope_prev;
    // colliders
    uint32_t n_colls = 0;
    DBuf<float4> CS, CL, CM, AABB, CREC;   // CREC: (CS, CL, CM, CI) interleaved, 64 B per collider, for the kernels that gather whole colliders
    DBuf<int4> CI;
    // constraints (raw columns in host order)
    uint32_t n_constraints = 0;
    DBuf<int4> KI;
    DBuf<float4> KA, KB;
    // ropes
    uint32_t n_ropes = 0, n_particles = 0, n_rope_units = 0, n_rope_damp = 0;
    uint32_t rope_colour_start[4] = {0, 0, 0, 0}, damp_colour_start[3] = {0, 0, 0};
    DBuf<int4> RU, RD;
    DBuf<float4> RPA, RPB;
    DBuf<int> rp_body, rp_rope;      // particle array: body slot (-1 = hole left by an edit) and rope index of every particle
    std::vector<SfgRope> rope_tab;   // host table: one range of the particle array + parameters per rope
    bool ropes_dirty = false;        // the unit lists have to be rebuilt from (rope_tab, rp_body) before the next tick
    DBuf<RopeBuildRec> rope_build;
    // staging
    DBuf<unsigned char> stage, stage2;   // stage2: pose read-back on the second stream (sfg_tick_poses)
    DBuf<unsigned long long> d_mask;
    DBuf<TickHeader> hdr;
    DBuf<uint32_t> tick_counts;   // [0] body-body pair units of the tick (visited twice), [2] / [3] touching units / entries (SFG_TUNE_COUNT_ENTRIES)
    // broadphase scratch
    GridParams grid{};
    uint32_t n_live = 0, n_entries = 0;
    DBuf<uint32_t> cell_cnt, cell_off;
    DBuf<uint32_t> keys0, keys1, vals0, vals1, sort_tmp, cell_start, scan_tmp;
    DBuf<float4> SA;
    DBuf<int4> SI;
    // pair units
    uint32_t n_units = 0, n_unit_colours = 0, prev_pairs = 0, stats_jp_rounds = 0;
    bool bouncy_seen = false;      // a collider with restitution != 0 was uploaded
    std::vector<uint32_t> unit_colour_start;   // n_unit_colours + 1
    DBuf<uint2> U_ids;
    DBuf<int2> U_ub;
    DBuf<uint2> adj2;   // body -> (unit, priority)
    DBuf<uint32_t> U_prio, U_colour, deg, adj_start, cursor, jp_counters, perm0, perm1, ckey0, ckey1;
    DBuf<float4> H, S, G0, G1, L0, L1, M0, M1, M2;
    DBuf<float2> LN;
    DBuf<float> RC;
    DBuf<uint32_t> WL, wl_count, near_count, far_min, steal;
    uint32_t last_substeps = 0;   // substeps of the last tick: tag_base + last_substeps = tag of a manifold entry written in its last substep
    uint32_t tick_seq = 0;        // ticks begun so far; its low 12 bits go into the manifold tags
    float4* m0_cleared = nullptr; // the M0 allocation that was last cleared
    // constraint units
    uint32_t n_kunits = 0, n_k_colours = 0;
    std::vector<uint32_t> k_colour_start;
    DBuf<uint2> KU_ids;
    DBuf<int2> KU_ub;
    DBuf<uint32_t> KU_prio, KU_colour;
    DBuf<float4> K0, K1, K2;
    // persistent solver
    DBuf<Stage> d_stages;
    DBuf<unsigned long long> trace, body_word;
    DBuf<uint32_t> adj_body;          // per list slot: body | side << 31
    DBuf<uint32_t> adj_colour;        // per list slot of a crowded body: the colour of that unit
    DBuf<uint2> U_slot;               // per unit: its slots in crowded lists
    DBuf<uint8_t> U_cls;              // per pair unit: order class inside a colour (k_pair_unit_setup)
    DBuf<int4> adj_rank;              // per unit: (body 0, body 1, rank in body 0's list, rank in body 1's list)
    std::vector<Stage> h_stages;
    // casts
    DBuf<SfgRay> d_rays;
    DBuf<SfgCastHit> d_hits;
    // islands + sleeping (sfg_islands.cuh)
    DBuf<uint32_t> isl_parent, isl_first, isl_flags, isl_count, sl_ticks, sl_valid, isl_counters;
    DBuf<unsigned long long> isl_sum, sl_sum;
    DBuf<ContactRec> ct[2];       // contact list of the last tick (with retention for sleeping islands), double-buffered
    uint32_t ct_n = 0, ct_cur = 0;
    bool islands_valid = false;   // the last tick ran the island pass
    bool island_flags_live = false;   // BF_ASLEEP / BF_FOREIGN may be set on bodies (an island pass ran since they were last cleared)
    // tick in flight (sfg_tick_begin .. sfg_tick_end)
    bool in_tick = false, tick_sleeping = false;
    uint32_t tick_substeps = 0, tick_next_substep = 0;
    SolverCtx tick_ctx;
    // slab decomposition (sfg_slab_*)
    bool slab_on = false, slab_was_on = false;
    float slab_lo = 0.f, slab_hi = 0.f, slab_halo = 0.f;
    uint32_t slab_epoch = 0;
    DBuf<uint32_t> slab_stamp, dbg, zone_list, zone_count;
    bool zone_valid = false;      // this tick's classification filled the zone lists
    // island sharding (sfg_island_shard_configure)
    uint32_t shard_rank = 0, shard_n = 1;
    // stats
    SfgStats stats{};
    uint32_t launches = 0;
};

namespace {

int fail(SfgWorld* w, int code, const char* what, cudaError_t e = cudaSuccess) {
    if (w) {
        w->err = what;
        if (e != cudaSuccess) { w->err += ": "; w->err += cudaGetErrorString(e); }
    }
    return code;
}
#define CK(call)                                                        \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return fail(w, SFG_E_CUDA, #call, e__); \
    } while (0)
#define CKL(w_)                                                                    \
    do {                                                                           \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) return fail(w_, SFG_E_CUDA, "kernel launch", e__); \
    } while (0)

inline uint32_t nblk(uint32_t n, uint32_t t = 256) { return (n + t - 1) / t; }

// ---- AoS <-> SoA on the device ----
__global__ void k_unpack_bodies(const SfgBody* in, uint32_t n, uint32_t first, float4* P, float4* Q, float4* V, float4* MI) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SfgBody b = in[i];
    const uint32_t s = first + i;
    P[s] = make_float4(b.x, b.y, 0.f, 0.f);
    Q[s] = make_float4(b.rot_s, b.rot_b, b.rot_s, b.rot_b);
    V[s] = make_float4(b.vx, b.vy, b.w, 0.f);
    MI[s] = make_float4((b.flags & SFG_BODY_MASS_FINITE) ? b.inv_mass : 0.f, (b.flags & SFG_BODY_INERTIA_FINITE) ? b.inv_inertia : 0.f,
                        __uint_as_float((b.flags & 15u) | ((b.flags & SFG_BODY_ALIVE) ? BF_EXISTS : 0u)), 0.f);
}
__global__ void k_pack_bodies(SfgBody* out, uint32_t n, uint32_t first, const float4* P, const float4* Q, const float4* V, const float4* MI) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = first + i;
    const float4 p = P[s], q = Q[s], v = V[s], m = MI[s];
    SfgBody b;
    b.x = p.x + p.z; b.y = p.y + p.w; b.rot_s = q.x; b.rot_b = q.y;
    b.vx = v.x; b.vy = v.y; b.w = v.z;
    const uint32_t z = __float_as_uint(m.z);
    b.flags = (z & ((z & BF_FOREIGN) ? 14u : 30u)) | ((z & BF_EXISTS) ? SFG_BODY_ALIVE : 0u) | ((z & BF_GHOST) ? SFG_BODY_GHOST : 0u) |
              (((z & BF_EXISTS) && !(z & BF_ALIVE)) ? SFG_BODY_INACTIVE : 0u);
    b.inv_mass = m.x; b.inv_inertia = m.y;
    b.reserved[0] = b.reserved[1] = 0;
    out[i] = b;
}
__global__ void k_set_poses(const float4* in, uint32_t n, uint32_t first, float4* P, float4* Q) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = in[i];
    P[first + i] = make_float4(v.x, v.y, 0.f, 0.f);
    Q[first + i] = make_float4(v.z, v.w, v.z, v.w);
}
__global__ void k_set_poses_indexed(const float4* in, const uint32_t* slots, uint32_t n, float4* P, float4* Q) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = in[i];
    const uint32_t s = slots[i];
    P[s] = make_float4(v.x, v.y, 0.f, 0.f);
    Q[s] = make_float4(v.z, v.w, v.z, v.w);
}
__global__ void k_get_poses_indexed(float4* out, const uint32_t* slots, uint32_t n, const float4* P, const float4* Q) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slots[i];
    const float4 p = P[s], q = Q[s];
    out[i] = make_float4(p.x + p.z, p.y + p.w, q.x, q.y);
}
__global__ void k_get_poses(float4* out, uint32_t n, uint32_t first, const float4* P, const float4* Q) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = P[first + i], q = Q[first + i];
    out[i] = make_float4(p.x + p.z, p.y + p.w, q.x, q.y);
}
__global__ void k_unpack_colliders(const SfgCollider* in, uint32_t n, uint32_t first, float4* CS, float4* CL, float4* CM, int4* CI, float4* CREC) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SfgCollider c = in[i];
    const uint32_t s = first + i;
    const float4 cs = make_float4(c.p0, c.p1, c.circle_r, __uint_as_float(c.kind));
    const float4 cl = make_float4(c.x, c.y, c.rot_s, c.rot_b);
    const float4 cm = make_float4(c.static_friction, c.dynamic_friction, c.restitution, __uint_as_float(c.flags & 15u));
    const int4 ci = make_int4(c.body, int(c.layer), int(c.domain), int(c.flags & 15u));
    CS[s] = cs; CL[s] = cl; CM[s] = cm; CI[s] = ci;
    // the same four columns once more as ONE 64-byte record per collider: the per-pair kernels of the graph phase gather whole
    // colliders at random (two per pair), and four 16-byte columns cost four half-used 32-byte sectors where the record costs two full ones
    float4* r = CREC + size_t(s) * 4;
    r[0] = cs; r[1] = cl; r[2] = cm; r[3] = make_float4(__int_as_float(ci.x), __int_as_float(ci.y), __int_as_float(ci.z), __int_as_float(ci.w));
}
__global__ void k_unpack_constraints(const SfgConstraint* in, uint32_t n, int4* KI, float4* KA, float4* KB) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SfgConstraint c = in[i];
    KI[i] = make_int4(c.owner, c.target, int(c.flags), 0);
    KA[i] = make_float4(c.compliance, c.linear_damping, c.angular_damping, c.distance);
    KB[i] = make_float4(c.off0x, c.off0y, c.off1x, c.off1y);
}
__global__ void k_collect_contacts(uint32_t n_entries, uint32_t n_units, const float2* LN, const float4* S, SfgContactInfo* out, uint32_t cap,
                                   uint32_t* count) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    const float2 n = LN[e];
    if (isnan(n.x) && isnan(n.y)) return;
    const uint32_t pos = atomicAdd(count, 1u);
    if (pos >= cap) return;
    const float4 s = S[e % n_units];
    SfgContactInfo ci;
    ci.collider0 = __float_as_uint(s.x); ci.collider1 = __float_as_uint(s.y); ci.nx = n.x; ci.ny = n.y;
    out[pos] = ci;
}
__global__ void k_export_manifolds(uint32_t n_entries, uint32_t n_units, uint32_t tag, const float4* S, const float4* M0, const float4* M1,
                                   const float4* M2, float* out) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    const float4 s = S[e % n_units], m0 = M0[e];
    const uint32_t bits = __float_as_uint(m0.x);
    const int count = (bits >> 8) == tag ? int(bits & M0_COUNT_MASK) : 0;   // entries not written in the last substep hold no contact
    float* o = out + 14 * size_t(e);
    o[0] = float(__float_as_uint(s.x)); o[1] = float(__float_as_uint(s.y));
    o[2] = float(count); o[3] = count ? m0.y : 0.f; o[4] = count ? m0.z : 0.f; o[5] = count ? m0.w : 0.f;
    float4 a = count >= 1 ? M1[e] : make_float4(0, 0, 0, 0), b = count >= 2 ? M2[e] : make_float4(0, 0, 0, 0);
    o[6] = a.x; o[7] = a.y; o[8] = a.z; o[9] = a.w; o[10] = b.x; o[11] = b.y; o[12] = b.z; o[13] = b.w;
}
__global__ void k_export_units(uint32_t n_units, const float4* H, const float4* S, uint32_t* hi, uint32_t* lo_mult) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    const float4 s = S[u];
    hi[u] = __float_as_uint(s.x);
    lo_mult[u] = __float_as_uint(s.y) | ((__float_as_uint(H[u].z) & UF_MULT2) ? 0x80000000u : 0u);
}

// standalone parity kernels
__global__ void k_dbg_isect(uint32_t n, const float* poses, const float* shapes, float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Iso p[2]; Shape s[2];
    for (int k = 0; k < 2; ++k) {
        const float* pp = poses + 8 * size_t(i) + 4 * k; const float* ss = shapes + 8 * size_t(i) + 4 * k;
        p[k] = mkiso(mk(pp[0], pp[1]), mkrot(pp[2], pp[3]));
        s[k].kind = uint32_t(ss[0]); s[k].p0 = ss[1]; s[k].p1 = ss[2]; s[k].r = ss[3];
    }
    Manifold m;
    intersection_check(p[0], p[1], s[0], s[1], m);
    float* o = out + 14 * size_t(i);
    for (int k = 0; k < 14; ++k) o[k] = 0.f;
    o[0] = float(m.count);
    if (m.count) { o[2] = m.n.x; o[3] = m.n.y; }
    for (int k = 0; k < m.count; ++k) { o[4 + 4 * k] = m.off[k][0].x; o[5 + 4 * k] = m.off[k][0].y; o[6 + 4 * k] = m.off[k][1].x; o[7 + 4 * k] = m.off[k][1].y; }
}
__global__ void k_dbg_cast(uint32_t n, const SfgRay* rays, const float* poses, const float* shapes, float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* pp = poses + 4 * size_t(i); const float* ss = shapes + 4 * size_t(i);
    Shape s; s.kind = uint32_t(ss[0]); s.p0 = ss[1]; s.p1 = ss[2]; s.r = ss[3];
    const SfgRay r = rays[i];
    Hit h = spherecast_collider(mk(r.sx, r.sy), mk(r.dx, r.dy), r.radius, mkiso(mk(pp[0], pp[1]), mkrot(pp[2], pp[3])), s);
    float* o = out + 6 * size_t(i);
    o[0] = h.hit ? 1.f : 0.f; o[1] = h.t; o[2] = h.n.x; o[3] = h.n.y; o[4] = h.p.x; o[5] = h.p.y;
}
__global__ void k_dbg_geometry(uint32_t op, uint32_t n, const float* args, const float* shapes, float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* ss = shapes + 4 * size_t(i);
    Shape s; s.kind = uint32_t(ss[0]); s.p0 = ss[1]; s.p1 = ss[2]; s.r = ss[3];
    const V2 a = mk(args[2 * size_t(i)], args[2 * size_t(i) + 1]);
    float* o = out + 7 * size_t(i);
    for (int k = 0; k < 7; ++k) o[k] = 0.f;
    if (op == 0) { bool in; V2 c = closest_boundary_point(s, a, &in); o[0] = c.x; o[1] = c.y; o[2] = in ? 1.f : 0.f; }
    else if (op == 1) o[0] = projected_extent(s, a);
    else if (op == 2) { PEdge e = supporting_edge(s, a); o[0] = e.n.x; o[1] = e.n.y; o[2] = e.e.start.x; o[3] = e.e.start.y; o[4] = e.e.dir.x; o[5] = e.e.dir.y; o[6] = e.e.len; }
    else if (op == 3) o[0] = point_in_collider(a, mkiso(mk(0.f, 0.f), mkrot(1.f, 0.f)), s) ? 1.f : 0.f;
}

// ---- slab decomposition: one world cut into x-intervals, one per GPU; bodies near a cut are exchanged as 64-byte records
struct HaloRec { uint32_t slot, pad[3]; float4 P, Q, V; };
// zone_list[side * zone_cap + k] / zone_count[side]: the owned bodies within `halo` of the cut on that side, i.e. what every
// owner -> ghost exchange of this tick sends (k_slab_pack_list walks the list instead of all bodies)
__global__ void k_slab_classify(uint32_t nb, const float4* P, float4* MI, const uint32_t* stamp, uint32_t epoch, float lo, float hi, float halo, int on,
                                uint32_t* zone_list, uint32_t* zone_count, uint32_t zone_cap) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    float4 m = MI[b];
    uint32_t fl = __float_as_uint(m.z) & ~(BF_ALIVE | BF_GHOST | BF_ZONE_L | BF_ZONE_R);
    if (fl & BF_EXISTS) {
        if (!on) fl |= BF_ALIVE;
        else {
            const float4 p = P[b];
            const float x = p.x + p.z;
            const bool owned = x >= lo && x < hi;
            if (owned) {
                fl |= BF_ALIVE;
                if (x < lo + halo) { fl |= BF_ZONE_L; const uint32_t k = atomicAdd(zone_count, 1u); if (k < zone_cap) zone_list[k] = b; }
                if (x >= hi - halo) { fl |= BF_ZONE_R; const uint32_t k = atomicAdd(zone_count + 1, 1u); if (k < zone_cap) zone_list[zone_cap + k] = b; }
            }
            else if (stamp[b] == epoch) fl |= BF_ALIVE | BF_GHOST;     // a neighbour sent it in this tick's refresh
        }
    }
    m.z = __uint_as_float(fl);
    MI[b] = m;
}
// mode 1 of k_slab_pack over this tick's zone list of one side (a few thousand bodies instead of all of them)
__global__ void k_slab_pack_list(const uint32_t* __restrict__ list, const uint32_t* __restrict__ count, uint32_t list_cap, const float4* P, const float4* Q,
                                 const float4* V, HaloRec* out, uint32_t cap, uint32_t* cursor) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = min(*count, list_cap);
    if (i == 0) *cursor = *count;                  // the header carries the count (an overflow of either capacity shows up there)
    if (i >= n || i >= cap) return;
    const uint32_t b = list[i];
    HaloRec r;
    r.slot = b; r.pad[0] = r.pad[1] = r.pad[2] = 0u; r.P = P[b]; r.Q = Q[b]; r.V = V[b];
    out[1 + i] = r;
}
// mode 0: by position (the refresh that opens a tick, before the classification exists); mode 1: by this tick's flags;
// mode 2: the bodies of this rank's islands (island sharding)
__global__ void k_slab_pack(uint32_t nb, int mode, int side, float lo, float hi, float halo, const float4* P, const float4* Q, const float4* V,
                            const float4* MI, HaloRec* out, uint32_t cap, uint32_t* cursor) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t fl = __float_as_uint(MI[b].z);
    const float4 p = P[b];
    bool send;
    if (mode == 0) {
        const float x = p.x + p.z;
        send = (fl & BF_EXISTS) && x >= lo && x < hi && (side == 0 ? x < lo + halo : x >= hi - halo);
    } else if (mode == 1) send = (fl & BF_ALIVE) && !(fl & BF_GHOST) && (fl & (side == 0 ? BF_ZONE_L : BF_ZONE_R));
    else send = (fl & BF_ALIVE) && !(fl & BF_FOREIGN);     // mode 2, island sharding: every body this rank solved (or keeps asleep)
    if (!send) return;
    const uint32_t pos = atomicAdd(cursor, 1u);
    if (pos >= cap) return;
    HaloRec r;
    r.slot = b; r.pad[0] = r.pad[1] = r.pad[2] = 0u; r.P = p; r.Q = Q[b]; r.V = V[b];
    out[1 + pos] = r;
}
__global__ void k_slab_unpack(const HaloRec* in, uint32_t cap, uint32_t nb, float4* P, float4* Q, float4* V, uint32_t* stamp, uint32_t epoch) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = min(in[0].slot, cap);
    if (i >= n) return;
    const HaloRec r = in[1 + i];
    if (r.slot >= nb) return;
    P[r.slot] = r.P; Q[r.slot] = r.Q; V[r.slot] = r.V;
    stamp[r.slot] = epoch;
}

// ---- rope topology on the device (sfg_set_ropes' comment): one thread per slot of the particle array
__global__ void k_rope_build(uint32_t n_slots, const int* __restrict__ rp_body, const int* __restrict__ rp_rope, const RopeBuildRec* __restrict__ recs,
                             int4* __restrict__ RU, int4* __restrict__ RD, int* __restrict__ rope_next, int* __restrict__ rope_prev) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_slots) return;
    const int r = rp_rope[p];
    if (r < 0) return;                                   // a hole
    const RopeBuildRec rec = recs[r];
    const uint32_t i = p - rec.first;
    if (i + 1 >= rec.count) return;                      // the last particle starts no link
    const int cur = rp_body[p], nxt = rp_body[p + 1], aft = i + 2 < rec.count ? rp_body[p + 2] : -1;
    RU[rec.ru_off[i % 3] + i / 3] = make_int4(cur, nxt, aft, r);
    RD[rec.rd_off[i % 2] + i / 2] = make_int4(cur, nxt, r, 0);
    rope_next[cur] = nxt; rope_prev[nxt] = cur;
}
// rp_rope from the rope table: one warp per rope labels its range of the particle array (the array was set to -1 = hole before)
__global__ void k_rope_label(uint32_t n_ropes, const RopeBuildRec* __restrict__ recs, int* __restrict__ rp_rope) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (r >= n_ropes) return;
    const uint32_t first = recs[r].first, count = recs[r].count;
    for (uint32_t i = lane; i < count; i += 32u) rp_rope[first + i] = int(r);
}
__global__ void k_rope_move(uint32_t from, uint32_t to, uint32_t count, int* rp_body) {   // to >= from + count (moves go to the end of the array)
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    rp_body[to + i] = rp_body[from + i];
    rp_body[from + i] = -1;
}
// rope.rs:226-286: which particles lost their body? appends their positions in the particle array to `out` (unordered)
__global__ void k_rope_find_dead(uint32_t n_slots, const int* __restrict__ rp_body, const int* __restrict__ rp_rope, const float4* __restrict__ MI, uint32_t n_bodies,
                                 uint32_t* count, uint32_t* out, uint32_t cap) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_slots || rp_rope[p] < 0) return;
    const int b = rp_body[p];
    const bool dead = b < 0 || uint32_t(b) >= n_bodies || !(__float_as_uint(MI[b].z) & BF_EXISTS);
    if (!dead) return;
    const uint32_t k = atomicAdd(count, 1u);
    if (k < cap) out[k] = p;
}

// debug (SFG_DEBUG_NAN=1 with SFG_TUNE_SEPARATE_LAUNCHES): first body whose state is not finite after a stage
__global__ void k_find_nonfinite(uint32_t nb, const float4* P, const float4* Q, const float4* V, const float4* MI, uint32_t* first) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb || !(__float_as_uint(MI[b].z) & BF_ALIVE)) return;
    const float4 p = P[b], q = Q[b], v = V[b];
    const float s = p.x + p.y + p.z + p.w + q.x + q.y + v.x + v.y + v.z;
    if (!isfinite(s)) atomicMin(first, b);
}

// The per-tick island flags on the bodies (asleep / solved by another rank) are rewritten by k_isl_init whenever the island pass
// runs. A tick that skips the pass (sleeping switched off, a slab world, island sharding switched off) must not inherit them:
// a body left "asleep" would never be integrated or solved again, and nothing could wake it.
__global__ void k_clear_island_flags(uint32_t nb, float4* MI) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t fl = __float_as_uint(MI[b].z);
    if (fl & (BF_ASLEEP | BF_FOREIGN)) MI[b].z = __uint_as_float(fl & ~(BF_ASLEEP | BF_FOREIGN));
}

int ensure_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return SFG_E_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return SFG_E_CUDA;
    return SFG_OK;
}

// Colour a unit graph and return the sorted permutation + colour starts. Shared by pairs and constraints.
int colour_units(SfgWorld* w, uint32_t n_units, const uint2* ids, const int2* ub, const uint32_t* prio, DBuf<uint32_t>& colour,
                 uint32_t** perm_out, std::vector<uint32_t>& colour_start, uint32_t* n_colours_out, bool pair_units) {
    cudaStream_t st = w->stream;
    colour_start.clear();
    *n_colours_out = 0;
    *perm_out = nullptr;
    if (n_units == 0) { colour_start.push_back(0); return SFG_OK; }
    const uint32_t nb = w->n_bodies;
    const size_t nr = pair_units ? std::max<size_t>(n_units, size_t(w->n_live) * RESERVE_PAIRS_PER_COLLIDER) : n_units;   // reserved size of pair-sized buffers
    // adjacency (deg was accumulated by the unit-setup kernel) + bucket keys of the colouring order
    constexpr uint32_t HIST_CAP = 65536;
    CK(w->adj_start.ensure(nb + 2)); CK(w->cursor.ensure(nb + 1)); CK(w->adj2.ensure(size_t(nr) * 2 + 1));
    CK(w->scan_tmp.ensure(scan_tmp_words(nb + 1)));
    CK(colour.ensure(nr)); CK(w->jp_counters.ensure(16 + HIST_CAP));   // [0, 16) control words, then units per colour
    CK(w->ckey0.ensure(nr)); CK(w->ckey1.ensure(nr)); CK(w->perm0.ensure(nr)); CK(w->perm1.ensure(nr));
    CK(w->sort_tmp.ensure(sort_tmp_words(nr)));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_colour_dataflow, SFG_COLOUR_THREADS, 0));
    const uint32_t grid = std::max(1u, std::min<uint32_t>(uint32_t(per_sm * w->n_sms), nblk(n_units, SFG_COLOUR_THREADS)));
    // a bucket must span fewer units than there are co-resident threads (progress argument in k_colour_dataflow): 2^bits buckets,
    // the `near` bit is the top bit of the priority and the rest is a hash, so a bucket holds at most n / 2^(bits - 1) units (all
    // units near, or all far) give or take Poisson noise; that worst case is kept under HALF of the resident threads. (It was an
    // eighth: the settled C3 pile then needs 16 bits = two radix passes over 4 M keys instead of one, 80 us of a 1.4 ms graph
    // phase, and the colouring itself is no slower with the wider buckets; measured in a same-box A/B, C5 -155 us.)
    // SFG_BUCKET_MARGIN (debug) overrides the factor.
    int bucket_bits = 8;
    static const int margin = [] { const char* e = getenv("SFG_BUCKET_MARGIN"); const int v = e ? atoi(e) : 2; return v < 1 ? 1 : v; }();
    while (bucket_bits < 24 && (size_t(n_units) >> (bucket_bits - 1)) * size_t(margin) > size_t(grid) * SFG_COLOUR_THREADS) bucket_bits += 8;
    w->launches += scan_exclusive(w->deg.p, w->adj_start.p, nb + 1, w->scan_tmp.p, nullptr, st);
    CK(cudaMemsetAsync(w->cursor.p, 0, size_t(nb + 1) * 4, st));
    CK(w->adj_body.ensure(size_t(nr) * 2 + 1)); CK(w->adj_rank.ensure(size_t(nr) + 1));
    CK(w->adj_colour.ensure(size_t(nr) * 2 + 1)); CK(w->U_slot.ensure(size_t(nr) + 1));
    CK(cudaMemsetAsync(w->adj_colour.p, 0xFF, (size_t(n_units) * 2 + 1) * 4, st));
    k_adj_fill<<<nblk(n_units), 256, 0, st>>>(n_units, ub, prio, w->adj_start.p, w->cursor.p, w->adj2.p, w->adj_body.p, w->adj_rank.p, 32 - bucket_bits, w->ckey0.p, w->perm0.p);
    CKL(w); w->launches++;
    const uint32_t* order;
    {
        int launches = 0;
        const int where = radix_sort_pairs(w->ckey0.p, w->perm0.p, w->ckey1.p, w->perm1.p, n_units, bucket_bits, w->sort_tmp.p, st, &launches);
        w->launches += launches;
        order = where ? w->perm1.p : w->perm0.p;
    }
    CK(cudaMemsetAsync(colour.p, 0xFF, size_t(n_units) * 4, st));
    CK(cudaMemsetAsync(w->jp_counters.p, 0, (16 + HIST_CAP) * 4, st));
    CK(w->body_word.ensure(nb + 1));
    CK(cudaMemsetAsync(w->body_word.p, 0, size_t(nb + 1) * 8, st));
    // list slots in use = sum of degrees <= 2 n_units; unused tail slots of the allocation are never read
    k_adj_rank<<<nblk(n_units * 2), 256, 0, st>>>(n_units * 2, w->adj2.p, w->adj_body.p, ids, w->adj_start.p, w->adj_rank.p, w->U_slot.p, w->jp_counters.p, nb);
    CKL(w); w->launches++;
    {
        uint32_t hist_cap = HIST_CAP;
        const uint32_t* adj_start = w->adj_start.p; const uint2* adj = w->adj2.p;
        uint32_t* col = colour.p; uint32_t* ctrl = w->jp_counters.p; uint32_t* hist = w->jp_counters.p + 16;
        unsigned long long* words = w->body_word.p;
        const int4* urec = w->adj_rank.p; const uint2* uslot = w->U_slot.p; uint32_t* adj_colour = w->adj_colour.p;
        void* args[] = {&n_units, (void*)&order, (void*)&ub, (void*)&prio, (void*)&ids, (void*)&adj_start, (void*)&adj, (void*)&urec, (void*)&uslot, &adj_colour, &col, &words, &ctrl, &hist, &hist_cap};
        static const bool dbg_slow = getenv("SFG_DEBUG_SLOW") != nullptr;   // debug: wall time of the colouring kernel alone
        std::chrono::steady_clock::time_point t0;
        if (dbg_slow) { CK(cudaStreamSynchronize(st)); t0 = std::chrono::steady_clock::now(); }
        CK(cudaLaunchCooperativeKernel((void*)k_colour_dataflow, dim3(grid), dim3(SFG_COLOUR_THREADS), args, 0, st));
        w->launches++;
        if (dbg_slow) {
            CK(cudaStreamSynchronize(st));
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (ms > 20.0) fprintf(stderr, "SFG_DEBUG_SLOW: k_colour_dataflow %.1f ms (%u units, grid %u)\n", ms, n_units, grid);
        }
    }
    // one read-back for the control words and the first 64 colours (a second one only if there are more)
    uint32_t* head = w->pin;                       // 80 words
    CK(cudaMemcpyAsync(head, w->jp_counters.p, (16 + 64) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t n_colours = head[1];
    w->stats_jp_rounds = head[4];
    if (n_colours == 0 || n_colours > HIST_CAP) return fail(w, SFG_E_CAPACITY, "colouring did not converge or needs > 65536 colours");
    if (head[5]) return fail(w, SFG_E_CAPACITY, "a body takes part in 2^24 or more pairs / constraints");
    if (head[6]) return fail(w, SFG_E_STATE, "colouring watchdog: a unit waited forever for its bodies' tickets");
    std::vector<uint32_t> hist(n_colours);
    memcpy(hist.data(), head + 16, std::min<size_t>(n_colours, 64) * 4);
    if (n_colours > 64) {
        CK(cudaMemcpyAsync(hist.data() + 64, w->jp_counters.p + 16 + 64, size_t(n_colours - 64) * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    colour_start.resize(n_colours + 1);
    uint32_t run = 0;
    for (uint32_t c = 0; c < n_colours; ++c) { colour_start[c] = run; run += hist[c]; }
    colour_start[n_colours] = run;
    if (run != n_units) return fail(w, SFG_E_CAPACITY, "colouring left units uncoloured");
    // stable sort by colour: perm = unit indices in colour-major order
    uint32_t* near_count = nullptr;
    if (pair_units) {
        CK(w->near_count.ensure(n_colours));
        CK(cudaMemsetAsync(w->near_count.p, 0, size_t(n_colours) * 4, st));
        near_count = w->near_count.p;
    }
    k_unit_sort_keys<<<nblk(n_units), 256, 0, st>>>(n_units, colour.p, pair_units ? w->U_cls.p : nullptr, w->ckey0.p, w->perm0.p, near_count);
    CKL(w); w->launches++;
    int bits = 7;
    while ((1u << bits) < n_colours * 128u) ++bits;
    int launches = 0;
    int where = radix_sort_pairs(w->ckey0.p, w->perm0.p, w->ckey1.p, w->perm1.p, n_units, bits, w->sort_tmp.p, st, &launches);
    w->launches += launches;
    *perm_out = where ? w->perm1.p : w->perm0.p;
    *n_colours_out = n_colours;
    return SFG_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

int sfg_abi_version(void) { return SFG_ABI_VERSION; }

void sfg_default_tuning(SfgTuning* t) {   // physics.rs:338-349
    memset(t, 0, sizeof(*t));
    t->substeps = 10; t->sleep_vel_threshold = 0.001f; t->fall_asleep_frames = 10; t->max_expected_acceleration = 10.0f;
}
void sfg_default_mask_matrix(uint64_t out[64]) {   // collision.rs:132-140
    for (int i = 0; i < 64; ++i) out[i] = ~0ull;
    out[63] &= ~(1ull << 63);
}

int sfg_world_create(const SfgTuning* tuning, const uint64_t mask_matrix[64], int device, SfgWorld** out) {
    if (!out) return SFG_E_INVALID;
    *out = nullptr;
    if (tuning && tuning->substeps == 0) return SFG_E_INVALID;      // same rule as sfg_set_tuning: dt = frame_dt / substeps
    int rc = ensure_device(device);
    if (rc != SFG_OK) return rc;
    Graveyard* yard = new Graveyard();
    tl_ctor_yard = yard;                     // every DBuf member of the new world retires its outgrown allocations here
    SfgWorld* w = new SfgWorld();
    tl_ctor_yard = nullptr;
    w->graveyard = yard;
    w->device = device;
    if (tuning) w->tuning = *tuning; else sfg_default_tuning(&w->tuning);
    if (mask_matrix) memcpy(w->mask, mask_matrix, sizeof(w->mask)); else sfg_default_mask_matrix(w->mask);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) w->n_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess) { delete w; delete yard; return SFG_E_CUDA; }
    if (cudaStreamCreateWithFlags(&w->stream2, cudaStreamNonBlocking) != cudaSuccess) { cudaStreamDestroy(w->stream); delete w; delete yard; return SFG_E_CUDA; }
    cudaEventCreateWithFlags(&w->ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&w->ev_join, cudaEventDisableTiming);
    for (auto& e : w->ev) cudaEventCreate(&e);
    for (auto& e : w->ev_cast) cudaEventCreate(&e);
    if (w->d_mask.ensure(64) != cudaSuccess || w->hdr.ensure(1) != cudaSuccess || w->tick_counts.ensure(4) != cudaSuccess) { delete w; delete yard; return SFG_E_CUDA; }
    if (cudaHostAlloc((void**)&w->pin, 1024, cudaHostAllocDefault) != cudaSuccess) { w->pin = nullptr; delete w; delete yard; return SFG_E_CUDA; }
    cudaMemcpyAsync(w->d_mask.p, w->mask, sizeof(w->mask), cudaMemcpyHostToDevice, w->stream);
    cudaStreamSynchronize(w->stream);
    *out = w;
    return SFG_OK;
}

void sfg_world_destroy(SfgWorld* w) {
    if (!w) return;
    cudaSetDevice(w->device);
    cudaStreamSynchronize(w->stream);
    if (w->stream2) cudaStreamSynchronize(w->stream2);
    w->graveyard->bury_all();
    DBuf<float4>* f4[] = {&w->P, &w->Q, &w->V, &w->OV, &w->MI, &w->CS, &w->CL, &w->CM, &w->AABB, &w->CREC, &w->KA, &w->KB, &w->RPA, &w->RPB, &w->SA,
                          &w->H, &w->S, &w->G0, &w->G1, &w->L0, &w->L1, &w->M0, &w->M1, &w->M2, &w->K0, &w->K1, &w->K2};
    for (auto* b : f4) b->release();
    DBuf<float2>* f2[] = {&w->EA, &w->PCD, &w->LAT, &w->LN};
    for (auto* b : f2) b->release();
    DBuf<uint32_t>* u1[] = {&w->keys0, &w->keys1, &w->vals0, &w->vals1, &w->sort_tmp, &w->cell_start, &w->scan_tmp, &w->U_prio,
                            &w->U_colour, &w->deg, &w->adj_start, &w->cursor, &w->jp_counters, &w->perm0, &w->perm1,
                            &w->ckey0, &w->ckey1, &w->KU_prio, &w->KU_colour, &w->cell_cnt, &w->cell_off, &w->isl_parent, &w->isl_first, &w->isl_flags, &w->isl_count, &w->sl_ticks, &w->sl_valid,
                            &w->isl_counters};
    for (auto* b : u1) b->release();
    DBuf<int>* i1[] = {&w->rope_next, &w->rope_prev, &w->rp_body, &w->rp_rope};
    for (auto* b : i1) b->release();
    DBuf<int4>* i4[] = {&w->CI, &w->KI, &w->RU, &w->RD, &w->SI};
    for (auto* b : i4) b->release();
    w->U_ids.release(); w->slab_stamp.release(); w->isl_sum.release(); w->sl_sum.release(); w->ct[0].release(); w->ct[1].release(); w->adj2.release(); w->KU_ids.release(); w->RC.release(); w->WL.release(); w->wl_count.release(); w->near_count.release(); w->far_min.release(); w->body_word.release(); w->trace.release(); w->adj_body.release(); w->adj_rank.release(); w->U_cls.release(); w->adj_colour.release(); w->U_slot.release(); w->U_ub.release(); w->KU_ub.release();
    w->stage.release(); w->stage2.release(); w->d_mask.release(); w->hdr.release(); w->d_stages.release(); w->d_rays.release(); w->d_hits.release();
    for (auto& e : w->ev) if (e) cudaEventDestroy(e);
    for (auto& e : w->ev_cast) if (e) cudaEventDestroy(e);
    if (w->ev_fence) cudaEventDestroy(w->ev_fence);
    if (w->ev_fork) cudaEventDestroy(w->ev_fork);
    if (w->ev_join) cudaEventDestroy(w->ev_join);
    if (w->stream2) cudaStreamDestroy(w->stream2);
    if (w->pin) cudaFreeHost(w->pin);
    if (w->stream) cudaStreamDestroy(w->stream);
    Graveyard* yard = w->graveyard;
    delete w;
    delete yard;
}

int sfg_world_clear(SfgWorld* w) {   // physics.rs:386-393
    if (!w) return SFG_E_INVALID;
    cudaSetDevice(w->device);
    cudaStreamSynchronize(w->stream);
    if (w->stream2) cudaStreamSynchronize(w->stream2);
    w->graveyard->bury_all();            // outgrown buffers of the run that ends here (this world's only)
    w->n_bodies = w->n_colls = w->n_constraints = w->n_ropes = w->n_particles = w->n_rope_units = w->n_rope_damp = 0;
    w->rope_tab.clear(); w->ropes_dirty = false;
    w->n_units = w->n_kunits = 0;
    w->ct_n = 0; w->islands_valid = false;
    w->ticked = false;
    memset(&w->stats, 0, sizeof(w->stats));
    return SFG_OK;
}
const char* sfg_last_error(const SfgWorld* w) { return w ? w->err.c_str() : "null world"; }

int sfg_set_tuning(SfgWorld* w, const SfgTuning* t) {
    if (!w || !t) return SFG_E_INVALID;
    if (t->substeps == 0) return fail(w, SFG_E_INVALID, "substeps must be > 0");
    if (t->substeps > SFG_MAX_SUBSTEPS) return fail(w, SFG_E_INVALID, "substeps exceed SFG_MAX_SUBSTEPS");
    w->tuning = *t;
    return SFG_OK;
}
int sfg_set_mask_matrix(SfgWorld* w, const uint64_t m[64]) {
    if (!w || !m) return SFG_E_INVALID;
    cudaSetDevice(w->device);
    memcpy(w->mask, m, sizeof(w->mask));
    CK(cudaMemcpyAsync(w->d_mask.p, w->mask, sizeof(w->mask), cudaMemcpyHostToDevice, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    return SFG_OK;
}

static int upload_bodies(SfgWorld* w, uint32_t first, uint32_t count, const SfgBody* b) {
    if (count == 0) return SFG_OK;
    cudaStream_t st = w->stream;
    CK(w->stage.ensure(size_t(count) * sizeof(SfgBody)));
    CK(cudaMemcpyAsync(w->stage.p, b, size_t(count) * sizeof(SfgBody), cudaMemcpyHostToDevice, st));
    k_unpack_bodies<<<nblk(count), 256, 0, st>>>((const SfgBody*)w->stage.p, count, first, w->P.p, w->Q.p, w->V.p, w->MI.p);
    CKL(w);
    CK(cudaStreamSynchronize(st));
    return SFG_OK;
}
int sfg_set_bodies(SfgWorld* w, uint32_t n, const SfgBody* b) {
    if (!w || (n && !b)) return SFG_E_INVALID;
    cudaSetDevice(w->device);
    CK(w->P.ensure(n)); CK(w->Q.ensure(n)); CK(w->V.ensure(n)); CK(w->OV.ensure(n)); CK(w->MI.ensure(n)); CK(w->EA.ensure(n));
    // a SHRINKING slot space may leave ropes / constraints pointing at slots that no longer exist: they have to be uploaded again.
    // A growing one (the host spawned bodies, e.g. Rope::extend_line) keeps them; the per-body link tables are rebuilt at the new size
    if (n < w->n_bodies) { w->n_ropes = w->n_particles = w->n_rope_units = w->n_rope_damp = 0; w->n_constraints = 0; w->rope_tab.clear(); w->ropes_dirty = false; }
    else if (n > w->n_bodies && !w->rope_tab.empty()) w->ropes_dirty = true;
    if (n < w->n_bodies) { w->n_colls = 0; w->ticked = false; }   // colliders already uploaded may point at slots that no longer exist: they have to be re-uploaded too
    // a new slot space starts a new world as far as sleeping is concerned: no island is being watched, no contact retained.
    // (Re-uploading the same slots every tick, as the PhysicsWorld mirror does, keeps the records: physics.rs:711-741 only
    // looks at velocities and at the island's identity.)
    if (n != w->n_bodies || !w->sl_valid.p) {
        CK(w->sl_valid.ensure(n + 1)); CK(w->sl_ticks.ensure(n + 1)); CK(w->sl_sum.ensure(n + 1));
        CK(cudaMemsetAsync(w->sl_valid.p, 0, size_t(n + 1) * 4, w->stream));
        w->ct_n = 0; w->islands_valid = false;
    }
    w->n_bodies = n;
    return upload_bodies(w, 0, n, b);
}
int sfg_update_bodies(SfgWorld* w, uint32_t first, uint32_t count, const SfgBody* b) {
    if (!w || (count && !b)) return SFG_E_INVALID;
    if (uint64_t(first) + count > w->n_bodies) return fail(w, SFG_E_INVALID, "body slot range out of bounds");
    cudaSetDevice(w->device);
    return upload_bodies(w, first, count, b);
}
static int upload_colliders(SfgWorld* w, uint32_t first, uint32_t count, const SfgCollider* c) {
    if (count == 0) return SFG_OK;
    for (uint32_t i = 0; i < count; ++i) {
        if (!(c[i].flags & SFG_COLL_ALIVE)) continue;
        if (c[i].kind > SFG_POLY_HEXAGON) return fail(w, SFG_E_INVALID, "unknown collider polygon kind");
        if (c[i].layer >= 64) return fail(w, SFG_E_INVALID, "collider layer must be < 64");
        if (c[i].body >= 0 && uint32_t(c[i].body) >= w->n_bodies) return fail(w, SFG_E_INVALID, "collider body slot out of range (set bodies first)");
        if (c[i].restitution != 0.0f) w->bouncy_seen = true;     // sticky until the next full upload: old velocities are kept only for worlds that bounce
    }
    cudaStream_t st = w->stream;
    CK(w->stage.ensure(size_t(count) * sizeof(SfgCollider)));
    CK(cudaMemcpyAsync(w->stage.p, c, size_t(count) * sizeof(SfgCollider), cudaMemcpyHostToDevice, st));
    k_unpack_colliders<<<nblk(count), 256, 0, st>>>((const SfgCollider*)w->stage.p, count, first, w->CS.p, w->CL.p, w->CM.p, w->CI.p, w->CREC.p);
    CKL(w);
    CK(cudaStreamSynchronize(st));
    return SFG_OK;
}
int sfg_set_colliders(SfgWorld* w, uint32_t n, const SfgCollider* c) {
    if (!w || (n && !c)) return SFG_E_INVALID;
    cudaSetDevice(w->device);
    CK(w->CS.ensure(n)); CK(w->CL.ensure(n)); CK(w->CM.ensure(n)); CK(w->CI.ensure(n)); CK(w->AABB.ensure(n)); CK(w->CREC.ensure(size_t(n) * 4));
    w->n_colls = n;
    w->ticked = false;
    w->bouncy_seen = false;
    return upload_colliders(w, 0, n, c);
}
int sfg_update_colliders(SfgWorld* w, uint32_t first, uint32_t count, const SfgCollider* c) {
    if (!w || (count && !c)) return SFG_E_INVALID;
    if (uint64_t(first) + count > w->n_colls) return fail(w, SFG_E_INVALID, "collider slot range out of bounds");
    cudaSetDevice(w->device);
    return upload_colliders(w, first, count, c);
}
int sfg_set_constraints(SfgWorld* w, uint32_t n, const SfgConstraint* c) {
    if (!w || (n && !c)) return SFG_E_INVALID;
    cudaSetDevice(w->device);
    for (uint32_t i = 0; i < n; ++i) {
        if (c[i].owner < 0 || uint32_t(c[i].owner) >= w->n_bodies) return fail(w, SFG_E_INVALID, "constraint owner slot out of range");
        if (c[i].target >= 0 && uint32_t(c[i].target) >= w->n_bodies) return fail(w, SFG_E_INVALID, "constraint target slot out of range");
        if ((c[i].flags & SFG_LIMIT_MASK) > SFG_LIMIT_GT) return fail(w, SFG_E_INVALID, "unknown constraint limit");
    }
    w->n_constraints = n;
    w->n_k_two = 0;
    for (uint32_t i = 0; i < n; ++i) if (c[i].target >= 0) w->n_k_two++;
    if (n == 0) return SFG_OK;
    cudaStream_t st = w->stream;
    CK(w->KI.ensure(n)); CK(w->KA.ensure(n)); CK(w->KB.ensure(n));
    CK(w->stage.ensure(size_t(n) * sizeof(SfgConstraint)));
    CK(cudaMemcpyAsync(w->stage.p, c, size_t(n) * sizeof(SfgConstraint), cudaMemcpyHostToDevice, st));
    k_unpack_constraints<<<nblk(n), 256, 0, st>>>((const SfgConstraint*)w->stage.p, n, w->KI.p, w->KA.p, w->KB.p);
    CKL(w);
    CK(cudaStreamSynchronize(st));
    return SFG_OK;
}

// rp_rope (particle -> rope, -1 = hole) is derived data: rebuilt from the host table after every edit, in one small upload and
// two launches whatever the number of ropes
static int relabel_rope_particles(SfgWorld* w) {
    cudaStream_t st = w->stream;
    const uint32_t nr = uint32_t(w->rope_tab.size());
    w->n_ropes = nr;
    w->ropes_dirty = true;
    if (w->n_particles == 0) return SFG_OK;
    CK(w->rp_rope.ensure(w->n_particles));
    CK(cudaMemsetAsync(w->rp_rope.p, 0xFF, size_t(w->n_particles) * 4, st));
    if (nr) {
        std::vector<RopeBuildRec> recs(nr);
        for (uint32_t r = 0; r < nr; ++r) { memset(&recs[r], 0, sizeof(RopeBuildRec)); recs[r].first = w->rope_tab[r].first_particle; recs[r].count = w->rope_tab[r].particle_count; }
        CK(w->rope_build.ensure(nr));
        CK(cudaMemcpyAsync(w->rope_build.p, recs.data(), size_t(nr) * sizeof(RopeBuildRec), cudaMemcpyHostToDevice, st));
        k_rope_label<<<nblk(nr * 32u), 256, 0, st>>>(nr, w->rope_build.p, w->rp_rope.p);
        CKL(w);
    }
    CK(cudaStreamSynchronize(st));        // recs is a stack-lifetime host buffer
    return SFG_OK;
}

// Rope topology lives in two places: a small table of ranges on the host (one entry per rope: where its particles sit in the
// device's particle array, and its parameters) and the particle array itself on the device (rp_body: particle -> body slot, -1 =
// hole). Everything the solver reads — step units (curr, next, after) coloured i mod 3 (solver.rs:118-179 touches particles i,
// i+1, i+2), damping pairs coloured i mod 2 (solver.rs:722-740), the next / prev links of the contact stage — is BUILT ON THE
// DEVICE from those two (k_rope_build), lazily before the next tick. So a topology edit (cut, extend, dead-particle split:
// sfg_rope_*) only changes the host table and a few device words; the particle array is never re-uploaded.
int sfg_set_ropes(SfgWorld* w, uint32_t n_ropes, const SfgRope* ropes, uint32_t n_particles, const int32_t* particle_bodies) {
    if (!w || (n_ropes && (!ropes || !particle_bodies))) return SFG_E_INVALID;
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    std::vector<int> rp_rope(n_particles, -1);
    std::vector<int> rp_body(n_particles, -1);
    for (uint32_t r = 0; r < n_ropes; ++r) {
        const SfgRope& rp = ropes[r];
        if (rp.particle_count < 2) return fail(w, SFG_E_INVALID, "Rope only had one particle");   // solver.rs:120 panics
        if (uint64_t(rp.first_particle) + rp.particle_count > n_particles) return fail(w, SFG_E_INVALID, "rope particle range out of bounds");
        const int32_t* p = particle_bodies + rp.first_particle;
        for (uint32_t i = 0; i < rp.particle_count; ++i) {
            if (p[i] < 0 || uint32_t(p[i]) >= w->n_bodies) return fail(w, SFG_E_INVALID, "rope particle body slot out of range");
            if (rp_rope[rp.first_particle + i] >= 0) return fail(w, SFG_E_INVALID, "rope particle ranges overlap");
            rp_rope[rp.first_particle + i] = int(r);
            rp_body[rp.first_particle + i] = p[i];
        }
    }
    w->rope_tab.assign(ropes, ropes + n_ropes);
    w->n_ropes = n_ropes; w->n_particles = n_particles;
    w->n_rope_units = w->n_rope_damp = 0;
    if (n_particles) {
        CK(w->rp_body.ensure(n_particles));
        CK(cudaMemcpyAsync(w->rp_body.p, rp_body.data(), size_t(n_particles) * 4, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
    }
    return relabel_rope_particles(w);
}

// Rope::cut_after (rope.rs:147-165), by position: the rope keeps particles [0, particle_index], the rest becomes a NEW rope with
// the same parameters, appended to the rope table (*new_rope_out = its index). Cutting after the last particle removes that
// particle from the rope (`self.particles.pop()`) and creates nothing (*new_rope_out = UINT32_MAX). No particle data moves.
int sfg_rope_cut_after(SfgWorld* w, uint32_t rope, uint32_t particle_index, uint32_t* new_rope_out) {
    if (!w) return SFG_E_INVALID;
    if (new_rope_out) *new_rope_out = 0xFFFFFFFFu;
    if (rope >= w->rope_tab.size()) return fail(w, SFG_E_INVALID, "rope index out of range");
    SfgRope& r = w->rope_tab[rope];
    if (particle_index >= r.particle_count) return fail(w, SFG_E_INVALID, "particle index beyond the rope");
    cudaSetDevice(w->device);
    if (particle_index + 1 == r.particle_count) {
        r.particle_count -= 1;                // the popped particle leaves the rope (its body lives on): its slot becomes a hole
    } else {
        SfgRope tail = r;
        tail.first_particle = r.first_particle + particle_index + 1;
        tail.particle_count = r.particle_count - particle_index - 1;
        r.particle_count = particle_index + 1;
        const uint32_t idx = uint32_t(w->rope_tab.size());
        w->rope_tab.push_back(tail);          // (invalidates r)
        if (new_rope_out) *new_rope_out = idx;
    }
    return relabel_rope_particles(w);
}

// Rope::extend_line (rope.rs:88-110): the host spawned `n_new` particle bodies (Rope::build_line) and uploaded them with the
// bodies / colliders; here their slots are appended to the rope. A rope that is not the last range of the particle array is
// moved to its end first (a device copy), so only the n_new slots cross the bus.
int sfg_rope_extend(SfgWorld* w, uint32_t rope, uint32_t n_new, const int32_t* particle_bodies) {
    if (!w || (n_new && !particle_bodies)) return SFG_E_INVALID;
    if (rope >= w->rope_tab.size()) return fail(w, SFG_E_INVALID, "rope index out of range");
    if (n_new == 0) return SFG_OK;
    for (uint32_t i = 0; i < n_new; ++i)
        if (particle_bodies[i] < 0 || uint32_t(particle_bodies[i]) >= w->n_bodies) return fail(w, SFG_E_INVALID, "rope particle body slot out of range (set bodies first)");
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    SfgRope& r = w->rope_tab[rope];
    const bool at_end = r.first_particle + r.particle_count == w->n_particles;
    const uint32_t new_len = at_end ? w->n_particles + n_new : w->n_particles + r.particle_count + n_new;
    if (new_len > w->rp_body.cap) {          // grow both columns, keeping their contents
        DBuf<int> nb;                       // (rp_rope is derived data: relabel_rope_particles re-creates it at the new size)
        nb.yard = w->graveyard;
        CK(nb.ensure(size_t(new_len) * 2));
        if (w->n_particles) CK(cudaMemcpyAsync(nb.p, w->rp_body.p, size_t(w->n_particles) * 4, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        std::swap(w->rp_body.p, nb.p); std::swap(w->rp_body.cap, nb.cap);
        if (nb.p) { w->graveyard->add(nb.p); nb.p = nullptr; nb.cap = 0; }      // never cudaFree in the middle of a run (DBuf's comment)
    }
    if (!at_end) {
        k_rope_move<<<nblk(r.particle_count), 256, 0, st>>>(r.first_particle, w->n_particles, r.particle_count, w->rp_body.p);
        CKL(w);
        r.first_particle = w->n_particles;
    }
    CK(cudaMemcpyAsync(w->rp_body.p + r.first_particle + r.particle_count, particle_bodies, size_t(n_new) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    r.particle_count += n_new;
    w->n_particles = new_len;
    return relabel_rope_particles(w);
}

// RopeSet::remove_dead_particles (rope.rs:226-286) against the bodies the device holds: a particle whose body slot is no longer
// alive is taken out of its rope and the rope is cut there; runs of dead particles are skipped; the first surviving piece keeps
// the rope's index, later pieces are appended as new ropes in the order the reference creates them; ropes left with no particle
// are deleted (the indices above them shift down, as thunderdome's iteration order does not matter to the solver). The device
// reports only the positions of the dead particles — a few words — and nothing is re-uploaded.
int sfg_rope_remove_dead_particles(SfgWorld* w, uint32_t* n_ropes_out) {
    if (!w) return SFG_E_INVALID;
    if (n_ropes_out) *n_ropes_out = uint32_t(w->rope_tab.size());
    if (w->rope_tab.empty() || w->n_particles == 0) return SFG_OK;
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    const uint32_t cap = 65536;
    CK(w->stage.ensure(size_t(cap + 1) * 4));
    uint32_t* cnt = (uint32_t*)w->stage.p;
    CK(cudaMemsetAsync(cnt, 0, 4, st));
    k_rope_find_dead<<<nblk(w->n_particles), 256, 0, st>>>(w->n_particles, w->rp_body.p, w->rp_rope.p, w->MI.p, w->n_bodies, cnt, cnt + 1, cap);
    CKL(w);
    uint32_t n_dead = 0;
    CK(cudaMemcpyAsync(&n_dead, cnt, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (n_dead == 0) return SFG_OK;
    if (n_dead > cap) return fail(w, SFG_E_CAPACITY, "more than 65536 rope particles died at once: upload the ropes again (sfg_set_ropes)");
    std::vector<uint32_t> dead(n_dead);
    CK(cudaMemcpyAsync(dead.data(), cnt + 1, size_t(n_dead) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::sort(dead.begin(), dead.end());
    std::vector<SfgRope> fresh;
    const size_t n_old = w->rope_tab.size();
    for (size_t ri = 0; ri < n_old; ++ri) {
        SfgRope& r = w->rope_tab[ri];
        const uint32_t lo = r.first_particle, hi = r.first_particle + r.particle_count;
        auto d0 = std::lower_bound(dead.begin(), dead.end(), lo), d1 = std::lower_bound(dead.begin(), dead.end(), hi);
        if (d0 == d1) continue;
        // live runs of [lo, hi) between the dead positions: the first one (possibly empty) stays in this rope, every later
        // non-empty one becomes a new rope (rope.rs:241-268)
        uint32_t run_start = lo;
        bool first = true;
        auto piece = [&](uint32_t a, uint32_t b) {
            if (first) { r.particle_count = b - a; r.first_particle = a; first = false; return; }
            if (b > a) { SfgRope t = r; t.first_particle = a; t.particle_count = b - a; fresh.push_back(t); }
        };
        for (auto d = d0; d != d1; ++d) { piece(run_start, *d); run_start = *d + 1; }
        piece(run_start, hi);
    }
    for (const SfgRope& t : fresh) w->rope_tab.push_back(t);
    // delete empty ropes (rope.rs:275-285); a rope left with ONE particle stays, and the next tick reports the reference's panic
    std::vector<SfgRope> kept;
    for (const SfgRope& r : w->rope_tab) if (r.particle_count > 0) kept.push_back(r);
    w->rope_tab.swap(kept);
    // the particle array: dead slots become holes, every piece carries its rope's (new) index
    int rc = relabel_rope_particles(w);
    if (n_ropes_out) *n_ropes_out = w->n_ropes;
    return rc;
}

// The rope table as the library holds it now, compacted: ropes_out[i].first_particle indexes particle_bodies_out (holes removed).
int sfg_get_ropes(SfgWorld* w, SfgRope* ropes_out, uint32_t rope_capacity, uint32_t* n_ropes_out, int32_t* particle_bodies_out, uint32_t particle_capacity,
                  uint32_t* n_particles_out) {
    if (!w || !n_ropes_out || !n_particles_out) return SFG_E_INVALID;
    uint32_t total = 0;
    for (const SfgRope& r : w->rope_tab) total += r.particle_count;
    *n_ropes_out = uint32_t(w->rope_tab.size()); *n_particles_out = total;
    if (!ropes_out || !particle_bodies_out) return SFG_OK;
    if (rope_capacity < w->rope_tab.size() || particle_capacity < total) return fail(w, SFG_E_CAPACITY, "sfg_get_ropes: output buffers too small");
    cudaSetDevice(w->device);
    std::vector<int32_t> all(w->n_particles);
    if (w->n_particles) {
        CK(cudaMemcpyAsync(all.data(), w->rp_body.p, size_t(w->n_particles) * 4, cudaMemcpyDeviceToHost, w->stream));
        CK(cudaStreamSynchronize(w->stream));
    }
    uint32_t o = 0;
    for (size_t i = 0; i < w->rope_tab.size(); ++i) {
        const SfgRope& r = w->rope_tab[i];
        ropes_out[i] = r; ropes_out[i].first_particle = o;
        for (uint32_t k = 0; k < r.particle_count; ++k) particle_bodies_out[o++] = all[r.first_particle + k];
    }
    return SFG_OK;
}

}  // extern "C"

namespace {

// Choose the hierarchical grid from the measured bounds (host side, a few dozen flops per tick).
void choose_grid(SfgWorld* w, const Bounds& b) {
    GridParams& g = w->grid;
    memset(&g, 0, sizeof(g));
    const float min_x = ord2f_host(b.min_x), min_y = ord2f_host(b.min_y), max_x = ord2f_host(b.max_x), max_y = ord2f_host(b.max_y);
    const float max_ext = ord2f_host(b.max_extent);
    const double mean_ext = b.n_live ? b.sum_extent / double(b.n_live) : 1.0;
    double c0 = w->tuning.cell_size > 0.f ? double(w->tuning.cell_size) : 2.0 * mean_ext;
    if (!(c0 > 1e-6)) c0 = 1.0;
    const double wx = std::max(double(max_x) - double(min_x), 0.0), wy = std::max(double(max_y) - double(min_y), 0.0);
    g.ox = min_x; g.oy = min_y;
    g.n_domains = b.max_domain + 1;
    const double cap = std::min(double(1u << 26), 8.0 * double(b.n_live) + 4096.0);
    for (;;) {
        int nl = 1;
        while (nl < MAX_LEVELS && c0 * std::pow(4.0, nl - 1) < double(max_ext)) ++nl;
        double total = 0;
        for (int l = 0; l < nl; ++l) {
            const double cell = c0 * std::pow(4.0, l);
            const double nx = std::floor(wx / cell) + 1.0, ny = std::floor(wy / cell) + 1.0;
            total += nx * ny * double(g.n_domains);
        }
        // the proxy grid repeats level 0 and the big-plain levels repeat the coarse ones
        if (total + total <= cap || c0 > 1e30) {
            g.n_levels = nl;
            uint32_t base = 0;
            for (int l = 0; l < nl; ++l) {
                const double cell = c0 * std::pow(4.0, l);
                g.cell[l] = float(cell); g.inv_cell[l] = float(1.0 / cell);
                g.nx[l] = int(std::floor(wx / cell) + 1.0); g.ny[l] = int(std::floor(wy / cell) + 1.0);
                g.base[l] = base;
                base += uint32_t(g.nx[l]) * uint32_t(g.ny[l]) * g.n_domains;
            }
            g.n_cells = base;
            base += 1;                                   // the huge list
            for (int l = 1; l < nl; ++l) { g.bp_base[l] = base; base += uint32_t(g.nx[l]) * uint32_t(g.ny[l]) * g.n_domains; }
            g.bp_base[0] = base;
            g.proxy_base = base;
            base += uint32_t(g.nx[0]) * uint32_t(g.ny[0]) * g.n_domains;
            g.n_keys = base;
            return;
        }
        c0 *= 1.5;
    }
}

// (rope_tab, rp_body) -> RU / RD / rope_next / rope_prev / per-rope parameters, all on the device (sfg_set_ropes' comment)
int rebuild_rope_units(SfgWorld* w) {
    cudaStream_t st = w->stream;
    w->ropes_dirty = false;
    const uint32_t nr = uint32_t(w->rope_tab.size());
    w->n_ropes = nr;
    w->n_rope_units = w->n_rope_damp = 0;
    for (int c = 0; c < 4; ++c) w->rope_colour_start[c] = 0;
    for (int c = 0; c < 3; ++c) w->damp_colour_start[c] = 0;
    CK(w->rope_next.ensure(w->n_bodies + 1)); CK(w->rope_prev.ensure(w->n_bodies + 1));
    if (w->n_bodies) {
        CK(cudaMemsetAsync(w->rope_next.p, 0xFF, size_t(w->n_bodies) * 4, st));
        CK(cudaMemsetAsync(w->rope_prev.p, 0xFF, size_t(w->n_bodies) * 4, st));
    }
    if (nr == 0) { w->n_particles = 0; return SFG_OK; }
    std::vector<RopeBuildRec> recs(nr);
    std::vector<float4> pa(nr), pb(nr);
    uint32_t ru_n[3] = {0, 0, 0}, rd_n[2] = {0, 0};
    for (uint32_t r = 0; r < nr; ++r) {
        const SfgRope& rp = w->rope_tab[r];
        if (rp.particle_count < 2) return fail(w, SFG_E_INVALID, "Rope only had one particle");   // solver.rs:120 panics
        if (uint64_t(rp.first_particle) + rp.particle_count > w->n_particles) return fail(w, SFG_E_STATE, "rope table out of step with the particle array");
        const uint32_t links = rp.particle_count - 1;
        for (uint32_t c = 0; c < 3; ++c) ru_n[c] += links > c ? (links - c + 2) / 3 : 0;
        for (uint32_t c = 0; c < 2; ++c) rd_n[c] += links > c ? (links - c + 1) / 2 : 0;
    }
    w->rope_colour_start[0] = 0; w->rope_colour_start[1] = ru_n[0]; w->rope_colour_start[2] = ru_n[0] + ru_n[1]; w->rope_colour_start[3] = ru_n[0] + ru_n[1] + ru_n[2];
    w->damp_colour_start[0] = 0; w->damp_colour_start[1] = rd_n[0]; w->damp_colour_start[2] = rd_n[0] + rd_n[1];
    uint32_t ru_cur[3] = {w->rope_colour_start[0], w->rope_colour_start[1], w->rope_colour_start[2]}, rd_cur[2] = {w->damp_colour_start[0], w->damp_colour_start[1]};
    for (uint32_t r = 0; r < nr; ++r) {
        const SfgRope& rp = w->rope_tab[r];
        const uint32_t links = rp.particle_count - 1;
        RopeBuildRec& o = recs[r];
        o.first = rp.first_particle; o.count = rp.particle_count; o.pad = 0;
        for (uint32_t c = 0; c < 3; ++c) { o.ru_off[c] = ru_cur[c]; ru_cur[c] += links > c ? (links - c + 2) / 3 : 0; }
        for (uint32_t c = 0; c < 2; ++c) { o.rd_off[c] = rd_cur[c]; rd_cur[c] += links > c ? (links - c + 1) / 2 : 0; }
        pa[r] = make_float4(rp.spacing, rp.compliance, rp.bending_max_angle, rp.bending_compliance);
        pb[r] = make_float4(rp.damping, 0.f, 0.f, 0.f);
    }
    w->n_rope_units = w->rope_colour_start[3]; w->n_rope_damp = w->damp_colour_start[2];
    CK(w->RU.ensure(w->n_rope_units + 1)); CK(w->RD.ensure(w->n_rope_damp + 1)); CK(w->RPA.ensure(nr)); CK(w->RPB.ensure(nr)); CK(w->rope_build.ensure(nr));
    CK(w->PCD.ensure(w->n_bodies)); CK(w->LAT.ensure(w->n_bodies));
    CK(cudaMemcpyAsync(w->rope_build.p, recs.data(), size_t(nr) * sizeof(RopeBuildRec), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(w->RPA.p, pa.data(), size_t(nr) * sizeof(float4), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(w->RPB.p, pb.data(), size_t(nr) * sizeof(float4), cudaMemcpyHostToDevice, st));
    k_rope_build<<<nblk(w->n_particles), 256, 0, st>>>(w->n_particles, w->rp_body.p, w->rp_rope.p, w->rope_build.p, w->RU.p, w->RD.p, w->rope_next.p, w->rope_prev.p);
    CKL(w);
    CK(cudaStreamSynchronize(st));        // recs / pa / pb are stack-lifetime host buffers
    return SFG_OK;
}

template <uint32_t KIND>
int launch_stage(SfgWorld* w, const SolverCtx& ctx, const Stage& s, bool last, uint32_t sub) {
    const uint32_t begin = s.begin, end = s.end;
    if (end <= begin) return SFG_OK;
    k_stage<KIND><<<nblk(end - begin), 256, 0, w->stream>>>(ctx, begin, end, last ? 1 : 0, sub, s.colour);
    CKL(w);
    w->launches++;
    return SFG_OK;
}

int launch_stage_dyn(SfgWorld* w, const SolverCtx& ctx, const Stage& s, bool last, uint32_t sub) {
    switch (s.kind) {
    case ST_INTEGRATE: return launch_stage<ST_INTEGRATE>(w, ctx, s, last, sub);
    case ST_ROPE_STEP: return launch_stage<ST_ROPE_STEP>(w, ctx, s, last, sub);
    case ST_CONSTRAINTS: return launch_stage<ST_CONSTRAINTS>(w, ctx, s, last, sub);
    case ST_PRECONTACT: return launch_stage<ST_PRECONTACT>(w, ctx, s, last, sub);
    case ST_CONTACTS: return launch_stage<ST_CONTACTS>(w, ctx, s, last, sub);
    case ST_DERIVE: return launch_stage<ST_DERIVE>(w, ctx, s, last, sub);
    case ST_CONTACT_VEL: return launch_stage<ST_CONTACT_VEL>(w, ctx, s, last, sub);
    case ST_CONSTRAINT_DAMP: return launch_stage<ST_CONSTRAINT_DAMP>(w, ctx, s, last, sub);
    case ST_ROPE_DAMP: return launch_stage<ST_ROPE_DAMP>(w, ctx, s, last, sub);
    case ST_ROPE_LATERAL: return launch_stage<ST_ROPE_LATERAL>(w, ctx, s, last, sub);
    default: return fail(w, SFG_E_STATE, "unknown solver stage");
    }
}

int broadphase(SfgWorld* w, float frame_dt) {
    cudaStream_t st = w->stream;
    const uint32_t nc = w->n_colls;
    w->n_units = 0; w->n_live = 0;
    if (nc == 0) return SFG_OK;
    // 1. boxes + bounds
    TickHeader h0;
    memset(&h0, 0, sizeof(h0));
    h0.bounds.min_x = h0.bounds.min_y = 0x7FFFFFFF; h0.bounds.max_x = h0.bounds.max_y = int(0x80000000u); h0.bounds.max_extent = int(0x80000000u);
    CK(cudaMemcpyAsync(w->hdr.p, &h0, sizeof(h0), cudaMemcpyHostToDevice, st));
    const float pad = w->tuning.max_expected_acceleration * frame_dt;
    k_aabb<<<nblk(nc), 256, 0, st>>>(nc, w->CS.p, w->CL.p, w->CI.p, w->P.p, w->Q.p, w->V.p, w->MI.p, frame_dt, pad, w->AABB.p, &w->hdr.p->bounds);
    CKL(w); w->launches++;
    static_assert(sizeof(TickHeader) <= 512, "pinned read-back block");
    TickHeader& h = *reinterpret_cast<TickHeader*>(w->pin);
    CK(cudaMemcpyAsync(&h, w->hdr.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    w->n_live = h.bounds.n_live;
    if (w->n_live == 0) { memset(&w->grid, 0, sizeof(w->grid)); return SFG_OK; }
    choose_grid(w, h.bounds);
    const GridParams& g = w->grid;
    // 2. entries (home + proxies / big-plain, sfg_state.cuh GridParams) + sort
    CK(w->cell_cnt.ensure(nc + 1)); CK(w->cell_off.ensure(nc + 1)); CK(w->scan_tmp.ensure(scan_tmp_words(nc + 1)));
    k_cell_count<<<nblk(nc), 256, 0, st>>>(g, nc, w->AABB.p, w->cell_cnt.p);
    CKL(w); w->launches++;
    w->launches += scan_exclusive(w->cell_cnt.p, w->cell_off.p, nc, w->scan_tmp.p, &w->hdr.p->n_entries, st);
    CK(cudaMemcpyAsync(w->pin + 200, &w->hdr.p->n_entries, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t ne = w->pin[200];
    w->n_entries = ne;
    if (ne == 0) { w->n_live = 0; return SFG_OK; }
    CK(w->keys0.ensure(ne)); CK(w->keys1.ensure(ne)); CK(w->vals0.ensure(ne)); CK(w->vals1.ensure(ne));
    CK(w->sort_tmp.ensure(sort_tmp_words(ne)));
    k_cell_emit<<<nblk(nc), 256, 0, st>>>(g, nc, w->AABB.p, w->CI.p, w->cell_off.p, w->keys0.p, w->vals0.p);
    CKL(w); w->launches++;
    int bits = 1;
    while (bits < 32 && (1ull << bits) <= uint64_t(g.n_keys) + 1) ++bits;
    int launches = 0;
    const int where = radix_sort_pairs(w->keys0.p, w->vals0.p, w->keys1.p, w->vals1.p, ne, bits, w->sort_tmp.p, st, &launches);
    w->launches += launches;
    const uint32_t* keys = where ? w->keys1.p : w->keys0.p;
    const uint32_t* vals = where ? w->vals1.p : w->vals0.p;
    // 3. sorted copies + cell starts
    // the key space is bounded by choose_grid's cap (cells of all levels <= cap / 2, repeated once for proxies and big-plain
    // levels): sized for the bound once, so a world that spreads out never reallocates its cell table in the middle of a run
    const size_t key_bound = size_t(std::min(double(1u << 26), 8.0 * double(w->n_live) + 4096.0)) + 2;
    CK(w->SA.ensure(ne)); CK(w->SI.ensure(ne)); CK(w->cell_start.ensure(std::max(size_t(g.n_keys) + 2, key_bound)));
    k_gather_sorted<<<nblk(ne), 256, 0, st>>>(g, ne, keys, vals, w->AABB.p, w->CI.p, w->SA.p, w->SI.p);
    CKL(w); w->launches++;
    k_cell_starts<<<nblk(ne + 1), 256, 0, st>>>(keys, ne, g.n_keys, w->cell_start.p);
    CKL(w); w->launches++;
    // 4. candidate pairs in one pass; capacity is last tick's count + 25 % (re-run once if it overflows)
    const uint32_t n_warps = (w->n_live + 31) / 32;
    size_t want = std::max<size_t>(size_t(w->prev_pairs) + w->prev_pairs / 4 + 4096, size_t(w->n_live) * RESERVE_PAIRS_PER_COLLIDER);
    for (int attempt = 0; attempt < 2; ++attempt) {
        CK(w->U_ids.ensure(want));
        const uint32_t cap = uint32_t(std::min<size_t>(w->U_ids.cap, 0xFFFFFFFFu));
        CK(cudaMemsetAsync(&w->hdr.p->n_pairs, 0, 4, st));
        k_pairs<<<nblk(n_warps * 32), 256, 0, st>>>(g, w->n_live, w->SA.p, w->SI.p, keys, w->cell_start.p, w->d_mask.p, &w->hdr.p->n_pairs, w->U_ids.p, cap);
        CKL(w); w->launches++;
        CK(cudaMemcpyAsync(w->pin + 201, &w->hdr.p->n_pairs, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const uint32_t n_pairs = w->pin[201];
        w->n_units = n_pairs;
        w->prev_pairs = n_pairs;
        if (n_pairs <= cap) break;
        if (attempt == 1) return fail(w, SFG_E_CAPACITY, "candidate pair buffer overflowed twice");
        want = size_t(n_pairs) + 1024;
    }
    return SFG_OK;
}

int join_islands(SfgWorld* w);
int build_graph(SfgWorld* w) {
    cudaStream_t st = w->stream;
    const uint32_t nb = w->n_bodies;
    w->n_unit_colours = 0; w->n_k_colours = 0; w->n_kunits = 0;
    w->unit_colour_start.assign(1, 0); w->k_colour_start.assign(1, 0);
    CK(w->deg.ensure(nb + 2));
    // ---- pair units
    if (w->n_units) {
        const uint32_t n = w->n_units;
        const size_t nr = std::max<size_t>(n, size_t(w->n_live) * RESERVE_PAIRS_PER_COLLIDER);
        CK(w->U_ub.ensure(nr)); CK(w->U_prio.ensure(nr)); CK(w->U_cls.ensure(nr));
        CK(cudaMemsetAsync(w->deg.p, 0, size_t(nb + 2) * 4, st));
        k_pair_unit_setup<<<nblk(n), 256, 0, st>>>(n, w->U_ids.p, w->CREC.p, w->MI.p, w->P.p, w->Q.p, w->V.p, w->frame_dt, w->U_ub.p, w->U_prio.p, w->deg.p, w->U_cls.p);
        CKL(w); w->launches++;
        uint32_t* perm = nullptr;
        int rc = colour_units(w, n, w->U_ids.p, w->U_ub.p, w->U_prio.p, w->U_colour, &perm, w->unit_colour_start, &w->n_unit_colours, true);
        if (rc != SFG_OK) return rc;
        CK(w->RC.ensure(nr)); CK(w->WL.ensure(nr));
        CK(w->H.ensure(nr)); CK(w->S.ensure(nr)); CK(w->G0.ensure(nr)); CK(w->G1.ensure(nr)); CK(w->L0.ensure(nr)); CK(w->L1.ensure(nr));
        CK(w->M0.ensure(nr * 2)); CK(w->M1.ensure(nr * 2)); CK(w->M2.ensure(nr * 2)); CK(w->LN.ensure(nr * 2));
        { int jr = join_islands(w); if (jr != SFG_OK) return jr; }
        k_build_pair_units<<<nblk(n), 256, 0, st>>>(n, perm, w->U_ids.p, w->CREC.p, w->MI.p,
                                                    w->n_ropes ? w->rope_next.p : nullptr, w->n_ropes ? w->rope_prev.p : nullptr, w->H.p, w->S.p,
                                                    w->G0.p, w->G1.p, w->L0.p, w->L1.p, w->RC.p, w->tick_counts.p);
        CKL(w); w->launches++;
        CK(cudaMemsetAsync(w->LN.p, 0xFF, size_t(n) * 2 * sizeof(float2), st));
        // manifold tags carry the tick number (sfg_state.cuh SolverCtx::tag_base): the column is cleared only when the 12-bit tick
        // counter wraps, or when the buffer is new memory
        if ((w->tick_seq & 0xFFFu) == 0u || w->M0.p != w->m0_cleared) {
            CK(cudaMemsetAsync(w->M0.p, 0, w->M0.cap * sizeof(float4), st));
            w->m0_cleared = w->M0.p;
        }
    }
    // ---- constraint units
    if (w->n_constraints) {
        const uint32_t n = w->n_constraints;
        w->n_kunits = n;
        CK(w->KU_ids.ensure(n)); CK(w->KU_ub.ensure(n)); CK(w->KU_prio.ensure(n));
        CK(cudaMemsetAsync(w->deg.p, 0, size_t(nb + 2) * 4, st));
        k_constraint_unit_setup<<<nblk(n), 256, 0, st>>>(n, w->KI.p, w->MI.p, w->KU_ids.p, w->KU_ub.p, w->KU_prio.p, w->deg.p);
        CKL(w); w->launches++;
        uint32_t* perm = nullptr;
        int rc = colour_units(w, n, w->KU_ids.p, w->KU_ub.p, w->KU_prio.p, w->KU_colour, &perm, w->k_colour_start, &w->n_k_colours, false);
        if (rc != SFG_OK) return rc;
        CK(w->K0.ensure(n)); CK(w->K1.ensure(n)); CK(w->K2.ensure(n));
        k_build_constraint_units<<<nblk(n), 256, 0, st>>>(n, perm, w->KI.p, w->KA.p, w->KB.p, w->K0.p, w->K1.p, w->K2.p);
        CKL(w); w->launches++;
    }
    return SFG_OK;
}

void fill_ctx(SfgWorld* w, SolverCtx& c, float dt, const SfgForceField* ff) {
    memset(&c, 0, sizeof(c));
    c.P = w->P.p; c.Q = w->Q.p; c.V = w->V.p; c.OV = w->OV.p; c.MI = w->MI.p; c.EA = w->EA.p; c.PCD = w->PCD.p;
    c.rope_next = w->rope_next.p; c.rope_prev = w->rope_prev.p; c.n_bodies = w->n_bodies; c.bouncy = w->bouncy_seen ? 1u : 0u;
    c.H = w->H.p; c.S = w->S.p; c.G0 = w->G0.p; c.G1 = w->G1.p; c.L0 = w->L0.p; c.L1 = w->L1.p;
    c.M0 = w->M0.p; c.M1 = w->M1.p; c.M2 = w->M2.p; c.LN = w->LN.p; c.n_units = w->n_units;
    c.stat_touch = (w->tuning.flags & SFG_TUNE_COUNT_ENTRIES) ? w->tick_counts.p + 2 : nullptr;
    c.tag_base = (w->tick_seq & 0xFFFu) << 12;
    c.RC = w->RC.p; c.WL = w->WL.p; c.wl_count = w->wl_count.p; c.near_count = w->near_count.p; c.n_colours = w->n_unit_colours;
    c.K0 = w->K0.p; c.K1 = w->K1.p; c.K2 = w->K2.p; c.n_kunits = w->n_kunits;
    c.RU = w->RU.p; c.RD = w->RD.p; c.RPA = w->RPA.p; c.RPB = w->RPB.p; c.rp_body = w->rp_body.p; c.rp_rope = w->rp_rope.p;
    c.LAT = w->LAT.p; c.n_particles = w->n_particles;
    c.dt = dt; c.inv_dt = float(1.0 / double(dt)); c.inv_dt_sq = c.inv_dt * c.inv_dt;
    c.ff.n = 0;
    if (ff)
        for (uint32_t i = 0; i < ff->n_terms && i < 4; ++i) {
            c.ff.kind[c.ff.n] = ff->terms[i].kind; c.ff.a[c.ff.n] = ff->terms[i].a; c.ff.b[c.ff.n] = ff->terms[i].b;
            c.ff.c[c.ff.n] = ff->terms[i].c; c.ff.d[c.ff.n] = ff->terms[i].d; c.ff.n++;
        }
}

int slab_classify(SfgWorld* w) {
    const uint32_t nb = w->n_bodies;
    if (nb == 0) return SFG_OK;
    CK(w->slab_stamp.ensure(nb + 1));
    CK(w->zone_list.ensure(size_t(2) * SLAB_ZONE_CAP)); CK(w->zone_count.ensure(2));
    CK(cudaMemsetAsync(w->zone_count.p, 0, 8, w->stream));
    k_slab_classify<<<nblk(nb), 256, 0, w->stream>>>(nb, w->P.p, w->MI.p, w->slab_stamp.p, w->slab_epoch, w->slab_lo, w->slab_hi, w->slab_halo, w->slab_on ? 1 : 0,
                                                   w->zone_list.p, w->zone_count.p, SLAB_ZONE_CAP);
    CKL(w); w->launches++;
    w->zone_valid = w->slab_on;
    w->slab_was_on = w->slab_on;
    return SFG_OK;
}
bool sleeping_enabled(const SfgWorld* w) {
    return !(w->tuning.flags & SFG_TUNE_NO_SLEEPING) && w->tuning.fall_asleep_frames != 0xFFFFFFFFu;
}
void fill_island_ctx(SfgWorld* w, IslandCtx& c) {
    memset(&c, 0, sizeof(c));
    c.n_bodies = w->n_bodies; c.n_units = w->n_units; c.n_constraints = w->n_constraints; c.n_links = w->n_rope_units;
    c.U_ids = w->U_ids.p; c.CI = w->CI.p; c.KI = w->KI.p; c.RU = w->RU.p; c.MI = w->MI.p; c.V = w->V.p;
    c.parent = w->isl_parent.p; c.isl_first = w->isl_first.p; c.isl_sum = w->isl_sum.p; c.isl_flags = w->isl_flags.p; c.isl_count = w->isl_count.p;
    c.sl_sum = w->sl_sum.p; c.sl_ticks = w->sl_ticks.p; c.sl_valid = w->sl_valid.p; c.counters = w->isl_counters.p;
    c.sleep_vel_threshold = w->tuning.sleep_vel_threshold; c.fall_asleep_frames = w->tuning.fall_asleep_frames;
    c.shard_rank = w->shard_rank; c.shard_n = w->shard_n; c.sleeping = sleeping_enabled(w) ? 1u : 0u;
}
// physics.rs:598-741: islands of the constraint graph, then which of them stay asleep this tick
int island_pass(SfgWorld* w, cudaStream_t st) {
    const uint32_t nb = w->n_bodies;
    if (nb == 0) return SFG_OK;
    CK(w->isl_parent.ensure(nb + 1)); CK(w->isl_first.ensure(nb + 1)); CK(w->isl_flags.ensure(nb + 1)); CK(w->isl_count.ensure(nb + 1)); CK(w->isl_sum.ensure(nb + 1));
    CK(w->isl_counters.ensure(8));
    IslandCtx c;
    fill_island_ctx(w, c);
    const uint32_t n_edges = c.n_units + c.n_constraints + c.n_links;
    k_isl_init<<<nblk(nb), 256, 0, st>>>(c); CKL(w); w->launches++;
    if (n_edges) {
        // SFG_ISL_PHASES (debug): 1 = every edge in one launch, 2 = halves, 3 = 1/4 1/4 1/2, 4 (default) = 1/8 1/8 1/4 1/2
        static const int n_phases = [] { const char* e = getenv("SFG_ISL_PHASES"); const int v = e ? atoi(e) : 4; return v < 1 ? 1 : (v > 4 ? 4 : v); }();
        static const uint32_t plan[4][4][2] = {{{0, 8}}, {{0, 4}, {4, 4}}, {{0, 2}, {2, 2}, {4, 4}}, {{0, 1}, {1, 1}, {2, 2}, {4, 4}}};
        for (int ph = 0; ph < n_phases; ++ph) {
            const uint32_t lo = plan[n_phases - 1][ph][0], width = plan[n_phases - 1][ph][1];
            const uint32_t n_threads = (n_edges + 7u) / 8u * width;
            k_isl_union<<<nblk(n_threads), 256, 0, st>>>(c, lo, width, n_edges, ph ? 1u : 0u); CKL(w); w->launches++;
            if (ph + 1 < n_phases) { k_isl_flatten<<<nblk(nb), 256, 0, st>>>(c); CKL(w); w->launches++; }   // k_isl_bodies_pre is the last one
        }
    }
    k_isl_bodies_pre<<<nblk(nb), 256, 0, st>>>(c); CKL(w); w->launches++;
    k_isl_decide<<<nblk(nb), 256, 0, st>>>(c); CKL(w); w->launches++;
    k_isl_mark<<<nblk(nb), 256, 0, st>>>(c); CKL(w); w->launches++;
    return SFG_OK;
}
// physics.rs:1097-1144: publish contacts (keeping those of sleeping islands), then watch slow islands
int island_post(SfgWorld* w) {
    cudaStream_t st = w->stream;
    const uint32_t nb = w->n_bodies;
    if (nb == 0) return SFG_OK;
    IslandCtx c;
    fill_island_ctx(w, c);
    const uint32_t n_entries = w->n_units * 2;
    DBuf<ContactRec>& old_list = w->ct[w->ct_cur];
    DBuf<ContactRec>& new_list = w->ct[w->ct_cur ^ 1];
    CK(new_list.ensure(std::max<size_t>(size_t(w->ct_n) + n_entries + 1, size_t(w->n_live) * RESERVE_PAIRS_PER_COLLIDER * 2)));
    uint32_t* cursor = w->isl_counters.p + 4;
    CK(cudaMemsetAsync(cursor, 0, 4, st));
    if (w->ct_n) {
        k_contacts_retain<<<nblk(w->ct_n), 256, 0, st>>>(w->ct_n, old_list.p, w->sl_sum.p, w->sl_ticks.p, w->sl_valid.p, nb, w->tuning.fall_asleep_frames,
                                                        new_list.p, cursor);
        CKL(w); w->launches++;
    }
    if (n_entries) {
        k_contacts_append<<<nblk(w->n_units), 256, 0, st>>>(w->n_units, w->LN.p, w->H.p, w->S.p, w->isl_parent.p, w->isl_first.p, w->isl_sum.p, new_list.p, cursor);
        CKL(w); w->launches++;
    }
    k_isl_bodies_post<<<nblk(nb), 256, 0, st>>>(c); CKL(w); w->launches++;
    k_isl_fall_asleep<<<nblk(nb), 256, 0, st>>>(c); CKL(w); w->launches++;
    w->ct_cur ^= 1;
    return SFG_OK;
}

// the island pass was forked onto stream2 (sfg_tick_begin); everything that reads its per-tick body flags waits here
int join_islands(SfgWorld* w) {
    if (!w->islands_in_flight) return SFG_OK;
    w->islands_in_flight = false;
    CK(cudaStreamWaitEvent(w->stream, w->ev_join, 0));
    return SFG_OK;
}

void build_stage_table(SfgWorld* w) {
    auto& s = w->h_stages;
    s.clear();
    auto add = [&](uint32_t kind, uint32_t b, uint32_t e, uint32_t colour = 0) { if (e > b) s.push_back(Stage{kind, b, e, colour}); };
    add(ST_INTEGRATE, 0, w->n_bodies);
    if (w->n_ropes) for (int c = 0; c < 3; ++c) add(ST_ROPE_STEP, w->rope_colour_start[c], w->rope_colour_start[c + 1]);
    for (uint32_t c = 0; c < w->n_k_colours; ++c) add(ST_CONSTRAINTS, w->k_colour_start[c], w->k_colour_start[c + 1]);
    if (w->n_ropes && w->n_units) add(ST_PRECONTACT, 0, w->n_particles);
    for (uint32_t c = 0; c < w->n_unit_colours; ++c) add(ST_CONTACTS, w->unit_colour_start[c], w->unit_colour_start[c + 1], c);
    add(ST_DERIVE, 0, w->n_bodies);
    for (uint32_t c = 0; c < w->n_unit_colours; ++c) add(ST_CONTACT_VEL, w->unit_colour_start[c], w->unit_colour_start[c + 1], c);
    for (uint32_t c = 0; c < w->n_k_colours; ++c) add(ST_CONSTRAINT_DAMP, w->k_colour_start[c], w->k_colour_start[c + 1]);
    if (w->n_ropes) {
        for (int c = 0; c < 2; ++c) add(ST_ROPE_DAMP, w->damp_colour_start[c], w->damp_colour_start[c + 1]);
        add(ST_ROPE_LATERAL, 0, w->n_particles);
    }
}

}  // namespace

extern "C" {

// A tick in three parts, so that a slab-decomposed world can exchange halo bodies between substeps (sfg_slab_*):
// begin = boxes, candidate pairs, islands, colouring; substeps = the persistent solver kernel over a range of substeps;
// end = contact list, sleeping bookkeeping, statistics. sfg_tick runs all three.
int sfg_tick_begin(SfgWorld* w, double frame_dt, int has_time_scale, double time_scale, const SfgForceField* ff) {
    if (!w) return SFG_E_INVALID;
    if (!(frame_dt > 0.0)) return fail(w, SFG_E_INVALID, "frame_dt must be > 0");
    if (ff && ff->n_terms > 4) return fail(w, SFG_E_INVALID, "force field: at most 4 terms");
    int rc = ensure_device(w->device);
    if (rc != SFG_OK) return fail(w, rc, "no CUDA device (there is no CPU fallback)");
    cudaStream_t st = w->stream;
    w->in_tick = false;
    // physics.rs:404-418
    uint32_t substeps;
    double dt;
    if (!has_time_scale) { substeps = w->tuning.substeps; dt = frame_dt / double(substeps); }
    else {
        // physics.rs:411-415: `(time_scale * substeps).ceil() as usize` — Rust's float-to-int cast saturates: negative and NaN
        // give 0 substeps (the tick then only runs the broadphase and the bookkeeping), it never wraps
        const double want = std::ceil(time_scale * double(w->tuning.substeps));
        if (!(want >= 1.0)) { substeps = 0; dt = 0.0; }
        else if (want > double(SFG_MAX_SUBSTEPS)) return fail(w, SFG_E_INVALID, "time_scale * substeps exceeds SFG_MAX_SUBSTEPS");
        else { substeps = uint32_t(want); dt = time_scale * frame_dt / double(substeps); }
    }
    if (substeps > SFG_MAX_SUBSTEPS) return fail(w, SFG_E_INVALID, "substeps exceed SFG_MAX_SUBSTEPS");
    w->launches = 0;
    w->islands_in_flight = false;
    w->frame_dt = float(frame_dt);
    w->tick_seq += 1;
    if (w->ropes_dirty) { rc = rebuild_rope_units(w); if (rc != SFG_OK) return rc; }
    CK(cudaEventRecord(w->ev[0], st));
    CK(cudaMemsetAsync(w->tick_counts.p, 0, 4 * sizeof(uint32_t), st));
    if (w->n_bodies) CK(cudaMemsetAsync(w->EA.p, 0, size_t(w->n_bodies) * sizeof(float2), st));
    if (w->n_ropes) CK(cudaMemsetAsync(w->LAT.p, 0xFF, size_t(w->n_bodies) * sizeof(float2), st));
    if (w->slab_on || w->slab_was_on) { rc = slab_classify(w); if (rc != SFG_OK) return rc; }
    rc = broadphase(w, float(frame_dt));
    if (rc != SFG_OK) return rc;
    CK(cudaEventRecord(w->ev[1], st));
    // islands across a slab seam are not tracked: slabs never sleep. Island sharding needs the islands even if nothing may sleep.
    w->tick_sleeping = (sleeping_enabled(w) || w->shard_n > 1) && !w->slab_on;
    // islands and colouring both only read the pair list: the island pass runs on a second stream next to the colouring
    // (union-find and the colouring dataflow are latency-bound and leave most of the machine idle on their own); the join is
    // in front of k_build_pair_units, the first reader of the asleep / foreign flags the island pass leaves on the bodies
    static const bool islands_inline = getenv("SFG_ISLANDS_INLINE") != nullptr;      // A/B switch: the island pass on the main stream, in front of the colouring
    if (w->tick_sleeping && islands_inline) {
        w->island_flags_live = true;
        rc = island_pass(w, st);
        if (rc != SFG_OK) return rc;
    } else if (w->tick_sleeping) {
        CK(cudaEventRecord(w->ev_fork, st));
        CK(cudaStreamWaitEvent(w->stream2, w->ev_fork, 0));
        w->island_flags_live = true;
        rc = island_pass(w, w->stream2);
        if (rc != SFG_OK) { cudaStreamSynchronize(w->stream2); return rc; }
        CK(cudaEventRecord(w->ev_join, w->stream2));
        w->islands_in_flight = true;
    } else if (w->island_flags_live && w->n_bodies) {
        // the island pass is off this tick: nobody stays asleep (or foreign) from a tick that ran it, and no island is being watched
        k_clear_island_flags<<<nblk(w->n_bodies), 256, 0, st>>>(w->n_bodies, w->MI.p);
        CKL(w); w->launches++;
        if (w->sl_valid.p) CK(cudaMemsetAsync(w->sl_valid.p, 0, size_t(w->n_bodies + 1) * 4, st));
        w->island_flags_live = false;
        w->ct_n = 0;
    }
    rc = build_graph(w);
    if (rc != SFG_OK) { if (w->islands_in_flight) { cudaStreamSynchronize(w->stream2); w->islands_in_flight = false; } return rc; }
    rc = join_islands(w);
    if (rc != SFG_OK) return rc;
    CK(cudaEventRecord(w->ev[2], st));
    w->tick_substeps = substeps; w->tick_next_substep = 0;
    if (substeps > 0 && w->n_bodies) {
        fill_ctx(w, w->tick_ctx, float(dt), ff);
        build_stage_table(w);
        if (w->n_unit_colours) {
            CK(w->wl_count.ensure(size_t(substeps) * w->n_unit_colours));
            CK(cudaMemsetAsync(w->wl_count.p, 0, size_t(substeps) * w->n_unit_colours * 4, st));
            w->tick_ctx.wl_count = w->wl_count.p;
            CK(w->far_min.ensure(size_t(substeps) * w->h_stages.size()));
            CK(cudaMemsetAsync(w->far_min.p, 0xFF, size_t(substeps) * w->h_stages.size() * 4, st));
            w->tick_ctx.far_min = w->far_min.p;
        }
        {   // one work-stealing counter per stage and substep (k_solve_persistent deals the tail of a big stage from it)
            const size_t n_ctr = size_t(substeps) * w->h_stages.size() + 1;
            CK(w->steal.ensure(n_ctr));
            CK(cudaMemsetAsync(w->steal.p, 0, n_ctr * 4, st));
            w->tick_ctx.steal = w->steal.p;
        }
        if (!(w->tuning.flags & SFG_TUNE_SEPARATE_LAUNCHES)) {
            const uint32_t ns = uint32_t(w->h_stages.size());
            CK(w->d_stages.ensure(ns));
            CK(cudaMemcpyAsync(w->d_stages.p, w->h_stages.data(), ns * sizeof(Stage), cudaMemcpyHostToDevice, st));
            CK(cudaMemsetAsync(&w->hdr.p->barrier, 0, 4, st));
        }
    }
    w->in_tick = true;
    return SFG_OK;
}

int sfg_tick_substeps(SfgWorld* w, uint32_t count) {
    if (!w) return SFG_E_INVALID;
    if (!w->in_tick) return fail(w, SFG_E_STATE, "sfg_tick_substeps outside sfg_tick_begin / sfg_tick_end");
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    const uint32_t substeps = w->tick_substeps, first = w->tick_next_substep;
    if (uint64_t(first) + count > substeps) return fail(w, SFG_E_INVALID, "more substeps than the tick has");
    w->tick_next_substep = first + count;
    if (count == 0 || w->n_bodies == 0) return SFG_OK;
    const SolverCtx& ctx = w->tick_ctx;
    int rc;
    if (w->tuning.flags & SFG_TUNE_SEPARATE_LAUNCHES) {
        static const bool debug_nan = getenv("SFG_DEBUG_NAN") != nullptr;
        bool reported = false;
        for (uint32_t s = first; s < first + count; ++s)
            for (size_t k = 0; k < w->h_stages.size(); ++k) {
                const Stage& sg = w->h_stages[k];
                rc = launch_stage_dyn(w, ctx, sg, s + 1 == substeps, s);
                if (rc != SFG_OK) return rc;
                if (debug_nan && !reported) {
                    uint32_t* flag = w->isl_counters.p ? w->isl_counters.p + 7 : nullptr;
                    if (!flag) { CK(w->isl_counters.ensure(8)); flag = w->isl_counters.p + 7; }
                    CK(cudaMemsetAsync(flag, 0xFF, 4, st));
                    k_find_nonfinite<<<nblk(w->n_bodies), 256, 0, st>>>(w->n_bodies, w->P.p, w->Q.p, w->V.p, w->MI.p, flag);
                    uint32_t fb = 0xFFFFFFFFu;
                    CK(cudaMemcpyAsync(&fb, flag, 4, cudaMemcpyDeviceToHost, st));
                    CK(cudaStreamSynchronize(st));
                    if (fb != 0xFFFFFFFFu) {
                        fprintf(stderr, "SFG_DEBUG_NAN: first non-finite body %u after substep %u stage %zu (kind %u, colour %u, range [%u, %u))\n", fb, s, k,
                                sg.kind, sg.colour, sg.begin, sg.end);
                        reported = true;
                    }
                }
            }
        return SFG_OK;
    }
    const uint32_t ns = uint32_t(w->h_stages.size());
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_persistent, SFG_SOLVER_THREADS, 0));
    if (per_sm < 1) return fail(w, SFG_E_CUDA, "persistent solver kernel does not fit on an SM");
    uint32_t max_range = 0;
    for (const Stage& sg : w->h_stages) max_range = std::max(max_range, sg.end - sg.begin);
    const uint32_t grid = std::max(1u, std::min(uint32_t(per_sm * w->n_sms), nblk(max_range, SFG_SOLVER_THREADS)));
    const Stage* dst = w->d_stages.p;
    uint32_t n_stages = ns, sub_begin = first, sub_end = first + count, n_sub = substeps;
    unsigned int* bar = &w->hdr.p->barrier;
    CK(cudaMemsetAsync(bar, 0, 4, st));
    // debug: SFG_TRACE=<file> dumps per-stage, per-CTA (arrival, release) %globaltimer stamps of the barrier
    const char* trace_path = getenv("SFG_TRACE");
    unsigned long long* trace = nullptr;
    const size_t trace_n = size_t(2) * count * ns * grid;
    if (trace_path) { CK(w->trace.ensure(trace_n)); trace = w->trace.p; CK(cudaMemsetAsync(trace, 0, trace_n * 8, st)); }   // skipped stages stay 0
    static const bool debug_nan = getenv("SFG_DEBUG_NAN") != nullptr;
    uint32_t* dbg = nullptr;
    if (debug_nan) { CK(w->dbg.ensure(4096 + 8)); CK(cudaMemsetAsync(w->dbg.p, 0xFF, (4096 + 8) * 4, st)); dbg = w->dbg.p; }
    void* args[] = {(void*)&ctx, (void*)&dst, &n_stages, &sub_begin, &sub_end, &n_sub, (void*)&bar, (void*)&trace, (void*)&dbg};
    CK(cudaLaunchCooperativeKernel((void*)k_solve_persistent, dim3(grid), dim3(SFG_SOLVER_THREADS), args, 0, st));
    w->launches++;
    if (debug_nan) {
        std::vector<uint32_t> h(4096 + 8);
        CK(cudaMemcpyAsync(h.data(), dbg, h.size() * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h[0] != 0xFFFFFFFFu) {
            const uint32_t idx = h[0], s = idx / ns, k = idx % ns;
            const Stage& sg = w->h_stages[k];
            fprintf(stderr, "SFG_DEBUG_NAN (persistent): first non-finite body %u after substep %u stage %u of %u (kind %u, colour %u, range [%u, %u))\n",
                    h[1 + std::min(idx, 4000u)], s, k, ns, sg.kind, sg.colour, sg.begin, sg.end);
        }
    }
#ifdef SFG_PROBE
    if (const char* pp = getenv("SFG_PROBE_OUT")) {
        std::vector<unsigned long long> h(1 + 8 * 65536);
        CK(cudaStreamSynchronize(st));
        CK(cudaMemcpyFromSymbol(h.data(), g_probe, h.size() * 8));
        if (FILE* f = fopen(pp, "wb")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
        unsigned long long zero = 0;
        CK(cudaMemcpyToSymbol(g_probe, &zero, 8));
    }
#endif
    if (trace_path) {
        std::vector<unsigned long long> h(trace_n);
        CK(cudaMemcpyAsync(h.data(), trace, trace_n * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (FILE* f = fopen(trace_path, "wb")) {
            uint32_t hdr4[4] = {count, ns, grid, 0};
            fwrite(hdr4, 4, 4, f);
            for (const Stage& sg : w->h_stages) { uint32_t r[4] = {sg.kind, sg.begin, sg.end, 0}; fwrite(r, 4, 4, f); }
            fwrite(h.data(), 8, trace_n, f);
            fclose(f);
        }
    }
    return SFG_OK;
}

static int tick_end_impl(SfgWorld* w, float* poses_out);
int sfg_tick_end(SfgWorld* w) { return tick_end_impl(w, nullptr); }
// poses_out != nullptr: the poses of all bodies are packed and copied to the host on the second stream while the contact
// list and the sleeping bookkeeping run on the first (they only read what the solver left behind)
static int tick_end_impl(SfgWorld* w, float* poses_out) {
    if (!w) return SFG_E_INVALID;
    if (!w->in_tick) return fail(w, SFG_E_STATE, "sfg_tick_end without sfg_tick_begin");
    if (w->tick_next_substep != w->tick_substeps) return fail(w, SFG_E_STATE, "sfg_tick_end before all substeps ran");
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    w->in_tick = false;
    const bool sleeping = w->tick_sleeping;
    const uint32_t substeps = w->tick_substeps;
    int rc;
    CK(cudaEventRecord(w->ev[3], st));
    if (poses_out && w->n_bodies) {
        CK(w->stage2.ensure(size_t(w->n_bodies) * 16));
        CK(cudaStreamWaitEvent(w->stream2, w->ev[3], 0));
        k_get_poses<<<nblk(w->n_bodies), 256, 0, w->stream2>>>((float4*)w->stage2.p, w->n_bodies, 0, w->P.p, w->Q.p);
        CKL(w); w->launches++;
        CK(cudaMemcpyAsync(poses_out, w->stage2.p, size_t(w->n_bodies) * 16, cudaMemcpyDeviceToHost, w->stream2));
    }
    if (sleeping) { rc = island_post(w); if (rc != SFG_OK) return rc; }
    uint32_t* isl_counts = w->pin + 208;           // 8 words
    uint32_t* tick_counts = w->pin + 216;          // 4 words
    memset(isl_counts, 0, 12 * 4);
    if (sleeping && w->n_bodies) CK(cudaMemcpyAsync(isl_counts, w->isl_counters.p, 8 * 4, cudaMemcpyDeviceToHost, st));
    if (w->tick_counts.p) CK(cudaMemcpyAsync(tick_counts, w->tick_counts.p, 4 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(w->ev[4], st));
    CK(cudaStreamSynchronize(st));
    if (poses_out) CK(cudaStreamSynchronize(w->stream2));
    w->islands_valid = sleeping;
    if (sleeping) w->ct_n = isl_counts[4];
    w->ticked = true;
    w->last_substeps = substeps;
    // stats
    SfgStats& s = w->stats;
    memset(&s, 0, sizeof(s));
    s.n_bodies = w->n_bodies; s.n_colliders = w->n_live; s.n_pairs = w->n_units;
    s.n_constraint_entries = w->n_constraints + w->n_k_two;      // two-body constraints are stored on both bodies (physics.rs:519-532)
    s.n_contact_colours = w->n_unit_colours; s.n_constraint_colours = w->n_k_colours;
    s.n_rope_particles = w->n_particles; s.n_ropes = w->n_ropes; s.substeps = substeps;
    s.grid_levels = uint32_t(w->grid.n_levels); s.grid_cells = w->grid.n_cells; s.cell_size = w->grid.cell[0];
    s.kernel_launches = w->launches;
    s.n_islands = isl_counts[0]; s.n_sleeping_bodies = isl_counts[1]; s.n_contacts = sleeping ? isl_counts[4] : 0u;
    cudaEventElapsedTime(&s.ms_broadphase, w->ev[0], w->ev[1]);
    cudaEventElapsedTime(&s.ms_graph, w->ev[1], w->ev[2]);
    cudaEventElapsedTime(&s.ms_solver, w->ev[2], w->ev[3]);
    cudaEventElapsedTime(&s.ms_total, w->ev[0], w->ev[4]);
    s.n_stages = uint32_t(w->h_stages.size());
    s.n_contact_entries = w->n_units + tick_counts[0];
    s.n_touching_units = tick_counts[2]; s.n_touching_entries = tick_counts[3];
    return SFG_OK;
}

int sfg_tick(SfgWorld* w, double frame_dt, int has_time_scale, double time_scale, const SfgForceField* ff) {
    int rc = sfg_tick_begin(w, frame_dt, has_time_scale, time_scale, ff);
    if (rc != SFG_OK) return rc;
    rc = sfg_tick_substeps(w, w->tick_substeps);
    if (rc != SFG_OK) return rc;
    return sfg_tick_end(w);
}

// Game::physics_tick in one call (game.rs:297-303): poses in (sync_hecs_to_physics), tick, poses out (sync_physics_to_hecs).
// The upload is queued in front of the tick instead of being waited for, and the download overlaps the end of the tick.
int sfg_tick_poses(SfgWorld* w, const float* poses_in, double frame_dt, int has_time_scale, double time_scale, const SfgForceField* ff,
                   float* poses_out) {
    if (!w) return SFG_E_INVALID;
    if (poses_in && w->n_bodies) {
        cudaSetDevice(w->device);
        cudaStream_t st = w->stream;
        CK(w->stage.ensure(size_t(w->n_bodies) * 16));
        CK(cudaMemcpyAsync(w->stage.p, poses_in, size_t(w->n_bodies) * 16, cudaMemcpyHostToDevice, st));
        k_set_poses<<<nblk(w->n_bodies), 256, 0, st>>>((const float4*)w->stage.p, w->n_bodies, 0, w->P.p, w->Q.p);
        CKL(w);
    }
    int rc = sfg_tick_begin(w, frame_dt, has_time_scale, time_scale, ff);
    if (rc != SFG_OK) return rc;
    rc = sfg_tick_substeps(w, w->tick_substeps);
    if (rc != SFG_OK) return rc;
    return tick_end_impl(w, poses_out);
}

// ---- slab decomposition (north_star "spatial slabs with boundary-body halo exchange")
int sfg_slab_configure(SfgWorld* w, float x_lo, float x_hi, float halo) {
    if (!w) return SFG_E_INVALID;
    if (x_hi > x_lo && !(halo > 0.f)) return fail(w, SFG_E_INVALID, "slab halo must be > 0");
    cudaSetDevice(w->device);
    w->slab_on = x_hi > x_lo;
    w->slab_lo = x_lo; w->slab_hi = x_hi; w->slab_halo = halo;
    CK(w->slab_stamp.ensure(size_t(w->n_bodies) + 1));
    CK(cudaMemsetAsync(w->slab_stamp.p, 0, (size_t(w->n_bodies) + 1) * 4, w->stream));
    w->slab_epoch = 0;
    return SFG_OK;
}
// ---- island sharding (SURVEY.md §8e "independent islands of one world"): every rank holds the whole world and runs the
// broadphase and the island pass on all of it, but colours and solves only the islands it owns; after the tick the ranks
// exchange the bodies they solved (sfg_slab_pack mode 2 / sfg_slab_unpack), so all copies agree again. Islands share no
// body, pair or constraint, so the result is bit-identical to one GPU.
int sfg_island_shard_configure(SfgWorld* w, uint32_t rank, uint32_t n_ranks) {
    if (!w) return SFG_E_INVALID;
    if (n_ranks > 1 && rank >= n_ranks) return fail(w, SFG_E_INVALID, "island shard rank out of range");
    if (n_ranks > 1 && w->slab_on) return fail(w, SFG_E_STATE, "island sharding and slab decomposition are exclusive");
    w->shard_rank = n_ranks > 1 ? rank : 0; w->shard_n = n_ranks > 1 ? n_ranks : 1;
    return SFG_OK;
}
int sfg_slab_begin(SfgWorld* w) {
    if (!w) return SFG_E_INVALID;
    if (!w->slab_on) return fail(w, SFG_E_STATE, "sfg_slab_configure first");
    w->slab_epoch += 1;
    return SFG_OK;
}
int sfg_slab_pack(SfgWorld* w, int mode, int side, void* device_out, uint32_t capacity_records, uint32_t* count_out) {
    if (!w || !device_out || !count_out || side < 0 || side > 1 || mode < 0 || mode > 2) return SFG_E_INVALID;
    if (mode < 2 && !w->slab_on) return fail(w, SFG_E_STATE, "sfg_slab_configure first");
    if (mode == 2 && w->shard_n < 2) return fail(w, SFG_E_STATE, "sfg_island_shard_configure first");
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    *count_out = 0;
    HaloRec* out = (HaloRec*)device_out;
    CK(cudaMemsetAsync(out, 0, sizeof(HaloRec), st));
    if (w->n_bodies) {
        k_slab_pack<<<nblk(w->n_bodies), 256, 0, st>>>(w->n_bodies, mode, side, w->slab_lo, w->slab_hi, w->slab_halo, w->P.p, w->Q.p, w->V.p, w->MI.p, out,
                                                     capacity_records, &out[0].slot);
        CKL(w); w->launches++;
    }
    uint32_t n = 0;
    CK(cudaMemcpyAsync(&n, &out[0].slot, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *count_out = n;
    if (n > capacity_records) return fail(w, SFG_E_CAPACITY, "halo buffer too small for the boundary zone");
    return SFG_OK;
}
int sfg_slab_unpack(SfgWorld* w, const void* device_in, uint32_t capacity_records) {
    if (!w || !device_in) return SFG_E_INVALID;
    if (!w->slab_on && w->shard_n < 2) return fail(w, SFG_E_STATE, "sfg_slab_configure or sfg_island_shard_configure first");
    if (capacity_records == 0 || w->n_bodies == 0) return SFG_OK;
    cudaSetDevice(w->device);
    CK(w->slab_stamp.ensure(size_t(w->n_bodies) + 1));
    k_slab_unpack<<<nblk(capacity_records), 256, 0, w->stream>>>((const HaloRec*)device_in, capacity_records, w->n_bodies, w->P.p, w->Q.p, w->V.p,
                                                               w->slab_stamp.p, w->slab_epoch);
    CKL(w); w->launches++;
    CK(cudaStreamSynchronize(w->stream));
    return SFG_OK;
}

int sfg_slab_pack_async(SfgWorld* w, int mode, int side, void* device_out, uint32_t capacity_records) {
    if (!w || !device_out || side < 0 || side > 1 || mode < 0 || mode > 2) return SFG_E_INVALID;
    if (mode < 2 && !w->slab_on) return fail(w, SFG_E_STATE, "sfg_slab_configure first");
    if (mode == 2 && w->shard_n < 2) return fail(w, SFG_E_STATE, "sfg_island_shard_configure first");
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    HaloRec* out = (HaloRec*)device_out;
    CK(cudaMemsetAsync(out, 0, sizeof(HaloRec), st));
    if (w->n_bodies && mode == 1 && w->zone_valid) {
        const uint32_t threads = std::min<uint32_t>(capacity_records, SLAB_ZONE_CAP);
        k_slab_pack_list<<<nblk(threads), 256, 0, st>>>(w->zone_list.p + size_t(side) * SLAB_ZONE_CAP, w->zone_count.p + side, SLAB_ZONE_CAP, w->P.p, w->Q.p, w->V.p, out,
                                                       capacity_records, &out[0].slot);
        CKL(w); w->launches++;
    } else if (w->n_bodies) {
        k_slab_pack<<<nblk(w->n_bodies), 256, 0, st>>>(w->n_bodies, mode, side, w->slab_lo, w->slab_hi, w->slab_halo, w->P.p, w->Q.p, w->V.p, w->MI.p, out,
                                                     capacity_records, &out[0].slot);
        CKL(w); w->launches++;
    }
    return SFG_OK;
}
int sfg_slab_unpack_async(SfgWorld* w, const void* device_in, uint32_t capacity_records) {
    if (!w || !device_in) return SFG_E_INVALID;
    if (!w->slab_on && w->shard_n < 2) return fail(w, SFG_E_STATE, "sfg_slab_configure or sfg_island_shard_configure first");
    if (capacity_records == 0 || w->n_bodies == 0) return SFG_OK;
    cudaSetDevice(w->device);
    CK(w->slab_stamp.ensure(size_t(w->n_bodies) + 1));
    k_slab_unpack<<<nblk(capacity_records), 256, 0, w->stream>>>((const HaloRec*)device_in, capacity_records, w->n_bodies, w->P.p, w->Q.p, w->V.p,
                                                               w->slab_stamp.p, w->slab_epoch);
    CKL(w); w->launches++;
    return SFG_OK;
}
int sfg_stream_fence(SfgWorld* w, void* external_stream, int direction) {
    if (!w || direction < 0 || direction > 1) return SFG_E_INVALID;
    cudaSetDevice(w->device);
    cudaStream_t ext = (cudaStream_t)external_stream;
    if (!w->ev_fence) CK(cudaEventCreateWithFlags(&w->ev_fence, cudaEventDisableTiming));
    if (direction == 0) { CK(cudaEventRecord(w->ev_fence, w->stream)); CK(cudaStreamWaitEvent(ext, w->ev_fence, 0)); }
    else { CK(cudaEventRecord(w->ev_fence, ext)); CK(cudaStreamWaitEvent(w->stream, w->ev_fence, 0)); }
    return SFG_OK;
}

int sfg_get_stats(SfgWorld* w, SfgStats* out) {
    if (!w || !out) return SFG_E_INVALID;
    *out = w->stats;
    return SFG_OK;
}

int sfg_get_bodies(SfgWorld* w, uint32_t first, uint32_t count, SfgBody* out) {
    if (!w || (count && !out)) return SFG_E_INVALID;
    if (uint64_t(first) + count > w->n_bodies) return fail(w, SFG_E_INVALID, "body slot range out of bounds");
    if (count == 0) return SFG_OK;
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    CK(w->stage.ensure(size_t(count) * sizeof(SfgBody)));
    k_pack_bodies<<<nblk(count), 256, 0, st>>>((SfgBody*)w->stage.p, count, first, w->P.p, w->Q.p, w->V.p, w->MI.p);
    CKL(w);
    CK(cudaMemcpyAsync(out, w->stage.p, size_t(count) * sizeof(SfgBody), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SFG_OK;
}

int sfg_set_poses(SfgWorld* w, uint32_t first, uint32_t count, const float* poses) {
    if (!w || (count && !poses)) return SFG_E_INVALID;
    if (uint64_t(first) + count > w->n_bodies) return fail(w, SFG_E_INVALID, "body slot range out of bounds");
    if (count == 0) return SFG_OK;
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    CK(w->stage.ensure(size_t(count) * 16));
    CK(cudaMemcpyAsync(w->stage.p, poses, size_t(count) * 16, cudaMemcpyHostToDevice, st));
    k_set_poses<<<nblk(count), 256, 0, st>>>((const float4*)w->stage.p, count, first, w->P.p, w->Q.p);
    CKL(w);
    CK(cudaStreamSynchronize(st));
    return SFG_OK;
}
int sfg_get_poses(SfgWorld* w, uint32_t first, uint32_t count, float* poses) {
    if (!w || (count && !poses)) return SFG_E_INVALID;
    if (uint64_t(first) + count > w->n_bodies) return fail(w, SFG_E_INVALID, "body slot range out of bounds");
    if (count == 0) return SFG_OK;
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    CK(w->stage.ensure(size_t(count) * 16));
    k_get_poses<<<nblk(count), 256, 0, st>>>((float4*)w->stage.p, count, first, w->P.p, w->Q.p);
    CKL(w);
    CK(cudaMemcpyAsync(poses, w->stage.p, size_t(count) * 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SFG_OK;
}

// pose sync for a list of slots (HecsSyncManager syncs only the entities whose options ask for that direction)
int sfg_set_poses_indexed(SfgWorld* w, uint32_t n, const uint32_t* slots, const float* poses) {
    if (!w || (n && (!slots || !poses))) return SFG_E_INVALID;
    if (n == 0) return SFG_OK;
    for (uint32_t i = 0; i < n; ++i) if (slots[i] >= w->n_bodies) return fail(w, SFG_E_INVALID, "body slot out of range");
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    CK(w->stage.ensure(size_t(n) * 20));
    float4* d_pose = (float4*)w->stage.p; uint32_t* d_slot = (uint32_t*)(w->stage.p + size_t(n) * 16);
    CK(cudaMemcpyAsync(d_pose, poses, size_t(n) * 16, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_slot, slots, size_t(n) * 4, cudaMemcpyHostToDevice, st));
    k_set_poses_indexed<<<nblk(n), 256, 0, st>>>(d_pose, d_slot, n, w->P.p, w->Q.p);
    CKL(w);
    CK(cudaStreamSynchronize(st));
    return SFG_OK;
}
int sfg_get_poses_indexed(SfgWorld* w, uint32_t n, const uint32_t* slots, float* poses) {
    if (!w || (n && (!slots || !poses))) return SFG_E_INVALID;
    if (n == 0) return SFG_OK;
    for (uint32_t i = 0; i < n; ++i) if (slots[i] >= w->n_bodies) return fail(w, SFG_E_INVALID, "body slot out of range");
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    CK(w->stage.ensure(size_t(n) * 20));
    float4* d_pose = (float4*)w->stage.p; uint32_t* d_slot = (uint32_t*)(w->stage.p + size_t(n) * 16);
    CK(cudaMemcpyAsync(d_slot, slots, size_t(n) * 4, cudaMemcpyHostToDevice, st));
    k_get_poses_indexed<<<nblk(n), 256, 0, st>>>(d_pose, d_slot, n, w->P.p, w->Q.p);
    CKL(w);
    CK(cudaMemcpyAsync(poses, d_pose, size_t(n) * 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SFG_OK;
}

int sfg_get_contacts(SfgWorld* w, SfgContactInfo* out, size_t capacity, size_t* n_out) {
    if (!w || !n_out) return SFG_E_INVALID;
    *n_out = 0;
    if (!w->ticked) return SFG_OK;
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    if (w->islands_valid) {   // the tick already published the list (with the contacts of sleeping islands retained)
        *n_out = w->ct_n;
        const uint32_t got = uint32_t(std::min<size_t>(capacity, w->ct_n));
        if (!got || !out) return SFG_OK;
        static_assert(sizeof(SfgContactInfo) == 16, "SfgContactInfo is 4 words");
        CK(w->stage.ensure(size_t(got) * sizeof(SfgContactInfo)));
        k_contacts_export<<<nblk(got), 256, 0, st>>>(got, w->ct[w->ct_cur].p, (uint32_t*)w->stage.p);
        CKL(w);
        CK(cudaMemcpyAsync(out, w->stage.p, size_t(got) * sizeof(SfgContactInfo), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return SFG_OK;
    }
    if (w->n_units == 0) return SFG_OK;
    const uint32_t n_entries = w->n_units * 2;
    const uint32_t cap = uint32_t(std::min<size_t>(capacity, n_entries));
    CK(w->stage.ensure(size_t(cap) * sizeof(SfgContactInfo) + 16));
    uint32_t* cnt = (uint32_t*)w->stage.p;
    SfgContactInfo* buf = (SfgContactInfo*)(w->stage.p + 16);
    CK(cudaMemsetAsync(cnt, 0, 4, st));
    k_collect_contacts<<<nblk(n_entries), 256, 0, st>>>(n_entries, w->n_units, w->LN.p, w->S.p, buf, cap, cnt);
    CKL(w);
    uint32_t total = 0;
    CK(cudaMemcpyAsync(&total, cnt, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t got = std::min(total, cap);
    if (got && out) { CK(cudaMemcpyAsync(out, buf, size_t(got) * sizeof(SfgContactInfo), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); }
    *n_out = total;   // total found; only min(total, capacity) were written
    return SFG_OK;
}

static int fill_cast_ctx(SfgWorld* w, CastCtx& c) {
    if (!w->ticked) return fail(w, SFG_E_STATE, "queries use the spatial index of the last tick: call sfg_tick first");
    c.g = w->grid; c.SA = w->SA.p; c.SI = w->SI.p; c.start = w->cell_start.p;
    c.CS = w->CS.p; c.CL = w->CL.p; c.CI = w->CI.p; c.P = w->P.p; c.Q = w->Q.p;
    return SFG_OK;
}

int sfg_cast_batch(SfgWorld* w, uint32_t n, const SfgRay* rays, SfgCastHit* out) {
    if (!w || (n && (!rays || !out))) return SFG_E_INVALID;
    if (n == 0) return SFG_OK;
    cudaSetDevice(w->device);
    CastCtx c;
    int rc = fill_cast_ctx(w, c);
    if (rc != SFG_OK) return rc;
    cudaStream_t st = w->stream;
    if (w->n_live == 0) { for (uint32_t i = 0; i < n; ++i) { memset(&out[i], 0, sizeof(SfgCastHit)); out[i].collider = -1; } return SFG_OK; }
    CK(w->d_rays.ensure(n)); CK(w->d_hits.ensure(n));
    CK(cudaMemcpyAsync(w->d_rays.p, rays, size_t(n) * sizeof(SfgRay), cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(w->ev_cast[0], st));
    k_cast_batch<<<nblk(n, 128), 128, 0, st>>>(c, n, w->d_rays.p, w->d_hits.p);
    CKL(w);
    CK(cudaEventRecord(w->ev_cast[1], st));
    CK(cudaMemcpyAsync(out, w->d_hits.p, size_t(n) * sizeof(SfgCastHit), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&w->last_cast_ms, w->ev_cast[0], w->ev_cast[1]);
    return SFG_OK;
}
int sfg_get_cast_kernel_ms(SfgWorld* w, float* out_ms) {
    if (!w || !out_ms) return SFG_E_INVALID;
    *out_ms = w->last_cast_ms;
    return SFG_OK;
}

int sfg_query_point_batch(SfgWorld* w, uint32_t n, const float* points_xy, const uint32_t* domains, uint32_t max_hits, int32_t* out_colliders,
                          uint32_t* out_counts) {
    if (!w || (n && (!points_xy || !out_colliders || !out_counts)) || max_hits == 0) return SFG_E_INVALID;
    if (n == 0) return SFG_OK;
    cudaSetDevice(w->device);
    CastCtx c;
    int rc = fill_cast_ctx(w, c);
    if (rc != SFG_OK) return rc;
    cudaStream_t st = w->stream;
    if (w->n_live == 0) { for (uint32_t i = 0; i < n; ++i) out_counts[i] = 0; for (size_t i = 0; i < size_t(n) * max_hits; ++i) out_colliders[i] = -1; return SFG_OK; }
    const size_t b_pts = size_t(n) * 8, b_dom = size_t(n) * 4, b_out = size_t(n) * max_hits * 4, b_cnt = size_t(n) * 4;
    CK(w->stage.ensure(b_pts + b_dom + b_out + b_cnt + 64));
    unsigned char* p = w->stage.p;
    float2* d_pts = (float2*)p; uint32_t* d_dom = (uint32_t*)(p + b_pts); int* d_out = (int*)(p + b_pts + b_dom); uint32_t* d_cnt = (uint32_t*)(p + b_pts + b_dom + b_out);
    CK(cudaMemcpyAsync(d_pts, points_xy, b_pts, cudaMemcpyHostToDevice, st));
    if (domains) CK(cudaMemcpyAsync(d_dom, domains, b_dom, cudaMemcpyHostToDevice, st));
    k_query_point<<<nblk(n, 128), 128, 0, st>>>(c, n, d_pts, domains ? d_dom : nullptr, max_hits, d_out, d_cnt);
    CKL(w);
    CK(cudaMemcpyAsync(out_colliders, d_out, b_out, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_counts, d_cnt, b_cnt, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SFG_OK;
}

int sfg_query_shape_batch(SfgWorld* w, uint32_t n, const SfgShapeQuery* queries, uint32_t max_hits, int32_t* out_colliders, uint32_t* out_counts) {
    if (!w || (n && (!queries || !out_colliders || !out_counts)) || max_hits == 0) return SFG_E_INVALID;
    if (n == 0) return SFG_OK;
    static_assert(sizeof(SfgShapeQuery) == sizeof(ShapeQuery), "SfgShapeQuery and the device record must match");
    for (uint32_t i = 0; i < n; ++i) if (queries[i].kind > SFG_POLY_HEXAGON) return fail(w, SFG_E_INVALID, "unknown collider polygon kind");
    cudaSetDevice(w->device);
    CastCtx c;
    int rc = fill_cast_ctx(w, c);
    if (rc != SFG_OK) return rc;
    cudaStream_t st = w->stream;
    if (w->n_live == 0) { for (uint32_t i = 0; i < n; ++i) out_counts[i] = 0; for (size_t i = 0; i < size_t(n) * max_hits; ++i) out_colliders[i] = -1; return SFG_OK; }
    const size_t b_q = size_t(n) * sizeof(ShapeQuery), b_out = size_t(n) * max_hits * 4, b_cnt = size_t(n) * 4;
    CK(w->stage.ensure(b_q + b_out + b_cnt + 64));
    unsigned char* p = w->stage.p;
    ShapeQuery* d_q = (ShapeQuery*)p; int* d_out = (int*)(p + b_q); uint32_t* d_cnt = (uint32_t*)(p + b_q + b_out);
    CK(cudaMemcpyAsync(d_q, queries, b_q, cudaMemcpyHostToDevice, st));
    k_query_shape<<<nblk(n, 128), 128, 0, st>>>(c, n, d_q, max_hits, d_out, d_cnt);
    CKL(w);
    CK(cudaMemcpyAsync(out_colliders, d_out, b_out, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_counts, d_cnt, b_cnt, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SFG_OK;
}

// ---------------------------------------------------------------- parity hooks
int sfg_debug_get_pairs(SfgWorld* w, uint32_t* out_pairs, size_t capacity_pairs, size_t* n_out) {
    if (!w || !n_out) return SFG_E_INVALID;
    *n_out = w->n_units;
    if (!out_pairs || w->n_units == 0) return SFG_OK;
    cudaSetDevice(w->device);
    const size_t n = std::min<size_t>(capacity_pairs, w->n_units);
    CK(cudaMemcpyAsync(out_pairs, w->U_ids.p, n * sizeof(uint2), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    return SFG_OK;
}
// the `near` bit of each pair's colouring priority, in the order of sfg_debug_get_pairs (include/sfg_hash.h)
int sfg_debug_get_pair_hints(SfgWorld* w, uint8_t* out_hints, size_t capacity_pairs, size_t* n_out) {
    if (!w || !n_out) return SFG_E_INVALID;
    *n_out = w->n_units;
    if (!out_hints || w->n_units == 0) return SFG_OK;
    cudaSetDevice(w->device);
    const size_t n = std::min<size_t>(capacity_pairs, w->n_units);
    std::vector<uint32_t> prio(n);
    CK(cudaMemcpyAsync(prio.data(), w->U_prio.p, n * 4, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    for (size_t i = 0; i < n; ++i) out_hints[i] = uint8_t(prio[i] >> 31);
    return SFG_OK;
}
// pair units in the solver's colour-major order: hi, lo | mult2 << 31, colour
int sfg_debug_get_contact_entries(SfgWorld* w, uint32_t* out_hi, uint32_t* out_lo_copy, uint32_t* out_colour, size_t capacity, size_t* n_out) {
    if (!w || !n_out) return SFG_E_INVALID;
    *n_out = w->n_units;
    if (w->n_units == 0 || !out_hi || !out_lo_copy || !out_colour) return SFG_OK;
    if (capacity < w->n_units) return fail(w, SFG_E_CAPACITY, "capacity < number of pair units");
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    const uint32_t n = w->n_units;
    CK(w->stage.ensure(size_t(n) * 8));
    uint32_t* d_hi = (uint32_t*)w->stage.p; uint32_t* d_lo = d_hi + n;
    k_export_units<<<nblk(n), 256, 0, st>>>(n, w->H.p, w->S.p, d_hi, d_lo);
    CKL(w);
    CK(cudaMemcpyAsync(out_hi, d_hi, size_t(n) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_lo_copy, d_lo, size_t(n) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (uint32_t c = 0; c < w->n_unit_colours; ++c)
        for (uint32_t u = w->unit_colour_start[c]; u < w->unit_colour_start[c + 1]; ++u) out_colour[u] = c;
    return SFG_OK;
}
// constraint units in colour-major order: index | mult2 << 31, colour
int sfg_debug_get_constraint_entries(SfgWorld* w, uint32_t* out_index_copy, uint32_t* out_colour, size_t capacity, size_t* n_out) {
    if (!w || !n_out) return SFG_E_INVALID;
    *n_out = w->n_kunits;
    if (w->n_kunits == 0 || !out_index_copy || !out_colour) return SFG_OK;
    if (capacity < w->n_kunits) return fail(w, SFG_E_CAPACITY, "capacity < number of constraint units");
    cudaSetDevice(w->device);
    std::vector<float4> k0(w->n_kunits);
    CK(cudaMemcpyAsync(k0.data(), w->K0.p, size_t(w->n_kunits) * sizeof(float4), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    for (uint32_t u = 0; u < w->n_kunits; ++u) {
        uint32_t idx; int32_t target;
        memcpy(&idx, &k0[u].w, 4); memcpy(&target, &k0[u].y, 4);
        out_index_copy[u] = idx | (target >= 0 ? 0x80000000u : 0u);
    }
    for (uint32_t c = 0; c < w->n_k_colours; ++c)
        for (uint32_t u = w->k_colour_start[c]; u < w->k_colour_start[c + 1]; ++u) out_colour[u] = c;
    return SFG_OK;
}
// per contact entry (entry = copy * n_units + unit): hi, lo, count, lambda_n, normal.xy, 2 x (offsets[0].xy, offsets[1].xy) = 14 floats
int sfg_debug_get_manifolds(SfgWorld* w, float* out, size_t capacity_entries, size_t* n_out) {
    if (!w || !n_out) return SFG_E_INVALID;
    const uint32_t n_entries = w->n_units * 2;
    *n_out = n_entries;
    if (n_entries == 0 || !out) return SFG_OK;
    if (capacity_entries < n_entries) return fail(w, SFG_E_CAPACITY, "capacity < number of contact entries");
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    CK(w->stage.ensure(size_t(n_entries) * 14 * 4));
    k_export_manifolds<<<nblk(n_entries), 256, 0, st>>>(n_entries, w->n_units, ((w->tick_seq & 0xFFFu) << 12) + w->last_substeps, w->S.p, w->M0.p, w->M1.p, w->M2.p, (float*)w->stage.p);
    CKL(w);
    CK(cudaMemcpyAsync(out, w->stage.p, size_t(n_entries) * 14 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SFG_OK;
}
// islands of the last tick: first_body (= smallest slot), edge_sum, body count, flags (1 cannot sleep, 2 moved since the last
// tick, 4 fast after the solve, 8 asleep = skipped this tick)
int sfg_debug_get_islands(SfgWorld* w, uint32_t* first_body, uint64_t* edge_sum, uint32_t* body_count, uint32_t* flags, size_t capacity,
                          size_t* n_out) {
    if (!w || !n_out) return SFG_E_INVALID;
    *n_out = 0;
    if (!w->islands_valid || w->n_bodies == 0) return SFG_OK;
    cudaSetDevice(w->device);
    cudaStream_t st = w->stream;
    const uint32_t cap = uint32_t(std::min<size_t>(capacity, w->n_bodies));
    const size_t bytes = 16 + size_t(cap) * (4 + 8 + 4 + 4) + 64;
    CK(w->stage.ensure(bytes));
    uint32_t* cursor = (uint32_t*)w->stage.p;
    unsigned long long* d_sum = (unsigned long long*)(w->stage.p + 16);
    uint32_t* d_first = (uint32_t*)(d_sum + cap); uint32_t* d_cnt = d_first + cap; uint32_t* d_fl = d_cnt + cap;
    CK(cudaMemsetAsync(cursor, 0, 4, st));
    IslandCtx c;
    fill_island_ctx(w, c);
    k_isl_export<<<nblk(w->n_bodies), 256, 0, st>>>(c, cap, cursor, d_first, d_sum, d_cnt, d_fl);
    CKL(w);
    uint32_t total = 0;
    CK(cudaMemcpyAsync(&total, cursor, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_out = total;
    const uint32_t got = std::min(total, cap);
    if (got && first_body && edge_sum && body_count && flags) {
        CK(cudaMemcpyAsync(first_body, d_first, size_t(got) * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(edge_sum, d_sum, size_t(got) * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(body_count, d_cnt, size_t(got) * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(flags, d_fl, size_t(got) * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return SFG_OK;
}
int sfg_debug_get_aabbs(SfgWorld* w, float* out, size_t capacity_colliders) {
    if (!w || !out) return SFG_E_INVALID;
    const size_t n = std::min<size_t>(capacity_colliders, w->n_colls);
    if (n == 0) return SFG_OK;
    cudaSetDevice(w->device);
    CK(cudaMemcpyAsync(out, w->AABB.p, n * sizeof(float4), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    return SFG_OK;
}

static int dbg_run(int device, size_t in_a_bytes, const void* in_a, size_t in_b_bytes, const void* in_b, size_t in_c_bytes, const void* in_c,
                   size_t out_bytes, void* out, void (*launch)(void*, void*, void*, void*, void*), void* user) {
    int rc = ensure_device(device);
    if (rc != SFG_OK) return rc;
    void *a = nullptr, *b = nullptr, *c = nullptr, *o = nullptr;
    cudaError_t e = cudaSuccess;
    if (in_a_bytes) e = cudaMalloc(&a, in_a_bytes);
    if (e == cudaSuccess && in_b_bytes) e = cudaMalloc(&b, in_b_bytes);
    if (e == cudaSuccess && in_c_bytes) e = cudaMalloc(&c, in_c_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&o, out_bytes);
    if (e == cudaSuccess && in_a_bytes) e = cudaMemcpy(a, in_a, in_a_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && in_b_bytes) e = cudaMemcpy(b, in_b, in_b_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && in_c_bytes) e = cudaMemcpy(c, in_c, in_c_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { launch(a, b, c, o, user); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out, o, out_bytes, cudaMemcpyDeviceToHost);
    cudaFree(a); cudaFree(b); cudaFree(c); cudaFree(o);
    return e == cudaSuccess ? SFG_OK : SFG_E_CUDA;
}
// out rows of 14: count, 0, normal.xy, 2 x (offsets[0].xy, offsets[1].xy), pad
int sfg_debug_intersection_check(int device, uint32_t n, const float* poses, const float* shapes, float* out) {
    if (!poses || !shapes || !out) return SFG_E_INVALID;
    if (n == 0) return SFG_OK;
    uint32_t nn = n;
    return dbg_run(device, size_t(n) * 32, poses, size_t(n) * 32, shapes, 0, nullptr, size_t(n) * 56, out,
                   [](void* a, void* b, void*, void* o, void* u) { uint32_t n = *(uint32_t*)u; k_dbg_isect<<<nblk(n), 256>>>(n, (const float*)a, (const float*)b, (float*)o); }, &nn);
}
int sfg_debug_spherecast_collider(int device, uint32_t n, const SfgRay* rays, const float* poses, const float* shapes, float* out) {
    if (!rays || !poses || !shapes || !out) return SFG_E_INVALID;
    if (n == 0) return SFG_OK;
    uint32_t nn = n;
    return dbg_run(device, size_t(n) * sizeof(SfgRay), rays, size_t(n) * 16, poses, size_t(n) * 16, shapes, size_t(n) * 24, out,
                   [](void* a, void* b, void* c, void* o, void* u) { uint32_t n = *(uint32_t*)u; k_dbg_cast<<<nblk(n), 256>>>(n, (const SfgRay*)a, (const float*)b, (const float*)c, (float*)o); }, &nn);
}
int sfg_debug_geometry(int device, uint32_t op, uint32_t n, const float* args, const float* shapes, float* out) {
    if (!args || !shapes || !out || op > 3) return SFG_E_INVALID;
    if (n == 0) return SFG_OK;
    uint32_t u[2] = {n, op};
    return dbg_run(device, size_t(n) * 8, args, size_t(n) * 16, shapes, 0, nullptr, size_t(n) * 28, out,
                   [](void* a, void* b, void*, void* o, void* uu) { uint32_t* u = (uint32_t*)uu; k_dbg_geometry<<<nblk(u[0]), 256>>>(u[1], u[0], (const float*)a, (const float*)b, (float*)o); }, u);
}

}  // extern "C"

// Batched raycasts / spherecasts and point queries over the tick-time hash grid.
// Semantics = closed form of PhysicsWorld::spherecast (physics.rs:1266-1311):
// min t over SOLID colliders whose tick-time AABB, padded by the radius, is entered at
// t_c < max_distance (strict) and whose exact hit has t <= max_distance; rays that start inside a
// shape miss it; ties keep the lowest collider slot. Candidate boxes are the ones stored at tick
// time (stale, velocity-extended, padded) while the exact test uses the current poses, as the
// reference does with the BVH left over from the last tick.
#pragma once
#include "../../include/starframe_gpu.h"
#include "sfg_geom.cuh"
#include "sfg_state.cuh"

namespace sfg {

struct CastCtx {
    GridParams g;
    const float4* SA; const int4* SI; const uint32_t* start;   // sorted boxes / (slot, body, layer|level<<8, domain) / cell starts
    const float4 *CS, *CL; const int4* CI;
    const float4 *P, *Q;
};

SFG_DI Iso collider_pose(const CastCtx& c, int slot, int body) {
    const float4 l = c.CL[slot];
    Iso loc = mkiso(mk(l.x, l.y), mkrot(l.z, l.w));
    if (body < 0) return loc;
    const float4 p = c.P[body], q = c.Q[body];
    return mkiso(mk(p.x + p.z, p.y + p.w), mkrot(q.x, q.y)) * loc;
}

SFG_DI void cast_try(const CastCtx& c, uint32_t cj, V2 start, V2 dir, float radius, float max_d, uint32_t domain,
                     int* best_c, Hit* best) {
    const int4 J = c.SI[cj];
    if (uint32_t(J.w) != domain) return;
    float4 box = c.SA[cj];
    box.x -= radius; box.y -= radius; box.z += radius; box.w += radius;
    float tc;
    if (!ray_aabb(start, dir, box, &tc)) return;
    if (tc >= max_d) return;
    const int4 ci = c.CI[J.x];
    if (uint32_t(ci.w) & CF_SENSOR) return;
    Hit h = spherecast_collider(start, dir, radius, collider_pose(c, J.x, J.y), mkshape(c.CS[J.x]));
    if (!h.hit || !(h.t <= max_d)) return;
    if (best->hit && (best->t < h.t || (best->t == h.t && *best_c < J.x))) return;
    *best = h; *best_c = J.x;
}

__global__ void __launch_bounds__(128) k_cast_batch(const __grid_constant__ CastCtx c, uint32_t n, const SfgRay* __restrict__ rays,
                                                    SfgCastHit* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SfgRay r = rays[i];
    const V2 start = mk(r.sx, r.sy), dir = mk(r.dx, r.dy);
    Hit best; best.hit = false; best.t = 0.f; best.n = mk(0.f, 0.f); best.p = mk(0.f, 0.f);
    int best_c = -1;
    const GridParams& g = c.g;
    const float len = fminf(r.max_distance, 4.0e6f);
    if (r.domain < g.n_domains) {
        for (int l = 0; l < g.n_levels; ++l) {
            // colliders of this level reach at most cell/2 from their centre cell's far edge => a ray
            // point p can only touch boxes whose centre cell is within +-1 of p's cell after padding by
            // the radius; walk the major axis cell by cell and cover the minor-axis span of each step.
            const float cs = g.cell[l], inv = g.inv_cell[l];
            const float margin = 0.5f * cs + r.radius;
            const bool xmajor = fabsf(dir.x) >= fabsf(dir.y);
            const float s_maj = xmajor ? start.x - g.ox : start.y - g.oy, d_maj = xmajor ? dir.x : dir.y;
            const float s_min = xmajor ? start.y - g.oy : start.x - g.ox, d_min = xmajor ? dir.y : dir.x;
            const int n_maj = xmajor ? g.nx[l] : g.ny[l], n_min = xmajor ? g.ny[l] : g.nx[l];
            const float e_maj = s_maj + len * d_maj;
            const float lo_maj = fminf(s_maj, e_maj) - margin, hi_maj = fmaxf(s_maj, e_maj) + margin;
            int c0 = max(int(floorf(lo_maj * inv)), 0), c1 = min(int(floorf(hi_maj * inv)), n_maj - 1);
            const int step = d_maj >= 0.f ? 1 : -1;
            if (step < 0) { int t = c0; c0 = c1; c1 = t; }
            for (int cm = c0; step > 0 ? cm <= c1 : cm >= c1; cm += step) {
                // slab of this column expanded by the margin; ray parameter range inside it
                const float slab_lo = cm * cs - margin, slab_hi = (cm + 1) * cs + margin;
                float t0 = 0.f, t1 = len;
                if (fabsf(d_maj) > 1e-12f) {
                    float ta = (slab_lo - s_maj) / d_maj, tb = (slab_hi - s_maj) / d_maj;
                    if (ta > tb) { float q = ta; ta = tb; tb = q; }
                    t0 = fmaxf(t0, ta); t1 = fminf(t1, tb);
                    if (t0 > t1) continue;
                }
                if (best.hit && t0 > best.t + margin) break;   // everything further along enters later than the best hit
                const float m0 = s_min + t0 * d_min, m1 = s_min + t1 * d_min;
                const int r0 = max(int(floorf((fminf(m0, m1) - margin) * inv)), 0), r1 = min(int(floorf((fmaxf(m0, m1) + margin) * inv)), n_min - 1);
                if (xmajor) {
                    for (int y = r0; y <= r1; ++y) {
                        const uint32_t cell = g.base[l] + (r.domain * uint32_t(g.ny[l]) + uint32_t(y)) * uint32_t(g.nx[l]) + uint32_t(cm);
                        for (uint32_t cj = c.start[cell]; cj < c.start[cell + 1]; ++cj) cast_try(c, cj, start, dir, r.radius, r.max_distance, r.domain, &best_c, &best);
                    }
                } else if (r0 <= r1) {
                    const uint32_t row = g.base[l] + (r.domain * uint32_t(g.ny[l]) + uint32_t(cm)) * uint32_t(g.nx[l]);
                    for (uint32_t cj = c.start[row + r0]; cj < c.start[row + r1 + 1]; ++cj) cast_try(c, cj, start, dir, r.radius, r.max_distance, r.domain, &best_c, &best);
                }
            }
        }
        for (uint32_t cj = c.start[g.n_cells]; cj < c.start[g.n_cells + 1]; ++cj) cast_try(c, cj, start, dir, r.radius, r.max_distance, r.domain, &best_c, &best);
    }
    SfgCastHit o;
    o.collider = best_c; o.t = best.t; o.px = best.p.x; o.py = best.p.y; o.nx = best.n.x; o.ny = best.n.y;
    o.reserved[0] = o.reserved[1] = 0;
    out[i] = o;
}

// physics.rs:1188-1210 — tick-time box contains the point (inclusive), then point_collider_bool on current poses
__global__ void __launch_bounds__(128) k_query_point(const __grid_constant__ CastCtx c, uint32_t n, const float2* __restrict__ pts,
                                                     const uint32_t* __restrict__ domains, uint32_t max_hits, int* __restrict__ out,
                                                     uint32_t* __restrict__ counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V2 p = mk(pts[i].x, pts[i].y);
    const uint32_t domain = domains ? domains[i] : 0u;
    const GridParams& g = c.g;
    uint32_t cnt = 0;
    auto visit = [&](uint32_t cj) {
        const int4 J = c.SI[cj];
        if (uint32_t(J.w) != domain) return;
        const float4 b = c.SA[cj];
        if (!(p.x >= b.x && p.x <= b.z && p.y >= b.y && p.y <= b.w)) return;
        if (!point_in_collider(p, collider_pose(c, J.x, J.y), mkshape(c.CS[J.x]))) return;
        if (cnt < max_hits) out[size_t(i) * max_hits + cnt] = J.x;
        ++cnt;
    };
    if (domain < g.n_domains) {
        for (int l = 0; l < g.n_levels; ++l) {
            int ix, iy;
            cell_of(g, l, p.x, p.y, &ix, &iy);
            for (int y = max(iy - 1, 0); y <= min(iy + 1, g.ny[l] - 1); ++y) {
                const uint32_t row = g.base[l] + (domain * uint32_t(g.ny[l]) + uint32_t(y)) * uint32_t(g.nx[l]);
                for (uint32_t cj = c.start[row + max(ix - 1, 0)]; cj < c.start[row + min(ix + 1, g.nx[l] - 1) + 1]; ++cj) visit(cj);
            }
        }
        for (uint32_t cj = c.start[g.n_cells]; cj < c.start[g.n_cells + 1]; ++cj) visit(cj);
    }
    for (uint32_t k = cnt; k < max_hits; ++k) out[size_t(i) * max_hits + k] = -1;
    counts[i] = cnt;
}

// physics.rs:1215-1245 query_shape: colliders whose tick-time box strictly overlaps the shape's box (bvh.rs:292-336, the same
// iterator the pair search uses), whose layer is in the caller's mask, and for which intersection_check on the CURRENT
// poses is non-zero. One thread per query; the query box may cover many cells.
struct ShapeQuery { float x, y, rot_s, rot_b; float p0, p1, circle_r; uint32_t kind; unsigned long long layer_mask; uint32_t domain, reserved; };
__global__ void __launch_bounds__(128) k_query_shape(const __grid_constant__ CastCtx c, uint32_t n, const ShapeQuery* __restrict__ qs, uint32_t max_hits,
                                                     int* __restrict__ out, uint32_t* __restrict__ counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ShapeQuery q = qs[i];
    Shape qs_shape; qs_shape.kind = q.kind; qs_shape.p0 = q.p0; qs_shape.p1 = q.p1; qs_shape.r = q.circle_r;
    const Iso qpose = mkiso(mk(q.x, q.y), mkrot(q.rot_s, q.rot_b));
    const float4 box = shape_aabb_exact(qs_shape, qpose.t, qpose.r);
    const GridParams& g = c.g;
    uint32_t cnt = 0;
    auto visit = [&](uint32_t cj) {
        const int4 J = c.SI[cj];
        if (uint32_t(J.w) != q.domain) return;
        if (!aabb_overlap(box, c.SA[cj])) return;
        if (!((q.layer_mask >> (J.z & 0xFF)) & 1ull)) return;
        Manifold m;
        intersection_check(qpose, collider_pose(c, J.x, J.y), qs_shape, mkshape(c.CS[J.x]), m);
        if (m.count == 0) return;
        if (cnt < max_hits) out[size_t(i) * max_hits + cnt] = J.x;
        ++cnt;
    };
    if (q.domain < g.n_domains && isfinite(box.x) && isfinite(box.y) && isfinite(box.z) && isfinite(box.w)) {
        for (int l = 0; l < g.n_levels; ++l) {
            int ix0, iy0, ix1, iy1;
            cell_of(g, l, box.x, box.y, &ix0, &iy0);
            cell_of(g, l, box.z, box.w, &ix1, &iy1);
            for (int y = max(iy0 - 1, 0); y <= min(iy1 + 1, g.ny[l] - 1); ++y) {
                const uint32_t row = g.base[l] + (q.domain * uint32_t(g.ny[l]) + uint32_t(y)) * uint32_t(g.nx[l]);
                for (uint32_t cj = c.start[row + max(ix0 - 1, 0)]; cj < c.start[row + min(ix1 + 1, g.nx[l] - 1) + 1]; ++cj) visit(cj);
            }
        }
        for (uint32_t cj = c.start[g.n_cells]; cj < c.start[g.n_cells + 1]; ++cj) visit(cj);
    }
    for (uint32_t k = cnt; k < max_hits; ++k) out[size_t(i) * max_hits + k] = -1;
    counts[i] = cnt;
}

}  // namespace sfg

// XPBD substep stages (one thread per body / per constraint unit of a colour) and the two ways of
// driving them: one launch per stage, or a single persistent cooperative kernel with grid-wide
// barriers between stages. Restates /root/reference/src/physics/solver.rs stage by stage; the
// Gauss-Seidel order is colour-major (include/sfg_hash.h) instead of the reference's DFS order.
#pragma once
#include <cooperative_groups.h>
#include "sfg_geom.cuh"
#include "sfg_state.cuh"

namespace sfg {
namespace cg = cooperative_groups;

// Body state is rewritten by other SMs between stages of the persistent kernel, so it always goes
// through L2 (ld.cg / st.cg); per-tick constant streams use the read-only path.
#ifndef SFG_L2_HINTS
#define SFG_L2_HINTS 1
#endif
// Body rows: every colour gathers them again and other SMs write them, so they go through L2 (.cg) with the default policy.
SFG_DI float4 ldb(const float4* p) { return __ldcg(p); }
SFG_DI float2 ldb(const float2* p) { return __ldcg(p); }
SFG_DI void stb(float4* p, float4 v) { __stcg(p, v); }
SFG_DI void stb(float2* p, float2 v) { __stcg(p, v); }
SFG_DI float ldc(const float* p) { return __ldg(p); }
SFG_DI uint32_t ldw(const uint32_t* p) { return __ldcg(p); }     // worklist slots
SFG_DI void stw(uint32_t* p, uint32_t v) { __stcg(p, v); }
#if SFG_L2_HINTS
// L2 residency hints: the unit columns (ldc) and the manifold entries (ldm / stm) stream past once per sweep — 0.4 GB per colour
// sweep of the settled pile against 126 MB of L2 — and are marked evict_first, so that the body rows stay resident between
// colours: settled C3 solver 3.58 -> 3.45 ms, C5 7.10 -> 6.88, landing 0.686 -> 0.664 in a same-box A/B. Marking the body rows
// evict_last on top, or hinting the 4-byte columns (reach, worklists) as well, changed nothing measurable.
SFG_DI unsigned long long l2_evict_first() { unsigned long long p; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
SFG_DI float4 ldc(const float4* p) {
    float4 v; asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(l2_evict_first())); return v;
}
SFG_DI float4 ldm(const float4* p) {
    float4 v; asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(l2_evict_first())); return v;
}
SFG_DI void stm(float4* p, float4 v) {
    asm volatile("st.global.cg.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(l2_evict_first()) : "memory");
}
SFG_DI void stm(float2* p, float2 v) {
    asm volatile("st.global.cg.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(l2_evict_first()) : "memory");
}
#else
SFG_DI float4 ldc(const float4* p) { return __ldg(p); }
SFG_DI float4 ldm(const float4* p) { return __ldcg(p); }
SFG_DI void stm(float4* p, float4 v) { __stcg(p, v); }
SFG_DI void stm(float2* p, float2 v) { __stcg(p, v); }
#endif

// forcefield.rs (out of line: one copy for the three places that integrate a body)
__device__ __noinline__ V2 force_at(const ForceParams& ff, V2 pos) {
    V2 acc = mk(0.f, 0.f);
    for (uint32_t i = 0; i < ff.n; ++i) {
        if (ff.kind[i] == 1u) acc = acc + mk(ff.a[i], ff.b[i]);
        else if (ff.kind[i] == 2u) {
            V2 dist = mk(ff.a[i], ff.b[i]) - pos;
            float strength = ff.c[i] / ((mag_sq(dist) + 1.0f) * ff.d[i]);
            acc = acc + strength * normalized(dist);
        }
    }
    return acc;
}

// solver.rs:57-74 + physics.rs:78-83. The columns are passed in so that the persistent kernel can fetch several bodies'
// worth of them before the first is used (a streaming stage is otherwise a chain of dependent round trips per warp).
SFG_DI void integrate_body(const SolverCtx& c, uint32_t i, float4 mi, float4 p, float4 q, float4 v) {
    const uint32_t fl = __float_as_uint(mi.z);
    if ((fl & (BF_ALIVE | BF_ASLEEP)) != BF_ALIVE) return;   // sleeping islands are skipped entirely (physics.rs:798-803)
    V2 pos = mk(p.x + p.z, p.y + p.w);
    if (!(fl & BF_NOGRAV) && (fl & BF_MASS)) {
        V2 a = force_at(c.ff, pos);
        v.x += a.x * c.dt; v.y += a.y * c.dt;
        if (c.bouncy) stb(c.EA + i, make_float2(a.x, a.y));
    }
    if (c.bouncy) stb(c.OV + i, v);                          // restitution is the only reader (stage_contact_vel)
    stb(c.V + i, v);
    Rot r = rot_from_angle(v.z * c.dt) * mkrot(q.x, q.y);
    stb(c.P + i, make_float4(pos.x, pos.y, v.x * c.dt, v.y * c.dt));
    stb(c.Q + i, make_float4(r.s, r.b, q.x, q.y));
}
SFG_DI void stage_integrate(const SolverCtx& c, uint32_t i) { integrate_body(c, i, ldb(c.MI + i), ldb(c.P + i), ldb(c.Q + i), ldb(c.V + i)); }

// solver.rs:91-97 (+ fold of the displacement into the position after the last substep)
SFG_DI void derive_body(const SolverCtx& c, uint32_t i, bool fold, float4 mi, float4 p, float4 q, float4 v) {
    if ((__float_as_uint(mi.z) & (BF_ALIVE | BF_ASLEEP)) != BF_ALIVE) return;
    v.x = p.z * c.inv_dt; v.y = p.w * c.inv_dt;
    Rot d = mkrot(q.x, q.y) * mkrot(q.z, -q.w);
    v.z = (-atan2f(d.b, d.s) * 2.0f) * c.inv_dt;
    stb(c.V + i, v);
    if (fold) stb(c.P + i, make_float4(p.x + p.z, p.y + p.w, 0.f, 0.f));
}
SFG_DI void stage_derive(const SolverCtx& c, uint32_t i, bool fold) { derive_body(c, i, fold, ldb(c.MI + i), ldb(c.P + i), ldb(c.Q + i), ldb(c.V + i)); }

// ---------------------------------------------------------------- ropes (solver.rs:114-181)
SFG_DI void stage_rope_step(const SolverCtx& c, uint32_t u) {
    const int4 ru = c.RU[u];
    const float4 pa = ldc(c.RPA + ru.w);
    float4 pc = ldb(c.P + ru.x), pn = ldb(c.P + ru.y);
    const float imc = ldb(c.MI + ru.x).x, imn = ldb(c.MI + ru.y).x;
    V2 dist = mk((pn.x - pc.x) + (pn.z - pc.z), (pn.y - pc.y) + (pn.w - pc.w));
    float dm = mag(dist);
    if (SFG_GUARD(3) && dm == 0.f) return;          // coincident particles: no direction to correct along (the reference gets NaN, sfg_geom.cuh circle_any)
    V2 dir = mk(dist.x / dm, dist.y / dm);
    float error = pa.x - dm;
    float lambda = -error / (imc + imn + pa.y * c.inv_dt_sq);
    pc.z += imc * lambda * dir.x; pc.w += imc * lambda * dir.y;
    pn.z += -imn * lambda * dir.x; pn.w += -imn * lambda * dir.y;
    stb(c.P + ru.x, pc);
    stb(c.P + ru.y, pn);
    if (ru.z < 0) return;
    float4 pf = ldb(c.P + ru.z);
    const float ima = ldb(c.MI + ru.z).x;
    V2 c2n = mk((pn.x - pc.x) + (pn.z - pc.z), (pn.y - pc.y) + (pn.w - pc.w));
    V2 n2a = mk((pf.x - pn.x) + (pf.z - pn.z), (pf.y - pn.y) + (pf.w - pn.w));
    float angle = acosf(dot(normalized(n2a), normalized(c2n)));
    float err = angle - pa.z;
    if (err > 0.f) {
        float lam = -err / (ima + pa.w * c.inv_dt_sq);
        float lo = dot(left_normal(c2n), n2a) > 0.f ? lam : -lam;
        Rot corr = rot_from_angle(lo * ima);
        V2 na = corr * n2a;
        // new position of `after`, expressed as a displacement from its own substep-start position
        float ndx = (pn.x - pf.x) + pn.z + na.x, ndy = (pn.y - pf.y) + pn.w + na.y;
        stb(c.LAT + ru.z, make_float2(ndx - pf.z, ndy - pf.w));
        pf.z = ndx; pf.w = ndy;
        stb(c.P + ru.z, pf);
    }
}
// solver.rs:722-740
SFG_DI void stage_rope_damp(const SolverCtx& c, uint32_t u) {
    const int4 rd = c.RD[u];
    const float damping = ldc(c.RPB + rd.z).x;
    float4 vc = ldb(c.V + rd.x), vn = ldb(c.V + rd.y);
    const float imc = ldb(c.MI + rd.x).x, imn = ldb(c.MI + rd.y).x;
    V2 rel = mk(vc.x - vn.x, vc.y - vn.y);
    float m = mag(rel);
    if (m != 0.f) {
        V2 dir = mk(rel.x / m, rel.y / m);
        float upd = -m * fminf(damping * c.dt, 1.0f);
        float imp = upd / (imc + imn);
        vc.x += imc * imp * dir.x; vc.y += imc * imp * dir.y;
        vn.x -= imn * imp * dir.x; vn.y -= imn * imp * dir.y;
        stb(c.V + rd.x, vc);
        stb(c.V + rd.y, vn);
    }
}
// solver.rs:750-765, one thread per rope particle
SFG_DI void stage_rope_lateral(const SolverCtx& c, uint32_t p) {
    if (c.rp_rope[p] < 0) return;              // a hole in the particle array (left by a topology edit, sfg_api.cu sfg_rope_*)
    const int b = c.rp_body[p];
    float2 corr = ldb(c.LAT + b);
    if (isnan(corr.x)) return;
    const float bend_c = ldc(c.RPA + c.rp_rope[p]).w;
    float4 v = ldb(c.V + b);
    const float im = ldb(c.MI + b).x;
    float cm = sqrtf(corr.x * corr.x + corr.y * corr.y);
    // Deviation (f32): the reference divides by the correction's length unguarded (solver.rs:752-756). In f64 a bending
    // correction is never exactly zero; in f32 a tiny rotation of the last segment can leave the particle's displacement
    // bit-identical, and 0 / 0 would poison the whole pile within the tick. A zero correction created no velocity, so the
    // clamp below is [0, 0] and the impulse is zero: skipping is the limit of the reference's formula.
    if (SFG_GUARD(0) && cm == 0.f) return;
    float vfc = cm * c.inv_dt;
    V2 dir = mk(corr.x / cm, corr.y / cm);
    float vid = v.x * dir.x + v.y * dir.y;
    float vc = fmaxf(fminf(vid, vfc), -vfc);
    float imp = -vc / (im + bend_c * c.inv_dt);
    v.x += im * imp * dir.x; v.y += im * imp * dir.y;
    stb(c.V + b, v);
}
// pre_contact_poses (solver.rs:83-85), only rope particles ever read them (solver.rs:337-362)
SFG_DI void stage_precontact(const SolverCtx& c, uint32_t p) {
    if (c.rp_rope[p] < 0) return;
    const int b = c.rp_body[p];
    float4 pp = ldb(c.P + b);
    stb(c.PCD + b, make_float2(pp.z, pp.w));
}

// ---------------------------------------------------------------- helpers
SFG_DI void apply_correction(float4& P, Rot& r, float im, float ii, float sgn, float lambda, V2 dir, float wdg) {
    const float k = sgn * im * lambda;
    P.z += k * dir.x; P.w += k * dir.y;
    if (ii != 0.f) r = rot_from_angle(sgn * ii * lambda * wdg) * r;
}

// ---------------------------------------------------------------- distance constraints (solver.rs:187-269)
SFG_DI void stage_constraints(const SolverCtx& c, uint32_t u) {
    const float4 k0 = ldc(c.K0 + u), k1 = ldc(c.K1 + u), k2 = ldc(c.K2 + u);
    const int owner = __float_as_int(k0.x), target = __float_as_int(k0.y);
    const uint32_t fl = __float_as_uint(k0.z);
    const uint32_t limit = fl & 3u;
    const int mult = target >= 0 ? 2 : 1;
    float4 P0 = ldb(c.P + owner), Q0 = ldb(c.Q + owner);
    float4 mi0 = ldb(c.MI + owner);
    if (__float_as_uint(mi0.z) & BF_ASLEEP) return;
    float4 P1 = make_float4(0, 0, 0, 0), Q1 = make_float4(1, 0, 1, 0), mi1 = make_float4(0, 0, 0, 0);
    if (target >= 0) { P1 = ldb(c.P + target); Q1 = ldb(c.Q + target); mi1 = ldb(c.MI + target); }
    Rot r0 = mkrot(Q0.x, Q0.y), r1 = mkrot(Q1.x, Q1.y);
    const V2 o0 = mk(k2.x, k2.y), o1 = mk(k2.z, k2.w);
    for (int m = 0; m < mult; ++m) {
        // world offsets relative to the owner's substep-start position
        V2 w0 = r0 * o0 + mk(P0.z, P0.w);
        V2 w1 = target >= 0 ? r1 * o1 + mk((P1.x - P0.x) + P1.z, (P1.y - P0.y) + P1.w) : mk(o1.x - P0.x, o1.y - P0.y);
        V2 actual = w1 - w0;
        float am = mag(actual);
        float error = k1.w - am;
        bool act = limit == 0u || (limit == 1u && error < 0.f) || (limit == 2u && error > 0.f);
        if (!act) continue;
        V2 dir = am != 0.f ? mk(actual.x / am, actual.y / am) : mk(0.f, 1.f);
        float wd0 = wedge(r0 * o0, dir);
        float wd1 = target >= 0 ? wedge(r1 * o1, dir) : 0.f;
        float e0 = mi0.x + (wd0 * wd0 * mi0.y), e1 = target >= 0 ? mi1.x + (wd1 * wd1 * mi1.y) : 0.f;
        float lambda = -error / (e0 + e1 + k1.x * c.inv_dt_sq);
        apply_correction(P0, r0, mi0.x, mi0.y, 1.f, lambda, dir, wd0);
        if (target >= 0) apply_correction(P1, r1, mi1.x, mi1.y, -1.f, lambda, dir, wd1);
    }
    if (mi0.x != 0.f || mi0.y != 0.f) { stb(c.P + owner, P0); Q0.x = r0.s; Q0.y = r0.b; stb(c.Q + owner, Q0); }
    if (target >= 0 && (mi1.x != 0.f || mi1.y != 0.f)) { stb(c.P + target, P1); Q1.x = r1.s; Q1.y = r1.b; stb(c.Q + target, Q1); }
}

// solver.rs:624-709
SFG_DI void stage_constraint_damp(const SolverCtx& c, uint32_t u) {
    const float4 k0 = ldc(c.K0 + u), k1 = ldc(c.K1 + u), k2 = ldc(c.K2 + u);
    const int owner = __float_as_int(k0.x), target = __float_as_int(k0.y);
    const int mult = target >= 0 ? 2 : 1;
    float4 Q0 = ldb(c.Q + owner), V0 = ldb(c.V + owner), mi0 = ldb(c.MI + owner);
    if (__float_as_uint(mi0.z) & BF_ASLEEP) return;
    float4 Q1 = make_float4(1, 0, 1, 0), V1 = make_float4(0, 0, 0, 0), mi1 = make_float4(0, 0, 0, 0);
    if (target >= 0) { Q1 = ldb(c.Q + target); V1 = ldb(c.V + target); mi1 = ldb(c.MI + target); }
    const V2 or0 = mkrot(Q0.x, Q0.y) * mk(k2.x, k2.y);
    const V2 or1 = target >= 0 ? mkrot(Q1.x, Q1.y) * mk(k2.z, k2.w) : mk(0.f, 0.f);
    const float lin = fminf(k1.y * c.dt, 1.0f), ang = fminf(k1.z * c.dt, 1.0f);
    for (int m = 0; m < mult; ++m) {
        V2 pv0 = mk(V0.x - or0.y * V0.z, V0.y + or0.x * V0.z);
        if (target >= 0) {
            V2 pv1 = mk(V1.x - or1.y * V1.z, V1.y + or1.x * V1.z);
            V2 rel = pv0 - pv1;
            float rm = mag(rel);
            if (rm == 0.f) continue;
            V2 dir = mk(rel.x / rm, rel.y / rm);
            float wd0 = wedge(or0, dir), wd1 = wedge(or1, dir);
            float e0 = mi0.x + (wd0 * wd0 * mi0.y), e1 = mi1.x + (wd1 * wd1 * mi1.y);
            float imp = (-rm * lin) / (e0 + e1);
            V0.x += mi0.x * imp * dir.x; V0.y += mi0.x * imp * dir.y; V0.z += mi0.y * imp * wd0;
            V1.x -= mi1.x * imp * dir.x; V1.y -= mi1.x * imp * dir.y; V1.z -= mi1.y * imp * wd1;
            if (k1.z > 0.f) {
                float relw = V0.z - V1.z;
                float ai = (-relw * ang) / (e0 + e1);
                V1.z -= mi1.y * ai;
                V0.z += mi0.y * ai;
            }
        } else {
            float pm = mag(pv0);
            if (pm == 0.f) continue;
            V2 dir = mk(pv0.x / pm, pv0.y / pm);
            float wd = wedge(or0, dir);
            float e = mi0.x + wd * wd * mi0.y;
            float imp = (-pm * lin) / e;
            V0.x += mi0.x * imp * dir.x; V0.y += mi0.x * imp * dir.y; V0.z += mi0.y * imp * wd;
            if (k1.z > 0.f) {
                float au = V0.z * ang;
                float ai = -au / e;
                V0.z += mi0.y * ai;
            }
        }
    }
    if (mi0.x != 0.f || mi0.y != 0.f) stb(c.V + owner, V0);
    if (target >= 0 && (mi1.x != 0.f || mi1.y != 0.f)) stb(c.V + target, V1);
}

// Debug build only (-DSFG_PROBE): cycle stamps along the contact path of units that found a contact, to see where a lone
// unit's latency goes. The stamp's predicate depends on a loaded / computed value, so it cannot issue before that value exists.
#ifdef SFG_PROBE
__device__ unsigned long long g_probe[1 + 8 * 65536];
#define SFG_STAMP(i, bits) asm volatile("{ .reg .pred p; setp.ne.u32 p, %1, 0x7fc12345; @p mov.u64 %0, %%clock64; }" : "+l"(pt[i]) : "r"(uint32_t(bits)) : "memory")
#else
#define SFG_STAMP(i, bits)
#endif

// ---------------------------------------------------------------- contacts (solver.rs:275-490, narrowphase inline)
SFG_DI void stage_contacts(const SolverCtx& c, uint32_t u, uint32_t colour, uint32_t colour_begin, uint32_t sub, bool head) {
#ifdef SFG_PROBE
    unsigned long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    SFG_STAMP(0, u);
#endif
    const float4 h = ldc(c.H + u);
    // First reject, on body positions alone: colliders sit within |local offset| of their body, so bodies farther apart
    // than RC = bounding radii + offsets (+ margin, and never below circle_circle's sqrt(0.001) rule) cannot produce a
    // contact in any branch of intersection_check. Neither visit of the unit moves anything then, and the manifold
    // entries keep an old tag = "no contact".
    // A unit's latency is a chain of memory round trips followed by the arithmetic, so the loads go out in as few batches
    // as their addresses allow: everything indexed by the unit itself together with the header, everything indexed by the
    // two bodies right after it. Units in the head of a colour (they may touch) fetch all of it at once; the rest of the
    // colour fetches only what the reject needs (one 16-byte gather per body) and nearly always stops there.
    const float reach = ldc(c.RC + u);
    float4 l0, l1, g0, g1;
    if (head) { l0 = ldc(c.L0 + u); l1 = ldc(c.L1 + u); g0 = ldc(c.G0 + u); g1 = ldc(c.G1 + u); }
    const int b0 = __float_as_int(h.x), b1 = __float_as_int(h.y);
    const uint32_t fl = __float_as_uint(h.z);
    if (fl & UF_ASLEEP) return;
    const float mu_s = h.w;
    const int i0 = max(b0, 0), i1 = max(b1, 0);     // a static side reads body 0's row and ignores it: no branch between the loads
    float4 P0 = ldb(c.P + i0), P1 = ldb(c.P + i1);
    float4 Q0, Q1, mi0, mi1;
    if (head) { Q0 = ldb(c.Q + i0); Q1 = ldb(c.Q + i1); mi0 = ldb(c.MI + i0); mi1 = ldb(c.MI + i1); }
    if (b0 < 0) P0 = make_float4(0, 0, 0, 0);
    if (b1 < 0) P1 = make_float4(0, 0, 0, 0);
    if (!head) {
        if (b0 < 0) l0 = ldc(c.L0 + u);
        if (b1 < 0) l1 = ldc(c.L1 + u);
    }
    {
        const V2 a = b0 >= 0 ? mk(P0.x, P0.y) : mk(l0.x, l0.y), da = mk(P0.z, P0.w);
        const V2 b = b1 >= 0 ? mk(P1.x, P1.y) : mk(l1.x, l1.y), db = mk(P1.z, P1.w);
        const V2 d = mk((b.x - a.x) + (db.x - da.x), (b.y - a.y) + (db.y - da.y));
        SFG_STAMP(2, __float_as_uint(d.x + d.y));
        if (mag_sq(d) > reach * reach) return;
    }
    if (!head) {
        if (b0 >= 0) l0 = ldc(c.L0 + u);
        if (b1 >= 0) l1 = ldc(c.L1 + u);
        g0 = ldc(c.G0 + u); g1 = ldc(c.G1 + u);
        Q0 = ldb(c.Q + i0); Q1 = ldb(c.Q + i1); mi0 = ldb(c.MI + i0); mi1 = ldb(c.MI + i1);
    }
    const Shape sh0 = mkshape(g0), sh1 = mkshape(g1);
    const Iso loc0 = mkiso(mk(l0.x, l0.y), mkrot(l0.z, l0.w)), loc1 = mkiso(mk(l1.x, l1.y), mkrot(l1.z, l1.w));
    if (b0 < 0) { Q0 = make_float4(1, 0, 1, 0); mi0 = make_float4(0, 0, 0, 0); }
    if (b1 < 0) { Q1 = make_float4(1, 0, 1, 0); mi1 = make_float4(0, 0, 0, 0); }
    const float im0 = mi0.x, ii0 = mi0.y, im1 = mi1.x, ii1 = mi1.y;
    // pair-local frame: origin at object 0's substep-start position (or its static translation)
    const V2 org = b0 >= 0 ? mk(P0.x, P0.y) : loc0.t;
    const V2 told0 = mk(0.f, 0.f);
    const V2 told1 = b1 >= 0 ? mk(P1.x - org.x, P1.y - org.y) : mk(0.f, 0.f);
    const Iso st0 = mkiso(mk(0.f, 0.f), loc0.r);                       // static collider 0 in the local frame
    const Iso st1 = mkiso(loc1.t - org, loc1.r);
    Rot r0 = mkrot(Q0.x, Q0.y), r1 = mkrot(Q1.x, Q1.y);
    const Rot ro0 = mkrot(Q0.z, Q0.w), ro1 = mkrot(Q1.z, Q1.w);
    const int mult = (fl & UF_MULT2) ? 2 : 1;
    const bool sees = (fl & (UF_SEES0 | UF_SEES1)) != 0;
    // Second reject, per visit, on the collider centres themselves: every branch of intersection_check that reports a
    // contact needs the rounded shapes to overlap, so centres farther apart than the two bounding radii (plus an f32
    // safety margin) give count == 0. The floor keeps circle_circle's "centres closer than sqrt(0.001) always collide"
    // rule (shape_shape.rs:95-103) out of the shortcut; a NaN distance fails the comparison and takes the exact path.
    const float reach2 = (bound_radius(sh0) + bound_radius(sh1)) * 1.001f + 1e-3f;
    const float far_sq = fmaxf(reach2 * reach2, 0.0011f);

    const uint32_t tag = (c.tag_base + sub + 1u) << 8;
    bool moved = false;
    SFG_STAMP(3, __float_as_uint(far_sq + l0.x + l1.x + Q0.x + Q1.x + im0 + im1));
    for (int m = 0; m < mult; ++m) {
        const uint32_t e = u + uint32_t(m) * c.n_units;
        const Iso pw0 = b0 >= 0 ? mkiso(mk(P0.z, P0.w), r0) * loc0 : st0;
        const Iso pw1 = b1 >= 0 ? mkiso(mk(told1.x + P1.z, told1.y + P1.w), r1) * loc1 : st1;
        Manifold mf;
        mf.count = 0;
        if (!(mag_sq(pw1.t - pw0.t) > far_sq)) intersection_check(pw0, pw1, sh0, sh1, mf);
#ifdef SFG_PROBE
        if (m == 0) SFG_STAMP(1, __float_as_uint(mf.n.x + mf.off[0][0].x + mf.off[0][1].x + mf.off[1][0].x) + uint32_t(mf.count));
#endif
        if (mf.count == 0) {
            // first visit found nothing: no pose changed, so the second visit finds the same nothing and the unit is done
            // without a single store. A second visit that finds nothing after a first that did must say so (fresh tag).
            if (m == 0) return;
            stm(c.M0 + e, make_float4(__int_as_float(int(tag)), 0.f, 0.f, 0.f));
            continue;
        }
        for (int k = 0; k < mf.count; ++k) {                            // solver.rs:308-315
            if (b0 >= 0) mf.off[k][0] = loc0 * mf.off[k][0];
            if (b1 >= 0) mf.off[k][1] = loc1 * mf.off[k][1];
        }
        stm(c.LN + e, make_float2(mf.n.x, mf.n.y));                          // solver.rs:318-320 (latch)
        float lambda_n = 0.f;
        if (sees) {
            if (mf.count == 1 && (fl & (UF_ROPE0 | UF_ROPE1))) {        // solver.rs:337-362
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (!(fl & (i ? UF_ROPE1 : UF_ROPE0))) continue;
                    const int bi = i ? b1 : b0;
                    const int prev = c.rope_prev[bi], next = c.rope_next[bi];
                    const float4 ps = i ? P1 : P0;
                    const float2 ds = ldb(c.PCD + bi);
                    const float4 pp = ldb(c.P + prev), pn = ldb(c.P + next);
                    const float2 dp = ldb(c.PCD + prev), dn = ldb(c.PCD + next);
                    V2 no = i ? -mf.n : mf.n;
                    V2 to_prev = mk((pp.x - ps.x) + (dp.x - ds.x), (pp.y - ps.y) + (dp.y - ds.y));
                    V2 to_next = mk((pn.x - ps.x) + (dn.x - ds.x), (pn.y - ps.y) + (dn.y - ds.y));
                    const V2 seg = dot(no, to_prev) > dot(no, to_next) ? to_prev : to_next;
                    if (SFG_GUARD(4) && seg.x == 0.f && seg.y == 0.f) continue;   // neighbour on top of the particle: keep the shape normal
                    V2 nn = normalized(left_normal(seg));
                    mf.n = dot(mf.n, nn) > 0.f ? nn : -nn;
                }
            }
            if (!(fl & UF_SENSOR)) {
                const V2 n = mf.n, tan = left_normal(mf.n);
                for (int k = 0; k < mf.count; ++k) {
                    V2 ow0, ow1, owo0, owo1;
                    float wn0 = 0.f, wn1 = 0.f, wt0 = 0.f, wt1 = 0.f, en0 = 0.f, en1 = 0.f, et0 = 0.f, et1 = 0.f;
                    if (b0 >= 0) {
                        V2 orot = r0 * mf.off[k][0];
                        wn0 = wedge(orot, n); wt0 = wedge(orot, tan);
                        ow0 = orot + mk(P0.z, P0.w);
                        en0 = im0 + (wn0 * wn0 * ii0); et0 = im0 + (wt0 * wt0 * ii0);
                        owo0 = ro0 * mf.off[k][0] + told0;
                    } else { ow0 = st0 * mf.off[k][0]; owo0 = ow0; }
                    if (b1 >= 0) {
                        V2 orot = r1 * mf.off[k][1];
                        wn1 = wedge(orot, n); wt1 = wedge(orot, tan);
                        ow1 = orot + mk(told1.x + P1.z, told1.y + P1.w);
                        en1 = im1 + (wn1 * wn1 * ii1); et1 = im1 + (wt1 * wt1 * ii1);
                        owo1 = ro1 * mf.off[k][1] + told1;
                    } else { ow1 = st1 * mf.off[k][1]; owo1 = ow1; }
                    float depth = dot(ow0 - ow1, n);
                    if (depth <= 0.f) { lambda_n = 0.f; continue; }
                    lambda_n = -depth / (en0 + en1);
                    moved = true;
                    apply_correction(P0, r0, im0, ii0, 1.f, lambda_n, n, wn0);
                    apply_correction(P1, r1, im1, ii1, -1.f, lambda_n, n, wn1);
                    if (fl & UF_HAS_SF) {                               // solver.rs:456-487
                        V2 odm = (ow0 - owo0) - (ow1 - owo1);
                        float mat = dot(odm, tan);
                        float lambda_t = -mat / (et0 + et1);
                        if (lambda_t < lambda_n * mu_s) {
                            apply_correction(P0, r0, im0, ii0, 1.f, lambda_t, tan, wt0);
                            apply_correction(P1, r1, im1, ii1, -1.f, lambda_t, tan, wt1);
                        }
                    }
                }
            }
        }
        SFG_STAMP(4 + m, __float_as_uint(lambda_n + P0.z + P1.z + r0.s + r1.s + mf.n.x));
        stm(c.M0 + e, make_float4(__int_as_float(mf.count | int(tag)), lambda_n, mf.n.x, mf.n.y));
        stm(c.M1 + e, make_float4(mf.off[0][0].x, mf.off[0][0].y, mf.off[0][1].x, mf.off[0][1].y));
        if (mf.count == 2) stm(c.M2 + e, make_float4(mf.off[1][0].x, mf.off[1][0].y, mf.off[1][1].x, mf.off[1][1].y));
    }
    if (moved) {
        if (fl & UF_SEES0) { stb(c.P + b0, P0); Q0.x = r0.s; Q0.y = r0.b; stb(c.Q + b0, Q0); }
        if (fl & UF_SEES1) { stb(c.P + b1, P1); Q1.x = r1.s; Q1.y = r1.b; stb(c.Q + b1, Q1); }
    }
    // the first visit found a contact: queue the unit for this substep's velocity sweep (solver.rs:496-618 skips sensors
    // and pairs that see no forces). One atomic per group of converged lanes.
    if (!(fl & UF_SENSOR) && sees) {
        const uint32_t act = __activemask(), lane = threadIdx.x & 31u;
        const int leader = __ffs(int(act)) - 1;
        uint32_t base = 0;
        if (int(lane) == leader) base = atomicAdd(c.wl_count + sub * c.n_colours + colour, uint32_t(__popc(act)));
        base = __shfl_sync(act, base, leader);
        stw(c.WL + colour_begin + base + uint32_t(__popc(act & ((1u << lane) - 1u))), u);
        if (c.stat_touch) {                                               // measurement only (SFG_TUNE_COUNT_ENTRIES)
            const uint32_t two = __ballot_sync(act, mult == 2);
            if (int(lane) == leader) { atomicAdd(c.stat_touch, uint32_t(__popc(act))); atomicAdd(c.stat_touch + 1, uint32_t(__popc(act) + __popc(two))); }
        }
        SFG_STAMP(6, base);
    }
#ifdef SFG_PROBE
    if (mult == 2 && pt[5] && c.n_colours && __popc(__activemask()) <= 2) {   // thinly dealt units only: their own latency, not a crowded warp's
        const unsigned long long slot = atomicAdd(&g_probe[0], 1ull);
        if (slot < 65536) for (int i = 0; i < 8; ++i) g_probe[1 + slot * 8 + i] = i == 7 ? (uint64_t(colour) << 32 | sub) : pt[i];
    }
#endif
}

// solver.rs:496-618
SFG_DI void stage_contact_vel(const SolverCtx& c, uint32_t u, uint32_t sub) {
    // Only units queued by this substep's position sweep come here, so they have a contact: everything indexed by the unit
    // (header, material, both manifold entries) is fetched in one batch, everything indexed by the bodies in a second one.
    const float4 h = ldc(c.H + u);
    const float4 s = ldc(c.S + u);
    const uint32_t e1 = u + c.n_units;
    float4 m0[2], m1[2], m2[2];
    m0[0] = ldm(c.M0 + u); m1[0] = ldm(c.M1 + u); m2[0] = ldm(c.M2 + u);
    m0[1] = ldm(c.M0 + e1); m1[1] = ldm(c.M1 + e1); m2[1] = ldm(c.M2 + e1);      // entries [n_units, 2 n_units) exist for every unit
    const uint32_t fl = __float_as_uint(h.z);
    const int mult = (fl & UF_MULT2) ? 2 : 1;
    const uint32_t tag = c.tag_base + sub + 1u;
    const int b0 = __float_as_int(h.x), b1 = __float_as_int(h.y);
    const int i0 = max(b0, 0), i1 = max(b1, 0);
    const float mu_d = s.z, rest_mat = s.w;
    // old velocities and external accelerations only feed the restitution term (solver.rs:555-576): e = 0 never reads them
    const bool bouncy = rest_mat != 0.f;
    float4 q0 = ldb(c.Q + i0), q1 = ldb(c.Q + i1), V0 = ldb(c.V + i0), V1 = ldb(c.V + i1), mi0 = ldb(c.MI + i0), mi1 = ldb(c.MI + i1);
    float4 OV0 = make_float4(0, 0, 0, 0), OV1 = OV0;
    V2 ea = mk(0.f, 0.f);
    if (bouncy) {
        if (b0 >= 0) { OV0 = ldb(c.OV + b0); float2 a = ldb(c.EA + b0); ea = ea + mk(a.x, a.y); }
        if (b1 >= 0) { OV1 = ldb(c.OV + b1); float2 a = ldb(c.EA + b1); ea = ea + mk(a.x, a.y); }
    }
    int cnt[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) { const uint32_t bits = __float_as_uint(m0[m].x); cnt[m] = (bits >> 8) == tag ? int(bits & M0_COUNT_MASK) : 0; }
    if (mult == 1) cnt[1] = 0;
    if (cnt[0] == 0 && cnt[1] == 0) return;
    if (b0 < 0) { q0 = make_float4(1, 0, 1, 0); V0 = make_float4(0, 0, 0, 0); mi0 = V0; }
    if (b1 < 0) { q1 = make_float4(1, 0, 1, 0); V1 = make_float4(0, 0, 0, 0); mi1 = V1; }
    const Rot r0 = mkrot(q0.x, q0.y), r1 = mkrot(q1.x, q1.y);
    const float im0 = mi0.x, ii0 = mi0.y, im1 = mi1.x, ii1 = mi1.y;
    const float rest_gate = c.dt * c.dt * mag_sq(ea);
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int count = cnt[m];
        if (count == 0) continue;
        const float lambda_n = m0[m].y;
        const V2 n = mk(m0[m].z, m0[m].w), tan = left_normal(n);
        for (int k = 0; k < count; ++k) {
            const float4 pt = k == 0 ? m1[m] : m2[m];
            const V2 or0 = b0 >= 0 ? r0 * mk(pt.x, pt.y) : mk(0.f, 0.f);
            const V2 or1 = b1 >= 0 ? r1 * mk(pt.z, pt.w) : mk(0.f, 0.f);
            V2 pv0 = mk(V0.x - or0.y * V0.z, V0.y + or0.x * V0.z), pv1 = mk(V1.x - or1.y * V1.z, V1.y + or1.x * V1.z);
            V2 opv0 = mk(OV0.x - or0.y * OV0.z, OV0.y + or0.x * OV0.z), opv1 = mk(OV1.x - or1.y * OV1.z, OV1.y + or1.x * OV1.z);
            V2 rel = pv0 - pv1;
            float nv = dot(rel, n);
            float onv = dot(opv0 - opv1, n);
            float rest = (!bouncy || onv * onv < rest_gate) ? 0.f : rest_mat;
            float dnv = -nv - rest * fmaxf(onv, 0.f);
            float dtv = 0.f;
            if (fl & UF_HAS_DF) {
                float tv = dot(rel, tan);
                float max_dv = c.inv_dt * lambda_n * mu_d;
                dtv = fminf(fabsf(tv), fabsf(max_dv)) * -signum_rs(tv);
            }
            V2 upd = dnv * n + dtv * tan;
            float um = mag(upd);
            if (um < 0.0001f) continue;
            V2 dir = mk(upd.x / um, upd.y / um);
            float wd0 = wedge(or0, dir), wd1 = wedge(or1, dir);
            float e0 = im0 + (wd0 * wd0 * ii0), e1d = im1 + (wd1 * wd1 * ii1);
            float imp = um / (e0 + e1d);
            V0.x += im0 * imp * dir.x; V0.y += im0 * imp * dir.y; V0.z += ii0 * imp * wd0;
            V1.x -= im1 * imp * dir.x; V1.y -= im1 * imp * dir.y; V1.z -= ii1 * imp * wd1;
        }
    }
    if (fl & UF_SEES0) stb(c.V + b0, V0);
    if (fl & UF_SEES1) stb(c.V + b1, V1);
}

// The two rejects of stage_contacts on their own: true = the unit cannot touch and stage_contacts would return from its first
// visit without a store. First the body-level test (header, reach, one 16-byte gather per body or the static collider's
// position); a unit that passes it (common for compounds: the body-level reach is loose) is tested on its collider centres
// like every visit of stage_contacts does, with a 1e-4 margin on the squared bound so that rounding differences between
// the two evaluations can only make this helper keep a unit, never drop one.
SFG_DI bool contacts_rejects(const SolverCtx& c, uint32_t u) {
    const float4 h = ldc(c.H + u);
    const int b0 = __float_as_int(h.x), b1 = __float_as_int(h.y);
    if (__float_as_uint(h.z) & UF_ASLEEP) return true;
    const float reach = ldc(c.RC + u);
    const float4 l0 = ldc(c.L0 + u), l1 = ldc(c.L1 + u);
    float4 A = make_float4(l0.x, l0.y, 0.f, 0.f), B = make_float4(l1.x, l1.y, 0.f, 0.f);
    if (b0 >= 0) A = ldb(c.P + b0);
    if (b1 >= 0) B = ldb(c.P + b1);
    const V2 d = mk((B.x - A.x) + (B.z - A.z), (B.y - A.y) + (B.w - A.w));
    if (mag_sq(d) > reach * reach) return true;
    // collider centres in the pair-local frame (origin = object 0's substep-start position or static translation)
    V2 t0 = mk(0.f, 0.f), t1 = d;
    if (b0 >= 0) { const float4 q = ldb(c.Q + b0); t0 = mkrot(q.x, q.y) * mk(l0.x, l0.y) + mk(A.z, A.w); }
    if (b1 >= 0) { const float4 q = ldb(c.Q + b1); t1 = mkrot(q.x, q.y) * mk(l1.x, l1.y) + mk((B.x - A.x) + B.z, (B.y - A.y) + B.w); }
    else t1 = mk(l1.x - A.x, l1.y - A.y);
    const float reach2 = (bound_radius(mkshape(ldc(c.G0 + u))) + bound_radius(mkshape(ldc(c.G1 + u)))) * 1.001f + 1e-3f;
    const float far_sq = fmaxf(reach2 * reach2, 0.0011f);
    return mag_sq(t1 - t0) > far_sq * 1.0001f;
}

// ---------------------------------------------------------------- drivers
SFG_DI void run_stage(const SolverCtx& c, uint32_t kind, uint32_t i, bool last_substep, uint32_t sub, uint32_t colour, uint32_t begin, bool head) {
    switch (kind) {
    case ST_INTEGRATE: stage_integrate(c, i); break;
    case ST_ROPE_STEP: stage_rope_step(c, i); break;
    case ST_CONSTRAINTS: stage_constraints(c, i); break;
    case ST_PRECONTACT: stage_precontact(c, i); break;
    case ST_CONTACTS: stage_contacts(c, i, colour, begin, sub, head); break;
    case ST_DERIVE: stage_derive(c, i, last_substep); break;
    case ST_CONTACT_VEL: stage_contact_vel(c, ldw(c.WL + i), sub); break;   // i indexes the colour's worklist
    case ST_CONSTRAINT_DAMP: stage_constraint_damp(c, i); break;
    case ST_ROPE_DAMP: stage_rope_damp(c, i); break;
    case ST_ROPE_LATERAL: stage_rope_lateral(c, i); break;
    default: break;
    }
}

template <uint32_t KIND>
__global__ void __launch_bounds__(256) k_stage(const __grid_constant__ SolverCtx c, uint32_t begin, uint32_t end, int last_substep, uint32_t sub,
                                               uint32_t colour) {
    if (KIND == ST_CONTACT_VEL) end = begin + __ldcg(c.wl_count + sub * c.n_colours + colour);   // only the units queued by the position sweep
    const uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < end) run_stage(c, KIND, i, last_substep != 0, sub, colour, begin, false);
}

// Grid-wide barrier of the persistent kernel: one release-atomic per CTA on a monotonically increasing counter, a
// relaxed polling load (no L1 invalidation per poll, unlike cooperative_groups' grid.sync) and one acquire fence at
// the end. All data that crosses SMs between stages is accessed with ld.cg / st.cg, so L1 never holds it.
SFG_DI unsigned long long global_timer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
SFG_DI void grid_barrier(unsigned int* counter, unsigned int& target, uint32_t* s_next, unsigned long long* trace = nullptr) {
    // s_next: the CTA's work counter of the stage that follows (k_solve_persistent deals a stage's items to the warps of a CTA
    // dynamically); it is reset here, between the two CTA-wide barriers, when no warp can be reading it
    if (gridDim.x == 1 && !trace) {   // a small world runs in one CTA: bar.sync orders its global accesses (all through L2)
        __syncthreads();
        if (threadIdx.x == 0) *s_next = 0u;
        __syncthreads();
        return;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *s_next = 0u;
        target += gridDim.x;
        if (trace) trace[0] = global_timer();
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int seen;
        do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); } while (seen < target);
        __threadfence();
        if (trace) trace[1] = global_timer();
    }
    __syncthreads();
}

// The whole substep loop of one tick in ONE launch: every CTA walks the stage table, and the units of a stage are dealt
// to warps round-robin ACROSS CTAs (chunk c of 32 units -> CTA c mod gridDim, warp slot c / gridDim), so a run of
// expensive units (they are grouped by narrowphase path) is spread over all SMs; then all CTAs meet at a barrier.
__global__ void __launch_bounds__(SFG_SOLVER_THREADS, 1) k_solve_persistent(const __grid_constant__ SolverCtx c, const Stage* __restrict__ stages,
                                                             uint32_t n_stages, uint32_t sub_begin, uint32_t sub_end, uint32_t substeps,
                                                             unsigned int* barrier_counter, unsigned long long* trace, uint32_t* dbg_nan) {
    __shared__ uint32_t s_next;                         // next item of the current stage that no warp of this CTA has taken yet
    if (threadIdx.x == 0) s_next = 0u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warps_per_cta = blockDim.x >> 5;
    const uint32_t first_chunk = (threadIdx.x >> 5) * gridDim.x + blockIdx.x, chunk_stride = gridDim.x * warps_per_cta;
    unsigned int target = 0;
    for (uint32_t s = sub_begin; s < sub_end; ++s) {   // substeps [sub_begin, sub_end) of a tick of `substeps`
        const bool last = s + 1 == substeps;
        uint32_t merged_until = 0;                      // contact stages below this index have been through the merged reject pass
        for (uint32_t k = 0; k < n_stages; ++k) {
            const Stage st = stages[k];
            // The high colours hold only pairs that were clearly apart at tick start (near_count == 0), and nearly always every
            // one of them is rejected by the distance test, stage after stage, each paying a barrier and two round trips for
            // nothing. A reject changes no state, so a run of such stages is tested in ONE pass; if every unit is rejected the
            // whole run is done. If some unit survives, the stages from the first survivor's on are run the ordinary way (the
            // stages before it changed nothing, so the order of the sweep is exactly the sequential one).
            if (st.kind == ST_CONTACTS && k >= merged_until && __ldg(c.near_count + st.colour) == 0u) {
                uint32_t k2 = k + 1;
                while (k2 < n_stages && stages[k2].kind == ST_CONTACTS && __ldg(c.near_count + stages[k2].colour) == 0u) ++k2;
                if (k2 - k >= 2u) {
                    const uint32_t lo = st.begin, hi = stages[k2 - 1u].end;
                    bool survivor = false;
                    uint32_t first = 0xFFFFFFFFu;
                    for (uint32_t i = lo + first_chunk * 32u + lane; i < hi; i += chunk_stride * 32u)
                        if (!contacts_rejects(c, i)) { survivor = true; first = min(first, i); }
                    if (survivor) {                     // rare: which stage is that unit in?
                        uint32_t ks = k;
                        while (ks + 1u < k2 && first >= stages[ks].end) ++ks;
                        atomicMin(c.far_min + size_t(s) * n_stages + k, ks);   // one word per run: a substep can hold several runs
                    }
                    grid_barrier(barrier_counter, target, &s_next, trace ? trace + 2 * (size_t((s - sub_begin) * n_stages + k) * gridDim.x + blockIdx.x) : nullptr);
                    const uint32_t fm = __ldcg(c.far_min + size_t(s) * n_stages + k);
                    merged_until = k2;
                    if (fm == 0xFFFFFFFFu) { k = k2 - 1u; continue; }       // every unit of the run rejected: all its stages are done
                    k = fm - 1u;                                             // resume at the first survivor's stage
                    continue;
                }
            }
            const uint32_t n = st.kind == ST_CONTACT_VEL ? __ldcg(c.wl_count + s * c.n_colours + st.colour) : st.end - st.begin;
            // an empty velocity worklist (the high colours hold only pairs that did not touch): no work and no barrier. The count
            // was final before the barrier of the derive stage, so every CTA takes the same branch.
            if (st.kind == ST_CONTACT_VEL && n == 0u) continue;
            // Units that may really touch (the head of a contact colour, or a whole velocity worklist) run the long,
            // branchy path; a stage lasts as long as its slowest warp, and a warp serialises the distinct paths of its lanes.
            // So those units are dealt out as thinly as the machine allows — ceil(n / warps) lanes per warp — instead of 32
            // to a warp; the cheap rest of the colour (one distance test per unit) follows in full warps.
            uint32_t n_head = 0;
            if (st.kind == ST_CONTACT_VEL) n_head = n;
            else if (st.kind == ST_CONTACTS) n_head = min(__ldg(c.near_count + st.colour), n);
            if (st.kind == ST_INTEGRATE || st.kind == ST_DERIVE) {
                // streaming stages: all four columns of BODY_ILP chunks are in flight before the first body is touched
                #ifndef SFG_BODY_ILP
#define SFG_BODY_ILP 2
#endif
                constexpr uint32_t BODY_ILP = SFG_BODY_ILP;
                const uint32_t chunks = (n + 31u) / 32u, items = (chunks + BODY_ILP - 1u) / BODY_ILP;
                for (uint32_t it = first_chunk; it < items; it += chunk_stride) {
                    float4 mi[BODY_ILP], p[BODY_ILP], q[BODY_ILP], v[BODY_ILP];
#pragma unroll
                    for (uint32_t j = 0; j < BODY_ILP; ++j) {
                        const uint32_t i = st.begin + min((it + j * items) * 32u + lane, n - 1u);   // out-of-range slots re-read the last body and are skipped below
                        mi[j] = ldb(c.MI + i); p[j] = ldb(c.P + i); q[j] = ldb(c.Q + i); v[j] = ldb(c.V + i);
                    }
#pragma unroll
                    for (uint32_t j = 0; j < BODY_ILP; ++j) {
                        const uint32_t i = (it + j * items) * 32u + lane;
                        if (i < n) {
                            if (st.kind == ST_INTEGRATE) integrate_body(c, st.begin + i, mi[j], p[j], q[j], v[j]);
                            else derive_body(c, st.begin + i, last, mi[j], p[j], q[j], v[j]);
                        }
                    }
                }
            } else {
            const uint32_t lanes = n_head ? min(32u, (n_head + chunk_stride - 1u) / chunk_stride) : 32u;
            const uint32_t groups = (n_head + lanes - 1u) / lanes, items = groups + (n - n_head + 31u) / 32u;
            // Items go to CTAs round-robin (item it -> CTA it mod gridDim: a run of expensive units is spread over all SMs) and,
            // inside a CTA, to whichever warp is free next (one shared-memory counter): the items of a contact colour differ
            // several-fold in cost — touching or not, one visit or two — and with a fixed deal the stage waited for the warp that
            // drew the expensive ones (17 % of the settled solver's time was spent between the first and the last CTA's arrival).
            // The last SFG_STEAL_PERCENT of a big stage's items are not dealt to CTAs at all: a warp whose CTA has run out takes them
            // one by one from a global counter, so the CTAs that drew cheap items help finish the stage (the sums of ~100 item
            // costs per CTA still differ by 10-15 % between the luckiest and the unluckiest of 148 CTAs). Only for stages with
            // SFG_STEAL_MIN_ITEMS items per CTA or more: in smaller ones the global atomics cost more than the balance returns.
            const uint32_t per_cta = items / gridDim.x;
            const uint32_t n_static = (SFG_STEAL_PERCENT && per_cta >= SFG_STEAL_MIN_ITEMS) ? (per_cta - per_cta * SFG_STEAL_PERCENT / 100u) * gridDim.x : items;
            uint32_t* steal = c.steal + size_t(s) * n_stages + k;
            bool tail = false;
            for (;;) {
                uint32_t it;
                if (!tail) {
                    uint32_t j = 0;
                    if (lane == 0) j = atomicAdd(&s_next, 1u);
                    j = __shfl_sync(0xFFFFFFFFu, j, 0);
                    it = j * gridDim.x + blockIdx.x;
                    if (it >= n_static) { if (n_static >= items) break; tail = true; continue; }
                } else {
                    uint32_t t = 0;
                    if (lane == 0) t = atomicAdd(steal, 1u);
                    t = __shfl_sync(0xFFFFFFFFu, t, 0);
                    it = n_static + t;
                    if (it >= items) break;
                }
                const bool head = it < groups;
                const uint32_t i = head ? it * lanes + lane : n_head + (it - groups) * 32u + lane;
                if (head ? (lane < lanes && i < n_head) : i < n) run_stage(c, st.kind, st.begin + i, last, s, st.colour, st.begin, head);
            }
            }
            grid_barrier(barrier_counter, target, &s_next, trace ? trace + 2 * (size_t((s - sub_begin) * n_stages + k) * gridDim.x + blockIdx.x) : nullptr);
            if (dbg_nan) {   // debug (SFG_DEBUG_NAN): the first stage after which some body is not finite
                for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < c.n_bodies; b += gridDim.x * blockDim.x) {
                    if (!(__float_as_uint(ldb(c.MI + b).z) & BF_ALIVE)) continue;
                    const float4 p = ldb(c.P + b), q = ldb(c.Q + b), v = ldb(c.V + b);
                    if (!isfinite(p.x + p.y + p.z + p.w + q.x + q.y + v.x + v.y + v.z)) { atomicMin(dbg_nan, s * n_stages + k); atomicMin(dbg_nan + 1 + min(s * n_stages + k, 4000u), b); }
                }
            }
        }
    }
}

}  // namespace sfg

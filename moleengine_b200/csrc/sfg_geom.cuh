// Device geometry for the Starframe physics tick (sm_100a): 2-D rotors / isometries, the closed
// collider-shape tables, the SAT + edge-clip + rounded-corner narrowphase and the ray / spherecast
// tests. f32 throughout; everything is evaluated in a frame local to the pair (callers subtract a
// reference origin first), which is what keeps f32 usable far from the world origin.
//
// Behaviour follows /root/reference/src/physics/collision/{collider.rs:505-892,
// shape_shape.rs:63-488, query.rs:7-308}; each function cites the item it replaces.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace sfg {

#define SFG_DI __device__ __forceinline__
#ifndef SFG_GUARD_MASK
#define SFG_GUARD_MASK 0xFFu            // zero-length guards (bisecting aid: build with -DSFG_GUARD_MASK=<bits> to switch some off)
#endif
#define SFG_GUARD(bit) ((SFG_GUARD_MASK >> (bit)) & 1u)

struct V2 { float x, y; };
SFG_DI V2 mk(float x, float y) { V2 v; v.x = x; v.y = y; return v; }
SFG_DI V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
SFG_DI V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
SFG_DI V2 operator-(V2 a) { return mk(-a.x, -a.y); }
SFG_DI V2 operator*(float s, V2 a) { return mk(s * a.x, s * a.y); }
SFG_DI V2 operator*(V2 a, float s) { return mk(s * a.x, s * a.y); }
SFG_DI float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
SFG_DI float wedge(V2 a, V2 b) { return a.x * b.y - b.x * a.y; }
SFG_DI float mag_sq(V2 a) { return a.x * a.x + a.y * a.y; }
SFG_DI float mag(V2 a) { return sqrtf(mag_sq(a)); }
SFG_DI V2 normalized(V2 a) { float r = 1.0f / mag(a); return mk(a.x * r, a.y * r); }
SFG_DI V2 left_normal(V2 a) { return mk(-a.y, a.x); }            // math.rs:393-396

// rotor {s, b = bv.xy}; from_angle(t) = {cos(t/2), -sin(t/2)}; positive angle = counter-clockwise
struct Rot { float s, b; };
SFG_DI Rot mkrot(float s, float b) { Rot r; r.s = s; r.b = b; return r; }
SFG_DI Rot rev(Rot r) { return mkrot(r.s, -r.b); }
SFG_DI Rot operator*(Rot p, Rot q) { return mkrot(p.s * q.s - p.b * q.b, p.s * q.b + q.s * p.b); }
// rotor * vector, written through the rotation's cosine and sine (the sandwich s*(s*v + b*Jv) + ... expands to exactly this):
// the two products depend on the rotor alone, so a rotor applied to several vectors — every pose in the narrowphase is — pays
// for them once and each further vector costs four operations instead of eight.
SFG_DI V2 operator*(Rot r, V2 v) {
    const float c = r.s * r.s - r.b * r.b, sn = 2.0f * r.s * r.b;
    return mk(c * v.x + sn * v.y, c * v.y - sn * v.x);
}
// from_angle: corrections inside the solver are rotations by milliradians, for which a three-term series is exact to f32
// rounding (the next terms are below 2^-40 relative). Everything else lives out of line — the h^9 series up to a quarter
// radian of half-angle, sincosf beyond (its argument reduction alone is ~130 instructions): inlined at every correction site
// that was 15 % of the solver kernel's code and it never runs there; the solver's lone warps are instruction-fetch bound.
__device__ __noinline__ void from_angle_general(float h, float* sn, float* cs) {
    const float h2 = h * h;
    if (h2 < 0.0625f) {
        *sn = h * (1.0f + h2 * (-1.0f / 6.0f + h2 * (1.0f / 120.0f + h2 * (-1.0f / 5040.0f + h2 * (1.0f / 362880.0f)))));
        *cs = 1.0f + h2 * (-0.5f + h2 * (1.0f / 24.0f + h2 * (-1.0f / 720.0f + h2 * (1.0f / 40320.0f + h2 * (-1.0f / 3628800.0f)))));
    } else sincosf(h, sn, cs);
}
SFG_DI Rot rot_from_angle(float angle) {
    const float h = 0.5f * angle, h2 = h * h;
    if (h2 < 1e-3f) return mkrot(1.0f + h2 * (-0.5f + h2 * (1.0f / 24.0f)), -(h * (1.0f + h2 * (-1.0f / 6.0f + h2 * (1.0f / 120.0f)))));
    float sn, cs;
    from_angle_general(h, &sn, &cs);
    return mkrot(cs, -sn);
}

struct Iso { V2 t; Rot r; };
SFG_DI Iso mkiso(V2 t, Rot r) { Iso i; i.t = t; i.r = r; return i; }
SFG_DI V2 operator*(Iso p, V2 v) { return p.r * v + p.t; }
SFG_DI Iso operator*(Iso a, Iso b) { return mkiso(a.r * b.t + a.t, a.r * b.r); }
SFG_DI Iso inversed(Iso a) { Rot rr = rev(a.r); return mkiso(rr * (-a.t), rr); }

enum : uint32_t { K_POINT = 0, K_SEGMENT = 1, K_RECT = 2, K_TRIANGLE = 3, K_HEXAGON = 4 };

// collider.rs:419-433 (constants truncated exactly as the reference has them)
#define SFG_PI6_COS 0.866025403784f
#define SFG_PI6_SIN 0.5f
#define SFG_PI6_TAN 0.57735026919f
#define SFG_SQRT_3 1.73205080757f

struct Shape { uint32_t kind; float p0, p1, r; };   // p0 = hl | hw | outer_r ; p1 = hh ; r = circle_r
SFG_DI Shape mkshape(float4 v) { Shape s; s.p0 = v.x; s.p1 = v.y; s.r = v.z; s.kind = __float_as_uint(v.w); return s; }

// radius of the circle around the collider's origin that contains the rounded shape (vertices of Triangle / Hexagon
// lie at distance p0 from the origin: collider.rs:600-665)
SFG_DI float bound_radius(const Shape& s) {
    const float core = s.kind == K_POINT ? 0.f : (s.kind == K_RECT ? sqrtf(s.p0 * s.p0 + s.p1 * s.p1) : s.p0);
    return core + s.r;
}

struct Edge { V2 start, dir; float len; };
struct PEdge { V2 n; Edge e; };

SFG_DI int edge_count(const Shape& s) { return s.kind == K_POINT ? 0 : (s.kind == K_SEGMENT ? 1 : (s.kind == K_RECT ? 2 : 3)); }   // collider.rs:558-566
SFG_DI bool symmetrical(const Shape& s) { return s.kind != K_TRIANGLE; }                                                             // :528

// collider.rs:572-665
SFG_DI PEdge get_edge(const Shape& s, int i) {      // the kinds are tested most common first (rectangles and their rounded variants)
    PEdge o;
    if (s.kind == K_RECT) {
        if (i == 0) { o.n = mk(1.f, 0.f); o.e.start = mk(s.p0, -s.p1); o.e.dir = mk(0.f, 1.f); o.e.len = 2.f * s.p1; }
        else { o.n = mk(0.f, 1.f); o.e.start = mk(s.p0, s.p1); o.e.dir = mk(-1.f, 0.f); o.e.len = 2.f * s.p0; }
    } else if (s.kind == K_SEGMENT) {
        o.n = mk(0.f, 1.f); o.e.start = mk(s.p0, 0.f); o.e.dir = mk(-1.f, 0.f); o.e.len = 2.f * s.p0;
    } else if (s.kind == K_TRIANGLE) {
        const float len = s.p0 / SFG_PI6_TAN;
        if (i == 0) { o.n = mk(0.f, -1.f); o.e.start = (-s.p0) * mk(SFG_PI6_COS, SFG_PI6_SIN); o.e.dir = mk(1.f, 0.f); }
        else if (i == 1) { o.n = mk(SFG_PI6_COS, SFG_PI6_SIN); o.e.start = (-s.p0) * mk(-SFG_PI6_COS, SFG_PI6_SIN); o.e.dir = mk(-SFG_PI6_SIN, SFG_PI6_COS); }
        else { o.n = mk(-SFG_PI6_COS, SFG_PI6_SIN); o.e.start = mk(0.f, s.p0); o.e.dir = mk(-SFG_PI6_SIN, -SFG_PI6_COS); }
        o.e.len = len;
    } else {  // hexagon
        if (i == 0) { o.n = mk(SFG_PI6_COS, SFG_PI6_SIN); o.e.start = mk(s.p0, 0.f); o.e.dir = mk(-SFG_PI6_SIN, SFG_PI6_COS); }
        else if (i == 1) { o.n = mk(0.f, 1.f); o.e.start = s.p0 * mk(SFG_PI6_SIN, SFG_PI6_COS); o.e.dir = mk(-1.f, 0.f); }
        else { o.n = mk(-SFG_PI6_COS, SFG_PI6_SIN); o.e.start = s.p0 * mk(-SFG_PI6_SIN, SFG_PI6_COS); o.e.dir = mk(-SFG_PI6_SIN, -SFG_PI6_COS); }
        o.e.len = s.p0;
    }
    return o;
}
SFG_DI PEdge mirrored(PEdge p) { p.n = -p.n; p.e.start = -p.e.start; p.e.dir = -p.e.dir; return p; }   // shape_shape.rs:378-384
SFG_DI Edge flipped(Edge e) { Edge o; o.start = e.start + e.len * e.dir; o.dir = -e.dir; o.len = e.len; return o; }

// collider.rs:668-686
SFG_DI float edge_extent(const Shape& s, int i) {
    if (s.kind == K_RECT) return i == 0 ? s.p0 : s.p1;
    if (s.kind == K_SEGMENT) return 0.f;
    if (s.kind == K_TRIANGLE) return s.p0 / 2.f;
    return SFG_PI6_COS * s.p0;
}
// collider.rs:692-714
SFG_DI float projected_extent(const Shape& s, V2 d) {
    if (s.kind == K_RECT) return fabsf(d.x) * s.p0 + fabsf(d.y) * s.p1;
    if (s.kind == K_SEGMENT) return fabsf(d.x) * s.p0;
    if (s.kind == K_POINT) return 0.f;
    if (s.kind == K_TRIANGLE) {
        float a = d.y, b = -(SFG_PI6_COS * d.x + SFG_PI6_SIN * d.y), c = -(-SFG_PI6_COS * d.x + SFG_PI6_SIN * d.y);
        return fmaxf(a, fmaxf(b, c)) * s.p0;
    }
    float a = fabsf(d.x), b = fabsf(SFG_PI6_SIN * d.x + SFG_PI6_COS * d.y), c = fabsf(-SFG_PI6_SIN * d.x + SFG_PI6_COS * d.y);
    return fmaxf(a, fmaxf(b, c)) * s.p0;
}
// collider.rs:722-790
SFG_DI PEdge supporting_edge(const Shape& s, V2 d) {
    PEdge o;
    if (s.kind == K_SEGMENT) {
        o.n = mk(0.f, copysignf(1.f, d.y));
        o.e.start = mk(copysignf(s.p0, d.x), 0.f); o.e.dir = mk(copysignf(1.f, -d.x), 0.f); o.e.len = 2.f * s.p0;
        return o;
    }
    if (s.kind == K_RECT) {
        o.e.start = mk(copysignf(s.p0, d.x), copysignf(s.p1, d.y));
        if (fabsf(d.x) > fabsf(d.y)) { o.n = mk(copysignf(1.f, d.x), 0.f); o.e.dir = mk(0.f, -copysignf(1.f, d.y)); o.e.len = s.p1 * 2.f; }
        else { o.n = mk(0.f, copysignf(1.f, d.y)); o.e.dir = mk(-copysignf(1.f, d.x), 0.f); o.e.len = s.p0 * 2.f; }
        return o;
    }
    // triangle / hexagon: Iterator::max_by keeps the LAST of equal maxima
    float best = 0.f;
    int code = 0;
    const bool sym = symmetrical(s);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const V2 n = get_edge(s, i).n;
        float v = dot(n, d);
        const bool mir = sym && v < 0.f;
        if (mir) v = -v;                          // dot(-n, d)
        if (i == 0 || v >= best) { code = i * 2 + (mir ? 1 : 0); best = v; }
    }
    o = get_edge(s, code >> 1);
    if (code & 1) o = mirrored(o);
    if (!(dot(o.e.dir, d) < 0.f)) o.e = flipped(o.e);
    return o;
}
// collider.rs:796-892
SFG_DI V2 closest_boundary_point(const Shape& s, V2 pt, bool* interior) {
    *interior = false;
    switch (s.kind) {
    case K_POINT: return mk(0.f, 0.f);
    case K_SEGMENT: return mk(fminf(fmaxf(pt.x, -s.p0), s.p0), 0.f);
    case K_RECT: {
        float xd = fabsf(pt.x) - s.p0, yd = fabsf(pt.y) - s.p1;
        bool ox = xd > 0.f, oy = yd > 0.f;
        float ex = copysignf(s.p0, pt.x), ey = copysignf(s.p1, pt.y);
        if (ox && oy) return mk(ex, ey);
        if (ox) return mk(ex, pt.y);
        if (oy) return mk(pt.x, ey);
        *interior = true;
        if (fabsf(xd) < fabsf(yd)) return mk(ex, pt.y);
        return mk(pt.x, ey);
    }
    default: {
        float min_d = 3.402823466e+38f;
        V2 cp = mk(0.f, 0.f);
        const bool sym = symmetrical(s);
#pragma unroll 1
        for (int i = 0; i < 3; ++i) {
            PEdge e = get_edge(s, i);
            if (sym && dot(e.n, pt) < 0.f) e = mirrored(e);
            V2 sp = pt - e.e.start;
            float t = dot(e.e.dir, sp);
            if (t < 0.f) {
                float d = mag(sp);
                if (d < min_d) { min_d = d; cp = e.e.start; *interior = false; }
            } else if (t <= e.e.len) {
                float nd = dot(e.n, sp);
                bool in = !(nd >= 0.f);
                if (in) nd = -nd;
                if (nd < min_d) { min_d = nd; cp = e.e.start + t * e.e.dir; *interior = in; }
            }
        }
        return cp;
    }
    }
}

// ---- AABB, bit-exact against the f32 oracle: explicit _rn ops so nothing is contracted to FMA.
// collider.rs:505-523 + :344-349 + collision.rs:30-63
SFG_DI V2 rot_apply_exact(Rot r, V2 v) {
    float fx = __fadd_rn(__fmul_rn(r.s, v.x), __fmul_rn(r.b, v.y));
    float fy = __fsub_rn(__fmul_rn(r.s, v.y), __fmul_rn(r.b, v.x));
    return mk(__fadd_rn(__fmul_rn(r.s, fx), __fmul_rn(r.b, fy)), __fsub_rn(__fmul_rn(r.s, fy), __fmul_rn(r.b, fx)));
}
SFG_DI Rot rot_mul_exact(Rot p, Rot q) {
    return mkrot(__fsub_rn(__fmul_rn(p.s, q.s), __fmul_rn(p.b, q.b)), __fadd_rn(__fmul_rn(p.s, q.b), __fmul_rn(q.s, p.b)));
}
SFG_DI float4 shape_aabb_exact(const Shape& s, V2 t, Rot r) {
    V2 ext;
    if (s.kind == K_POINT) ext = mk(0.f, 0.f);
    else if (s.kind == K_SEGMENT) { V2 a = rot_apply_exact(r, mk(s.p0, 0.f)); ext = mk(fabsf(a.x), fabsf(a.y)); }
    else if (s.kind == K_RECT) {
        V2 a = rot_apply_exact(r, mk(s.p0, 0.f)), b = rot_apply_exact(r, mk(0.f, s.p1));
        ext = mk(__fadd_rn(fabsf(a.x), fabsf(b.x)), __fadd_rn(fabsf(a.y), fabsf(b.y)));
    } else ext = mk(s.p0, s.p0);
    float4 o;
    o.x = __fadd_rn(__fsub_rn(-ext.x, s.r), t.x); o.y = __fadd_rn(__fsub_rn(-ext.y, s.r), t.y);
    o.z = __fadd_rn(__fadd_rn(ext.x, s.r), t.x); o.w = __fadd_rn(__fadd_rn(ext.y, s.r), t.y);
    return o;
}
// collision.rs:91-99 (strict: touching boxes do not intersect)
SFG_DI bool aabb_overlap(float4 a, float4 b) {
    float mnx = fmaxf(a.x, b.x), mny = fmaxf(a.y, b.y), mxx = fminf(a.z, b.z), mxy = fminf(a.w, b.w);
    return !(mnx >= mxx || mny >= mxy);
}

// ---- narrowphase ----
struct Manifold {            // shape_shape.rs:6-10,55-60 — both points of a Two share one normal
    int count;
    V2 n;                    // world space, away from object 0
    V2 off[2][2];            // [point][object], object-local
};
SFG_DI void flip(Manifold& m) {                                   // shape_shape.rs:76-81
    m.n = -m.n;
#pragma unroll
    for (int k = 0; k < 2; ++k) { V2 t = m.off[k][0]; m.off[k][0] = m.off[k][1]; m.off[k][1] = t; }
}

// shape_shape.rs:87-112
SFG_DI void circle_circle(Iso p1, float r1, Iso p2, float r2, Manifold& m) {
    V2 d = p2.t - p1.t;
    float dsq = mag_sq(d), rs = r1 + r2;
    V2 n;
    if (dsq < 0.001f) n = mk(1.f, 0.f);
    else if (dsq < rs * rs) n = normalized(d);
    else { m.count = 0; return; }
    m.count = 1; m.n = n;
    m.off[0][0] = rev(p1.r) * (r1 * n);
    m.off[0][1] = rev(p2.r) * ((-r2) * n);
}
// shape_shape.rs:114-154; returned manifold has the circle as object 0
SFG_DI void circle_any(Iso pc, float rc, Iso po, const Shape& so, Manifold& m) {
    Iso local = inversed(po) * pc;
    V2 dist = local.t;
    bool interior;
    V2 cp = closest_boundary_point(so, dist, &interior);
    V2 fc = dist - cp;
    // Deviation (f32 robustness): where the reference normalises a vector that can be exactly zero it gets NaN and
    // the NaN poisons every body the solver reaches in that tick. In f64 that takes an exact coincidence; in f32 it happens
    // in crowded scenes (a rope pile lost 482 bodies at tick 256). A centre exactly on the boundary reports no contact for
    // this visit; the next visit sees it a rounding error to one side.
    if (SFG_GUARD(1) && fc.x == 0.f && fc.y == 0.f) { m.count = 0; return; }
    if (interior) {
        V2 dir = normalized(fc);
        m.count = 1; m.n = po.r * dir;
        m.off[0][0] = rev(local.r) * (rc * dir);
        m.off[0][1] = cp - so.r * dir;
        return;
    }
    float rs = rc + so.r, dsq = mag_sq(fc);
    if (dsq >= rs * rs) { m.count = 0; return; }
    V2 dir = fc * (1.0f / sqrtf(dsq));
    m.count = 1; m.n = po.r * (-dir);
    m.off[0][0] = rev(local.r) * (rc * (-dir));
    m.off[0][1] = cp + so.r * dir;
}

// shape_shape.rs:458-488; returns 0 intersects, 1 passes, 2 misses
SFG_DI int clip_edge(const Edge& target, const Edge& edge, float* enters, float* exits) {
    V2 sd = target.start - edge.start;
    float denom = edge.dir.x * (-target.dir.y) - (-target.dir.x) * edge.dir.y;
    float t0 = (sd.x * (-target.dir.y) - (-target.dir.x) * sd.y) / denom;
    float t1 = (edge.dir.x * sd.y - sd.x * edge.dir.y) / denom;
    if (t0 >= 0.f && t0 <= edge.len && t1 >= 0.f && t1 <= target.len) return 0;
    float ddd = dot(sd, target.dir), dirs = dot(edge.dir, target.dir);
    float st = ddd / dirs, en = (target.len + ddd) / dirs;
    if ((st <= 0.f && en <= 0.f) || (st >= edge.len && en >= edge.len)) return 2;
    if (st < en) { *enters = fmaxf(st, 0.f); *exits = fminf(en, edge.len); }
    else { *enters = fmaxf(en, 0.f); *exits = fminf(st, edge.len); }
    return 1;
}

// shape_shape.rs:160-349. Everything is kept in named registers (no dynamically indexed arrays): the SAT loop is
// unrolled over the two owners, and the winning owner's data is picked with selects afterwards.
// the outward normal of edge i alone (what get_edge(s, i).n is): the SAT loop needs nothing else of the edge
SFG_DI V2 edge_normal(const Shape& s, int i) {
    if (s.kind == K_RECT) return i == 0 ? mk(1.f, 0.f) : mk(0.f, 1.f);
    if (s.kind == K_SEGMENT) return mk(0.f, 1.f);
    if (s.kind == K_TRIANGLE) return i == 0 ? mk(0.f, -1.f) : (i == 1 ? mk(SFG_PI6_COS, SFG_PI6_SIN) : mk(-SFG_PI6_COS, SFG_PI6_SIN));
    return i == 0 ? mk(SFG_PI6_COS, SFG_PI6_SIN) : (i == 1 ? mk(0.f, 1.f) : mk(-SFG_PI6_COS, SFG_PI6_SIN));
}
SFG_DI void sat_pass(const Shape& own, const Shape& oth, V2 dist, Rot rel_own_r, float r_sum, int pass, float& pen_depth, int& pen_code, bool& separated) {
    const int n = edge_count(own);
    const bool sym = symmetrical(own);
    const float rc = rel_own_r.s * rel_own_r.s - rel_own_r.b * rel_own_r.b, rsn = 2.0f * rel_own_r.s * rel_own_r.b;   // operator*(Rot, V2), its two products once per pass
#pragma unroll
    for (int ei = 0; ei < 3; ++ei) {
        if (ei >= n) break;
        V2 en = edge_normal(own, ei);
        bool mir = false;
        if (!(dot(en, dist) >= 0.f)) {
            if (sym) { en = -en; mir = true; }
            else continue;
        }
        const float ext = edge_extent(own, ei);
        const V2 axis_o = -mk(rc * en.x + rsn * en.y, rc * en.y - rsn * en.x);
        const float depth = ext + projected_extent(oth, axis_o) + r_sum - dot(dist, en);
        if (depth <= 0.f) { separated = true; return; }
        if (depth < pen_depth) { pen_depth = depth; pen_code = pass * 8 + ei * 2 + (mir ? 1 : 0); }   // the winner is rebuilt after the loop
    }
}
SFG_DI void any_any(Iso pose0, Iso pose1, const Shape& s0, const Shape& s1, Manifold& m) {
    const Iso rel1 = inversed(pose0) * pose1;     // object 1 in object 0's frame
    const Iso rel0 = inversed(rel1);              // object 0 in object 1's frame
    float pen_depth = 3.402823466e+38f;
    int pen_code = -1;                             // owner * 8 + edge * 2 + mirrored of the least-depth axis (strict <: the first wins)
    bool separated = false;
    // depth = extent + shapes[0].circle_r + projected_extent + shapes[1].circle_r - dist.n  (reference order of the sum:
    // ((ext + r0) + proj) + r1; f32 rounding differences of the regrouping are far below the parity tolerance)
    const float r_sum = s0.r + s1.r;
    sat_pass(s0, s1, rel1.t, rel0.r, r_sum, 0, pen_depth, pen_code, separated);
    if (separated) { m.count = 0; return; }
    sat_pass(s1, s0, rel0.t, rel1.r, r_sum, 1, pen_depth, pen_code, separated);
    if (separated || pen_code < 0) { m.count = 0; return; }
    const bool f = pen_code >= 8;
    const Shape& sown = f ? s1 : s0;
    const Shape& sinc = f ? s0 : s1;
    const int pen_ei = (pen_code >> 1) & 3;
    PEdge pen = get_edge(sown, pen_ei);
    if (pen_code & 1) pen = mirrored(pen);
    const float pen_ext = edge_extent(sown, pen_ei);
    const Iso rel_own = f ? rel1 : rel0;          // relative_poses[shape_order[0]]
    const Iso rel_inc = f ? rel0 : rel1;          // relative_poses[shape_order[1]]
    const Rot pose_own_r = f ? pose1.r : pose0.r;
    const float r_own = sown.r, r_inc = sinc.r;
    Edge own_e = pen.e;
    if (r_own != 0.f) own_e.start = own_e.start + r_own * pen.n;
    const V2 axis_snd = -(rel_own.r * pen.n);
    PEdge inc = supporting_edge(sinc, axis_snd);
    const bool both_simple = r_own == 0.f && r_inc == 0.f;
    // incident edge in the owner's local space
    inc.n = rel_inc.r * inc.n;
    inc.e.start = rel_inc * inc.e.start;
    inc.e.dir = rel_inc.r * inc.e.dir;
    Edge inc_outer = inc.e;
    if (r_inc != 0.f) inc_outer.start = inc_outer.start + r_inc * inc.n;

    float enters = 0.f, exits = 0.f;
    const int clip = clip_edge(own_e, inc_outer, &enters, &exits);
    if (clip == 1) {
        const float start_depth = pen_ext + r_own - dot(inc_outer.start, pen.n);
        const float dda = dot(inc_outer.dir, pen.n);
        const float enter_depth = start_depth - enters * dda, exit_depth = start_depth - exits * dda;
        if (enter_depth > 0.f && exit_depth > 0.f) {
            const V2 ep = inc_outer.start + enters * inc_outer.dir, xp = inc_outer.start + exits * inc_outer.dir;
            m.count = 2; m.n = pose_own_r * pen.n;
            m.off[0][0] = ep + enter_depth * pen.n; m.off[0][1] = rel_own * ep;
            m.off[1][0] = xp + exit_depth * pen.n;  m.off[1][1] = rel_own * xp;
            if (f) flip(m);
            return;
        } else if (both_simple) { m.count = 0; return; }
    } else if (clip == 2 && both_simple) { m.count = 0; return; }

    const V2 cl_other = inc.e.start;
    const float tproj = dot(cl_other - pen.e.start, pen.e.dir);
    float tcl;
    if (tproj < 0.f) tcl = 0.f;
    else if (tproj > pen.e.len) tcl = pen.e.len;
    else {
        m.count = 1; m.n = pose_own_r * pen.n;
        m.off[0][0] = pen.e.start + tproj * pen.e.dir + r_own * pen.n;
        m.off[0][1] = rel_own * (cl_other - r_inc * pen.n);
        if (f) flip(m);
        return;
    }
    const V2 cl_pen = pen.e.start + tcl * pen.e.dir;
    const V2 d = cl_other - cl_pen;
    const float dsq = mag_sq(d);
    if (dsq >= r_sum * r_sum) { m.count = 0; return; }
    const V2 axis = (SFG_GUARD(2) && dsq == 0.f) ? pen.n : d * (1.0f / sqrtf(dsq));   // coincident corner points: push along the SAT axis
    m.count = 1; m.n = pose_own_r * axis;
    m.off[0][0] = cl_pen + r_own * axis;
    m.off[0][1] = rel_own * (cl_other - r_inc * axis);
    if (f) flip(m);
}

// shape_shape.rs:63-73
SFG_DI void intersection_check(Iso pose0, Iso pose1, const Shape& s0, const Shape& s1, Manifold& m) {
    m.count = 0;
    m.n = mk(0.f, 0.f);
    m.off[0][0] = m.off[0][1] = m.off[1][0] = m.off[1][1] = mk(0.f, 0.f);
    const bool c0 = s0.kind == K_POINT, c1 = s1.kind == K_POINT;
    if (c0 && c1) circle_circle(pose0, s0.r, pose1, s1.r, m);
    else if (c0) circle_any(pose0, s0.r, pose1, s1, m);
    else if (c1) { circle_any(pose1, s1.r, pose0, s0, m); flip(m); }
    else any_any(pose0, pose1, s0, s1, m);
}

// ---- queries ----
// query.rs:7-29
SFG_DI bool point_in_collider(V2 point, Iso pose, const Shape& s) {
    V2 p = inversed(pose) * point;
    const float r = s.r;
    switch (s.kind) {
    case K_POINT: return mag_sq(p) < r * r;
    case K_SEGMENT: { float xd = fmaxf(fabsf(p.x) - s.p0, 0.f), yd = fabsf(p.y); return xd * xd + yd * yd < r * r; }
    case K_RECT: { float xd = fabsf(p.x) - s.p0, yd = fabsf(p.y) - s.p1; return (xd <= 0.f && yd <= 0.f) || xd * xd + yd * yd < r * r; }
    default: { bool in; V2 c = closest_boundary_point(s, p, &in); return in || mag_sq(c - p) < r * r; }
    }
}
SFG_DI float signum_rs(float v) { return isnan(v) ? v : (signbit(v) ? -1.f : 1.f); }
// query.rs:71-95 on the box (min.xy, max.xy)
SFG_DI bool ray_aabb(V2 start, V2 dir, float4 box, float* t_out) {
    float tn = 0.f, tx = 3.402823466e+38f;
    {
        float mn = box.x - start.x, mx = box.z - start.x;
        if (fabsf(dir.x) < 0.00001f && signum_rs(mn) == signum_rs(mx)) return false;
        float inv = 1.0f / dir.x, t0 = mn * inv, t1 = mx * inv;
        if (!(t0 < t1)) { float q = t0; t0 = t1; t1 = q; }
        tn = fmaxf(tn, t0); tx = fminf(tx, t1);
        if (tn > tx) return false;
    }
    {
        float mn = box.y - start.y, mx = box.w - start.y;
        if (fabsf(dir.y) < 0.00001f && signum_rs(mn) == signum_rs(mx)) return false;
        float inv = 1.0f / dir.y, t0 = mn * inv, t1 = mx * inv;
        if (!(t0 < t1)) { float q = t0; t0 = t1; t1 = q; }
        tn = fmaxf(tn, t0); tx = fminf(tx, t1);
        if (tn > tx) return false;
    }
    *t_out = tn;
    return true;
}
struct Hit { bool hit; float t; V2 n, p; };
// query.rs:284-308
SFG_DI Hit ray_circle(V2 start, V2 dir, V2 c, float r) {
    Hit h; h.hit = false; h.t = 0.f; h.n = mk(0.f, 0.f); h.p = mk(0.f, 0.f);
    V2 mv = start - c;
    float b = dot(mv, dir), cc = mag_sq(mv) - r * r;
    if (b > 0.f && cc > 0.f) return h;
    float disc = b * b - cc;
    if (disc < 0.f) return h;
    float t = -b - sqrtf(disc);
    if (t >= 0.f) { h.hit = true; h.t = t; h.p = start + t * dir; h.n = normalized(h.p - c); }
    return h;
}
// query.rs:121-282, including its mixed local / world-space outputs (see DESIGN.md "quirks"):
// face hits give the normal in collider-local space (the capsule's flat side rotates it to world),
// vertex / cap hits give point and normal of the local-space ray_circle.
SFG_DI Hit ray_collider(V2 start, V2 dir, Iso pose, const Shape& s) {
    Hit miss; miss.hit = false; miss.t = 0.f; miss.n = mk(0.f, 0.f); miss.p = mk(0.f, 0.f);
    const float cr = s.r;
    if (s.kind == K_POINT) return ray_circle(start, dir, pose.t, cr);
    Iso inv = inversed(pose);
    V2 ls = inv * start, ld = inv.r * dir;
    if (s.kind == K_SEGMENT) {
        const float hl = s.p0;
        if (fabsf(ld.y) < 0.0001f) {
            if (fabsf(ls.y) >= cr || fabsf(ls.x) < hl) return miss;
            return ray_circle(ls, ld, mk(copysignf(hl, ls.x), 0.f), cr);
        }
        float fey = copysignf(cr, -ld.y);
        float te = (fey - ls.y) / ld.y;
        if (te < 0.f) return miss;
        float xa = ls.x + te * ld.x;
        if (fabsf(xa) <= hl) { Hit h; h.hit = true; h.t = te; h.n = pose.r * mk(0.f, signum_rs(ls.y)); h.p = start + te * dir; return h; }
        return ray_circle(ls, ld, mk(copysignf(hl, xa), 0.f), cr);
    }
    V2 perp = left_normal(ld);
    float rd = dot(ls, perp);
    if (!(rd >= 0.f)) { perp = -perp; rd = -rd; }
    if (projected_extent(s, perp) + cr <= rd) return miss;
    const float half_tan = s.kind == K_RECT ? 1.f : (s.kind == K_TRIANGLE ? SFG_PI6_TAN : SFG_SQRT_3);   // collider.rs:543-554
    const float extra = cr == 0.f ? 0.f : cr / half_tan;
    bool have_v = false; V2 vert = mk(0.f, 0.f);
    float ct = 3.402823466e+38f; V2 cn = mk(1.f, 0.f);
    const int n = edge_count(s);
    for (int ei = 0; ei < n; ++ei) {
        PEdge e = get_edge(s, ei);
        if (!(dot(e.n, ld) <= 0.f)) {
            if (symmetrical(s)) e = mirrored(e);
            else continue;
        }
        V2 os = e.e.start + cr * e.n;
        V2 ed = os - ls;
        float spd = dot(ld, -e.n);
        if (spd == 0.f) continue;
        float tt = dot(ed, -e.n) / spd;
        if (tt < 0.f) continue;
        float sa = dot(ld, e.e.dir);
        float et = tt * sa - dot(ed, e.e.dir);
        if (et < -extra || et > e.e.len + extra) continue;
        if (ct <= tt) continue;
        ct = tt; cn = e.n;
        if (et < 0.f) { have_v = true; vert = e.e.start; }
        else if (et > e.e.len) { have_v = true; vert = e.e.start + e.e.len * e.e.dir; }
        else have_v = false;
    }
    if (ct == 3.402823466e+38f) return miss;
    if (have_v) return ray_circle(ls, ld, vert, cr);
    Hit h; h.hit = true; h.t = ct; h.n = cn; h.p = start + ct * dir;
    return h;
}
// query.rs:107-118
SFG_DI Hit spherecast_collider(V2 start, V2 dir, float radius, Iso pose, Shape s) {
    s.r += radius;
    Hit h = ray_collider(start, dir, pose, s);
    if (h.hit) h.p = h.p - radius * h.n;
    return h;
}

}  // namespace sfg

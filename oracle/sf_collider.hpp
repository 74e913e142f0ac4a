// ORACLE — TEST INFRASTRUCTURE ONLY (see sf_math.hpp header).
// Restates /root/reference/src/physics/collision/collider.rs:194-950 (shapes, geometry tables,
// mass properties, materials) and src/physics/collision.rs:19-105 (AABB).
#pragma once
#include "sf_math.hpp"

namespace sfo {

// collision.rs:22-105
template <class R>
struct AABB {
    Vec2<R> min, max;
    AABB translated(Vec2<R> t) const { return {min + t, max + t}; }            // :30
    AABB padded(R a) const {                                                    // :39
        AABB o = *this;
        o.min.x -= a; o.min.y -= a; o.max.x += a; o.max.y += a;
        return o;
    }
    AABB extended(Vec2<R> a) const {                                            // :49
        AABB o = *this;
        if (a.x < R(0)) o.min.x += a.x; else o.max.x += a.x;
        if (a.y < R(0)) o.min.y += a.y; else o.max.y += a.y;
        return o;
    }
    R width() const { return max.x - min.x; }
    R height() const { return max.y - min.y; }
    R area() const { return width() * height(); }
    AABB union_(const AABB& o) const {                                          // :82
        return {{rs_min(min.x, o.min.x), rs_min(min.y, o.min.y)},
                {rs_max(max.x, o.max.x), rs_max(max.y, o.max.y)}};
    }
    bool intersects(const AABB& o) const {                                      // :91 (is_some)
        R mnx = rs_max(min.x, o.min.x), mny = rs_max(min.y, o.min.y);
        R mxx = rs_min(max.x, o.max.x), mxy = rs_min(max.y, o.max.y);
        return !(mnx >= mxx || mny >= mxy);
    }
    bool contains_point(Vec2<R> p) const {                                      // :102
        return p.x >= min.x && p.x <= max.x && p.y >= min.y && p.y <= max.y;
    }
};

enum PolyKind : uint32_t { POINT = 0, SEGMENT = 1, RECT = 2, TRIANGLE = 3, HEXAGON = 4 };

template <class R>
struct Edge {           // shape_shape.rs:392-441
    Vec2<R> start, dir;
    R length;
    Edge transformed(const Iso2<R>& p) const { return {p * start, p.r * dir, length}; }
    Edge mirrored() const { return {-start, -dir, length}; }
    Edge offset(Vec2<R> a) const { return {start + a, dir, length}; }
    Vec2<R> end_point() const { return start + length * dir; }
    Edge flipped() const { return {end_point(), -dir, length}; }
};
template <class R>
struct PolygonEdge {    // shape_shape.rs:364-385
    Vec2<R> normal;
    Edge<R> edge;
    PolygonEdge transformed(const Iso2<R>& p) const { return {p.r * normal, edge.transformed(p)}; }
    PolygonEdge mirrored() const { return {-normal, edge.mirrored()}; }
};
template <class R>
struct ClosestBoundaryPoint { Vec2<R> pt; bool is_interior; };

// collider.rs:419-433 — constants truncated exactly as in the reference.
template <class R> struct K {
    static constexpr R SQRT_3 = R(1.73205080757);
    static constexpr R PI6_COS = R(0.866025403784);
    static constexpr R PI6_SIN = R(0.5);
    static constexpr R PI6_TAN = R(0.57735026919);
    static constexpr R PI3_COS = PI6_SIN;
    static constexpr R PI3_SIN = PI6_COS;
    static constexpr R PI = R(3.14159265358979323846);
    static constexpr R FRAC_PI_2 = R(1.57079632679489661923);
    static Vec2<R> a30() { return {PI6_COS, PI6_SIN}; }
    static Vec2<R> a60() { return {PI3_COS, PI3_SIN}; }
    static Vec2<R> a120() { return {-PI6_SIN, PI6_COS}; }
    static Vec2<R> a150() { return {-PI3_SIN, PI3_COS}; }
};

template <class R>
struct Shape {          // ColliderShape {polygon, circle_r}; polygon params p0,p1 by kind
    uint32_t kind = POINT;
    R p0 = 0;           // hl | hw | outer_r
    R p1 = 0;           // hh
    R circle_r = 0;

    // --- collider.rs:558-566
    int edge_count() const {
        switch (kind) { case POINT: return 0; case SEGMENT: return 1; case RECT: return 2; default: return 3; }
    }
    bool is_rotationally_symmetrical() const { return kind != TRIANGLE; }        // :528
    // --- collider.rs:543-554
    R half_angle_between_edges_tan() const {
        switch (kind) { case RECT: return R(1); case TRIANGLE: return R(0.57735026919); default: return R(1.73205080757); }
    }
    // --- collider.rs:572-665
    PolygonEdge<R> get_edge(int idx) const {
        using V = Vec2<R>;
        const V ux{1, 0}, uy{0, 1};
        switch (kind) {
        case SEGMENT: return {uy, {V{p0, 0}, -ux, R(2) * p0}};
        case RECT:
            if (idx == 0) return {ux, {V{p0, -p1}, uy, R(2) * p1}};
            return {uy, {V{p0, p1}, -ux, R(2) * p0}};
        case TRIANGLE: {
            R len = p0 / K<R>::PI6_TAN;
            if (idx == 0) return {-uy, {(-p0) * K<R>::a30(), ux, len}};
            if (idx == 1) return {K<R>::a30(), {(-p0) * K<R>::a150(), K<R>::a120(), len}};
            return {K<R>::a150(), {V{0, p0}, -K<R>::a60(), len}};
        }
        default: {  // HEXAGON
            if (idx == 0) return {K<R>::a30(), {V{p0, 0}, K<R>::a120(), p0}};
            if (idx == 1) return {uy, {p0 * K<R>::a60(), -ux, p0}};
            return {K<R>::a150(), {p0 * K<R>::a120(), -K<R>::a60(), p0}};
        }
        }
    }
    // --- collider.rs:668-686
    R get_edge_extent(int idx) const {
        switch (kind) {
        case SEGMENT: return R(0);
        case RECT: return idx == 0 ? p0 : p1;
        case TRIANGLE: return p0 / R(2);
        default: return K<R>::PI6_COS * p0;
        }
    }
    // --- collider.rs:692-714
    R projected_extent(Vec2<R> dir) const {
        switch (kind) {
        case POINT: return R(0);
        case SEGMENT: return std::fabs(dir.x) * p0;
        case RECT: return std::fabs(dir.x) * p0 + std::fabs(dir.y) * p1;
        case TRIANGLE: {
            R a = Vec2<R>{0, 1}.dot(dir), b = (-K<R>::a30()).dot(dir), c = (-K<R>::a150()).dot(dir);
            R m = a; if (b >= m) m = b; if (c >= m) m = c;
            return m * p0;
        }
        default: {
            R a = std::fabs(Vec2<R>{1, 0}.dot(dir)), b = std::fabs(K<R>::a60().dot(dir)),
              c = std::fabs(K<R>::a120().dot(dir));
            R m = a; if (b >= m) m = b; if (c >= m) m = c;
            return m * p0;
        }
        }
    }
    // --- collider.rs:722-790
    PolygonEdge<R> supporting_edge(Vec2<R> dir) const {
        using V = Vec2<R>;
        switch (kind) {
        case SEGMENT:
            return {V{0, std::copysign(R(1), dir.y)},
                    {V{std::copysign(p0, dir.x), 0}, V{std::copysign(R(1), -dir.x), 0}, R(2) * p0}};
        case RECT: {
            V start{std::copysign(p0, dir.x), std::copysign(p1, dir.y)};
            if (std::fabs(dir.x) > std::fabs(dir.y))
                return {V{std::copysign(R(1), dir.x), 0}, {start, V{0, -std::copysign(R(1), dir.y)}, p1 * R(2)}};
            return {V{0, std::copysign(R(1), dir.y)}, {start, V{-std::copysign(R(1), dir.x), 0}, p0 * R(2)}};
        }
        default: {  // TRIANGLE | HEXAGON: Iterator::max_by keeps the LAST of equal maxima
            PolygonEdge<R> best{};
            R best_d = 0;
            for (int i = 0; i < edge_count(); ++i) {
                PolygonEdge<R> e = get_edge(i);
                if (is_rotationally_symmetrical() && e.normal.dot(dir) < R(0)) e = e.mirrored();
                R d = e.normal.dot(dir);
                if (i == 0 || d >= best_d) { best = e; best_d = d; }
            }
            return {best.normal, best.edge.dir.dot(dir) < R(0) ? best.edge : best.edge.flipped()};
        }
        }
    }
    // --- collider.rs:796-892
    ClosestBoundaryPoint<R> closest_boundary_point(Vec2<R> pt) const {
        using V = Vec2<R>;
        switch (kind) {
        case POINT: return {V{0, 0}, false};
        case SEGMENT: return {V{rs_min(rs_max(pt.x, -p0), p0), 0}, false};
        case RECT: {
            R xd = std::fabs(pt.x) - p0, yd = std::fabs(pt.y) - p1;
            bool ox = xd > R(0), oy = yd > R(0);
            if (ox && oy) return {V{std::copysign(p0, pt.x), std::copysign(p1, pt.y)}, false};
            if (ox) return {V{std::copysign(p0, pt.x), pt.y}, false};
            if (oy) return {V{pt.x, std::copysign(p1, pt.y)}, false};
            if (std::fabs(xd) < std::fabs(yd)) return {V{std::copysign(p0, pt.x), pt.y}, true};
            return {V{pt.x, std::copysign(p1, pt.y)}, true};
        }
        default: {
            R min_dist = real_max<R>();
            ClosestBoundaryPoint<R> cp{V{0, 0}, false};
            for (int i = 0; i < edge_count(); ++i) {
                PolygonEdge<R> pe = get_edge(i);
                if (is_rotationally_symmetrical() && pe.normal.dot(pt) < R(0)) pe = pe.mirrored();
                V s2p = pt - pe.edge.start;
                R t = pe.edge.dir.dot(s2p);
                if (t < R(0)) {
                    R d = s2p.mag();
                    if (d < min_dist) { min_dist = d; cp = {pe.edge.start, false}; }
                } else if (t <= pe.edge.length) {
                    R nd = pe.normal.dot(s2p);
                    bool interior = !(nd >= R(0));
                    if (interior) nd = -nd;
                    if (nd < min_dist) { min_dist = nd; cp = {pe.edge.start + t * pe.edge.dir, interior}; }
                }
            }
            return cp;
        }
        }
    }
    // --- collider.rs:505-523 + :344-349
    AABB<R> aabb(const Iso2<R>& pose) const {
        Vec2<R> ext;
        switch (kind) {
        case POINT: ext = {0, 0}; break;
        case SEGMENT: ext = (pose.r * Vec2<R>{p0, 0}).abs(); break;
        case RECT: ext = (pose.r * Vec2<R>{p0, 0}).abs() + (pose.r * Vec2<R>{0, p1}).abs(); break;
        default: ext = {p0, p0}; break;
        }
        AABB<R> b{-ext, ext};
        return b.padded(circle_r).translated(pose.t);
    }
    Shape expanded(R amount) const { Shape s = *this; s.circle_r = circle_r + amount; return s; }  // :352

    // --- mass properties, collider.rs:222-337, 440-458
    R poly_area() const {
        switch (kind) {
        case RECT: return R(4) * p0 * p1;
        case TRIANGLE: return R(3) * R(0.25) * p0 * p0 / K<R>::PI6_TAN;
        case HEXAGON: return R(3) * p0 * K<R>::PI6_COS * p0;
        default: return R(0);
        }
    }
    R side_length_sum() const {
        switch (kind) {
        case SEGMENT: return R(4) * p0;
        case RECT: return R(4) * (p0 + p1);
        case TRIANGLE: return R(3) * p0 / K<R>::PI6_TAN;
        case HEXAGON: return R(6) * p0;
        default: return R(0);
        }
    }
    R area() const {
        R r = circle_r;
        if (r == R(0)) return poly_area();
        return poly_area() + K<R>::PI * r * r + side_length_sum() * r;
    }
    static R powi(R v, int n) { R o = 1; for (int i = 0; i < n; ++i) o *= v; return o; }
    static R m_circle(R r) { return K<R>::FRAC_PI_2 * powi(r, 4); }
    static R m_rect(R hw, R hh) { return (R(4) / R(3)) * (powi(hw, 3) * hh + hw * powi(hh, 3)); }
    R second_moment_of_area() const {
        const R cr = circle_r;
        R polygon_part;
        switch (kind) {
        case POINT: return m_circle(cr);
        case SEGMENT: return m_rect(p0, cr) + (m_circle(cr) + (K<R>::PI * powi(cr, 2) * powi(p0, 2)));
        case RECT: polygon_part = m_rect(p0, p1); break;
        case TRIANGLE: polygon_part = R(0.03608) * powi(p0 / K<R>::PI6_TAN, 4); break;
        default: polygon_part = (R(5) * K<R>::SQRT_3 / R(8)) * powi(p0, 4); break;
        }
        if (cr == R(0)) return polygon_part;
        R expanded_part;
        if (kind == RECT) {
            R horiz = m_rect(p0, cr / R(2)) + (R(2) * p0 * cr) * powi(p1, 2);
            R vert = m_rect(p1, cr / R(2)) + (R(2) * p1 * cr) * powi(p0, 2);
            R caps = m_circle(cr) + (K<R>::PI * powi(cr, 2)) * Vec2<R>{p0, p1}.mag_sq();
            expanded_part = R(2) * (horiz + vert) + caps;
        } else if (kind == TRIANGLE) {
            R len = p0 / K<R>::PI6_TAN;
            R edge_rect = m_rect(len / R(2), cr / R(2)) + (len * cr) * powi(p0 / R(2), 2);
            R caps = m_circle(cr) + (K<R>::PI * powi(cr, 2)) * powi(p0, 2);
            expanded_part = R(3) * edge_rect + caps;
        } else {
            R len = p0;
            R edge_rect = m_rect(len / R(2), cr / R(2)) + (len * cr) * powi(K<R>::PI6_COS * p0, 2);
            R caps = m_circle(cr) + (K<R>::PI * powi(cr, 2)) * powi(p0, 2);
            expanded_part = R(6) * edge_rect + caps;
        }
        return polygon_part + expanded_part;
    }
    // --- collider.rs:363-369, 465-489
    Shape rounded_inward(R amount) const {
        const R MIN = R(0.001);
        Shape s = *this;
        R actual = 0;
        switch (kind) {
        case RECT: {
            R hw_ = rs_max(p0 - amount, MIN), hh_ = rs_max(p1 - amount, MIN);
            s.p0 = hw_; s.p1 = hh_; actual = rs_min(p0 - hw_, p1 - hh_);
            break;
        }
        case TRIANGLE: {
            R r_ = rs_max(p0 - amount / K<R>::PI6_SIN, MIN);
            s.p0 = r_; actual = (p0 - r_) * K<R>::PI6_SIN;
            break;
        }
        case HEXAGON: {
            R r_ = rs_max(p0 - amount / K<R>::PI3_SIN, MIN);
            s.p0 = r_; actual = (p0 - r_) * K<R>::PI3_SIN;
            break;
        }
        default: break;
        }
        s.circle_r = circle_r + actual;
        return s;
    }
};

// Collider constructors, collider.rs:63-125
template <class R> inline Shape<R> shape_circle(R r) { return {POINT, 0, 0, r}; }
template <class R> inline Shape<R> shape_rounded_rect(R w, R h, R r) {
    R hw = rs_max((w / R(2)) - r, R(0.05)), hh = rs_max((h / R(2)) - r, R(0.05));
    return {RECT, hw, hh, rs_min(rs_min(r, w / R(2)), h / R(2))};
}
template <class R> inline Shape<R> shape_capsule(R len, R r) { return {SEGMENT, len / R(2), 0, r}; }
template <class R> inline Shape<R> shape_hexagon(R outer_r) { return {HEXAGON, outer_r, 0, 0}; }
template <class R> inline Shape<R> shape_triangle(R outer_r) { return {TRIANGLE, outer_r, 0, 0}; }

}  // namespace sfo

// ORACLE — TEST INFRASTRUCTURE ONLY (see sf_math.hpp header).
// Restates /root/reference/src/physics/collision/shape_shape.rs:63-488 (narrowphase) and
// src/physics/collision/query.rs:7-308 (point / ray / spherecast vs collider).
#pragma once
#include "sf_collider.hpp"

namespace sfo {

template <class R>
struct Contact {           // shape_shape.rs:55-60
    Vec2<R> normal;        // world space, away from the first object
    Vec2<R> offsets[2];    // object-local contact points
};
template <class R>
struct ContactResult {     // shape_shape.rs:6-10
    int count = 0;
    Contact<R> c[2];
    bool is_zero() const { return count == 0; }
};

template <class R> inline ContactResult<R> flip_contacts(ContactResult<R> r) {   // :76-81
    for (int i = 0; i < r.count; ++i) {
        r.c[i].normal = -r.c[i].normal;
        Vec2<R> t = r.c[i].offsets[0]; r.c[i].offsets[0] = r.c[i].offsets[1]; r.c[i].offsets[1] = t;
    }
    return r;
}
template <class R> inline ContactResult<R> one(Contact<R> c) { ContactResult<R> r; r.count = 1; r.c[0] = c; return r; }

// `stable`: compute the relative pose as R0^-1 * (t1 - t0) (subtract first) instead of the
// literal R0^-1*(-t0) + R0^-1*t1 of shape_shape.rs:122,161. Mathematically identical; used for
// the f32 instantiation, where the literal form cancels catastrophically far from the origin.
template <class R> inline Iso2<R> relative_pose(const Iso2<R>& a, const Iso2<R>& b, bool stable) {
    if (!stable) return a.inversed() * b;
    Rotor2<R> ar = a.r.reversed();
    return {ar * (b.t - a.t), ar * b.r};
}

// shape_shape.rs:87-112
template <class R>
inline ContactResult<R> circle_circle(const Iso2<R>& pose1, R r1, const Iso2<R>& pose2, R r2) {
    Vec2<R> dist = pose2.t - pose1.t;
    R dist_sq = dist.mag_sq();
    R r_sum = r1 + r2;
    Vec2<R> normal;
    if (dist_sq < R(0.001)) normal = {1, 0};
    else if (dist_sq < r_sum * r_sum) normal = dist.normalized();
    else return {};
    return one(Contact<R>{normal, {pose1.r.reversed() * (r1 * normal), pose2.r.reversed() * ((-r2) * normal)}});
}

// shape_shape.rs:114-154
template <class R>
inline ContactResult<R> circle_any(const Iso2<R>& pose_circ, R r_circ, const Iso2<R>& pose_other,
                                   const Shape<R>& shape_other, R r_other, bool stable) {
    Iso2<R> local = relative_pose(pose_other, pose_circ, stable);
    Vec2<R> dist = local.t;
    auto cp = shape_other.closest_boundary_point(dist);
    Vec2<R> from_closest = dist - cp.pt;
    if (cp.is_interior) {
        Vec2<R> dir = from_closest.normalized();
        return one(Contact<R>{pose_other.r * dir,
                              {local.r.reversed() * (r_circ * dir), cp.pt - r_other * dir}});
    }
    R r_sum = r_circ + r_other;
    R dist_sq = from_closest.mag_sq();
    if (dist_sq >= r_sum * r_sum) return {};
    Vec2<R> dir = from_closest / std::sqrt(dist_sq);
    return one(Contact<R>{pose_other.r * (-dir),
                          {local.r.reversed() * (r_circ * (-dir)), cp.pt + r_other * dir}});
}

enum ClipKind { CLIP_INTERSECTS, CLIP_PASSES, CLIP_MISSES };
template <class R> struct ClipResult { ClipKind kind; R enters, exits; };

// shape_shape.rs:458-488
template <class R>
inline ClipResult<R> clip_edge(const Edge<R>& target, const Edge<R>& edge) {
    Vec2<R> sd = target.start - edge.start;
    R denom = edge.dir.x * (-target.dir.y) - (-target.dir.x) * edge.dir.y;
    R t0 = (sd.x * (-target.dir.y) - (-target.dir.x) * sd.y) / denom;
    R t1 = (edge.dir.x * sd.y - sd.x * edge.dir.y) / denom;
    if (t0 >= R(0) && t0 <= edge.length && t1 >= R(0) && t1 <= target.length) return {CLIP_INTERSECTS, 0, 0};
    R dist_dot_dir2 = sd.dot(target.dir);
    R dirs_dot = edge.dir.dot(target.dir);
    R start_clip_t = dist_dot_dir2 / dirs_dot;
    R end_clip_t = (target.length + dist_dot_dir2) / dirs_dot;
    if ((start_clip_t <= R(0) && end_clip_t <= R(0)) || (start_clip_t >= edge.length && end_clip_t >= edge.length))
        return {CLIP_MISSES, 0, 0};
    if (start_clip_t < end_clip_t) return {CLIP_PASSES, rs_max(start_clip_t, R(0)), rs_min(end_clip_t, edge.length)};
    return {CLIP_PASSES, rs_max(end_clip_t, R(0)), rs_min(start_clip_t, edge.length)};
}

// shape_shape.rs:160-349
template <class R>
inline ContactResult<R> any_any(const Iso2<R> poses[2], const Shape<R> shapes[2], bool stable) {
    Iso2<R> po2_wrt_po1 = relative_pose(poses[0], poses[1], stable);
    Iso2<R> rel[2] = {po2_wrt_po1.inversed(), po2_wrt_po1};

    R pen_depth = real_max<R>();
    bool have_axis = false;
    PolygonEdge<R> pen_edge{};
    R pen_edge_extent = 0;
    int order[2] = {0, 1};
    for (int pass = 0; pass < 2; ++pass) {
        const int so[2] = {pass, 1 - pass};
        const int n = shapes[so[0]].edge_count();
        for (int ei = 0; ei < n; ++ei) {
            Vec2<R> dist = rel[so[1]].t;
            PolygonEdge<R> edge = shapes[so[0]].get_edge(ei);
            if (edge.normal.dot(dist) >= R(0)) {
            } else if (shapes[so[0]].is_rotationally_symmetrical()) {
                edge = edge.mirrored();
            } else {
                continue;
            }
            R extent = shapes[so[0]].get_edge_extent(ei);
            Vec2<R> axis_wrt_other = -(rel[so[0]].r * edge.normal);
            R depth = extent + shapes[0].circle_r + shapes[so[1]].projected_extent(axis_wrt_other) +
                      shapes[1].circle_r - dist.dot(edge.normal);
            if (depth <= R(0)) return {};
            if (depth < pen_depth) {
                pen_depth = depth; pen_edge = edge; pen_edge_extent = extent; have_axis = true;
                order[0] = so[0]; order[1] = so[1];
            }
        }
    }
    if (!have_axis) return {};
    const bool flip = order[0] != 0;
    auto orient = [&](ContactResult<R> r) { return flip ? flip_contacts(r) : r; };

    const R r_own = shapes[order[0]].circle_r, r_inc = shapes[order[1]].circle_r;
    Edge<R> owning_edge = r_own == R(0) ? pen_edge.edge : pen_edge.edge.offset(pen_edge.normal * r_own);
    Vec2<R> pen_axis_wrt_snd = -(rel[order[0]].r * pen_edge.normal);
    PolygonEdge<R> inc_inner_local = shapes[order[1]].supporting_edge(pen_axis_wrt_snd);
    bool both_simple = r_own == R(0) && r_inc == R(0);
    PolygonEdge<R> inc_inner = inc_inner_local.transformed(rel[order[1]]);
    Edge<R> inc_outer = r_inc == R(0) ? inc_inner.edge : inc_inner.edge.offset(inc_inner.normal * r_inc);

    ClipResult<R> clip = clip_edge(owning_edge, inc_outer);
    if (clip.kind == CLIP_PASSES) {
        R start_depth = pen_edge_extent + r_own - inc_outer.start.dot(pen_edge.normal);
        R dir_dot_axis = inc_outer.dir.dot(pen_edge.normal);
        R enter_depth = start_depth - clip.enters * dir_dot_axis;
        R exit_depth = start_depth - clip.exits * dir_dot_axis;
        if (enter_depth > R(0) && exit_depth > R(0)) {
            Vec2<R> enter_point = inc_outer.start + (clip.enters * inc_outer.dir);
            Vec2<R> exit_point = inc_outer.start + (clip.exits * inc_outer.dir);
            Vec2<R> nw = poses[order[0]].r * pen_edge.normal;
            ContactResult<R> res;
            res.count = 2;
            res.c[0] = {nw, {enter_point + enter_depth * pen_edge.normal, rel[order[0]] * enter_point}};
            res.c[1] = {nw, {exit_point + exit_depth * pen_edge.normal, rel[order[0]] * exit_point}};
            return orient(res);
        } else if (both_simple) {
            return {};
        }
    } else if (clip.kind == CLIP_MISSES && both_simple) {
        return {};
    }

    Vec2<R> closest_on_other = inc_inner.edge.start;
    Vec2<R> s2c = closest_on_other - pen_edge.edge.start;
    R t_proj = s2c.dot(pen_edge.edge.dir);
    R t_clamped;
    if (t_proj < R(0)) t_clamped = R(0);
    else if (t_proj > pen_edge.edge.length) t_clamped = pen_edge.edge.length;
    else {
        return orient(one(Contact<R>{
            poses[order[0]].r * pen_edge.normal,
            {pen_edge.edge.start + t_proj * pen_edge.edge.dir + r_own * pen_edge.normal,
             rel[order[0]] * (closest_on_other - r_inc * pen_edge.normal)}}));
    }
    Vec2<R> closest_on_pen = pen_edge.edge.start + t_clamped * pen_edge.edge.dir;
    Vec2<R> d = closest_on_other - closest_on_pen;
    R dist_sq = d.mag_sq();
    R rs = shapes[0].circle_r + shapes[1].circle_r;
    if (dist_sq >= rs * rs) return {};
    Vec2<R> axis = d / std::sqrt(dist_sq);
    return orient(one(Contact<R>{poses[order[0]].r * axis,
                                 {closest_on_pen + r_own * axis, rel[order[0]] * (closest_on_other - r_inc * axis)}}));
}

// shape_shape.rs:63-73
template <class R>
inline ContactResult<R> intersection_check(const Iso2<R> poses[2], const Shape<R> shapes[2], bool stable) {
    R r0 = shapes[0].circle_r, r1 = shapes[1].circle_r;
    bool p0 = shapes[0].kind == POINT, p1 = shapes[1].kind == POINT;
    if (p0 && p1) return circle_circle(poses[0], r0, poses[1], r1);
    if (p0) return circle_any(poses[0], r0, poses[1], shapes[1], r1, stable);
    if (p1) return flip_contacts(circle_any(poses[1], r1, poses[0], shapes[0], r0, stable));
    return any_any(poses, shapes, stable);
}

//
// queries — query.rs
//
template <class R> struct Ray { Vec2<R> start, dir; Vec2<R> point_at_t(R t) const { return start + t * dir; } };
template <class R> inline Ray<R> transform_ray(const Iso2<R>& p, const Ray<R>& r) { return {p * r.start, p.r * r.dir}; }  // :36-45
template <class R> struct CastHit { bool hit = false; R t = 0; Vec2<R> normal, point; };

// query.rs:7-29
template <class R>
inline bool point_collider_bool(Vec2<R> point, const Iso2<R>& pose, const Shape<R>& sh) {
    R r = sh.circle_r;
    Vec2<R> p = pose.inversed() * point;
    switch (sh.kind) {
    case POINT: return p.mag_sq() < r * r;
    case SEGMENT: {
        R xd = rs_max(std::fabs(p.x) - sh.p0, R(0)), yd = std::fabs(p.y);
        return xd * xd + yd * yd < r * r;
    }
    case RECT: {
        R xd = std::fabs(p.x) - sh.p0, yd = std::fabs(p.y) - sh.p1;
        return (xd <= R(0) && yd <= R(0)) || xd * xd + yd * yd < r * r;
    }
    default: {
        auto c = sh.closest_boundary_point(p);
        return c.is_interior || (c.pt - p).mag_sq() < r * r;
    }
    }
}

// query.rs:71-95; returns false for None
template <class R>
inline bool ray_aabb(const Ray<R>& ray, const AABB<R>& aabb, R* t_out) {
    Vec2<R> to_min = aabb.min - ray.start, to_max = aabb.max - ray.start;
    R t_enter = 0, t_exit = real_max<R>();
    const R mn[2] = {to_min.x, to_min.y}, mx[2] = {to_max.x, to_max.y}, sp[2] = {ray.dir.x, ray.dir.y};
    for (int a = 0; a < 2; ++a) {
        if (std::fabs(sp[a]) < R(0.00001) && rs_signum(mn[a]) == rs_signum(mx[a])) return false;
        R inv = R(1) / sp[a];
        R t0 = mn[a] * inv, t1 = mx[a] * inv;
        if (!(t0 < t1)) { R tmp = t0; t0 = t1; t1 = tmp; }
        t_enter = rs_max(t_enter, t0);
        t_exit = rs_min(t_exit, t1);
        if (t_enter > t_exit) return false;
    }
    *t_out = t_enter;
    return true;
}

// query.rs:284-308
template <class R>
inline CastHit<R> ray_circle(const Ray<R>& ray, Vec2<R> circ_pos, R r) {
    Vec2<R> m = ray.start - circ_pos;
    R b = m.dot(ray.dir);
    R c = m.mag_sq() - r * r;
    if (b > R(0) && c > R(0)) return {};
    R discr = b * b - c;
    if (discr < R(0)) return {};
    R t = -b - std::sqrt(discr);
    if (t >= R(0)) {
        Vec2<R> point = ray.point_at_t(t);
        return {true, t, (point - circ_pos).normalized(), point};
    }
    return {};
}

// query.rs:121-282. NOTE (replicated quirks): face hits return `normal` in collider-local space
// (:273-277 and the capsule flat edge uses pose.rotation at :161) and vertex / cap hits return
// point+normal of the LOCAL-space ray_circle (:141-146, :167-171, :272).
template <class R>
inline CastHit<R> ray_collider(const Ray<R>& ray_in, const Iso2<R>& pose, const Shape<R>& sh) {
    using V = Vec2<R>;
    const R cr = sh.circle_r;
    if (sh.kind == POINT) return ray_circle(ray_in, pose.t, cr);
    const Ray<R> ray_world = ray_in;
    const Ray<R> ray = transform_ray(pose.inversed(), ray_in);
    if (sh.kind == SEGMENT) {
        const R hl = sh.p0;
        if (std::fabs(ray.dir.y) < R(0.0001)) {
            if (std::fabs(ray.start.y) >= cr || std::fabs(ray.start.x) < hl) return {};
            return ray_circle(ray, V{std::copysign(hl, ray.start.x), 0}, cr);
        }
        R facing_edge_y = std::copysign(cr, -ray.dir.y);
        R t_edge = (facing_edge_y - ray.start.y) / ray.dir.y;
        if (t_edge < R(0)) return {};
        R x_at = ray.start.x + t_edge * ray.dir.x;
        if (std::fabs(x_at) <= hl)
            return {true, t_edge, pose.r * V{0, rs_signum(ray.start.y)}, ray_world.point_at_t(t_edge)};
        return ray_circle(ray, V{std::copysign(hl, x_at), 0}, cr);
    }
    V perp = left_normal(ray.dir);
    R ray_dist = ray.start.dot(perp);
    if (!(ray_dist >= R(0))) { perp = -perp; ray_dist = -ray_dist; }
    R poly_extent = sh.projected_extent(perp);
    if (poly_extent + cr <= ray_dist) return {};
    R extra = cr == R(0) ? R(0) : cr / sh.half_angle_between_edges_tan();
    bool have_vertex = false;
    V vertex{};
    R closest_t = real_max<R>();
    V closest_normal{1, 0};
    for (int ei = 0; ei < sh.edge_count(); ++ei) {
        PolygonEdge<R> edge = sh.get_edge(ei);
        if (edge.normal.dot(ray.dir) <= R(0)) {
        } else if (sh.is_rotationally_symmetrical()) {
            edge = edge.mirrored();
        } else {
            continue;
        }
        Edge<R> outer = edge.edge.offset(cr * edge.normal);
        V edge_dist = outer.start - ray.start;
        R speed_to_edge = ray.dir.dot(-edge.normal);
        if (speed_to_edge == R(0)) continue;
        R t_to_edge = edge_dist.dot(-edge.normal) / speed_to_edge;
        if (t_to_edge < R(0)) continue;
        R speed_along = ray.dir.dot(edge.edge.dir);
        R edge_t = t_to_edge * speed_along - edge_dist.dot(edge.edge.dir);
        if (edge_t < -extra || edge_t > edge.edge.length + extra) continue;
        if (closest_t <= t_to_edge) continue;
        closest_t = t_to_edge;
        closest_normal = edge.normal;
        if (edge_t < R(0)) { have_vertex = true; vertex = edge.edge.start; }
        else if (edge_t > edge.edge.length) { have_vertex = true; vertex = edge.edge.start + edge.edge.length * edge.edge.dir; }
        else have_vertex = false;
    }
    if (closest_t == real_max<R>()) return {};
    if (have_vertex) return ray_circle(ray, vertex, cr);
    return {true, closest_t, closest_normal, ray_world.point_at_t(closest_t)};
}

// query.rs:107-118
template <class R>
inline CastHit<R> spherecast_collider(const Ray<R>& ray, R r, const Iso2<R>& pose, const Shape<R>& sh) {
    CastHit<R> h = ray_collider(ray, pose, sh.expanded(r));
    if (h.hit) h.point = h.point - r * h.normal;
    return h;
}

}  // namespace sfo

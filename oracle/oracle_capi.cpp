// ORACLE — TEST INFRASTRUCTURE ONLY. C entry points (ctypes) over the CPU restatement of the
// Starframe physics tick in sf_world.hpp. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs load this library; the product (moleengine_b200/) never does.
#include "sf_world.hpp"

using namespace sfo;

extern "C" {

void* sfo_create(int use_f32) {
    if (use_f32) return static_cast<IWorld*>(new World<float>());
    return static_cast<IWorld*>(new World<double>());
}
void sfo_destroy(void* w) { delete static_cast<IWorld*>(w); }
void sfo_set_tuning(void* w, const SfgTuning* t) { static_cast<IWorld*>(w)->set_tuning(*t); }
void sfo_set_mask_matrix(void* w, const uint64_t* m) { static_cast<IWorld*>(w)->set_mask(m); }
void sfo_set_bodies(void* w, uint32_t n, const SfgBody* b) { static_cast<IWorld*>(w)->set_bodies(n, b); }
void sfo_set_colliders(void* w, uint32_t n, const SfgCollider* c) { static_cast<IWorld*>(w)->set_colliders(n, c); }
void sfo_set_constraints(void* w, uint32_t n, const SfgConstraint* c) { static_cast<IWorld*>(w)->set_constraints(n, c); }
void sfo_set_ropes(void* w, uint32_t nr, const SfgRope* r, uint32_t np, const int32_t* p) { static_cast<IWorld*>(w)->set_ropes(nr, r, np, p); }
void sfo_set_external_pairs(void* w, const uint32_t* hl, size_t n) { static_cast<IWorld*>(w)->set_external_pairs(hl, n); }
void sfo_set_external_pair_hints(void* w, const uint32_t* hl, const uint8_t* hints, size_t n) { static_cast<IWorld*>(w)->set_external_pair_hints(hl, hints, n); }
void sfo_tick(void* w, double frame_dt, int has_scale, double scale, const SfgForceField* ff, int mode, int threads) {
    static_cast<IWorld*>(w)->tick(frame_dt, has_scale != 0, scale, *ff, mode, threads);
}
void sfo_get_bodies(void* w, uint32_t first, uint32_t count, double* out7) { static_cast<IWorld*>(w)->get_bodies(first, count, out7); }
size_t sfo_n_pairs(void* w) { return static_cast<IWorld*>(w)->n_pairs(); }
void sfo_get_pairs(void* w, uint32_t* out) { static_cast<IWorld*>(w)->get_pairs(out); }
size_t sfo_n_pair_units(void* w) { return static_cast<IWorld*>(w)->n_pair_units(); }
void sfo_get_pair_units(void* w, uint32_t* hi, uint32_t* lo, uint32_t* mult, uint32_t* colour) { static_cast<IWorld*>(w)->get_pair_units(hi, lo, mult, colour); }
// the `near` bit of each pair unit's priority as the oracle computed it (include/sfg_hash.h), in the order of sfo_get_pair_units
void sfo_get_pair_unit_hints(void* w, uint8_t* near_hint) { static_cast<IWorld*>(w)->get_pair_unit_hints(near_hint); }
size_t sfo_n_constraint_units(void* w) { return static_cast<IWorld*>(w)->n_constraint_units(); }
void sfo_get_constraint_units(void* w, uint32_t* idx, uint32_t* mult, uint32_t* colour) { static_cast<IWorld*>(w)->get_constraint_units(idx, mult, colour); }
size_t sfo_n_entries(void* w) { return static_cast<IWorld*>(w)->n_entries(); }
void sfo_get_manifolds(void* w, double* out14) { static_cast<IWorld*>(w)->get_manifolds(out14); }
size_t sfo_n_contacts(void* w) { return static_cast<IWorld*>(w)->n_contacts(); }
void sfo_get_contacts(void* w, uint32_t* c01, double* nxy) { static_cast<IWorld*>(w)->get_contacts(c01, nxy); }
void sfo_get_aabbs(void* w, double* out4) { static_cast<IWorld*>(w)->get_aabbs(out4); }
void sfo_cast_batch(void* w, uint32_t n, const SfgRay* rays, double* out6, int use_bvh) { static_cast<IWorld*>(w)->cast_batch(n, rays, out6, use_bvh); }
uint32_t sfo_query_point(void* w, double x, double y, uint32_t domain, int32_t* out, uint32_t cap) { return static_cast<IWorld*>(w)->query_point(x, y, domain, out, cap); }
uint32_t sfo_query_shape(void* w, const SfgShapeQuery* q, int32_t* out, uint32_t cap) { return static_cast<IWorld*>(w)->query_shape(*q, out, cap); }
void sfo_get_stats(void* w, uint64_t* out8) { static_cast<IWorld*>(w)->get_stats(out8); }
size_t sfo_penetrations(void* w, double* out, size_t cap) { return static_cast<IWorld*>(w)->penetrations(out, cap); }
size_t sfo_n_islands(void* w) { return static_cast<IWorld*>(w)->n_all_islands(); }
void sfo_get_islands(void* w, uint64_t* first_body, uint64_t* edge_sum, uint64_t* body_count, uint64_t* flags) {
    static_cast<IWorld*>(w)->get_all_islands(first_body, edge_sum, body_count, flags);
}

// ---- standalone primitives (known-answer tests + device parity) ----
}  // extern "C"
template <class R> static Shape<R> mk_shape(const double* s) { return {uint32_t(s[0]), R(s[1]), R(s[2]), R(s[3])}; }
extern "C" {
}  // extern "C"
template <class R> static Iso2<R> mk_pose(const double* p) { return {{R(p[0]), R(p[1])}, {R(p[2]), R(p[3])}}; }
extern "C" {

// n pairs; poses: 2n rows (x, y, s, b); shapes: 2n rows (kind, p0, p1, r); out: n rows of 12:
// count, normal.xy, 2 x (offsets[0].xy, offsets[1].xy), pad
}  // extern "C"
template <class R> static void isect(uint32_t n, const double* poses, const double* shapes, double* out) {
    for (uint32_t i = 0; i < n; ++i) {
        Iso2<R> p[2] = {mk_pose<R>(poses + 8 * size_t(i)), mk_pose<R>(poses + 8 * size_t(i) + 4)};
        Shape<R> s[2] = {mk_shape<R>(shapes + 8 * size_t(i)), mk_shape<R>(shapes + 8 * size_t(i) + 4)};
        ContactResult<R> c = intersection_check(p, s, false);
        double* o = out + 12 * size_t(i);
        for (int k = 0; k < 12; ++k) o[k] = 0;
        o[0] = c.count;
        if (c.count) { o[1] = c.c[0].normal.x; o[2] = c.c[0].normal.y; }
        for (int k = 0; k < c.count; ++k) {
            o[3 + 4 * k] = c.c[k].offsets[0].x; o[4 + 4 * k] = c.c[k].offsets[0].y;
            o[5 + 4 * k] = c.c[k].offsets[1].x; o[6 + 4 * k] = c.c[k].offsets[1].y;
        }
    }
}
extern "C" {
void sfo_intersection_check(int use_f32, uint32_t n, const double* poses, const double* shapes, double* out) {
    if (use_f32) isect<float>(n, poses, shapes, out); else isect<double>(n, poses, shapes, out);
}

// rays: n rows (sx, sy, dx, dy, radius); poses n rows; shapes n rows; out rows (hit, t, nx, ny, px, py)
}  // extern "C"
template <class R> static void scast(uint32_t n, const double* rays, const double* poses, const double* shapes, double* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const double* r = rays + 5 * size_t(i);
        Ray<R> ray{{R(r[0]), R(r[1])}, {R(r[2]), R(r[3])}};
        CastHit<R> h = spherecast_collider(ray, R(r[4]), mk_pose<R>(poses + 4 * size_t(i)), mk_shape<R>(shapes + 4 * size_t(i)));
        double* o = out + 6 * size_t(i);
        o[0] = h.hit; o[1] = h.t; o[2] = h.normal.x; o[3] = h.normal.y; o[4] = h.point.x; o[5] = h.point.y;
    }
}
extern "C" {
void sfo_spherecast_collider(int use_f32, uint32_t n, const double* rays, const double* poses, const double* shapes, double* out) {
    if (use_f32) scast<float>(n, rays, poses, shapes, out); else scast<double>(n, rays, poses, shapes, out);
}

// geometry tables; op as in sfg_debug_geometry plus: 4 = get_edge(idx = args.x) rows of 7,
// 5 = edge_count / get_edge_extent(idx) (out: count, extent), 6 = area / second moment (out: area, moment),
// 7 = rounded_inward(amount = args.x) (out: kind, p0, p1, r), 8 = ray_aabb(args = ray; shape row = aabb) (out: hit, t)
}  // extern "C"
template <class R> static void geom(uint32_t op, uint32_t n, const double* args, const double* shapes, double* out) {
    for (uint32_t i = 0; i < n; ++i) {
        Shape<R> s = mk_shape<R>(shapes + 4 * size_t(i));
        Vec2<R> a{R(args[2 * size_t(i)]), R(args[2 * size_t(i) + 1])};
        double* o = out + 7 * size_t(i);
        for (int k = 0; k < 7; ++k) o[k] = 0;
        switch (op) {
        case 0: { auto c = s.closest_boundary_point(a); o[0] = c.pt.x; o[1] = c.pt.y; o[2] = c.is_interior; break; }
        case 1: o[0] = s.projected_extent(a); break;
        case 2: case 4: {
            PolygonEdge<R> e = op == 2 ? s.supporting_edge(a) : s.get_edge(int(a.x));
            o[0] = e.normal.x; o[1] = e.normal.y; o[2] = e.edge.start.x; o[3] = e.edge.start.y;
            o[4] = e.edge.dir.x; o[5] = e.edge.dir.y; o[6] = e.edge.length; break;
        }
        case 3: { Iso2<R> id{{0, 0}, {1, 0}}; o[0] = point_collider_bool(a, id, s); break; }
        case 5: o[0] = s.edge_count(); o[1] = s.edge_count() ? s.get_edge_extent(int(a.x)) : 0; break;
        case 6: o[0] = s.area(); o[1] = s.second_moment_of_area(); break;
        case 7: { Shape<R> r = s.rounded_inward(a.x); o[0] = r.kind; o[1] = r.p0; o[2] = r.p1; o[3] = r.circle_r; break; }
        default: break;
        }
    }
}
extern "C" {
void sfo_geometry(int use_f32, uint32_t op, uint32_t n, const double* args, const double* shapes, double* out) {
    if (use_f32) geom<float>(op, n, args, shapes, out); else geom<double>(op, n, args, shapes, out);
}

// clip_edge(target, edge): each edge = (start.xy, dir.xy, length); out = (kind, enters, exits)
void sfo_clip_edge(const double* target, const double* edge, double* out) {
    Edge<double> t{{target[0], target[1]}, {target[2], target[3]}, target[4]};
    Edge<double> e{{edge[0], edge[1]}, {edge[2], edge[3]}, edge[4]};
    ClipResult<double> r = clip_edge(t, e);
    out[0] = r.kind; out[1] = r.enters; out[2] = r.exits;
}
// ray_aabb(ray = sx, sy, dx, dy; aabb = min.xy, max.xy) -> (hit, t)
void sfo_ray_aabb(const double* ray, const double* aabb, double* out) {
    Ray<double> r{{ray[0], ray[1]}, {ray[2], ray[3]}};
    AABB<double> b{{aabb[0], aabb[1]}, {aabb[2], aabb[3]}};
    double t = 0;
    out[0] = ray_aabb(r, b, &t); out[1] = t;
}
// from_angle / pose helpers so tests can build inputs with the oracle's own convention
void sfo_from_angle(double angle, double* out_sb) { auto r = Rotor2<double>::from_angle(angle); out_sb[0] = r.s; out_sb[1] = r.b; }

}  // extern "C"

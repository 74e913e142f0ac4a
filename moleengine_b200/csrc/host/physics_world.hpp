// starframe::PhysicsWorld on top of the C ABI (include/starframe_gpu.h) — the host side of the drop-in,
// written in C++ because the reference's toolchain (Rust) does not exist in this image. Names, argument
// meaning, ownership and housekeeping mirror the reference's Rust API so a user of
//   /root/reference/src/physics.rs:352-1311 (PhysicsWorld), physics/entity_set.rs, constraint_set.rs,
//   constraint.rs (ConstraintBuilder), rope.rs, body.rs, collision/collider.rs (constructors, mass properties),
//   collision.rs (CollisionMaskMatrix), forcefield.rs
// finds the same calls here. Host state is f64 like the reference; records are rounded to f32 once at upload.
// The host owns arenas + generations; only slots cross the ABI. No physics is computed on the host.
#pragma once
#include <cmath>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/starframe_gpu.h"

namespace starframe {

struct DVec2 { double x = 0, y = 0; };
inline DVec2 operator+(DVec2 a, DVec2 b) { return {a.x + b.x, a.y + b.y}; }
inline DVec2 operator-(DVec2 a, DVec2 b) { return {a.x - b.x, a.y - b.y}; }
inline DVec2 operator*(double s, DVec2 a) { return {s * a.x, s * a.y}; }
inline double mag(DVec2 a) { return std::sqrt(a.x * a.x + a.y * a.y); }

struct DRotor2 {   // uv::DRotor2 {s, bv.xy}; from_angle(t) = {cos(t/2), -sin(t/2)}
    double s = 1, b = 0;
    static DRotor2 from_angle(double a) { return {std::cos(a / 2), -std::sin(a / 2)}; }
    DVec2 operator*(DVec2 v) const { double fx = s * v.x + b * v.y, fy = s * v.y - b * v.x; return {s * fx + b * fy, s * fy - b * fx}; }
};
struct PhysicsPose {   // math.rs:12 DIsometry2
    DVec2 translation; DRotor2 rotation;
    PhysicsPose() = default;
    PhysicsPose(DVec2 t, double angle) : translation(t), rotation(DRotor2::from_angle(angle)) {}
    DVec2 operator*(DVec2 v) const { return rotation * v + translation; }
};
struct Velocity { DVec2 linear; double angular = 0; };   // physics.rs:46-53

// ---- thunderdome-like arena: slots + generations, LIFO free list, iteration in slot order ----
struct Index { uint32_t slot = 0, generation = 0; bool operator==(const Index& o) const { return slot == o.slot && generation == o.generation; } };
template <class T>
class Arena {
    struct Entry { bool live = false; uint32_t generation = 0; T value{}; };
    std::vector<Entry> e_;
    std::vector<uint32_t> free_;
    size_t len_ = 0;
public:
    Index insert(const T& v) {
        uint32_t s;
        if (!free_.empty()) { s = free_.back(); free_.pop_back(); }
        else { s = uint32_t(e_.size()); e_.emplace_back(); }
        e_[s].live = true; e_[s].generation++; e_[s].value = v; ++len_;
        return {s, e_[s].generation};
    }
    void insert_at(Index i, const T& v) {   // coll_bodies shares the collider arena's indices (entity_set.rs:112)
        if (i.slot >= e_.size()) e_.resize(i.slot + 1);
        if (!e_[i.slot].live) ++len_;
        e_[i.slot].live = true; e_[i.slot].generation = i.generation; e_[i.slot].value = v;
    }
    T* get(Index i) { return (i.slot < e_.size() && e_[i.slot].live && e_[i.slot].generation == i.generation) ? &e_[i.slot].value : nullptr; }
    const T* get(Index i) const { return const_cast<Arena*>(this)->get(i); }
    bool contains(Index i) const { return get(i) != nullptr; }
    T* get_by_slot(uint32_t s) { return (s < e_.size() && e_[s].live) ? &e_[s].value : nullptr; }
    std::optional<Index> index_of_slot(uint32_t s) const { if (s < e_.size() && e_[s].live) return Index{s, e_[s].generation}; return std::nullopt; }
    std::optional<T> remove(Index i) {
        T* p = get(i);
        if (!p) return std::nullopt;
        T out = *p; e_[i.slot].live = false; free_.push_back(i.slot); --len_;
        return out;
    }
    void clear() { e_.clear(); free_.clear(); len_ = 0; }
    size_t len() const { return len_; }
    uint32_t slot_count() const { return uint32_t(e_.size()); }
    template <class F> void for_each(F&& f) { for (uint32_t s = 0; s < e_.size(); ++s) if (e_[s].live) f(Index{s, e_[s].generation}, e_[s].value); }
    template <class F> void retain(F&& keep) { for (uint32_t s = 0; s < e_.size(); ++s) if (e_[s].live && !keep(Index{s, e_[s].generation}, e_[s].value)) { e_[s].live = false; free_.push_back(s); --len_; } }
};
struct BodyKey { Index i; bool operator==(const BodyKey& o) const { return i == o.i; } };
struct ColliderKey { Index i; bool operator==(const ColliderKey& o) const { return i == o.i; } };
struct RopeKey { Index i; };
struct ConstraintKey { Index i; };

// ---- collision/collider.rs ----
struct PhysicsMaterial {   // :904-921
    std::optional<double> static_friction_coef = 1.6, dynamic_friction_coef = 1.5;
    double restitution_coef = 0.0;
};
enum class Polygon : uint32_t { Point = SFG_POLY_POINT, LineSegment = SFG_POLY_SEGMENT, Rect = SFG_POLY_RECT, Triangle = SFG_POLY_TRIANGLE, Hexagon = SFG_POLY_HEXAGON };
struct ColliderInfo { double area, second_moment_of_area; };   // body.rs:15-19
struct ColliderShape {   // :197-200 + :378-410 (p0 = hl | hw | outer_r, p1 = hh)
    Polygon polygon = Polygon::Point; double p0 = 0, p1 = 0, circle_r = 1.0;
    static constexpr double SQRT_3 = 1.73205080757, PI6_COS = 0.866025403784, PI6_TAN = 0.57735026919, PI = 3.14159265358979323846;
    double poly_area() const {
        switch (polygon) { case Polygon::Rect: return 4.0 * p0 * p1; case Polygon::Triangle: return 3.0 * 0.25 * p0 * p0 / PI6_TAN;
                           case Polygon::Hexagon: return 3.0 * p0 * PI6_COS * p0; default: return 0.0; }
    }
    double side_length_sum() const {
        switch (polygon) { case Polygon::LineSegment: return 4.0 * p0; case Polygon::Rect: return 4.0 * (p0 + p1);
                           case Polygon::Triangle: return 3.0 * p0 / PI6_TAN; case Polygon::Hexagon: return 6.0 * p0; default: return 0.0; }
    }
    double area() const { double r = circle_r; return r == 0.0 ? poly_area() : poly_area() + PI * r * r + side_length_sum() * r; }   // :222-233
    static double m_circle(double r) { return (PI / 2) * r * r * r * r; }
    static double m_rect(double hw, double hh) { return (4.0 / 3.0) * (hw * hw * hw * hh + hw * hh * hh * hh); }
    double second_moment_of_area() const {   // :238-337
        const double cr = circle_r;
        auto caps = [&](double d2) { return m_circle(cr) + (PI * cr * cr) * d2; };
        switch (polygon) {
        case Polygon::Point: return m_circle(cr);
        case Polygon::LineSegment: return m_rect(p0, cr) + (m_circle(cr) + (PI * cr * cr * p0 * p0));
        case Polygon::Rect: {
            double poly = m_rect(p0, p1);
            if (cr == 0.0) return poly;
            double horiz = m_rect(p0, cr / 2) + (2 * p0 * cr) * p1 * p1, vert = m_rect(p1, cr / 2) + (2 * p1 * cr) * p0 * p0;
            return poly + 2 * (horiz + vert) + caps(p0 * p0 + p1 * p1);
        }
        case Polygon::Triangle: {
            double len = p0 / PI6_TAN, poly = 0.03608 * len * len * len * len;
            if (cr == 0.0) return poly;
            return poly + 3 * (m_rect(len / 2, cr / 2) + (len * cr) * (p0 / 2) * (p0 / 2)) + caps(p0 * p0);
        }
        default: {
            double poly = (5 * SQRT_3 / 8) * p0 * p0 * p0 * p0;
            if (cr == 0.0) return poly;
            return poly + 6 * (m_rect(p0 / 2, cr / 2) + (p0 * cr) * (PI6_COS * p0) * (PI6_COS * p0)) + caps(p0 * p0);
        }
        }
    }
};
struct Collider {   // :19-29
    ColliderShape shape;
    bool is_sensor_ = false; PhysicsMaterial material;
    PhysicsPose pose;
    size_t layer = 0;
    static Collider new_circle(double r) { Collider c; c.shape = {Polygon::Point, 0, 0, r}; return c; }
    static Collider new_rounded_rect(double w, double h, double r) {
        Collider c; c.shape = {Polygon::Rect, std::fmax(w / 2 - r, 0.05), std::fmax(h / 2 - r, 0.05), std::fmin(std::fmin(r, w / 2), h / 2)}; return c;
    }
    static Collider new_rect(double w, double h) { return new_rounded_rect(w, h, 0.0); }
    static Collider new_square(double s) { return new_rect(s, s); }
    static Collider new_capsule(double len, double r) { Collider c; c.shape = {Polygon::LineSegment, len / 2, 0, r}; return c; }
    static Collider new_hexagon(double outer_r) { Collider c; c.shape = {Polygon::Hexagon, outer_r, 0, 0}; return c; }
    static Collider new_triangle(double outer_r) { Collider c; c.shape = {Polygon::Triangle, outer_r, 0, 0}; return c; }
    Collider with_pose(PhysicsPose p) const { Collider c = *this; c.pose = p; return c; }
    Collider with_material(PhysicsMaterial m) const { Collider c = *this; c.is_sensor_ = false; c.material = m; return c; }
    Collider sensor() const { Collider c = *this; c.is_sensor_ = true; return c; }
    Collider with_layer(size_t l) const { Collider c = *this; c.layer = l; return c; }
    bool is_solid() const { return !is_sensor_; }
    bool is_sensor() const { return is_sensor_; }
    ColliderInfo info() const { return {shape.area(), shape.second_moment_of_area()}; }
};
struct CompoundColliderSetup {   // collision/compound_collider.rs
    const std::vector<Collider>& colliders;
    DVec2 center_of_mass() const {
        DVec2 s; for (auto& c : colliders) s = s + c.shape.area() * c.pose.translation;
        return (1.0 / double(colliders.size())) * s;   // sic: divided by the collider COUNT (:17-23)
    }
    ColliderInfo info_around_point(DVec2 center) const {
        double a = 0, m = 0;
        for (auto& c : colliders) { double ar = c.shape.area(); a += ar; DVec2 o = c.pose.translation - center; m += c.shape.second_moment_of_area() + ar * (o.x * o.x + o.y * o.y); }
        return {a, m};
    }
};

// ---- body.rs ----
struct Mass { bool finite = false; double mass = 0, inverse = 0; static Mass from(double m) { return {true, m, 1.0 / m}; } static Mass infinite() { return {}; } double inv() const { return finite ? inverse : 0.0; } };
struct Body {
    PhysicsPose pose; Velocity velocity; Mass mass, moment_of_inertia; bool ignores_gravity = false;
    static Body new_particle(double m) { Body b; b.mass = Mass::from(m); return b; }
    static Body new_dynamic(ColliderInfo ci, double density) { Body b; b.mass = Mass::from(ci.area * density); b.moment_of_inertia = Mass::from(ci.second_moment_of_area * density); return b; }
    static Body new_dynamic_const_mass(ColliderInfo ci, double m) { Body b; double d = m / ci.area; b.mass = Mass::from(m); b.moment_of_inertia = Mass::from(ci.second_moment_of_area * d); return b; }
    static Body new_kinematic() { return Body{}; }
    Body with_pose(PhysicsPose p) const { Body b = *this; b.pose = p; return b; }
    Body with_velocity(Velocity v) const { Body b = *this; b.velocity = v; return b; }
    Body ignore_gravity() const { Body b = *this; b.ignores_gravity = true; return b; }
    bool sees_forces() const { return mass.finite || moment_of_inertia.finite; }
};

// ---- entity_set.rs ----
class EntitySet {
public:
    Arena<Body> bodies; Arena<Collider> colliders; Arena<BodyKey> coll_bodies;
    Body* get_body(BodyKey k) { return bodies.get(k.i); }
    Collider* get_collider(ColliderKey k) { return colliders.get(k.i); }
    std::optional<BodyKey> get_collider_body_key(ColliderKey k) const { const BodyKey* b = coll_bodies.get(k.i); return b ? std::optional<BodyKey>(*b) : std::nullopt; }
    BodyKey insert_body(const Body& b) { return {bodies.insert(b)}; }
    ColliderKey insert_collider(const Collider& c) { return {colliders.insert(c)}; }
    ColliderKey attach_collider(BodyKey b, const Collider& c) { ColliderKey k = insert_collider(c); coll_bodies.insert_at(k.i, b); return k; }
    void attach_existing_collider(BodyKey b, ColliderKey c) { coll_bodies.insert_at(c.i, b); }
    std::optional<Body> remove_body(BodyKey b) { return bodies.remove(b.i); }
    std::optional<Collider> remove_collider(ColliderKey c) { coll_bodies.remove(c.i); return colliders.remove(c.i); }
    void remove_orphan_colliders() {   // :168-176
        colliders.retain([&](Index k, Collider&) { const BodyKey* b = coll_bodies.get(k); return b ? bodies.contains(b->i) : true; });
    }
    void clear() { bodies.clear(); colliders.clear(); coll_bodies.clear(); }
};

// ---- constraint.rs / constraint_set.rs ----
enum class ConstraintLimit : uint32_t { Eq = SFG_LIMIT_EQ, Lt = SFG_LIMIT_LT, Gt = SFG_LIMIT_GT };
struct Constraint {
    BodyKey owner; std::optional<BodyKey> target; double compliance = 0, linear_damping = 0.1, angular_damping = 0;
    DVec2 offsets[2]; ConstraintLimit limit = ConstraintLimit::Eq; double distance = 0; bool can_sleep = true;
};
class ConstraintBuilder {   // constraint.rs:60-178
    Constraint c_;
public:
    explicit ConstraintBuilder(BodyKey owner) { c_.owner = owner; }
    ConstraintBuilder& with_target(BodyKey t) { c_.target = t; return *this; }
    ConstraintBuilder& with_origin(DVec2 p) { c_.offsets[0] = p; return *this; }
    ConstraintBuilder& with_target_origin(DVec2 p) { c_.offsets[1] = p; return *this; }
    ConstraintBuilder& with_compliance(double v) { c_.compliance = v; return *this; }
    ConstraintBuilder& with_linear_damping(double v) { c_.linear_damping = v; return *this; }
    ConstraintBuilder& with_angular_damping(double v) { c_.angular_damping = v; return *this; }
    ConstraintBuilder& with_limit(ConstraintLimit l) { c_.limit = l; return *this; }
    ConstraintBuilder& disable_sleeping() { c_.can_sleep = false; return *this; }
    Constraint build_distance(double d) const { Constraint c = c_; c.distance = d; return c; }
    Constraint build_attachment() const { return build_distance(0.0); }
};
class ConstraintSet {
public:
    Arena<Constraint> constraints;
    ConstraintKey insert(const Constraint& c) { return {constraints.insert(c)}; }
    Constraint* get(ConstraintKey k) { return constraints.get(k.i); }
    void remove(ConstraintKey k) { constraints.remove(k.i); }
    void clear() { constraints.clear(); }
};

// ---- rope.rs ----
constexpr size_t ROPE_LAYER = 63;
struct RopeParameters {   // :16-43
    double spacing = 0.1, thickness = 0.12, compliance = 0.0000001, bending_max_angle = double(float(30.0f * 3.14159265358979323846f / 180.0f)),
           bending_compliance = 0.2, damping = 20.0;
    PhysicsMaterial material{std::nullopt, 1.5, 0.0};
    double particle_mass = 0.02;
};
struct RopeParticle { BodyKey body; ColliderKey collider; };
struct Rope {
    RopeParameters params; std::vector<RopeParticle> particles;
    static void build_line(std::vector<RopeParticle>& out, const RopeParameters& p, DVec2 start, DVec2 step, size_t count, EntitySet& es) {   // :113-141
        Collider proto = Collider::new_circle(p.thickness / 2.0).with_layer(ROPE_LAYER).with_material(p.material);
        DVec2 pos = start;
        for (size_t i = 0; i < count; ++i) {
            Body b = Body::new_particle(p.particle_mass); b.pose.translation = pos;
            BodyKey bk = es.insert_body(b);
            out.push_back({bk, es.attach_collider(bk, proto)});
            pos = pos + step;
        }
    }
    static Rope spawn_line(RopeParameters params, DVec2 start, DVec2 end, EntitySet& es) {   // :61-85
        DVec2 dist = end - start; double dm = mag(dist);
        size_t segs = size_t(std::round(dm / params.spacing));
        params.spacing = dm / double(segs);
        DVec2 step = (params.spacing / dm) * dist;
        Rope r; r.params = params;
        build_line(r.particles, params, start, step, segs + 1, es);
        return r;
    }
    void extend_line(DVec2 dir, size_t count, EntitySet& es) {   // :88-110
        if (particles.empty()) return;
        Body* last = es.get_body(particles.back().body);
        if (!last) return;
        DVec2 step = params.spacing * dir;
        build_line(particles, params, last->pose.translation + step, step, count, es);
    }
    std::optional<Rope> cut_after(BodyKey particle) {   // :147-165
        size_t idx = particles.size();
        for (size_t i = 0; i < particles.size(); ++i) if (particles[i].body == particle) { idx = i; break; }
        if (idx == particles.size()) return std::nullopt;
        if (idx + 1 == particles.size()) { particles.pop_back(); return std::nullopt; }
        Rope cut; cut.params = params; cut.particles.assign(particles.begin() + long(idx) + 1, particles.end());
        particles.resize(idx + 1);
        return cut;
    }
};
class RopeSet {
public:
    Arena<Rope> ropes;
    RopeKey insert(const Rope& r) { return {ropes.insert(r)}; }
    Rope* get(RopeKey k) { return ropes.get(k.i); }
    void remove(RopeKey k, EntitySet& es) { auto r = ropes.remove(k.i); if (r) for (auto& p : r->particles) es.remove_body(p.body); }
    void clear() { ropes.clear(); }
    void remove_dead_particles(EntitySet& es) {   // :226-286: cut ropes at dead particles, drop empty ropes
        std::vector<Rope> fresh;
        ropes.for_each([&](Index, Rope& rope) {
            std::vector<RopeParticle> all = rope.particles, cur;
            bool first = true;
            rope.particles.clear();
            auto flush = [&]() {
                if (cur.empty()) return;
                if (first) rope.particles = cur; else { Rope n; n.params = rope.params; n.particles = cur; fresh.push_back(n); }
                cur.clear();
            };
            for (auto& p : all) {
                if (es.bodies.contains(p.body.i)) cur.push_back(p);
                else { flush(); first = false; }
            }
            flush();
        });
        for (auto& r : fresh) insert(r);
        ropes.retain([](Index, Rope& r) { return !r.particles.empty(); });
    }
};

// ---- collision.rs:108-176 ----
struct CollisionMaskMatrix {
    uint64_t m[64];
    CollisionMaskMatrix() { for (auto& v : m) v = ~0ull; ignore_within(ROPE_LAYER); }
    void ignore(size_t a, size_t b) { m[a] &= ~(1ull << b); m[b] &= ~(1ull << a); }
    void ignore_within(size_t l) { m[l] &= ~(1ull << l); }
    void ignore_all(size_t l) { for (size_t o = 0; o < 64; ++o) ignore(l, o); }
    void unignore(size_t a, size_t b) { m[a] |= 1ull << b; m[b] |= 1ull << a; }
    bool get(size_t a, size_t b) const { return (m[a] >> b) & 1ull; }
};

// ---- forcefield.rs: the generic trait becomes data (a sum of up to four terms) ----
struct ForceField {
    SfgForceField f{};
    static ForceField none() { return {}; }
    static ForceField gravity(DVec2 g) { ForceField o; o.f.n_terms = 1; o.f.terms[0] = {SFG_FF_GRAVITY, float(g.x), float(g.y), 0, 0}; return o; }
    static ForceField point_gravity(DVec2 pos, double strength, double falloff) { ForceField o; o.f.n_terms = 1; o.f.terms[0] = {SFG_FF_POINT_GRAVITY, float(pos.x), float(pos.y), float(strength), float(falloff)}; return o; }
    ForceField sum(const ForceField& other) const {   // forcefield.rs:18-24 Sum<F1, F2>
        ForceField o = *this;
        for (uint32_t i = 0; i < other.f.n_terms; ++i) { if (o.f.n_terms >= 4) throw std::runtime_error("ForceField: at most 4 terms cross the ABI"); o.f.terms[o.f.n_terms++] = other.f.terms[i]; }
        return o;
    }
};

struct TuningConstants { size_t substeps = 10; double sleep_vel_threshold = 0.001; size_t fall_asleep_frames = 10; double max_expected_acceleration = 10.0; size_t min_bodies_per_thread = 64; };   // physics.rs:311-349
struct ContactInfo { ColliderKey colliders[2]; DVec2 normal; };   // physics.rs:112-118
struct Ray { DVec2 start, dir; };
struct CastHit { ColliderKey collider; DVec2 point, normal; double t; };   // physics.rs:131-149

class PhysicsWorld {
    SfgWorld* raw_ = nullptr;
    std::vector<ContactInfo> contacts_;
    void ck(int rc) const { if (rc != SFG_OK) throw std::runtime_error(std::string("starframe-gpu: ") + sfg_last_error(raw_)); }   // the reference panics
public:
    TuningConstants consts; CollisionMaskMatrix mask_matrix; EntitySet entity_set; RopeSet rope_set; ConstraintSet constraint_set;   // physics.rs:353-357

    explicit PhysicsWorld(TuningConstants c = {}, CollisionMaskMatrix m = {}, int device = 0) : consts(c), mask_matrix(m) {   // physics.rs:366
        int rc = sfg_world_create(nullptr, m.m, device, &raw_);
        if (rc != SFG_OK) throw std::runtime_error("starframe-gpu: no CUDA device (there is no CPU fallback)");
    }
    ~PhysicsWorld() { sfg_world_destroy(raw_); }
    PhysicsWorld(const PhysicsWorld&) = delete;
    PhysicsWorld& operator=(const PhysicsWorld&) = delete;
    void clear() { entity_set.clear(); rope_set.clear(); constraint_set.clear(); contacts_.clear(); sfg_world_clear(raw_); }   // physics.rs:386-393

    // physics.rs:396-1158
    void tick(double frame_dt, std::optional<double> time_scale, const ForceField& forcefield) {
        entity_set.remove_orphan_colliders();
        rope_set.remove_dead_particles(entity_set);
        constraint_set.constraints.retain([&](Index, Constraint& c) { return entity_set.bodies.contains(c.owner.i) && (!c.target || entity_set.bodies.contains(c.target->i)); });   // :423-428
        SfgTuning t; sfg_default_tuning(&t);
        t.substeps = uint32_t(consts.substeps); t.sleep_vel_threshold = float(consts.sleep_vel_threshold);
        t.fall_asleep_frames = uint32_t(consts.fall_asleep_frames); t.max_expected_acceleration = float(consts.max_expected_acceleration);
        ck(sfg_set_tuning(raw_, &t));
        ck(sfg_set_mask_matrix(raw_, mask_matrix.m));
        std::vector<SfgBody> b(entity_set.bodies.slot_count());
        entity_set.bodies.for_each([&](Index i, Body& v) {
            SfgBody& o = b[i.slot];
            o.x = float(v.pose.translation.x); o.y = float(v.pose.translation.y); o.rot_s = float(v.pose.rotation.s); o.rot_b = float(v.pose.rotation.b);
            o.vx = float(v.velocity.linear.x); o.vy = float(v.velocity.linear.y); o.w = float(v.velocity.angular);
            o.flags = SFG_BODY_ALIVE | (v.ignores_gravity ? SFG_BODY_IGNORES_GRAVITY : 0) | (v.mass.finite ? SFG_BODY_MASS_FINITE : 0) | (v.moment_of_inertia.finite ? SFG_BODY_INERTIA_FINITE : 0);
            o.inv_mass = float(v.mass.inv()); o.inv_inertia = float(v.moment_of_inertia.inv());
        });
        std::vector<SfgCollider> c(entity_set.colliders.slot_count());
        entity_set.colliders.for_each([&](Index i, Collider& v) {
            SfgCollider& o = c[i.slot];
            o.kind = uint32_t(v.shape.polygon); o.p0 = float(v.shape.p0); o.p1 = float(v.shape.p1); o.circle_r = float(v.shape.circle_r);
            o.layer = uint32_t(v.layer);
            o.flags = SFG_COLL_ALIVE | (v.is_sensor_ ? SFG_COLL_SENSOR : 0) | (v.material.static_friction_coef ? SFG_COLL_HAS_STATIC_FRICTION : 0) | (v.material.dynamic_friction_coef ? SFG_COLL_HAS_DYNAMIC_FRICTION : 0);
            o.static_friction = float(v.material.static_friction_coef.value_or(0)); o.dynamic_friction = float(v.material.dynamic_friction_coef.value_or(0));
            o.restitution = float(v.material.restitution_coef);
            o.x = float(v.pose.translation.x); o.y = float(v.pose.translation.y); o.rot_s = float(v.pose.rotation.s); o.rot_b = float(v.pose.rotation.b);
            const BodyKey* bk = entity_set.coll_bodies.get(i);
            o.body = (bk && entity_set.bodies.contains(bk->i)) ? int32_t(bk->i.slot) : -1;
        });
        std::vector<SfgConstraint> k;
        constraint_set.constraints.for_each([&](Index, Constraint& v) {   // arena iteration order = bufs.user_constraints (:429-431)
            SfgConstraint o{};
            o.owner = int32_t(v.owner.i.slot); o.target = v.target ? int32_t(v.target->i.slot) : -1;
            o.flags = uint32_t(v.limit) | (v.can_sleep ? SFG_CONSTRAINT_CAN_SLEEP : 0);
            o.compliance = float(v.compliance); o.linear_damping = float(v.linear_damping); o.angular_damping = float(v.angular_damping); o.distance = float(v.distance);
            o.off0x = float(v.offsets[0].x); o.off0y = float(v.offsets[0].y); o.off1x = float(v.offsets[1].x); o.off1y = float(v.offsets[1].y);
            k.push_back(o);
        });
        std::vector<SfgRope> r; std::vector<int32_t> rp;
        rope_set.ropes.for_each([&](Index, Rope& v) {
            if (v.particles.size() < 2) throw std::runtime_error("Rope only had one particle");   // solver.rs:120
            SfgRope o{};
            o.spacing = float(v.params.spacing); o.compliance = float(v.params.compliance); o.bending_max_angle = float(v.params.bending_max_angle);
            o.bending_compliance = float(v.params.bending_compliance); o.damping = float(v.params.damping);
            o.first_particle = uint32_t(rp.size()); o.particle_count = uint32_t(v.particles.size());
            for (auto& p : v.particles) rp.push_back(int32_t(p.body.i.slot));
            r.push_back(o);
        });
        ck(sfg_set_bodies(raw_, uint32_t(b.size()), b.data()));
        ck(sfg_set_colliders(raw_, uint32_t(c.size()), c.data()));
        ck(sfg_set_constraints(raw_, uint32_t(k.size()), k.data()));
        ck(sfg_set_ropes(raw_, uint32_t(r.size()), r.data(), uint32_t(rp.size()), rp.data()));
        ck(sfg_tick(raw_, frame_dt, time_scale ? 1 : 0, time_scale.value_or(0.0), &forcefield.f));
        if (!b.empty()) ck(sfg_get_bodies(raw_, 0, uint32_t(b.size()), b.data()));
        entity_set.bodies.for_each([&](Index i, Body& v) {   // physics.rs:1150-1157
            const SfgBody& o = b[i.slot];
            v.pose.translation = {o.x, o.y}; v.pose.rotation = {o.rot_s, o.rot_b};
            v.velocity.linear = {o.vx, o.vy}; v.velocity.angular = o.w;
        });
        size_t n = 0;
        ck(sfg_get_contacts(raw_, nullptr, 0, &n));
        std::vector<SfgContactInfo> ci(n);
        if (n) ck(sfg_get_contacts(raw_, ci.data(), n, &n));
        contacts_.clear();
        for (auto& x : ci) {
            auto i0 = entity_set.colliders.index_of_slot(x.collider0), i1 = entity_set.colliders.index_of_slot(x.collider1);
            if (i0 && i1) contacts_.push_back({{ColliderKey{*i0}, ColliderKey{*i1}}, {x.nx, x.ny}});
        }
    }

    // physics.rs:1165-1178: the queried collider comes first and the normal faces away from it
    std::vector<ContactInfo> contacts_for_collider(ColliderKey coll) const {
        std::vector<ContactInfo> out;
        for (auto& c : contacts_) {
            if (c.colliders[0] == coll) out.push_back(c);
            else if (c.colliders[1] == coll) out.push_back({{c.colliders[1], c.colliders[0]}, {-c.normal.x, -c.normal.y}});
        }
        return out;
    }
    std::optional<CastHit> spherecast(double radius, Ray ray, double max_distance) {   // physics.rs:1266-1311
        SfgRay r{float(ray.start.x), float(ray.start.y), float(ray.dir.x), float(ray.dir.y), float(radius), float(max_distance), 0, 0};
        SfgCastHit h{};
        ck(sfg_cast_batch(raw_, 1, &r, &h));
        if (h.collider < 0) return std::nullopt;
        auto idx = entity_set.colliders.index_of_slot(uint32_t(h.collider));
        if (!idx) return std::nullopt;
        return CastHit{ColliderKey{*idx}, {h.px, h.py}, {h.nx, h.ny}, h.t};
    }
    std::optional<CastHit> raycast(Ray ray, double max_distance) { return spherecast(0.0, ray, max_distance); }   // physics.rs:1255-1257
    std::vector<std::pair<ColliderKey, std::optional<BodyKey>>> query_point(DVec2 point) {   // physics.rs:1188-1210
        float p[2] = {float(point.x), float(point.y)};
        int32_t out[32]; uint32_t cnt = 0;
        ck(sfg_query_point_batch(raw_, 1, p, nullptr, 32, out, &cnt));
        std::vector<std::pair<ColliderKey, std::optional<BodyKey>>> res;
        for (uint32_t i = 0; i < cnt && i < 32; ++i) {
            auto idx = entity_set.colliders.index_of_slot(uint32_t(out[i]));
            if (idx) res.push_back({ColliderKey{*idx}, entity_set.get_collider_body_key(ColliderKey{*idx})});
        }
        return res;
    }
    // physics.rs:1215-1245; mask: bit l set = layer l is reported (CollisionLayerMask, collision.rs:112-128)
    std::vector<std::pair<ColliderKey, std::optional<BodyKey>>> query_shape(PhysicsPose pose, ColliderShape shape, uint64_t mask) {
        SfgShapeQuery q{};
        q.x = float(pose.translation.x); q.y = float(pose.translation.y); q.rot_s = float(pose.rotation.s); q.rot_b = float(pose.rotation.b);
        q.kind = uint32_t(shape.polygon); q.p0 = float(shape.p0); q.p1 = float(shape.p1); q.circle_r = float(shape.circle_r);
        q.layer_mask = mask;
        int32_t out[64]; uint32_t cnt = 0;
        ck(sfg_query_shape_batch(raw_, 1, &q, 64, out, &cnt));
        std::vector<std::pair<ColliderKey, std::optional<BodyKey>>> res;
        for (uint32_t i = 0; i < cnt && i < 64; ++i) {
            auto idx = entity_set.colliders.index_of_slot(uint32_t(out[i]));
            if (idx) res.push_back({ColliderKey{*idx}, entity_set.get_collider_body_key(ColliderKey{*idx})});
        }
        return res;
    }
    SfgWorld* raw() { return raw_; }
};

}  // namespace starframe

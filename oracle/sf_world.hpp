// ORACLE — TEST INFRASTRUCTURE ONLY (see sf_math.hpp header). Never used by the product path.
//
// CPU restatement of Starframe's physics tick:
//   /root/reference/src/physics.rs:396-1158  (PhysicsWorld::tick)
//   /root/reference/src/physics/solver.rs    (solve + every stage)
//   /root/reference/src/physics.rs:1255-1311 (raycast / spherecast), :1188-1210 (query_point)
//   /root/reference/src/physics/forcefield.rs, body.rs, constraint.rs, rope.rs (data model)
//
// Two orders of Gauss-Seidel are offered:
//   MODE_REFERENCE — the reference's own order: candidate pairs in BVH emission order,
//     islands by DFS over the constraint graph (physics.rs:598-705) including the duplicate-edge
//     quirk (every body-body contact / constraint is met from both endpoints and pushed twice),
//     islands put to sleep (physics.rs:711-741, 1128-1144), sequential sweep per island.
//   MODE_COLOURED — the order the CUDA path uses ("the CPU reference is fed the same coloured
//     constraint order"): the same multiset of constraint applications per substep, visited
//     colour-major. The colouring is defined in include/sfg_hash.h. No sleeping.
// Parity status: the reference's own tests do not pin tick / solver results —
// "parity unpinned" for everything above the collision primitives; this restatement is the pin
// (held to closed forms of the reference's update rules by tests/test_oracle_physics_kat.py).
#pragma once
#include <algorithm>
#include <array>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <vector>
#include "../include/sfg_hash.h"
#include "../include/starframe_gpu.h"
#include "sf_bvh.hpp"

namespace sfo {

enum Mode { MODE_REFERENCE = 0, MODE_COLOURED = 1 };

template <class R>
struct BodyS {                       // body.rs:7-13
    Iso2<R> pose;
    Vec2<R> lin{0, 0};
    R ang = 0;
    bool mass_finite = false, inertia_finite = false;
    R inv_mass = 0, inv_inertia = 0;  // Mass::inv() (body.rs:121-128)
    bool ignores_gravity = false;
    bool alive = false;
    bool sees_forces() const { return mass_finite || inertia_finite; }     // body.rs:92-97
    R vel_mag_sq() const { return lin.mag_sq() + ang * ang; }             // physics.rs:57-59
    Vec2<R> point_velocity(Vec2<R> off) const { return lin + left_normal(off) * ang; }  // :63-66
};

template <class R>
struct ColliderS {                   // collider.rs:19-29
    Shape<R> shape;
    uint32_t layer = 0;
    bool sensor = false;
    bool has_sf = true, has_df = true;
    R sf = R(1.6), df = R(1.5), rest = 0;
    Iso2<R> pose;
    int32_t body = -1;
    uint32_t domain = 0;
    bool alive = false;
};

template <class R>
struct ConstraintS {                 // constraint.rs:12-38
    int32_t owner = 0, target = -1;
    R compliance = 0, linear_damping = 0, angular_damping = 0, distance = 0;
    Vec2<R> off[2];
    uint32_t limit = 0;
    bool can_sleep = true;
};

template <class R>
struct RopeS {                       // rope.rs:16-50 (solver-visible part)
    R spacing, compliance, bending_max_angle, bending_compliance, damping;
    std::vector<int32_t> particles;  // body slots
};

struct PairUnit { uint32_t hi, lo, mult, colour, near_hint; };   // one candidate pair, solved `mult` times; near_hint: include/sfg_hash.h
struct ConstraintUnit { uint32_t idx, mult, colour; };

template <class R>
struct ContactInfoS { uint32_t c0, c1; Vec2<R> normal; uint64_t isl_first, isl_sum; };

struct IWorld {
    virtual ~IWorld() {}
    virtual void set_tuning(const SfgTuning& t) = 0;
    virtual void set_mask(const uint64_t* m) = 0;
    virtual void set_bodies(uint32_t n, const SfgBody* b) = 0;
    virtual void set_colliders(uint32_t n, const SfgCollider* c) = 0;
    virtual void set_constraints(uint32_t n, const SfgConstraint* c) = 0;
    virtual void set_ropes(uint32_t nr, const SfgRope* r, uint32_t np, const int32_t* p) = 0;
    virtual void set_external_pairs(const uint32_t* hi_lo, size_t n) = 0;
    virtual void set_external_pair_hints(const uint32_t* hi_lo, const uint8_t* hints, size_t n) = 0;
    virtual void tick(double frame_dt, bool has_scale, double scale, const SfgForceField& ff, int mode,
                      int threads) = 0;
    virtual void get_bodies(uint32_t first, uint32_t count, double* out7) const = 0;
    virtual size_t n_pairs() const = 0;
    virtual void get_pairs(uint32_t* out) const = 0;
    virtual size_t n_pair_units() const = 0;
    virtual void get_pair_units(uint32_t* hi, uint32_t* lo, uint32_t* mult, uint32_t* colour) const = 0;
    virtual void get_pair_unit_hints(uint8_t* near_hint) const = 0;
    virtual size_t n_constraint_units() const = 0;
    virtual void get_constraint_units(uint32_t* idx, uint32_t* mult, uint32_t* colour) const = 0;
    virtual size_t n_entries() const = 0;
    virtual void get_manifolds(double* out12) const = 0;   // per entry, see capi
    virtual size_t n_contacts() const = 0;
    virtual void get_contacts(uint32_t* c01, double* nxy) const = 0;
    virtual void get_aabbs(double* out4) const = 0;
    virtual void cast_batch(uint32_t n, const SfgRay* rays, double* out6, int use_bvh) const = 0;
    virtual uint32_t query_point(double x, double y, uint32_t domain, int32_t* out, uint32_t cap) const = 0;
    virtual uint32_t query_shape(const SfgShapeQuery& q, int32_t* out, uint32_t cap) const = 0;
    virtual void get_stats(uint64_t* out8) const = 0;
    virtual size_t penetrations(double* out, size_t cap) = 0;
    virtual size_t n_all_islands() const = 0;
    virtual void get_all_islands(uint64_t* first_body, uint64_t* edge_sum, uint64_t* body_count, uint64_t* flags) const = 0;
};

template <class R>
class World final : public IWorld {
public:
    using V = Vec2<R>;
    std::vector<BodyS<R>> bodies;
    std::vector<ColliderS<R>> colliders;
    std::vector<ConstraintS<R>> constraints;
    std::vector<RopeS<R>> ropes;
    SfgTuning tuning{};
    uint64_t mask[64];

    World() {
        tuning.substeps = 10; tuning.sleep_vel_threshold = 0.001f; tuning.fall_asleep_frames = 10;
        tuning.max_expected_acceleration = 10.0f;
        for (int i = 0; i < 64; ++i) mask[i] = ~0ull;
        mask[63] &= ~(1ull << 63);                                        // collision.rs:134-140
    }

    // ------------------------------------------------------------------ uploads
    void set_tuning(const SfgTuning& t) override { tuning = t; }
    void set_mask(const uint64_t* m) override { std::memcpy(mask, m, sizeof(mask)); }
    void set_bodies(uint32_t n, const SfgBody* b) override {
        bodies.assign(n, BodyS<R>{});
        for (uint32_t i = 0; i < n; ++i) {
            BodyS<R>& o = bodies[i];
            o.alive = b[i].flags & SFG_BODY_ALIVE;
            o.pose = {{R(b[i].x), R(b[i].y)}, {R(b[i].rot_s), R(b[i].rot_b)}};
            o.lin = {R(b[i].vx), R(b[i].vy)};
            o.ang = R(b[i].w);
            o.ignores_gravity = b[i].flags & SFG_BODY_IGNORES_GRAVITY;
            o.mass_finite = b[i].flags & SFG_BODY_MASS_FINITE;
            o.inertia_finite = b[i].flags & SFG_BODY_INERTIA_FINITE;
            o.inv_mass = o.mass_finite ? R(b[i].inv_mass) : R(0);
            o.inv_inertia = o.inertia_finite ? R(b[i].inv_inertia) : R(0);
        }
    }
    void set_colliders(uint32_t n, const SfgCollider* c) override {
        colliders.assign(n, ColliderS<R>{});
        for (uint32_t i = 0; i < n; ++i) {
            ColliderS<R>& o = colliders[i];
            o.alive = c[i].flags & SFG_COLL_ALIVE;
            o.shape = {c[i].kind, R(c[i].p0), R(c[i].p1), R(c[i].circle_r)};
            o.layer = c[i].layer;
            o.sensor = c[i].flags & SFG_COLL_SENSOR;
            o.has_sf = c[i].flags & SFG_COLL_HAS_STATIC_FRICTION;
            o.has_df = c[i].flags & SFG_COLL_HAS_DYNAMIC_FRICTION;
            o.sf = R(c[i].static_friction); o.df = R(c[i].dynamic_friction); o.rest = R(c[i].restitution);
            o.pose = {{R(c[i].x), R(c[i].y)}, {R(c[i].rot_s), R(c[i].rot_b)}};
            o.body = c[i].body;
            o.domain = c[i].domain;
        }
    }
    void set_constraints(uint32_t n, const SfgConstraint* c) override {
        constraints.assign(n, ConstraintS<R>{});
        for (uint32_t i = 0; i < n; ++i) {
            ConstraintS<R>& o = constraints[i];
            o.owner = c[i].owner; o.target = c[i].target;
            o.compliance = R(c[i].compliance); o.linear_damping = R(c[i].linear_damping);
            o.angular_damping = R(c[i].angular_damping); o.distance = R(c[i].distance);
            o.off[0] = {R(c[i].off0x), R(c[i].off0y)}; o.off[1] = {R(c[i].off1x), R(c[i].off1y)};
            o.limit = c[i].flags & SFG_LIMIT_MASK;
            o.can_sleep = c[i].flags & SFG_CONSTRAINT_CAN_SLEEP;
        }
    }
    void set_ropes(uint32_t nr, const SfgRope* r, uint32_t np, const int32_t* p) override {
        (void)np;
        ropes.clear();
        for (uint32_t i = 0; i < nr; ++i) {
            RopeS<R> o{R(r[i].spacing), R(r[i].compliance), R(r[i].bending_max_angle),
                       R(r[i].bending_compliance), R(r[i].damping), {}};
            o.particles.assign(p + r[i].first_particle, p + r[i].first_particle + r[i].particle_count);
            ropes.push_back(std::move(o));
        }
    }
    void set_external_pairs(const uint32_t* hl, size_t n) override {
        ext_pairs_.assign(hl, hl + 2 * n);
        use_ext_pairs_ = true;
    }
    // the `near` bit of each pair's colouring priority as the device computed it (include/sfg_hash.h); pairs without a hint count as near
    void set_external_pair_hints(const uint32_t* hl, const uint8_t* hints, size_t n) override {
        ext_hints_.clear();
        for (size_t i = 0; i < n; ++i) ext_hints_[(uint64_t(hl[2 * i]) << 32) | hl[2 * i + 1]] = hints[i];
    }

    // ------------------------------------------------------------------ helpers
    Iso2<R> collider_world_pose(const ColliderS<R>& c) const {
        if (c.body >= 0 && bodies[size_t(c.body)].alive) return bodies[size_t(c.body)].pose * c.pose;
        return c.pose;
    }
    static V ff_value(const SfgForceField& ff, V pos) {                  // forcefield.rs
        V acc{0, 0};
        for (uint32_t i = 0; i < ff.n_terms && i < 4; ++i) {
            const SfgForceTerm& t = ff.terms[i];
            if (t.kind == SFG_FF_GRAVITY) acc += V{R(t.a), R(t.b)};
            else if (t.kind == SFG_FF_POINT_GRAVITY) {
                V dist = V{R(t.a), R(t.b)} - pos;
                R strength = R(t.c) / ((dist.mag_sq() + R(1)) * R(t.d));
                acc += strength * dist.normalized();
            }
        }
        return acc;
    }
    bool mask_get(uint32_t l1, uint32_t l2) const { return (mask[l1] >> l2) & 1ull; }

    // ------------------------------------------------------------------ tick
    struct Island {
        uint64_t first_body = 0, edge_sum = 0;
        size_t body_start = 0, body_count = 0, rope_start = 0, rope_count = 0;
        size_t constr_start = 0, constr_count = 0, pair_start = 0, pair_count = 0;
        bool can_sleep = true;
    };
    struct Sleeping { uint64_t first_body, edge_sum; bool continues; uint64_t ticks; };

    // per-tick working state (slot-indexed where the reference uses island-sorted buffers;
    // results do not depend on the storage order)
    std::vector<AABB<R>> aabbs;
    std::vector<uint32_t> pair_keys;                 // 2 per candidate pair: [later, earlier]
    std::vector<Iso2<R>> old_poses, pre_contact_poses;
    std::vector<V> old_lin; std::vector<R> old_ang;
    std::vector<V> ext_acc;
    std::vector<int32_t> rope_next, rope_prev;
    std::vector<uint8_t> lateral_set; std::vector<V> lateral;
    // entry lists (an "entry" = one visit of a pair / constraint inside a substep)
    std::vector<uint32_t> e_pair;                    // pair index per contact entry
    std::vector<ContactResult<R>> contacts, last_contacts;
    std::vector<R> lambdas;
    std::vector<uint32_t> e_constr;                  // constraint index per constraint entry
    std::vector<uint32_t> e_bodies, e_ropes;         // island-ordered body slots / rope indices
    std::vector<Island> islands;
    struct IslandRec { uint64_t first_body, edge_sum, body_count, flags; };   // every island of the last tick; flags: 1 cannot sleep, 8 asleep
    std::vector<IslandRec> all_islands;
    std::vector<Sleeping> sleeping;
    std::vector<ContactInfoS<R>> contact_infos;
    std::vector<PairUnit> pair_units;
    std::vector<ConstraintUnit> constraint_units;
    std::vector<uint32_t> kept_pairs;                // parity object: pairs with >=1 body, distinct bodies
    Bvh<R> bvh;
    uint64_t stat_colours_pairs = 0, stat_colours_constr = 0, stat_substeps = 0;
    R dt = 0, inv_dt = 0, inv_dt_sq = 0;

    void housekeeping() {
        // entity_set.rs:168-176 remove_orphan_colliders
        for (auto& c : colliders)
            if (c.alive && c.body >= 0 && !(size_t(c.body) < bodies.size() && bodies[size_t(c.body)].alive)) c.alive = false;
        // rope.rs:226-286 remove_dead_particles (split ropes at dead particles)
        std::vector<RopeS<R>> out;
        for (auto& r : ropes) {
            std::vector<int32_t> cur;
            for (int32_t p : r.particles) {
                bool ok = p >= 0 && size_t(p) < bodies.size() && bodies[size_t(p)].alive;
                if (ok) cur.push_back(p);
                else if (!cur.empty()) { RopeS<R> n = r; n.particles = cur; out.push_back(n); cur.clear(); }
            }
            if (!cur.empty()) { RopeS<R> n = r; n.particles = cur; out.push_back(n); }
        }
        ropes.swap(out);
        // physics.rs:423-428
        std::vector<ConstraintS<R>> kc;
        for (auto& c : constraints) {
            bool ok = c.owner >= 0 && size_t(c.owner) < bodies.size() && bodies[size_t(c.owner)].alive;
            if (c.target >= 0) ok = ok && size_t(c.target) < bodies.size() && bodies[size_t(c.target)].alive;
            if (ok) kc.push_back(c);
        }
        constraints.swap(kc);
    }

    void compute_aabbs(R frame_dt) {                                      // physics.rs:443-459
        const R pad = R(tuning.max_expected_acceleration) * frame_dt;
        aabbs.assign(colliders.size(), AABB<R>{});
        for (size_t i = 0; i < colliders.size(); ++i) {
            const auto& c = colliders[i];
            if (!c.alive) continue;
            if (c.body >= 0) {
                const auto& b = bodies[size_t(c.body)];
                aabbs[i] = c.shape.aabb(b.pose * c.pose).extended(b.lin * frame_dt).padded(pad);
            } else {
                aabbs[i] = c.shape.aabb(c.pose);
            }
        }
    }

    void broadphase_bvh() {                                               // physics.rs:443-475
        bvh.clear();
        pair_keys.clear();
        for (size_t i = 0; i < colliders.size(); ++i) {
            const auto& c = colliders[i];
            if (!c.alive) continue;
            bvh.test_aabb(aabbs[i], [&](uint32_t other) {
                if (colliders[other].domain != c.domain) return;          // batched worlds never meet
                if (mask_get(c.layer, colliders[other].layer)) { pair_keys.push_back(uint32_t(i)); pair_keys.push_back(other); }
            });
            bvh.insert(uint32_t(i), aabbs[i]);
        }
    }
    // Closed form of the same set by sort-and-sweep; emission order is
    // (later slot asc, earlier slot asc). Used by MODE_COLOURED where only the set matters.
    void broadphase_sweep() {
        pair_keys.clear();
        std::vector<uint32_t> idx;
        for (size_t i = 0; i < colliders.size(); ++i) if (colliders[i].alive) idx.push_back(uint32_t(i));
        std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {
            if (colliders[a].domain != colliders[b].domain) return colliders[a].domain < colliders[b].domain;
            if (aabbs[a].min.x != aabbs[b].min.x) return aabbs[a].min.x < aabbs[b].min.x;
            return a < b;
        });
        std::vector<std::pair<uint32_t, uint32_t>> found;
        for (size_t a = 0; a < idx.size(); ++a) {
            const uint32_t i = idx[a];
            for (size_t b = a + 1; b < idx.size(); ++b) {
                const uint32_t j = idx[b];
                if (colliders[j].domain != colliders[i].domain) break;
                if (aabbs[j].min.x >= aabbs[i].max.x) break;
                if (!aabbs[i].intersects(aabbs[j])) continue;
                uint32_t hi = std::max(i, j), lo = std::min(i, j);
                if (mask_get(colliders[hi].layer, colliders[lo].layer)) found.push_back({hi, lo});
            }
        }
        std::sort(found.begin(), found.end());
        for (auto& p : found) { pair_keys.push_back(p.first); pair_keys.push_back(p.second); }
    }

    // A statistic of the current STATE, not part of the tick (long-run parity gates compare its distribution between the device
    // and the reference order): the depth of every contact point of every overlapping solid pair, depth as solve_contacts
    // measures it (solver.rs:425-431: (p0 - p1) . normal with the points on the two rounded surfaces). Returns how many were
    // found; the first `cap` are written.
    size_t penetrations(double* out, size_t cap) override {
        compute_aabbs(R(0));                                               // tight boxes: overlapping shapes have overlapping boxes
        broadphase_sweep();
        size_t n = 0;
        for (size_t p = 0; p + 1 < pair_keys.size(); p += 2) {
            const uint32_t ck[2] = {pair_keys[p], pair_keys[p + 1]};
            if (!pair_kept(ck[0], ck[1])) continue;
            const ColliderS<R>* cl[2] = {&colliders[ck[0]], &colliders[ck[1]]};
            if (cl[0]->sensor || cl[1]->sensor) continue;
            Iso2<R> pw[2];
            for (int i = 0; i < 2; ++i) pw[i] = cl[i]->body >= 0 ? bodies[size_t(cl[i]->body)].pose * cl[i]->pose : cl[i]->pose;
            Shape<R> sh[2] = {cl[0]->shape, cl[1]->shape};
            ContactResult<R> c = intersection_check(pw, sh, false);
            for (int k = 0; k < c.count; ++k) {
                const R depth = (pw[0] * c.c[k].offsets[0] - pw[1] * c.c[k].offsets[1]).dot(c.c[k].normal);
                if (!(depth > R(0))) continue;
                if (n < cap) out[n] = double(depth);
                ++n;
            }
        }
        return n;
    }

    // body slot of a collider or -1 (entity_set.coll_bodies)
    int32_t cbody(uint32_t coll) const { return colliders[coll].body; }
    bool pair_kept(uint32_t hi, uint32_t lo) const {
        int32_t b0 = cbody(hi), b1 = cbody(lo);
        if (b0 < 0 && b1 < 0) return false;                               // physics.rs:579
        if (b0 >= 0 && b1 >= 0 && b0 == b1) return false;                 // physics.rs:551-555
        return true;
    }

    void tick(double frame_dt_d, bool has_scale, double scale, const SfgForceField& ff, int mode,
              int threads) override {
        housekeeping();
        // physics.rs:404-418
        size_t substeps;
        double dtd;
        if (!has_scale) { substeps = tuning.substeps; dtd = frame_dt_d / double(substeps); }
        else {
            substeps = size_t(std::ceil(scale * double(tuning.substeps)));
            dtd = scale * frame_dt_d / double(substeps);
        }
        stat_substeps = substeps;
        dt = R(dtd); inv_dt = R(1.0 / dtd); inv_dt_sq = R((1.0 / dtd) * (1.0 / dtd));
        const R frame_dt = R(frame_dt_d);
        frame_dt_f32_ = float(frame_dt_d);

        compute_aabbs(frame_dt);
        if (use_ext_pairs_) pair_keys = ext_pairs_;
        else if (mode == MODE_REFERENCE) broadphase_bvh();
        else broadphase_sweep();
        kept_pairs.clear();
        for (size_t p = 0; p + 1 < pair_keys.size(); p += 2)
            if (pair_kept(pair_keys[p], pair_keys[p + 1])) { kept_pairs.push_back(pair_keys[p]); kept_pairs.push_back(pair_keys[p + 1]); }

        const size_t nb = bodies.size();
        // rope links (physics.rs:823-838)
        rope_next.assign(nb, -1); rope_prev.assign(nb, -1);
        for (auto& r : ropes)
            for (size_t i = 0; i + 1 < r.particles.size(); ++i) {
                rope_next[size_t(r.particles[i])] = r.particles[i + 1];
                rope_prev[size_t(r.particles[i + 1])] = r.particles[i];
            }
        lateral_set.assign(nb, 0); lateral.assign(nb, V{});                // physics.rs:841-843
        old_poses.resize(nb); pre_contact_poses.resize(nb); old_lin.resize(nb); old_ang.resize(nb);
        for (size_t i = 0; i < nb; ++i) {                                  // physics.rs:845-854
            old_poses[i] = bodies[i].pose; pre_contact_poses[i] = bodies[i].pose;
            old_lin[i] = bodies[i].lin; old_ang[i] = bodies[i].ang;
        }
        ext_acc.assign(nb, V{});                                           // physics.rs:857-859

        if (mode == MODE_REFERENCE) tick_reference(ff, substeps, threads);
        else tick_coloured(ff, substeps);
    }

    // ------------------------------------------------------------------ reference order
    struct GEdge { uint8_t kind; uint32_t body, idx; };  // 0 rope,1 constraint,2 contact,3 static constraint,4 static contact
    void tick_reference(const SfgForceField& ff, size_t substeps, int threads) {
        const size_t nb = bodies.size();
        // constraint graph (physics.rs:488-581); adjacency in insertion order
        std::vector<std::vector<GEdge>> adj(nb);
        for (size_t ri = 0; ri < ropes.size(); ++ri) {
            auto& r = ropes[ri];
            for (size_t i = 0; i + 1 < r.particles.size(); ++i) {
                uint32_t a = uint32_t(r.particles[i]), b = uint32_t(r.particles[i + 1]);
                adj[a].push_back({0, b, uint32_t(ri)});
                adj[b].push_back({0, a, uint32_t(ri)});
            }
        }
        for (size_t ci = 0; ci < constraints.size(); ++ci) {
            auto& c = constraints[ci];
            if (c.target >= 0) {
                adj[size_t(c.owner)].push_back({1, uint32_t(c.target), uint32_t(ci)});
                adj[size_t(c.target)].push_back({1, uint32_t(c.owner), uint32_t(ci)});
            } else adj[size_t(c.owner)].push_back({3, 0, uint32_t(ci)});
        }
        for (size_t p = 0; p + 1 < pair_keys.size(); p += 2) {
            int32_t b1 = cbody(pair_keys[p]), b2 = cbody(pair_keys[p + 1]);
            uint32_t pi = uint32_t(p / 2);
            if (b1 >= 0 && b2 >= 0) {
                if (b1 == b2) continue;
                adj[size_t(b1)].push_back({2, uint32_t(b2), pi});
                adj[size_t(b2)].push_back({2, uint32_t(b1), pi});
            } else if (b1 >= 0) adj[size_t(b1)].push_back({4, 0, pi});
            else if (b2 >= 0) adj[size_t(b2)].push_back({4, 0, pi});
        }
        // islands (physics.rs:598-705); explicit-stack DFS with the recursive visit order
        std::vector<uint8_t> assigned(nb, 0);
        std::vector<uint32_t> fb, fr, fc, fp;   // sorted_first_pass
        std::vector<Island> isl;
        std::vector<std::pair<uint32_t, size_t>> stack;
        for (size_t root = 0; root < nb; ++root) {
            if (!bodies[root].alive || assigned[root]) continue;
            Island is;
            is.first_body = root;
            is.body_start = fb.size(); is.rope_start = fr.size(); is.constr_start = fc.size(); is.pair_start = fp.size();
            auto enter = [&](uint32_t b) {
                if (assigned[b]) return;
                assigned[b] = 1; fb.push_back(b); is.body_count++;
                stack.push_back({b, 0});
            };
            enter(uint32_t(root));
            while (!stack.empty()) {
                uint32_t b = stack.back().first;
                size_t& pos = stack.back().second;
                if (pos >= adj[b].size()) { stack.pop_back(); continue; }
                GEdge e = adj[b][pos++];
                switch (e.kind) {
                case 0: {
                    is.can_sleep = false;
                    bool seen = false;
                    for (size_t k = is.rope_start; k < is.rope_start + is.rope_count; ++k) if (fr[k] == e.idx) { seen = true; break; }
                    if (!seen) { fr.push_back(e.idx); is.rope_count++; }
                    is.edge_sum += uint64_t(b + 1) * uint64_t(e.body + 1);
                    enter(e.body);
                    break;
                }
                case 1:
                    fc.push_back(e.idx); is.constr_count++;
                    if (!constraints[e.idx].can_sleep) is.can_sleep = false;
                    is.edge_sum += uint64_t(b + 1) * uint64_t(e.body + 1);
                    enter(e.body);
                    break;
                case 2:
                    fp.push_back(e.idx); is.pair_count++;
                    is.edge_sum += uint64_t(b + 1) * uint64_t(e.body + 1);
                    enter(e.body);
                    break;
                case 3:
                    fc.push_back(e.idx); is.constr_count++;
                    if (!constraints[e.idx].can_sleep) is.can_sleep = false;
                    is.edge_sum += uint64_t(b + 1);
                    break;
                default:
                    fp.push_back(e.idx); is.pair_count++;
                    is.edge_sum += uint64_t(b + 1);
                    break;
                }
            }
            isl.push_back(is);
        }
        // sleeping filter (physics.rs:711-741)
        const bool sleeping_on = !(tuning.flags & SFG_TUNE_NO_SLEEPING) && tuning.fall_asleep_frames != UINT32_MAX;
        for (auto& s : sleeping) s.continues = false;
        all_islands.clear();
        std::vector<Island> kept;
        for (auto& is : isl) {
            bool keep = true;
            if (sleeping_on)
                for (auto& s : sleeping)
                    if (s.first_body == is.first_body && s.edge_sum == is.edge_sum) {
                        bool moving = false;
                        for (size_t k = is.body_start; k < is.body_start + is.body_count; ++k)
                            if (bodies[fb[k]].vel_mag_sq() >= R(tuning.sleep_vel_threshold)) { moving = true; break; }
                        if (moving) { keep = true; break; }
                        s.continues = true;
                        keep = s.ticks < tuning.fall_asleep_frames;
                        break;
                    }
            if (keep) kept.push_back(is);
            all_islands.push_back({is.first_body, is.edge_sum, is.body_count, uint64_t(is.can_sleep ? 0 : 1) | uint64_t(keep ? 0 : 8)});
        }
        sleeping.erase(std::remove_if(sleeping.begin(), sleeping.end(), [](const Sleeping& s) { return !s.continues; }), sleeping.end());
        std::stable_sort(kept.begin(), kept.end(), [](const Island& a, const Island& b) { return a.body_count > b.body_count; });
        // second pass: compact (physics.rs:746-774) + entry buffers (physics.rs:784-887)
        e_bodies.clear(); e_ropes.clear(); e_constr.clear(); e_pair.clear();
        for (auto& is : kept) {
            size_t s;
            s = e_bodies.size(); e_bodies.insert(e_bodies.end(), fb.begin() + long(is.body_start), fb.begin() + long(is.body_start + is.body_count)); is.body_start = s;
            s = e_ropes.size(); e_ropes.insert(e_ropes.end(), fr.begin() + long(is.rope_start), fr.begin() + long(is.rope_start + is.rope_count)); is.rope_start = s;
            s = e_constr.size(); e_constr.insert(e_constr.end(), fc.begin() + long(is.constr_start), fc.begin() + long(is.constr_start + is.constr_count)); is.constr_start = s;
            s = e_pair.size(); e_pair.insert(e_pair.end(), fp.begin() + long(is.pair_start), fp.begin() + long(is.pair_start + is.pair_count)); is.pair_start = s;
        }
        islands = kept;
        contacts.assign(e_pair.size(), ContactResult<R>{});
        last_contacts.assign(e_pair.size(), ContactResult<R>{});
        lambdas.assign(e_pair.size(), R(0));
        pair_units.clear(); constraint_units.clear();

        // island groups per thread (physics.rs:895-947), then the hot loop (:1070-1082)
        std::vector<size_t> group_sizes;
        size_t thread_count = threads > 1 ? size_t(threads) : 1;
        if (thread_count > 1 && !islands.empty()) {
            size_t live = 0; for (auto& b : bodies) live += b.alive;
            size_t ideal = (live + thread_count - 1) / thread_count;
            ideal = std::max<size_t>(ideal, 64);                          // min_bodies_per_thread
            size_t covered = 0, in_group = 0, next_split = ideal;
            for (size_t i = 0; i < islands.size(); ++i) {
                size_t after = covered + islands[i].body_count;
                if (i + 1 == islands.size()) { group_sizes.push_back(in_group + 1); continue; }
                if (after < next_split) in_group++;
                else {
                    if (after - next_split < next_split - covered) { group_sizes.push_back(in_group + 1); in_group = 0; }
                    else { group_sizes.push_back(in_group); in_group = 1; }
                    while (after > next_split) next_split += ideal;
                }
                covered = after;
            }
        } else group_sizes.push_back(islands.size());

        auto run_group = [&](size_t first, size_t count) {
            for (size_t i = first; i < first + count; ++i) {
                const Island& is = islands[i];
                View v{e_bodies.data() + is.body_start, is.body_count, e_ropes.data() + is.rope_start, is.rope_count,
                       e_constr.data() + is.constr_start, is.constr_count, e_pair.data() + is.pair_start, is.pair_count,
                       is.pair_start, 0};
                for (size_t s = 0; s < substeps; ++s) solve(ff, v);
            }
        };
        if (group_sizes.size() <= 1) run_group(0, islands.size());
        else {
            std::vector<std::thread> th;
            size_t first = 0;
            for (size_t g : group_sizes) { if (g) th.emplace_back(run_group, first, g); first += g; }
            for (auto& t : th) t.join();
        }

        // publish contacts (physics.rs:1097-1122)
        std::vector<ContactInfoS<R>> keep_c;
        for (auto& c : contact_infos)
            for (auto& s : sleeping)
                if (s.first_body == c.isl_first && s.edge_sum == c.isl_sum && s.ticks >= tuning.fall_asleep_frames) { keep_c.push_back(c); break; }
        contact_infos.swap(keep_c);
        for (auto& is : islands)
            for (size_t k = is.pair_start; k < is.pair_start + is.pair_count; ++k)
                if (last_contacts[k].count > 0)
                    contact_infos.push_back({pair_keys[2 * e_pair[k]], pair_keys[2 * e_pair[k] + 1], last_contacts[k].c[0].normal, is.first_body, is.edge_sum});
        // fall asleep (physics.rs:1128-1144)
        if (sleeping_on)
            for (auto& is : islands) {
                if (!is.can_sleep) continue;
                bool all_slow = true;
                for (size_t k = is.body_start; k < is.body_start + is.body_count; ++k)
                    if (!(bodies[e_bodies[k]].vel_mag_sq() < R(tuning.sleep_vel_threshold))) { all_slow = false; break; }
                if (!all_slow) continue;
                bool found = false;
                for (auto& s : sleeping) if (s.first_body == is.first_body && s.edge_sum == is.edge_sum) { s.ticks++; found = true; break; }
                if (!found) sleeping.push_back({is.first_body, is.edge_sum, false, 0});
            }
    }

    // ------------------------------------------------------------------ coloured order
    // Greedy colouring in descending priority == Jones-Plassmann fixed point (include/sfg_hash.h).
    template <class Unit, class BodiesOf, class Prio>
    static uint32_t colour_units(std::vector<Unit>& units, size_t n_bodies, BodiesOf bodies_of, Prio prio) {
        std::vector<uint32_t> order(units.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = uint32_t(i);
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return prio(units[b]) < prio(units[a]); });
        std::vector<std::vector<uint64_t>> used(n_bodies);
        uint32_t n_colours = 0;
        for (uint32_t ui : order) {
            int32_t bs[2]; bodies_of(units[ui], bs);
            uint32_t c = 0;
            for (;; ++c) {
                bool taken = false;
                for (int k = 0; k < 2; ++k)
                    if (bs[k] >= 0) { auto& u = used[size_t(bs[k])]; if (c / 64 < u.size() && ((u[c / 64] >> (c % 64)) & 1)) taken = true; }
                if (!taken) break;
            }
            units[ui].colour = c;
            for (int k = 0; k < 2; ++k)
                if (bs[k] >= 0) { auto& u = used[size_t(bs[k])]; if (u.size() <= c / 64) u.resize(c / 64 + 1, 0); u[c / 64] |= 1ull << (c % 64); }
            n_colours = std::max(n_colours, c + 1);
        }
        return n_colours;
    }
    int32_t dyn_body(int32_t b) const { return (b >= 0 && bodies[size_t(b)].sees_forces()) ? b : -1; }
    // include/sfg_hash.h: f32 on purpose, whatever R is (the state a tick starts from is the f32 upload or an f32 read-back)
    uint32_t compute_near_hint(uint32_t hi, uint32_t lo) const {
        float c[2][2], r[2], v[2][2];
        const uint32_t ids[2] = {hi, lo};
        for (int k = 0; k < 2; ++k) {
            const ColliderS<R>& co = colliders[ids[k]];
            r[k] = sfg_bound_radius(co.shape.kind, float(co.shape.p0), float(co.shape.p1), float(co.shape.circle_r));
            v[k][0] = v[k][1] = 0.0f;
            if (co.body < 0) { c[k][0] = float(co.pose.t.x); c[k][1] = float(co.pose.t.y); }
            else {
                const BodyS<R>& b = bodies[size_t(co.body)];
                sfg_collider_centre(float(b.pose.t.x), float(b.pose.t.y), float(b.pose.r.s), float(b.pose.r.b), float(co.pose.t.x), float(co.pose.t.y),
                                    &c[k][0], &c[k][1]);
                v[k][0] = float(b.lin.x); v[k][1] = float(b.lin.y);
            }
        }
        return sfg_near_hint(c[0][0], c[0][1], r[0], c[1][0], c[1][1], r[1], v[0][0], v[0][1], v[1][0], v[1][1], float(frame_dt_f32_));
    }

    void tick_coloured(const SfgForceField& ff, size_t substeps) {
        // units
        pair_units.clear();
        for (size_t p = 0; p + 1 < kept_pairs.size(); p += 2) {
            uint32_t hi = kept_pairs[p], lo = kept_pairs[p + 1];
            // the `near` bit of the priority: computed here from the tick-start state with the shared exactly-rounded f32
            // definition (include/sfg_hash.h sfg_near_hint); hints set from outside (tests of the priority rule) override it
            uint32_t near_hint = compute_near_hint(hi, lo);
            if (!ext_hints_.empty()) { auto it = ext_hints_.find((uint64_t(hi) << 32) | lo); if (it != ext_hints_.end()) near_hint = it->second ? 1u : 0u; }
            pair_units.push_back({hi, lo, (cbody(hi) >= 0 && cbody(lo) >= 0) ? 2u : 1u, 0, near_hint});
        }
        constraint_units.clear();
        for (size_t i = 0; i < constraints.size(); ++i) constraint_units.push_back({uint32_t(i), constraints[i].target >= 0 ? 2u : 1u, 0});
        typedef std::array<uint32_t, 4> P4;
        stat_colours_pairs = colour_units(pair_units, bodies.size(),
            [&](const PairUnit& u, int32_t* bs) { bs[0] = dyn_body(cbody(u.hi)); bs[1] = dyn_body(cbody(u.lo)); },
            [](const PairUnit& u) { return P4{{sfg_pair_priority(u.hi, u.lo, u.near_hint), u.hi, u.lo, 0}}; });
        stat_colours_constr = colour_units(constraint_units, bodies.size(),
            [&](const ConstraintUnit& u, int32_t* bs) { bs[0] = dyn_body(constraints[u.idx].owner); bs[1] = dyn_body(constraints[u.idx].target); },
            [](const ConstraintUnit& u) { return P4{{sfg_hash3(u.idx, SFG_CONSTRAINT_TAG, 0), u.idx, 0, 0}}; });
        // colour-major order; inside a colour the order cannot matter (no shared dynamic body);
        // we keep (hi, lo) ascending.
        std::stable_sort(pair_units.begin(), pair_units.end(), [](const PairUnit& a, const PairUnit& b) { return a.colour < b.colour; });
        std::stable_sort(constraint_units.begin(), constraint_units.end(), [](const ConstraintUnit& a, const ConstraintUnit& b) { return a.colour < b.colour; });
        // entries: unit u is visited `mult` times back to back
        pair_keys.clear();
        e_pair.clear();
        for (size_t u = 0; u < pair_units.size(); ++u) {
            pair_keys.push_back(pair_units[u].hi); pair_keys.push_back(pair_units[u].lo);
            for (uint32_t m = 0; m < pair_units[u].mult; ++m) e_pair.push_back(uint32_t(u));
        }
        e_constr.clear();
        for (auto& u : constraint_units) for (uint32_t m = 0; m < u.mult; ++m) e_constr.push_back(u.idx);
        e_bodies.clear();
        for (size_t i = 0; i < bodies.size(); ++i) if (bodies[i].alive) e_bodies.push_back(uint32_t(i));
        e_ropes.clear();
        for (size_t i = 0; i < ropes.size(); ++i) e_ropes.push_back(uint32_t(i));
        contacts.assign(e_pair.size(), ContactResult<R>{});
        last_contacts.assign(e_pair.size(), ContactResult<R>{});
        lambdas.assign(e_pair.size(), R(0));
        islands.clear();
        View v{e_bodies.data(), e_bodies.size(), e_ropes.data(), e_ropes.size(), e_constr.data(), e_constr.size(),
               e_pair.data(), e_pair.size(), 0, 1};
        for (size_t s = 0; s < substeps; ++s) solve(ff, v);
        contact_infos.clear();
        for (size_t k = 0; k < e_pair.size(); ++k)
            if (last_contacts[k].count > 0)
                contact_infos.push_back({pair_keys[2 * e_pair[k]], pair_keys[2 * e_pair[k] + 1], last_contacts[k].c[0].normal, 0, 0});
    }

    // ------------------------------------------------------------------ solver (solver.rs)
    struct View {
        const uint32_t* body_list; size_t n_bodies;
        const uint32_t* rope_list; size_t n_ropes;
        const uint32_t* constr_list; size_t n_constr;
        const uint32_t* pair_list; size_t n_pairs;
        size_t entry_base;       // offset of this view's first contact entry in contacts/lambdas
        int rope_colouring;      // 0 = sequential along the chain (reference), 1 = colour-major (i mod 3 / i mod 2)
    };

    void solve(const SfgForceField& ff, const View& v) {                  // solver.rs:55-108
        for (size_t k = 0; k < v.n_bodies; ++k) {                          // :57-74
            const uint32_t i = v.body_list[k];
            BodyS<R>& b = bodies[i];
            if (!b.ignores_gravity && b.mass_finite) {
                V a = ff_value(ff, b.pose.t);
                b.lin += a * dt;
                ext_acc[i] = a;
            }
            old_lin[i] = b.lin; old_ang[i] = b.ang;
            old_poses[i] = b.pose;
            b.pose.append_translation(b.lin * dt);                        // physics.rs:78-83
            b.pose.prepend_rotation(Rotor2<R>::from_angle(b.ang * dt));
        }
        if (v.n_ropes) solve_ropes(v);
        if (v.n_constr) solve_constraints(v);
        for (size_t k = 0; k < v.n_bodies; ++k) pre_contact_poses[v.body_list[k]] = bodies[v.body_list[k]].pose;
        if (v.n_pairs) solve_contacts(v);
        for (size_t k = 0; k < v.n_bodies; ++k) {                          // :91-97
            const uint32_t i = v.body_list[k];
            BodyS<R>& b = bodies[i];
            b.lin = (b.pose.t - old_poses[i].t) * inv_dt;
            Rotor2<R> d = b.pose.r * old_poses[i].r.reversed();
            R angle_diff = -std::atan2(d.b, d.s) * R(2);
            b.ang = angle_diff * inv_dt;
        }
        if (v.n_pairs) contact_velocity_step(v);
        if (v.n_constr) constraint_damping(v);
        if (v.n_ropes) rope_velocity_step(v);
    }

    // one step of the chain walk of solver.rs:118-179: distance(curr,next) then bending(curr,next,after)
    void rope_step(const RopeS<R>& rope, int32_t curr, int32_t next) {
        BodyS<R>&bc = bodies[size_t(curr)], &bn = bodies[size_t(next)];
        V dist = bn.pose.t - bc.pose.t;
        R dist_mag = dist.mag();
        V dir = dist / dist_mag;
        R error = rope.spacing - dist_mag;
        R lambda = -error / (bc.inv_mass + bn.inv_mass + rope.compliance * inv_dt_sq);
        bc.pose.append_translation(bc.inv_mass * lambda * dir);
        bn.pose.append_translation(-bn.inv_mass * lambda * dir);
        int32_t after = rope_next[size_t(next)];
        if (after < 0) return;
        BodyS<R>& ba = bodies[size_t(after)];
        V curr_to_next = bn.pose.t - bc.pose.t;
        V next_to_after = ba.pose.t - bn.pose.t;
        R angle = std::acos(next_to_after.normalized().dot(curr_to_next.normalized()));
        R err = angle - rope.bending_max_angle;
        if (err > R(0)) {
            R lam = -err / (ba.inv_mass + rope.bending_compliance * inv_dt_sq);
            R lam_o = left_normal(curr_to_next).dot(next_to_after) > R(0) ? lam : -lam;
            Rotor2<R> corr = Rotor2<R>::from_angle(lam_o * ba.inv_mass);
            V old_pos = ba.pose.t;
            ba.pose.t = bn.pose.t + corr * next_to_after;
            lateral_set[size_t(after)] = 1;
            lateral[size_t(after)] = ba.pose.t - old_pos;
        }
    }
    void solve_ropes(const View& v) {                                      // solver.rs:114-181
        for (size_t k = 0; k < v.n_ropes; ++k) {
            const RopeS<R>& rope = ropes[v.rope_list[k]];
            const size_t n = rope.particles.size();
            if (n < 2) continue;   // reference: expect("Rope only had one particle") panics
            if (!v.rope_colouring) {
                for (size_t i = 0; i + 1 < n; ++i) rope_step(rope, rope.particles[i], rope.particles[i + 1]);
            }
        }
        if (v.rope_colouring)
            for (size_t c = 0; c < 3; ++c)
                for (size_t k = 0; k < v.n_ropes; ++k) {
                    const RopeS<R>& rope = ropes[v.rope_list[k]];
                    for (size_t i = c; i + 1 < rope.particles.size(); i += 3) rope_step(rope, rope.particles[i], rope.particles[i + 1]);
                }
    }

    void solve_constraints(const View& v) {                                // solver.rs:187-269
        for (size_t k = 0; k < v.n_constr; ++k) {
            const ConstraintS<R>& c = constraints[v.constr_list[k]];
            BodyS<R>& b0 = bodies[size_t(c.owner)];
            BodyS<R>* b1 = c.target >= 0 ? &bodies[size_t(c.target)] : nullptr;
            R im[2] = {b0.inv_mass, b1 ? b1->inv_mass : R(0)};
            R ii[2] = {b0.inv_inertia, b1 ? b1->inv_inertia : R(0)};
            V ow[2] = {b0.pose * c.off[0], b1 ? b1->pose * c.off[1] : c.off[1]};
            V actual = ow[1] - ow[0];
            R mag = actual.mag();
            R error = c.distance - mag;
            bool act = c.limit == SFG_LIMIT_EQ || (c.limit == SFG_LIMIT_LT && error < R(0)) || (c.limit == SFG_LIMIT_GT && error > R(0));
            if (!act) continue;
            V dir = mag != R(0) ? actual / mag : V{0, 1};
            if (b1) {
                V orot[2] = {b0.pose.r * c.off[0], b1->pose.r * c.off[1]};
                R wd[2] = {orot[0].wedge(dir), orot[1].wedge(dir)};
                R eim[2] = {im[0] + (wd[0] * wd[0] * ii[0]), im[1] + (wd[1] * wd[1] * ii[1])};
                R lambda = -error / (eim[0] + eim[1] + c.compliance * inv_dt_sq);
                b0.pose.append_translation(im[0] * lambda * dir);
                b0.pose.prepend_rotation(Rotor2<R>::from_angle(ii[0] * lambda * wd[0]));
                b1->pose.append_translation(-im[1] * lambda * dir);
                b1->pose.prepend_rotation(Rotor2<R>::from_angle(-ii[1] * lambda * wd[1]));
            } else {
                V orot = b0.pose.r * c.off[0];
                R wd = orot.wedge(dir);
                R eim = im[0] + wd * wd * ii[0];
                R lambda = -error / (eim + c.compliance * inv_dt_sq);
                b0.pose.append_translation(im[0] * lambda * dir);
                b0.pose.prepend_rotation(Rotor2<R>::from_angle(ii[0] * lambda * wd));
            }
        }
    }

    void apply_pos(int32_t bi, R sign, R lambda, V dirn, R wedge) {       // solver.rs:435-452 pattern
        if (bi < 0) return;
        BodyS<R>& b = bodies[size_t(bi)];
        b.pose.append_translation(sign * b.inv_mass * lambda * dirn);
        b.pose.prepend_rotation(Rotor2<R>::from_angle(sign * b.inv_inertia * lambda * wedge));
    }

    void solve_contacts(const View& v) {                                   // solver.rs:275-490
        for (size_t k = 0; k < v.n_pairs; ++k) {
            const size_t e = v.entry_base + k;
            const uint32_t ck[2] = {pair_keys[2 * v.pair_list[k]], pair_keys[2 * v.pair_list[k] + 1]};
            const int32_t bd[2] = {cbody(ck[0]), cbody(ck[1])};
            ContactResult<R>& contact = contacts[e];
            R& lambda_n = lambdas[e];
            if (bd[0] < 0 && bd[1] < 0) { contact = {}; continue; }
            const ColliderS<R>* cl[2] = {&colliders[ck[0]], &colliders[ck[1]]};
            Iso2<R> pw[2];
            for (int i = 0; i < 2; ++i) pw[i] = bd[i] >= 0 ? bodies[size_t(bd[i])].pose * cl[i]->pose : cl[i]->pose;
            Shape<R> sh[2] = {cl[0]->shape, cl[1]->shape};
            contact = intersection_check(pw, sh, false);
            for (int c = 0; c < contact.count; ++c)                        // :308-315
                for (int i = 0; i < 2; ++i)
                    if (bd[i] >= 0) contact.c[c].offsets[i] = cl[i]->pose * contact.c[c].offsets[i];
            if (contact.count > 0) last_contacts[e] = contact;             // :318-320
            bool sf0 = bd[0] >= 0 && bodies[size_t(bd[0])].sees_forces();
            bool sf1 = bd[1] >= 0 && bodies[size_t(bd[1])].sees_forces();
            if (!sf0 && !sf1) continue;                                    // :323-331
            if (contact.count == 1) {                                      // :337-362
                const R ndir[2] = {R(1), R(-1)};
                for (int i = 0; i < 2; ++i) {
                    if (bd[i] < 0) continue;
                    int32_t prev = rope_prev[size_t(bd[i])], next = rope_next[size_t(bd[i])];
                    if (prev < 0 || next < 0) continue;
                    V no = contact.c[0].normal * ndir[i];
                    V to_prev = pre_contact_poses[size_t(prev)].t - pre_contact_poses[size_t(bd[i])].t;
                    V to_next = pre_contact_poses[size_t(next)].t - pre_contact_poses[size_t(bd[i])].t;
                    V nn = no.dot(to_prev) > no.dot(to_next) ? left_normal(to_prev).normalized() : left_normal(to_next).normalized();
                    contact.c[0].normal = contact.c[0].normal.dot(nn) > R(0) ? nn : -nn;
                }
            }
            if (cl[0]->sensor || cl[1]->sensor) continue;                  // :364-370
            const bool has_sf = cl[0]->has_sf && cl[1]->has_sf;
            const R mu_s = (cl[0]->sf + cl[1]->sf) / R(2);
            for (int c = 0; c < contact.count; ++c) {
                const Contact<R>& ct = contact.c[c];
                V tangent = left_normal(ct.normal);
                V ow[2], owo[2]; R wn[2], en[2], wt[2], et[2];
                for (int i = 0; i < 2; ++i) {
                    if (bd[i] < 0) {
                        ow[i] = cl[i]->pose * ct.offsets[i];
                        wn[i] = 0; en[i] = 0; owo[i] = ow[i]; wt[i] = 0; et[i] = 0;
                    } else {
                        const BodyS<R>& b = bodies[size_t(bd[i])];
                        V orot = b.pose.r * ct.offsets[i];
                        wn[i] = orot.wedge(ct.normal);
                        wt[i] = orot.wedge(tangent);
                        ow[i] = b.pose * ct.offsets[i];
                        en[i] = b.inv_mass + (wn[i] * wn[i] * b.inv_inertia);
                        owo[i] = old_poses[size_t(bd[i])] * ct.offsets[i];
                        et[i] = b.inv_mass + (wt[i] * wt[i] * b.inv_inertia);
                    }
                }
                R depth = (ow[0] - ow[1]).dot(ct.normal);
                if (depth <= R(0)) { lambda_n = R(0); continue; }
                lambda_n = -depth / (en[0] + en[1]);
                apply_pos(bd[0], R(1), lambda_n, ct.normal, wn[0]);
                apply_pos(bd[1], R(-1), lambda_n, ct.normal, wn[1]);
                if (has_sf) {                                              // :456-487
                    V odm = (ow[0] - owo[0]) - (ow[1] - owo[1]);
                    R mat = odm.dot(tangent);
                    R max_dx = lambda_n * mu_s;
                    R lambda_t = -mat / (et[0] + et[1]);
                    if (lambda_t < max_dx) {
                        apply_pos(bd[0], R(1), lambda_t, tangent, wt[0]);
                        apply_pos(bd[1], R(-1), lambda_t, tangent, wt[1]);
                    }
                }
            }
        }
    }

    void contact_velocity_step(const View& v) {                            // solver.rs:496-618
        for (size_t k = 0; k < v.n_pairs; ++k) {
            const size_t e = v.entry_base + k;
            const uint32_t ck[2] = {pair_keys[2 * v.pair_list[k]], pair_keys[2 * v.pair_list[k] + 1]};
            const ColliderS<R>* cl[2] = {&colliders[ck[0]], &colliders[ck[1]]};
            if (cl[0]->sensor || cl[1]->sensor) continue;
            const int32_t bd[2] = {cbody(ck[0]), cbody(ck[1])};
            bool sf0 = bd[0] >= 0 && bodies[size_t(bd[0])].sees_forces();
            bool sf1 = bd[1] >= 0 && bodies[size_t(bd[1])].sees_forces();
            if (!sf0 && !sf1) continue;
            const ContactResult<R>& contact = contacts[e];
            const R lambda_n = lambdas[e];
            for (int c = 0; c < contact.count; ++c) {
                const Contact<R>& ct = contact.c[c];
                R im[2], ii[2]; V orot[2], pv[2], opv[2], ea[2];
                for (int i = 0; i < 2; ++i) {
                    if (bd[i] < 0) {
                        im[i] = 0; ii[i] = 0; orot[i] = cl[i]->pose.r * ct.offsets[i];
                        pv[i] = {}; opv[i] = {}; ea[i] = {};
                    } else {
                        const BodyS<R>& b = bodies[size_t(bd[i])];
                        orot[i] = b.pose.r * ct.offsets[i];
                        im[i] = b.inv_mass; ii[i] = b.inv_inertia;
                        pv[i] = b.point_velocity(orot[i]);
                        opv[i] = old_lin[size_t(bd[i])] + left_normal(orot[i]) * old_ang[size_t(bd[i])];
                        ea[i] = ext_acc[size_t(bd[i])];
                    }
                }
                V rel = pv[0] - pv[1];
                R normal_vel = rel.dot(ct.normal);
                R old_normal_vel = (opv[0] - opv[1]).dot(ct.normal);
                R rest = (old_normal_vel * old_normal_vel < dt * dt * (ea[0] + ea[1]).mag_sq())
                             ? R(0) : rs_max(cl[0]->rest, cl[1]->rest);
                R dnv = -normal_vel - rest * rs_max(old_normal_vel, R(0));
                V tangent = left_normal(ct.normal);
                R dtv = 0;
                if (cl[0]->has_df && cl[1]->has_df) {
                    R mu = (cl[0]->df + cl[1]->df) / R(2);
                    R tv = rel.dot(tangent);
                    R max_dv = inv_dt * lambda_n * mu;
                    dtv = rs_min(std::fabs(tv), std::fabs(max_dv)) * -rs_signum(tv);
                }
                V upd = dnv * ct.normal + dtv * tangent;
                R mag = upd.mag();
                if (mag < R(0.0001)) continue;
                V dirn = upd / mag;
                R wd[2] = {orot[0].wedge(dirn), orot[1].wedge(dirn)};
                R eim[2] = {im[0] + (wd[0] * wd[0] * ii[0]), im[1] + (wd[1] * wd[1] * ii[1])};
                R imp = mag / (eim[0] + eim[1]);
                if (bd[0] >= 0) { BodyS<R>& b = bodies[size_t(bd[0])]; b.lin += im[0] * imp * dirn; b.ang += ii[0] * imp * wd[0]; }
                if (bd[1] >= 0) { BodyS<R>& b = bodies[size_t(bd[1])]; b.lin -= im[1] * imp * dirn; b.ang -= ii[1] * imp * wd[1]; }
            }
        }
    }

    void constraint_damping(const View& v) {                               // solver.rs:624-709
        for (size_t k = 0; k < v.n_constr; ++k) {
            const ConstraintS<R>& c = constraints[v.constr_list[k]];
            BodyS<R>& b0 = bodies[size_t(c.owner)];
            if (c.target >= 0) {
                BodyS<R>& b1 = bodies[size_t(c.target)];
                V orot[2] = {b0.pose.r * c.off[0], b1.pose.r * c.off[1]};
                V rel = b0.point_velocity(orot[0]) - b1.point_velocity(orot[1]);
                R mag = rel.mag();
                if (mag == R(0)) continue;
                V dir = rel / mag;
                R wd[2] = {orot[0].wedge(dir), orot[1].wedge(dir)};
                R eim[2] = {b0.inv_mass + (wd[0] * wd[0] * b0.inv_inertia), b1.inv_mass + (wd[1] * wd[1] * b1.inv_inertia)};
                R upd = -mag * rs_min(c.linear_damping * dt, R(1));
                R imp = upd / (eim[0] + eim[1]);
                b0.lin += b0.inv_mass * imp * dir; b0.ang += b0.inv_inertia * imp * wd[0];
                b1.lin -= b1.inv_mass * imp * dir; b1.ang -= b1.inv_inertia * imp * wd[1];
                if (c.angular_damping > R(0)) {
                    R relw = b0.ang - b1.ang;
                    R au = -relw * rs_min(c.angular_damping * dt, R(1));
                    R ai = au / (eim[0] + eim[1]);
                    b1.ang -= b1.inv_inertia * ai;
                    b0.ang += b0.inv_inertia * ai;
                }
            } else {
                V orot = b0.pose.r * c.off[0];
                V pv = b0.point_velocity(orot);
                R mag = pv.mag();
                if (mag == R(0)) continue;
                V dir = pv / mag;
                R wd = orot.wedge(dir);
                R eim = b0.inv_mass + wd * wd * b0.inv_inertia;
                R upd = -mag * rs_min(c.linear_damping * dt, R(1));
                R imp = upd / eim;
                b0.lin += b0.inv_mass * imp * dir; b0.ang += b0.inv_inertia * imp * wd;
                if (c.angular_damping > R(0)) {
                    R au = b0.ang * rs_min(c.angular_damping * dt, R(1));
                    R ai = -au / eim;
                    b0.ang += b0.inv_inertia * ai;
                }
            }
        }
    }

    void rope_damp_pair(const RopeS<R>& rope, int32_t curr, int32_t next) {  // solver.rs:722-740
        BodyS<R>&bc = bodies[size_t(curr)], &bn = bodies[size_t(next)];
        V rel = bc.lin - bn.lin;
        R mag = rel.mag();
        if (mag != R(0)) {
            V dir = rel / mag;
            R upd = -mag * rs_min(rope.damping * dt, R(1));
            R imp = upd / (bc.inv_mass + bn.inv_mass);
            bc.lin += bc.inv_mass * imp * dir;
            bn.lin -= bn.inv_mass * imp * dir;
        }
    }
    void rope_lateral_fix(const RopeS<R>& rope, int32_t p) {               // solver.rs:750-765
        if (!lateral_set[size_t(p)]) return;
        BodyS<R>& b = bodies[size_t(p)];
        V corr = lateral[size_t(p)];
        R cm = corr.mag();
        R vfc = cm * inv_dt;
        V dir = corr / cm;
        R vid = b.lin.dot(dir);
        R vc = rs_max(rs_min(vid, vfc), -vfc);
        R imp = -vc / (b.inv_mass + rope.bending_compliance * inv_dt);
        b.lin += b.inv_mass * imp * dir;
    }
    void rope_velocity_step(const View& v) {                               // solver.rs:715-772
        if (!v.rope_colouring) {
            for (size_t k = 0; k < v.n_ropes; ++k) {
                const RopeS<R>& rope = ropes[v.rope_list[k]];
                for (size_t i = 0; i + 1 < rope.particles.size(); ++i) rope_damp_pair(rope, rope.particles[i], rope.particles[i + 1]);
                for (int32_t p : rope.particles) rope_lateral_fix(rope, p);
            }
            return;
        }
        for (size_t c = 0; c < 2; ++c)
            for (size_t k = 0; k < v.n_ropes; ++k) {
                const RopeS<R>& rope = ropes[v.rope_list[k]];
                for (size_t i = c; i + 1 < rope.particles.size(); i += 2) rope_damp_pair(rope, rope.particles[i], rope.particles[i + 1]);
            }
        for (size_t k = 0; k < v.n_ropes; ++k) {
            const RopeS<R>& rope = ropes[v.rope_list[k]];
            for (int32_t p : rope.particles) rope_lateral_fix(rope, p);
        }
    }

    // ------------------------------------------------------------------ queries
    // physics.rs:1266-1311. use_bvh = 1 walks the tick-time BVH best-first like the reference
    // (requires a MODE_REFERENCE tick); 0 evaluates the closed form:
    // min t over solid colliders whose tick-time AABB padded by radius is entered at
    // t_c < max_distance, exact hit t <= max_distance; ties keep the lowest collider slot.
    CastHit<R> cast_one(const SfgRay& sr, int32_t* coll_out, int use_bvh) const {
        Ray<R> ray{{R(sr.sx), R(sr.sy)}, {R(sr.dx), R(sr.dy)}};
        const R radius = R(sr.radius), max_d = R(sr.max_distance);
        CastHit<R> best; int32_t best_c = -1;
        auto try_coll = [&](uint32_t ci) {
            const auto& c = colliders[ci];
            if (!c.alive || c.sensor || c.domain != sr.domain) return;
            CastHit<R> h = spherecast_collider(ray, radius, collider_world_pose(c), c.shape);
            if (!h.hit || !(h.t <= max_d)) return;
            if (best.hit && best.t <= h.t) return;
            best = h; best_c = int32_t(ci);
        };
        if (use_bvh) {
            bvh.sweep_aabb(radius, ray, max_d, [&](R t, uint32_t ci) {
                if (t >= max_d || (best.hit && t >= best.t)) return false;
                try_coll(ci);
                return true;
            });
        } else {
            for (size_t ci = 0; ci < colliders.size() && ci < aabbs.size(); ++ci) {
                if (!colliders[ci].alive) continue;
                R t;
                if (!ray_aabb(ray, aabbs[ci].padded(radius), &t)) continue;
                if (t >= max_d) continue;
                try_coll(uint32_t(ci));
            }
        }
        *coll_out = best_c;
        return best;
    }
    void cast_batch(uint32_t n, const SfgRay* rays, double* out, int use_bvh) const override {
        for (uint32_t i = 0; i < n; ++i) {
            int32_t c;
            CastHit<R> h = cast_one(rays[i], &c, use_bvh);
            double* o = out + 6 * size_t(i);
            o[0] = c; o[1] = h.t; o[2] = h.point.x; o[3] = h.point.y; o[4] = h.normal.x; o[5] = h.normal.y;
        }
    }
    uint32_t query_point(double x, double y, uint32_t domain, int32_t* out, uint32_t cap) const override {  // physics.rs:1188-1210
        uint32_t n = 0;
        V p{R(x), R(y)};
        for (size_t ci = 0; ci < colliders.size() && ci < aabbs.size(); ++ci) {
            const auto& c = colliders[ci];
            if (!c.alive || c.domain != domain) continue;
            if (!aabbs[ci].contains_point(p)) continue;
            if (point_collider_bool(p, collider_world_pose(c), c.shape)) { if (n < cap) out[n] = int32_t(ci); ++n; }
        }
        return n;
    }

    uint32_t query_shape(const SfgShapeQuery& q, int32_t* out, uint32_t cap) const override {   // physics.rs:1215-1245
        uint32_t n = 0;
        Iso2<R> poses[2];
        Shape<R> shapes[2];
        poses[0] = {{R(q.x), R(q.y)}, {R(q.rot_s), R(q.rot_b)}};
        shapes[0] = {q.kind, R(q.p0), R(q.p1), R(q.circle_r)};
        const AABB<R> box = shapes[0].aabb(poses[0]);
        for (size_t ci = 0; ci < colliders.size() && ci < aabbs.size(); ++ci) {
            const auto& c = colliders[ci];
            if (!c.alive || c.domain != q.domain) continue;
            if (!aabbs[ci].intersects(box)) continue;                    // bvh.test_aabb: strict overlap with the tick-time box
            if (!((q.layer_mask >> c.layer) & 1ull)) continue;             // mask.get(coll.layer)
            poses[1] = collider_world_pose(c); shapes[1] = c.shape;
            if (intersection_check(poses, shapes, false).is_zero()) continue;
            if (n < cap) out[n] = int32_t(ci);
            ++n;
        }
        return n;
    }

    // ------------------------------------------------------------------ readback
    void get_bodies(uint32_t first, uint32_t count, double* o) const override {
        for (uint32_t i = 0; i < count; ++i) {
            const BodyS<R>& b = bodies[first + i];
            double* p = o + 7 * size_t(i);
            p[0] = b.pose.t.x; p[1] = b.pose.t.y; p[2] = b.pose.r.s; p[3] = b.pose.r.b; p[4] = b.lin.x; p[5] = b.lin.y; p[6] = b.ang;
        }
    }
    size_t n_pairs() const override { return kept_pairs.size() / 2; }
    void get_pairs(uint32_t* out) const override { std::memcpy(out, kept_pairs.data(), kept_pairs.size() * 4); }
    size_t n_pair_units() const override { return pair_units.size(); }
    void get_pair_units(uint32_t* hi, uint32_t* lo, uint32_t* mult, uint32_t* colour) const override {
        for (size_t i = 0; i < pair_units.size(); ++i) { hi[i] = pair_units[i].hi; lo[i] = pair_units[i].lo; mult[i] = pair_units[i].mult; colour[i] = pair_units[i].colour; }
    }
    void get_pair_unit_hints(uint8_t* near_hint) const override { for (size_t i = 0; i < pair_units.size(); ++i) near_hint[i] = uint8_t(pair_units[i].near_hint); }
    size_t n_constraint_units() const override { return constraint_units.size(); }
    void get_constraint_units(uint32_t* idx, uint32_t* mult, uint32_t* colour) const override {
        for (size_t i = 0; i < constraint_units.size(); ++i) { idx[i] = constraint_units[i].idx; mult[i] = constraint_units[i].mult; colour[i] = constraint_units[i].colour; }
    }
    size_t n_entries() const override { return e_pair.size(); }
    // per entry 14 doubles: hi, lo, count, lambda, normal.xy, 2 x (offsets[0].xy, offsets[1].xy)
    void get_manifolds(double* o) const override {
        for (size_t e = 0; e < e_pair.size(); ++e) {
            double* p = o + 14 * e;
            const ContactResult<R>& c = contacts[e];
            p[0] = pair_keys[2 * e_pair[e]]; p[1] = pair_keys[2 * e_pair[e] + 1];
            p[2] = c.count; p[3] = lambdas[e];
            p[4] = c.count ? c.c[0].normal.x : 0; p[5] = c.count ? c.c[0].normal.y : 0;
            for (int k = 0; k < 2; ++k) {
                bool on = k < c.count;
                p[6 + 4 * k] = on ? c.c[k].offsets[0].x : 0; p[7 + 4 * k] = on ? c.c[k].offsets[0].y : 0;
                p[8 + 4 * k] = on ? c.c[k].offsets[1].x : 0; p[9 + 4 * k] = on ? c.c[k].offsets[1].y : 0;
            }
        }
    }
    size_t n_contacts() const override { return contact_infos.size(); }
    void get_contacts(uint32_t* c01, double* nxy) const override {
        for (size_t i = 0; i < contact_infos.size(); ++i) {
            c01[2 * i] = contact_infos[i].c0; c01[2 * i + 1] = contact_infos[i].c1;
            nxy[2 * i] = contact_infos[i].normal.x; nxy[2 * i + 1] = contact_infos[i].normal.y;
        }
    }
    void get_aabbs(double* o) const override {
        for (size_t i = 0; i < aabbs.size(); ++i) { o[4 * i] = aabbs[i].min.x; o[4 * i + 1] = aabbs[i].min.y; o[4 * i + 2] = aabbs[i].max.x; o[4 * i + 3] = aabbs[i].max.y; }
    }
    size_t n_all_islands() const override { return all_islands.size(); }
    void get_all_islands(uint64_t* first_body, uint64_t* edge_sum, uint64_t* body_count, uint64_t* flags) const override {
        for (size_t i = 0; i < all_islands.size(); ++i) {
            first_body[i] = all_islands[i].first_body; edge_sum[i] = all_islands[i].edge_sum;
            body_count[i] = all_islands[i].body_count; flags[i] = all_islands[i].flags;
        }
    }
    void get_stats(uint64_t* o) const override {
        o[0] = kept_pairs.size() / 2; o[1] = e_pair.size(); o[2] = e_constr.size(); o[3] = stat_colours_pairs;
        o[4] = stat_colours_constr; o[5] = islands.size(); o[6] = sleeping.size(); o[7] = stat_substeps;
    }

private:
    std::vector<uint32_t> ext_pairs_;
    bool use_ext_pairs_ = false;
    std::unordered_map<uint64_t, uint8_t> ext_hints_;
    float frame_dt_f32_ = 1.0f / 60.0f;
};

}  // namespace sfo

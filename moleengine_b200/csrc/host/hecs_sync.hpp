// starframe::HecsSyncManager on top of the C++ PhysicsWorld mirror — the ECS <-> physics pose sync that brackets every tick in
// the reference's game loop (Game::physics_tick, /root/reference/src/game.rs:297-303):
//     sync_hecs_to_physics  (physics/hecs_sync.rs:121-180)   auto-register, auto-delete physics objects, poses ECS -> physics
//     PhysicsWorld::tick
//     sync_physics_to_hecs  (physics/hecs_sync.rs:185-227)   auto-delete ECS entities, poses physics -> ECS
// Same registration maps (indexed by the body / collider key), the same four options and presets (hecs_sync.rs:9-56), the same
// rules: a bodied collider's entity drives / follows the BODY's pose composed with the collider's local pose, entities without a
// Pose component are skipped, a mapping whose physics object is gone is dropped even when auto_delete_hecs is off.
// The reference's ECS is the `hecs` crate; here it is the smallest interface the sync needs (EcsWorld below), so that the logic
// can be compiled and tested on its own (tests/host_cpp/hecs_sync_test.cpp runs without a GPU: it only needs `entity_set`).
#pragma once
#include <cstdint>
#include <map>
#include <optional>

#include "physics_world.hpp"

namespace starframe {

// math.rs:96-143 Pose: the f32 pose component games keep on their entities
struct Pose {
    float x = 0, y = 0, rot_s = 1, rot_b = 0;
    static Pose from_physics(const PhysicsPose& p) { return {float(p.translation.x), float(p.translation.y), float(p.rotation.s), float(p.rotation.b)}; }   // math.rs:124-129
    PhysicsPose to_physics() const { PhysicsPose p; p.translation = {x, y}; p.rotation = {rot_s, rot_b}; return p; }                                   // math.rs:131-138
};
inline PhysicsPose compose(const PhysicsPose& a, const PhysicsPose& b) {   // uv::DIsometry2 * DIsometry2
    PhysicsPose o;
    o.translation = a.rotation * b.translation + a.translation;
    o.rotation = {a.rotation.s * b.rotation.s - a.rotation.b * b.rotation.b, a.rotation.s * b.rotation.b + b.rotation.s * a.rotation.b};
    return o;
}

// what the sync needs from an ECS world (hecs::World): liveness, despawn, an optional Pose per entity, and the entities that
// carry a BodyKey / ColliderKey component (for auto-registration)
using Entity = uint64_t;
class EcsWorld {
public:
    struct Record { bool alive = true; std::optional<Pose> pose; std::optional<BodyKey> body; std::optional<ColliderKey> collider; };
    Entity spawn(std::optional<Pose> pose = std::nullopt, std::optional<BodyKey> body = std::nullopt, std::optional<ColliderKey> coll = std::nullopt) {
        Entity e = next_++;
        records_[e] = Record{true, pose, body, coll};
        return e;
    }
    bool contains(Entity e) const { auto it = records_.find(e); return it != records_.end() && it->second.alive; }
    void despawn(Entity e) { records_.erase(e); }
    Pose* pose(Entity e) { auto it = records_.find(e); return (it != records_.end() && it->second.pose) ? &*it->second.pose : nullptr; }
    template <class F> void for_each(F&& f) { for (auto& kv : records_) f(kv.first, kv.second); }
private:
    std::map<Entity, Record> records_;
    Entity next_ = 1;
};

struct HecsSyncOptions {   // hecs_sync.rs:9-56
    bool hecs_to_physics = false, physics_to_hecs = false, auto_delete_physics = false, auto_delete_hecs = false;
    static HecsSyncOptions both_ways() { return {true, true, true, true}; }
    static HecsSyncOptions hecs_to_physics_only() { return {true, false, true, false}; }
    static HecsSyncOptions physics_to_hecs_only() { return {false, true, false, true}; }
    static HecsSyncOptions do_not_sync() { return {false, false, false, false}; }
};

class HecsSyncManager {   // hecs_sync.rs:62-227
public:
    std::optional<HecsSyncOptions> default_opts;   // :66-69: auto-register entities that carry a BodyKey / ColliderKey
    HecsSyncManager() = default;
    static HecsSyncManager new_autosync(HecsSyncOptions o) { HecsSyncManager m; m.default_opts = o; return m; }
    void clear() { body_entity_map_.clear(); collider_entity_map_.clear(); }
    void register_body(BodyKey b, Entity e, HecsSyncOptions o) { body_entity_map_.insert_at(b.i, {e, o}); }
    void register_collider(ColliderKey c, Entity e, HecsSyncOptions o) { collider_entity_map_.insert_at(c.i, {e, o}); }
    std::optional<Entity> get_body_entity(BodyKey b) const { auto* r = body_entity_map_.get(b.i); return r ? std::optional<Entity>(r->entity) : std::nullopt; }
    std::optional<Entity> get_collider_entity(ColliderKey c) const { auto* r = collider_entity_map_.get(c.i); return r ? std::optional<Entity>(r->entity) : std::nullopt; }
    size_t n_bodies() const { return body_entity_map_.len(); }
    size_t n_colliders() const { return collider_entity_map_.len(); }

    // hecs_sync.rs:121-180 — `World` only needs an `entity_set` member (PhysicsWorld has one)
    template <class World>
    void sync_hecs_to_physics(World& physics, EcsWorld& ecs) {
        if (default_opts) {   // :127-140
            ecs.for_each([&](Entity e, EcsWorld::Record& r) {
                if (r.body && !body_entity_map_.contains(r.body->i)) body_entity_map_.insert_at(r.body->i, {e, *default_opts});
                if (r.collider && !collider_entity_map_.contains(r.collider->i)) collider_entity_map_.insert_at(r.collider->i, {e, *default_opts});
            });
        }
        body_entity_map_.retain([&](Index key, Mapping& m) {
            if (m.opts.auto_delete_physics && !ecs.contains(m.entity)) { physics.entity_set.remove_body(BodyKey{key}); return false; }   // :142-148
            if (m.opts.hecs_to_physics) {
                Pose* pose = ecs.pose(m.entity);
                if (!pose) return true;
                Body* body = physics.entity_set.get_body(BodyKey{key});
                if (!body) return true;
                body->pose = pose->to_physics();
            }
            return true;
        });
        collider_entity_map_.retain([&](Index key, Mapping& m) {   // :161-179
            ColliderKey ck{key};
            if (m.opts.auto_delete_physics && !ecs.contains(m.entity)) { physics.entity_set.remove_collider(ck); return false; }
            if (m.opts.hecs_to_physics) {
                Pose* pose = ecs.pose(m.entity);
                if (!pose) return true;
                auto bk = physics.entity_set.get_collider_body_key(ck);
                Body* body = bk ? physics.entity_set.get_body(*bk) : nullptr;
                if (body) body->pose = pose->to_physics();                       // the collider's entity drives the BODY's pose
                else if (Collider* c = physics.entity_set.get_collider(ck)) c->pose = pose->to_physics();
            }
            return true;
        });
    }

    // hecs_sync.rs:185-227
    template <class World>
    void sync_physics_to_hecs(World& physics, EcsWorld& ecs) {
        body_entity_map_.retain([&](Index key, Mapping& m) {
            if (!physics.entity_set.bodies.contains(key)) {
                if (m.opts.auto_delete_hecs) ecs.despawn(m.entity);
                return false;                                                    // dropped from the map either way (:192-194)
            }
            if (m.opts.physics_to_hecs) {
                if (Pose* pose = ecs.pose(m.entity)) *pose = Pose::from_physics(physics.entity_set.get_body(BodyKey{key})->pose);
            }
            return true;
        });
        collider_entity_map_.retain([&](Index key, Mapping& m) {
            if (!physics.entity_set.colliders.contains(key)) {
                if (m.opts.auto_delete_hecs) ecs.despawn(m.entity);
                return false;
            }
            if (m.opts.physics_to_hecs) {
                Pose* pose = ecs.pose(m.entity);
                if (!pose) return true;
                ColliderKey ck{key};
                Collider* c = physics.entity_set.get_collider(ck);
                auto bk = physics.entity_set.get_collider_body_key(ck);
                Body* body = bk ? physics.entity_set.get_body(*bk) : nullptr;
                *pose = Pose::from_physics(body ? compose(body->pose, c->pose) : c->pose);   // :216-221: the collider's GLOBAL pose
            }
            return true;
        });
    }

private:
    struct Mapping { Entity entity = 0; HecsSyncOptions opts; };
    Arena<Mapping> body_entity_map_, collider_entity_map_;
};

// Game::physics_tick (game.rs:297-303)
template <class World, class FF>
void physics_tick(HecsSyncManager& sync, World& physics, EcsWorld& ecs, double dt_fixed, const FF& ff, std::optional<double> time_scale) {
    sync.sync_hecs_to_physics(physics, ecs);
    physics.tick(dt_fixed, time_scale, ff);
    sync.sync_physics_to_hecs(physics, ecs);
}

}  // namespace starframe

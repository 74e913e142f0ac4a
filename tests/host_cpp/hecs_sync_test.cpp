// HecsSyncManager mirror (moleengine_b200/csrc/host/hecs_sync.hpp) against the rules of /root/reference/src/physics/hecs_sync.rs.
// Host logic only: `Physics` below has nothing but an entity_set, so this runs without a GPU. argv[1] == "tick" also drives
// physics_tick() on a real PhysicsWorld (needs the GPU).
#include <cstdio>
#include <cstring>
#include "../../moleengine_b200/csrc/host/hecs_sync.hpp"

namespace sf = starframe;
static int fails = 0;
#define CHECK(c) do { if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); ++fails; } } while (0)

struct Physics { sf::EntitySet entity_set; };

int main(int argc, char** argv) {
    Physics ph;
    sf::EcsWorld ecs;
    auto mk_body = [&](double x, double y) { return ph.entity_set.insert_body(sf::Body::new_particle(1.0).with_pose(sf::PhysicsPose({x, y}, 0.0))); };

    // ---- manual registration, the four presets (hecs_sync.rs:17-56)
    sf::HecsSyncManager sync;
    sf::BodyKey b_both = mk_body(0, 0), b_in = mk_body(1, 0), b_out = mk_body(2, 0), b_none = mk_body(3, 0);
    sf::Entity e_both = ecs.spawn(sf::Pose{10, 10, 1, 0}), e_in = ecs.spawn(sf::Pose{11, 10, 1, 0}), e_out = ecs.spawn(sf::Pose{12, 10, 1, 0}), e_none = ecs.spawn(sf::Pose{13, 10, 1, 0});
    sync.register_body(b_both, e_both, sf::HecsSyncOptions::both_ways());
    sync.register_body(b_in, e_in, sf::HecsSyncOptions::hecs_to_physics_only());
    sync.register_body(b_out, e_out, sf::HecsSyncOptions::physics_to_hecs_only());
    sync.register_body(b_none, e_none, sf::HecsSyncOptions::do_not_sync());
    CHECK(sync.get_body_entity(b_in) == e_in && !sync.get_body_entity(sf::BodyKey{{99, 1}}));
    sync.sync_hecs_to_physics(ph, ecs);
    CHECK(ph.entity_set.get_body(b_both)->pose.translation.x == 10 && ph.entity_set.get_body(b_in)->pose.translation.x == 11);   // ECS -> physics
    CHECK(ph.entity_set.get_body(b_out)->pose.translation.x == 2 && ph.entity_set.get_body(b_none)->pose.translation.x == 3);    // not these
    for (sf::BodyKey k : {b_both, b_in, b_out, b_none}) ph.entity_set.get_body(k)->pose.translation.y = -7;                      // "the tick"
    sync.sync_physics_to_hecs(ph, ecs);
    CHECK(ecs.pose(e_both)->y == -7 && ecs.pose(e_out)->y == -7);          // physics -> ECS
    CHECK(ecs.pose(e_in)->y == 10 && ecs.pose(e_none)->y == 10);           // one-way in / no sync: untouched

    // ---- auto-delete, both directions (hecs_sync.rs:142-148, 187-194)
    ecs.despawn(e_both);                                                   // entity gone, auto_delete_physics on -> body removed
    ecs.despawn(e_out);                                                    // entity gone, auto_delete_physics OFF -> body stays, mapping stays
    sync.sync_hecs_to_physics(ph, ecs);
    CHECK(!ph.entity_set.bodies.contains(b_both.i) && !sync.get_body_entity(b_both));
    CHECK(ph.entity_set.bodies.contains(b_out.i) && sync.get_body_entity(b_out));
    ph.entity_set.remove_body(b_in);                                       // body gone, auto_delete_hecs OFF -> entity stays, mapping dropped (:192-194)
    ph.entity_set.remove_body(b_none);
    sf::BodyKey b_x = mk_body(5, 5);
    sf::Entity e_x = ecs.spawn(sf::Pose{}, std::nullopt, std::nullopt);
    sync.register_body(b_x, e_x, sf::HecsSyncOptions::both_ways());
    ph.entity_set.remove_body(b_x);                                        // body gone, auto_delete_hecs on -> entity despawned
    sync.sync_physics_to_hecs(ph, ecs);
    CHECK(ecs.contains(e_in) && !sync.get_body_entity(b_in));
    CHECK(ecs.contains(e_none) && !sync.get_body_entity(b_none));
    CHECK(!ecs.contains(e_x) && !sync.get_body_entity(b_x));

    // ---- auto-registration (hecs_sync.rs:127-140) and colliders (a bodied collider's entity drives / follows the body: :161-179, :216-221)
    Physics ph2;
    sf::EcsWorld ecs2;
    sf::HecsSyncManager autosync = sf::HecsSyncManager::new_autosync(sf::HecsSyncOptions::both_ways());
    sf::BodyKey body = ph2.entity_set.insert_body(sf::Body::new_particle(1.0).with_pose(sf::PhysicsPose({1, 2}, 0.0)));
    sf::ColliderKey attached = ph2.entity_set.attach_collider(body, sf::Collider::new_circle(0.5).with_pose(sf::PhysicsPose({0.5, 0}, 0.0)));
    sf::ColliderKey fixed = ph2.entity_set.insert_collider(sf::Collider::new_circle(1.0).with_pose(sf::PhysicsPose({-4, 0}, 0.0)));
    sf::Entity e_body = ecs2.spawn(sf::Pose{1, 2, 1, 0}, body);
    sf::Entity e_fixed = ecs2.spawn(sf::Pose{-4, 1, 1, 0}, std::nullopt, fixed);
    sf::Entity e_nopose = ecs2.spawn(std::nullopt, ph2.entity_set.insert_body(sf::Body::new_kinematic()));
    autosync.sync_hecs_to_physics(ph2, ecs2);
    CHECK(autosync.get_body_entity(body) == e_body && autosync.get_collider_entity(fixed) == e_fixed && autosync.n_bodies() == 2);
    CHECK(ph2.entity_set.get_collider(fixed)->pose.translation.y == 1);                        // static collider pose driven by its entity
    (void)e_nopose;                                                                             // registered, but without a Pose nothing is synced
    sf::Entity e_att = ecs2.spawn(sf::Pose{7, 7, 1, 0}, std::nullopt, attached);
    autosync.sync_hecs_to_physics(ph2, ecs2);
    CHECK(ph2.entity_set.get_body(body)->pose.translation.x == 7);                              // the attached collider's entity moved the BODY
    ph2.entity_set.get_body(body)->pose = sf::PhysicsPose({3, 4}, 3.14159265358979323846 / 2);   // "the tick": quarter turn
    autosync.sync_physics_to_hecs(ph2, ecs2);
    CHECK(std::fabs(ecs2.pose(e_body)->x - 3) < 1e-6 && std::fabs(ecs2.pose(e_body)->y - 4) < 1e-6);
    CHECK(std::fabs(ecs2.pose(e_att)->x - 3.0) < 1e-6 && std::fabs(ecs2.pose(e_att)->y - 4.5) < 1e-6);   // body pose * collider pose: (0.5, 0) turned to (0, 0.5)
    autosync.clear();
    CHECK(autosync.n_bodies() == 0 && autosync.n_colliders() == 0);

    // ---- Game::physics_tick on a real PhysicsWorld (GPU)
    if (argc > 1 && !strcmp(argv[1], "tick")) {
        sf::TuningConstants consts; consts.substeps = 4;
        sf::PhysicsWorld world(consts);
        sf::EcsWorld game;
        sf::HecsSyncManager gs = sf::HecsSyncManager::new_autosync(sf::HecsSyncOptions::both_ways());
        world.entity_set.insert_collider(sf::Collider::new_rounded_rect(20, 0.2, 0.0).with_pose(sf::PhysicsPose({0, -1}, 0.0)));
        sf::Collider box = sf::Collider::new_rounded_rect(1, 1, 0.0);
        sf::BodyKey bk = world.entity_set.insert_body(sf::Body::new_dynamic(box.info(), 0.5).with_pose(sf::PhysicsPose({0, 3}, 0.0)));
        world.entity_set.attach_collider(bk, box);
        sf::Entity e = game.spawn(sf::Pose{0, 3, 1, 0}, bk);
        sf::ForceField gravity = sf::ForceField::gravity({0.0, -9.81});
        for (int t = 0; t < 90; ++t) sf::physics_tick(gs, world, game, 1.0 / 60.0, gravity, 1.0);
        printf("box_y %.4f\n", game.pose(e)->y);
        CHECK(game.pose(e)->y < 0.0 && game.pose(e)->y > -0.6);            // fell and came to rest on the ground, seen through the ECS
        game.pose(e)->y = 5.0f;                                            // the game teleports the entity: the physics body follows
        sf::physics_tick(gs, world, game, 1.0 / 60.0, gravity, 1.0);
        CHECK(world.entity_set.get_body(bk)->pose.translation.y > 4.9);
        game.despawn(e);
        sf::physics_tick(gs, world, game, 1.0 / 60.0, gravity, 1.0);
        CHECK(world.entity_set.bodies.len() == 0);
    }
    printf(fails ? "FAILED %d\n" : "ok\n", fails);
    return fails ? 1 : 0;
}

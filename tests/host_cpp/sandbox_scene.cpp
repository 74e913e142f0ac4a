// Drives the C++ PhysicsWorld mirror (moleengine_b200/csrc/host/physics_world.hpp) the way the reference's sandbox
// drives starframe::PhysicsWorld (examples/sandbox/recipes.rs:131-153 spawn_block, :181-214 spawn_body,
// :406-481 RopeConnectedBlocks; scenes/test.ron) and prints the resulting body poses, one per line.
#include <cstdio>
#include "../../moleengine_b200/csrc/host/physics_world.hpp"

namespace sf = starframe;

static sf::BodyKey spawn_block(sf::PhysicsWorld& w, double x, double y, double width, double height, bool is_static) {
    sf::Collider coll = sf::Collider::new_rounded_rect(width, height, 0.0);
    if (is_static) { w.entity_set.insert_collider(coll.with_pose(sf::PhysicsPose({x, y}, 0.0))); return {}; }
    sf::ColliderKey ck = w.entity_set.insert_collider(coll);
    sf::BodyKey bk = w.entity_set.insert_body(sf::Body::new_dynamic(coll.info(), 0.5).with_pose(sf::PhysicsPose({x, y}, 0.0)));
    w.entity_set.attach_existing_collider(bk, ck);
    return bk;
}

int main(int argc, char** argv) {
    int ticks = argc > 1 ? atoi(argv[1]) : 30;
    sf::TuningConstants consts;
    consts.substeps = 4;
    sf::PhysicsWorld world(consts);
    spawn_block(world, 0, -5, 40, 0.2, true);
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 6; ++i) spawn_block(world, (i - 2.5) * 1.07 + 0.01 * j, -4.38 + 1.02 * j, 1, 1, false);
    // a ball through the compound-collider path (spawn_body)
    std::vector<sf::Collider> parts = {sf::Collider::new_circle(0.5)};
    sf::CompoundColliderSetup setup{parts};
    sf::DVec2 com = setup.center_of_mass();
    sf::BodyKey ball = world.entity_set.insert_body(sf::Body::new_dynamic(setup.info_around_point(com), 0.5).with_pose(sf::PhysicsPose({0.3, 1.5}, 0.0)));
    sf::ColliderKey ball_coll = world.entity_set.attach_collider(ball, parts[0]);
    // RopeConnectedBlocks
    sf::BodyKey b1 = spawn_block(world, -6, 3, 1, 1, false), b2 = spawn_block(world, -3, 3, 1, 1, false);
    sf::Rope rope = sf::Rope::spawn_line(sf::RopeParameters{}, {-6 + 0.55, 3 + 0.55}, {-3 - 0.55, 3 + 0.55}, world.entity_set);
    world.constraint_set.insert(sf::ConstraintBuilder(rope.particles.front().body).with_target(b1).with_target_origin({0.55, 0.55}).build_attachment());
    world.constraint_set.insert(sf::ConstraintBuilder(rope.particles.back().body).with_target(b2).with_target_origin({-0.55, 0.55}).build_attachment());
    sf::RopeKey rk = world.rope_set.insert(rope);
    sf::ForceField gravity = sf::ForceField::gravity({0.0, -9.81});
    for (int t = 0; t < ticks; ++t) {
        if (t == ticks / 2) {   // user code edits the world between ticks: remove a rope particle (splits the rope) and a block
            sf::Rope* r = world.rope_set.get(rk);
            world.entity_set.remove_body(r->particles[r->particles.size() / 2].body);
            world.entity_set.remove_body(b2);
        }
        world.tick(1.0 / 60.0, 1.0, gravity);
    }
    printf("bodies %zu colliders %zu ropes %zu constraints %zu\n", world.entity_set.bodies.len(), world.entity_set.colliders.len(), world.rope_set.ropes.len(),
           world.constraint_set.constraints.len());
    world.entity_set.bodies.for_each([&](sf::Index i, sf::Body& b) { printf("body %u %.6f %.6f %.6f %.6f\n", i.slot, b.pose.translation.x, b.pose.translation.y, b.velocity.linear.x, b.velocity.linear.y); });
    auto hit = world.raycast({{0.3, 8.0}, {0.0, -1.0}}, 30.0);
    if (hit) printf("ray t %.5f collider_slot %u is_ball %d\n", hit->t, hit->collider.i.slot, int(hit->collider == ball_coll));
    printf("ball_contacts %zu\n", world.contacts_for_collider(ball_coll).size());
    printf("query_point_hits %zu\n", world.query_point(world.entity_set.get_body(ball)->pose.translation).size());
    sf::ColliderShape probe; probe.polygon = sf::Polygon::Rect; probe.p0 = 0.6; probe.p1 = 0.6; probe.circle_r = 0.0;
    printf("query_shape_hits %zu\n", world.query_shape(world.entity_set.get_body(ball)->pose, probe, ~0ull).size());
    return 0;
}

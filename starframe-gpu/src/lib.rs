//! `GpuPhysicsWorld`: the reference's `PhysicsWorld` API (src/physics.rs:352-1311) on top of `libstarframe_gpu`.
//!
//! UNCOMPILED in this repository's image (no Rust toolchain). The same wrapper is written — compiled and tested — in C++:
//! `moleengine_b200/csrc/host/physics_world.hpp` (PhysicsWorld) and `hecs_sync.hpp` (HecsSyncManager); this file is the Rust
//! a maintainer of the reference adds, kept in step with the header by `tests/test_rust_binding.py` (through the -sys crate).
//!
//! What stays on the host, untouched: the reference's own `EntitySet`, `ConstraintSet`, `RopeSet` (thunderdome arenas with
//! generations) and the pre-tick housekeeping (physics.rs:399-431). Arrays cross the ABI indexed by `Index::slot()`;
//! generations never leave the host. f64 is rounded to f32 once at upload (the reference already round-trips poses through
//! the f32 `Pose` component every frame, math.rs:124-129).
use starframe::{
    math as m,
    physics::{
        Body, BodyKey, CastHit, Collider, ColliderKey, CollisionLayerMask, CollisionMaskMatrix, ConstraintSet, ContactInfo, EntitySet, Mass,
        Ray, RopeSet, TuningConstants,
    },
};
use starframe_gpu_sys as sys;
use thunderdome as td;

/// forcefield.rs:5-7 is a generic trait; across the ABI a force field is data: a sum of up to four terms.
pub trait GpuForceField {
    fn as_sfg(&self) -> sys::SfgForceField;
}
pub struct Gravity(pub m::DVec2);
pub struct PointGravity { pub position: m::DVec2, pub strength: f64, pub falloff: f64 }
pub struct NoneField;
impl GpuForceField for NoneField {
    fn as_sfg(&self) -> sys::SfgForceField { sys::SfgForceField::default() }
}
impl GpuForceField for Gravity {
    fn as_sfg(&self) -> sys::SfgForceField {
        let mut ff = sys::SfgForceField::default();
        ff.n_terms = 1;
        ff.terms[0] = sys::SfgForceTerm { kind: sys::SFG_FF_GRAVITY, a: self.0.x as f32, b: self.0.y as f32, c: 0.0, d: 0.0 };
        ff
    }
}
impl GpuForceField for PointGravity {
    fn as_sfg(&self) -> sys::SfgForceField {
        let mut ff = sys::SfgForceField::default();
        ff.n_terms = 1;
        ff.terms[0] = sys::SfgForceTerm { kind: sys::SFG_FF_POINT_GRAVITY, a: self.position.x as f32, b: self.position.y as f32,
                                          c: self.strength as f32, d: self.falloff as f32 };
        ff
    }
}

pub struct GpuPhysicsWorld {
    pub consts: TuningConstants,            // physics.rs:353
    pub mask_matrix: CollisionMaskMatrix,   // :354
    pub entity_set: EntitySet,              // :355
    pub rope_set: RopeSet,                  // :356
    pub constraint_set: ConstraintSet,      // :357
    raw: *mut sys::SfgWorld,
    contacts: Vec<ContactInfo>,
    body_records: Vec<sys::SfgBody>,
    collider_records: Vec<sys::SfgCollider>,
}

fn check(raw: *mut sys::SfgWorld, rc: i32) {
    // the reference panics on invariant violations (solver.rs:120, physics.rs:469,726); the ABI never unwinds, so the
    // wrapper turns a status code back into the reference's contract
    if rc != sys::SFG_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(sys::sfg_last_error(raw)) }.to_string_lossy().into_owned();
        panic!("starframe_gpu error {rc}: {msg}");
    }
}

fn tuning_record(c: &TuningConstants) -> sys::SfgTuning {
    sys::SfgTuning {
        substeps: c.substeps as u32,
        sleep_vel_threshold: c.sleep_vel_threshold as f32,
        fall_asleep_frames: if c.fall_asleep_frames == usize::MAX { u32::MAX } else { c.fall_asleep_frames as u32 },
        max_expected_acceleration: c.max_expected_acceleration as f32,
        cell_size: 0.0,
        flags: 0,
        reserved: [0; 2],
    }
}

fn body_record(b: Option<&Body>) -> sys::SfgBody {
    let Some(b) = b else { return sys::SfgBody::default() };   // a free arena slot: flags = 0 (not alive)
    let (mf, im) = match b.mass { Mass::Finite { inverse, .. } => (sys::SFG_BODY_MASS_FINITE, inverse), Mass::Infinite => (0, 0.0) };
    let (inf, ii) = match b.moment_of_inertia { Mass::Finite { inverse, .. } => (sys::SFG_BODY_INERTIA_FINITE, inverse), Mass::Infinite => (0, 0.0) };
    sys::SfgBody {
        x: b.pose.translation.x as f32, y: b.pose.translation.y as f32, rot_s: b.pose.rotation.s as f32, rot_b: b.pose.rotation.bv.xy as f32,
        vx: b.velocity.linear.x as f32, vy: b.velocity.linear.y as f32, w: b.velocity.angular as f32,
        flags: sys::SFG_BODY_ALIVE | mf | inf | if b.ignores_gravity { sys::SFG_BODY_IGNORES_GRAVITY } else { 0 },
        inv_mass: im as f32, inv_inertia: ii as f32, reserved: [0; 2],
    }
}

impl GpuPhysicsWorld {
    /// PhysicsWorld::new (physics.rs:366)
    pub fn new(consts: TuningConstants, mask_matrix: CollisionMaskMatrix, device: i32) -> Self {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { sys::sfg_world_create(&tuning_record(&consts), mask_matrix.as_rows().as_ptr(), device, &mut raw) };
        assert!(rc == sys::SFG_OK, "sfg_world_create failed with {rc} (no CUDA device? there is no CPU fallback)");
        Self { consts, mask_matrix, entity_set: EntitySet::default(), rope_set: RopeSet::default(), constraint_set: ConstraintSet::default(),
               raw, contacts: Vec::new(), body_records: Vec::new(), collider_records: Vec::new() }
    }

    /// PhysicsWorld::clear (physics.rs:386-393)
    pub fn clear(&mut self) {
        self.entity_set.clear();
        self.rope_set.clear();
        self.constraint_set.clear();
        self.contacts.clear();
        check(self.raw, unsafe { sys::sfg_world_clear(self.raw) });
    }

    /// PhysicsWorld::tick (physics.rs:396-1158)
    pub fn tick(&mut self, frame_dt: f64, time_scale: Option<f64>, forcefield: &impl GpuForceField) {
        // 0. the reference's housekeeping stays on the host (physics.rs:399-431)
        self.entity_set.remove_orphan_colliders();                     // entity_set.rs:168-176
        self.rope_set.remove_dead_particles(&self.entity_set);         // rope.rs:226-286
        self.constraint_set.retain_live(&self.entity_set);             // physics.rs:423-428
        // 1. upload by slot
        self.body_records.clear();
        for slot in 0..self.entity_set.bodies.capacity() as u32 {
            self.body_records.push(body_record(self.entity_set.bodies.get_by_slot(slot).map(|(_, b)| b)));
        }
        self.collider_records.clear();
        for slot in 0..self.entity_set.colliders.capacity() as u32 {
            self.collider_records.push(collider_record(&self.entity_set, slot));
        }
        let (constraints, ropes, rope_particles) = constraint_and_rope_records(&self.constraint_set, &self.rope_set);
        unsafe {
            check(self.raw, sys::sfg_set_tuning(self.raw, &tuning_record(&self.consts)));
            check(self.raw, sys::sfg_set_mask_matrix(self.raw, self.mask_matrix.as_rows().as_ptr()));
            check(self.raw, sys::sfg_set_bodies(self.raw, self.body_records.len() as u32, self.body_records.as_ptr()));
            check(self.raw, sys::sfg_set_colliders(self.raw, self.collider_records.len() as u32, self.collider_records.as_ptr()));
            check(self.raw, sys::sfg_set_constraints(self.raw, constraints.len() as u32, constraints.as_ptr()));
            check(self.raw, sys::sfg_set_ropes(self.raw, ropes.len() as u32, ropes.as_ptr(), rope_particles.len() as u32, rope_particles.as_ptr()));
            // 2. the tick
            let ff = forcefield.as_sfg();
            let (has, scale) = time_scale.map_or((0, 0.0), |s| (1, s));
            check(self.raw, sys::sfg_tick(self.raw, frame_dt, has, scale, &ff));
            // 3. write-back (physics.rs:1150-1157) and contacts (physics.rs:1097-1122)
            check(self.raw, sys::sfg_get_bodies(self.raw, 0, self.body_records.len() as u32, self.body_records.as_mut_ptr()));
        }
        for (key, body) in self.entity_set.bodies.iter_mut() {
            let r = &self.body_records[key.slot() as usize];
            body.pose.translation = m::DVec2::new(r.x as f64, r.y as f64);
            body.pose.rotation.s = r.rot_s as f64;
            body.pose.rotation.bv.xy = r.rot_b as f64;
            body.velocity.linear = m::DVec2::new(r.vx as f64, r.vy as f64);
            body.velocity.angular = r.w as f64;
        }
        self.fetch_contacts();
    }

    fn fetch_contacts(&mut self) {
        let mut n = 0usize;
        unsafe { check(self.raw, sys::sfg_get_contacts(self.raw, std::ptr::null_mut(), 0, &mut n)) };
        let mut raw = vec![sys::SfgContactInfo::default(); n];
        unsafe { check(self.raw, sys::sfg_get_contacts(self.raw, raw.as_mut_ptr(), n, &mut n)) };
        self.contacts.clear();
        for c in &raw[..n.min(raw.len())] {
            let (Some(k0), Some(k1)) = (self.entity_set.colliders.contains_slot(c.collider0), self.entity_set.colliders.contains_slot(c.collider1)) else { continue };
            self.contacts.push(ContactInfo { colliders: [ColliderKey(k0), ColliderKey(k1)], normal: m::DVec2::new(c.nx as f64, c.ny as f64), ..Default::default() });
        }
    }

    /// PhysicsWorld::contacts_for_collider (physics.rs:1165-1178): oriented so that the queried collider is first, normal away from it
    pub fn contacts_for_collider(&self, coll: ColliderKey) -> impl Iterator<Item = ContactInfo> + '_ {
        self.contacts.iter().filter_map(move |c| {
            if c.colliders[0] == coll { Some(*c) }
            else if c.colliders[1] == coll { Some(ContactInfo { colliders: [c.colliders[1], c.colliders[0]], normal: -c.normal, ..*c }) }
            else { None }
        })
    }

    /// PhysicsWorld::raycast (physics.rs:1255)
    pub fn raycast(&mut self, ray: Ray, max_distance: f64) -> Option<CastHit> { self.spherecast(0.0, ray, max_distance) }

    /// PhysicsWorld::spherecast (physics.rs:1266-1311)
    pub fn spherecast(&mut self, radius: f64, ray: Ray, max_distance: f64) -> Option<CastHit> {
        let r = sys::SfgRay { sx: ray.start.x as f32, sy: ray.start.y as f32, dx: ray.dir.x as f32, dy: ray.dir.y as f32, radius: radius as f32,
                              max_distance: max_distance as f32, domain: 0, reserved: 0 };
        let mut hit = sys::SfgCastHit::default();
        unsafe { check(self.raw, sys::sfg_cast_batch(self.raw, 1, &r, &mut hit)) };
        if hit.collider < 0 { return None; }
        let key = self.entity_set.colliders.contains_slot(hit.collider as u32)?;
        Some(CastHit { collider: ColliderKey(key), t: hit.t as f64, point: m::DVec2::new(hit.px as f64, hit.py as f64), normal: m::DVec2::new(hit.nx as f64, hit.ny as f64) })
    }

    /// PhysicsWorld::query_point (physics.rs:1188-1210)
    pub fn query_point(&mut self, point: m::DVec2) -> Vec<(ColliderKey, Option<BodyKey>)> {
        let (xy, mut out, mut count) = ([point.x as f32, point.y as f32], [-1i32; 64], 0u32);
        unsafe { check(self.raw, sys::sfg_query_point_batch(self.raw, 1, xy.as_ptr(), std::ptr::null(), 64, out.as_mut_ptr(), &mut count)) };
        self.keys_of(&out[..(count as usize).min(64)])
    }

    /// PhysicsWorld::query_shape (physics.rs:1215-1245)
    pub fn query_shape(&mut self, pose: m::PhysicsPose, shape: starframe::physics::ColliderShape, mask: CollisionLayerMask) -> Vec<(ColliderKey, Option<BodyKey>)> {
        let (kind, p0, p1) = polygon_record(&shape.polygon);
        let q = sys::SfgShapeQuery { x: pose.translation.x as f32, y: pose.translation.y as f32, rot_s: pose.rotation.s as f32, rot_b: pose.rotation.bv.xy as f32,
                                     p0, p1, circle_r: shape.circle_r as f32, kind, layer_mask: mask.0, domain: 0, reserved: 0 };
        let (mut out, mut count) = ([-1i32; 64], 0u32);
        unsafe { check(self.raw, sys::sfg_query_shape_batch(self.raw, 1, &q, 64, out.as_mut_ptr(), &mut count)) };
        self.keys_of(&out[..(count as usize).min(64)])
    }

    fn keys_of(&self, slots: &[i32]) -> Vec<(ColliderKey, Option<BodyKey>)> {
        slots.iter().filter_map(|&s| {
            let key = ColliderKey(self.entity_set.colliders.contains_slot(s as u32)?);
            Some((key, self.entity_set.get_collider_body_key(key)))
        }).collect()
    }

    pub(crate) fn raw(&self) -> *mut sys::SfgWorld { self.raw }
}

impl Drop for GpuPhysicsWorld {
    fn drop(&mut self) { unsafe { sys::sfg_world_destroy(self.raw) } }
}

fn polygon_record(p: &starframe::physics::ColliderPolygon) -> (u32, f32, f32) {
    use starframe::physics::ColliderPolygon as P;
    match *p {
        P::Point => (sys::SFG_POLY_POINT, 0.0, 0.0),
        P::LineSegment { hl } => (sys::SFG_POLY_SEGMENT, hl as f32, 0.0),
        P::Rect { hw, hh } => (sys::SFG_POLY_RECT, hw as f32, hh as f32),
        P::Triangle { outer_r } => (sys::SFG_POLY_TRIANGLE, outer_r as f32, 0.0),
        P::Hexagon { outer_r } => (sys::SFG_POLY_HEXAGON, outer_r as f32, 0.0),
    }
}

fn collider_record(set: &EntitySet, slot: u32) -> sys::SfgCollider {
    let Some((key, c)) = set.colliders.get_by_slot(slot) else { return sys::SfgCollider { body: -1, ..Default::default() } };
    let c: &Collider = c;
    let (kind, p0, p1) = polygon_record(&c.shape.polygon);
    let mat = c.material;
    sys::SfgCollider {
        kind, layer: c.layer as u32,
        flags: sys::SFG_COLL_ALIVE | if c.is_sensor() { sys::SFG_COLL_SENSOR } else { 0 }
            | if mat.static_friction_coef.is_some() { sys::SFG_COLL_HAS_STATIC_FRICTION } else { 0 }
            | if mat.dynamic_friction_coef.is_some() { sys::SFG_COLL_HAS_DYNAMIC_FRICTION } else { 0 },
        body: set.get_collider_body_key(ColliderKey(key)).map_or(-1, |b| b.0.slot() as i32),
        p0, p1, circle_r: c.shape.circle_r as f32, domain: 0,
        x: c.pose.translation.x as f32, y: c.pose.translation.y as f32, rot_s: c.pose.rotation.s as f32, rot_b: c.pose.rotation.bv.xy as f32,
        static_friction: mat.static_friction_coef.unwrap_or(0.0) as f32, dynamic_friction: mat.dynamic_friction_coef.unwrap_or(0.0) as f32,
        restitution: mat.restitution_coef as f32, reserved: 0,
    }
}

fn constraint_and_rope_records(cs: &ConstraintSet, rs: &RopeSet) -> (Vec<sys::SfgConstraint>, Vec<sys::SfgRope>, Vec<i32>) {
    use starframe::physics::ConstraintLimit as L;
    let constraints = cs.iter().map(|c| sys::SfgConstraint {
        owner: c.owner.0.slot() as i32, target: c.target.map_or(-1, |t| t.0.slot() as i32),
        flags: match c.limit { L::Eq => sys::SFG_LIMIT_EQ, L::Lt => sys::SFG_LIMIT_LT, L::Gt => sys::SFG_LIMIT_GT } | if c.can_sleep { sys::SFG_CONSTRAINT_CAN_SLEEP } else { 0 },
        reserved: 0, compliance: c.compliance as f32, linear_damping: c.linear_damping as f32, angular_damping: c.angular_damping as f32,
        distance: c.distance() as f32, off0x: c.offsets[0].x as f32, off0y: c.offsets[0].y as f32, off1x: c.offsets[1].x as f32, off1y: c.offsets[1].y as f32,
    }).collect();
    let (mut ropes, mut particles) = (Vec::new(), Vec::new());
    for rope in rs.iter() {
        let p = &rope.params;
        ropes.push(sys::SfgRope { spacing: p.spacing as f32, compliance: p.compliance as f32, bending_max_angle: p.bending_max_angle.rad() as f32,
                                  bending_compliance: p.bending_compliance as f32, damping: p.damping as f32, first_particle: particles.len() as u32,
                                  particle_count: rope.particles.len() as u32, reserved: 0 });
        particles.extend(rope.particles.iter().map(|q| q.body.0.slot() as i32));
    }
    (constraints, ropes, particles)
}

// ---------------------------------------------------------------------------------------------------------------------------
/// hecs_sync.rs:9-56
#[derive(Clone, Copy, Debug)]
pub struct HecsSyncOptions { pub hecs_to_physics: bool, pub physics_to_hecs: bool, pub auto_delete_physics: bool, pub auto_delete_hecs: bool }
impl HecsSyncOptions {
    pub fn both_ways() -> Self { Self { hecs_to_physics: true, physics_to_hecs: true, auto_delete_physics: true, auto_delete_hecs: true } }
    pub fn hecs_to_physics_only() -> Self { Self { hecs_to_physics: true, physics_to_hecs: false, auto_delete_physics: true, auto_delete_hecs: false } }
    pub fn physics_to_hecs_only() -> Self { Self { hecs_to_physics: false, physics_to_hecs: true, auto_delete_physics: false, auto_delete_hecs: true } }
    pub fn do_not_sync() -> Self { Self { hecs_to_physics: false, physics_to_hecs: false, auto_delete_physics: false, auto_delete_hecs: false } }
}

/// HecsSyncManager (hecs_sync.rs:62-227) for a GpuPhysicsWorld. Same registration, auto-register, auto-delete and direction
/// rules as the reference. The pose traffic is what differs: instead of copying every pose through the host-side Body structs
/// and re-uploading whole bodies, `sync_hecs_to_physics` gathers the f32 `Pose` of the entities whose options say
/// hecs_to_physics into one (slot, pose) list for `sfg_set_poses_indexed`, and `sync_physics_to_hecs` reads back exactly the
/// slots whose options say physics_to_hecs with `sfg_get_poses_indexed` — 16 bytes per synced body each way.
#[derive(Default)]
pub struct GpuHecsSync {
    pub default_opts: Option<HecsSyncOptions>,
    body_entity_map: td::Arena<(hecs::Entity, HecsSyncOptions)>,
    collider_entity_map: td::Arena<(hecs::Entity, HecsSyncOptions)>,
    slots: Vec<u32>,
    poses: Vec<[f32; 4]>,
}

impl GpuHecsSync {
    pub fn new() -> Self { Self::default() }
    pub fn new_autosync(opts: HecsSyncOptions) -> Self { Self { default_opts: Some(opts), ..Self::default() } }
    pub fn clear(&mut self) { self.body_entity_map.clear(); self.collider_entity_map.clear(); }
    pub fn register_body(&mut self, body: BodyKey, entity: hecs::Entity, opts: HecsSyncOptions) { self.body_entity_map.insert_at(body.0, (entity, opts)); }
    pub fn register_collider(&mut self, coll: ColliderKey, entity: hecs::Entity, opts: HecsSyncOptions) { self.collider_entity_map.insert_at(coll.0, (entity, opts)); }
    pub fn get_body_entity(&self, body: BodyKey) -> Option<hecs::Entity> { self.body_entity_map.get(body.0).map(|(e, _)| *e) }
    pub fn get_collider_entity(&self, coll: ColliderKey) -> Option<hecs::Entity> { self.collider_entity_map.get(coll.0).map(|(e, _)| *e) }

    /// hecs_sync.rs:121-180. Call before `GpuPhysicsWorld::tick`.
    pub fn sync_hecs_to_physics(&mut self, physics: &mut GpuPhysicsWorld, hecs_world: &mut hecs::World) {
        if let Some(opts) = self.default_opts {                       // auto-register new entities (:127-140)
            for (entity, body_key) in hecs_world.query_mut::<&BodyKey>() {
                if !self.body_entity_map.contains(body_key.0) { self.body_entity_map.insert_at(body_key.0, (entity, opts)); }
            }
            for (entity, coll_key) in hecs_world.query_mut::<&ColliderKey>() {
                if !self.collider_entity_map.contains(coll_key.0) { self.collider_entity_map.insert_at(coll_key.0, (entity, opts)); }
            }
        }
        let (slots, poses) = (&mut self.slots, &mut self.poses);
        slots.clear();
        poses.clear();
        self.body_entity_map.retain(|body_key, (entity, opts)| {
            if opts.auto_delete_physics && !hecs_world.contains(*entity) {   // :142-148
                physics.entity_set.remove_body(BodyKey(body_key));
                return false;
            }
            if opts.hecs_to_physics {
                let Ok(pose) = hecs_world.query_one_mut::<&m::Pose>(*entity) else { return true };
                let Some(body) = physics.entity_set.get_body_mut(BodyKey(body_key)) else { return true };
                body.pose = m::PhysicsPose::from(*pose);
                let r = body.pose.rotation;
                slots.push(body_key.slot());
                poses.push([pose.translation.x, pose.translation.y, r.s as f32, r.bv.xy as f32]);
            }
            true
        });
        self.collider_entity_map.retain(|coll_key, (entity, opts)| {       // :161-179: a bodied collider's entity drives the BODY's pose
            let coll_key = ColliderKey(coll_key);
            if opts.auto_delete_physics && !hecs_world.contains(*entity) {
                physics.entity_set.remove_collider(coll_key);
                return false;
            }
            if opts.hecs_to_physics {
                let Ok(pose) = hecs_world.query_one_mut::<&m::Pose>(*entity) else { return true };
                if let Some(body_key) = physics.entity_set.get_collider_body_key(coll_key) {
                    let body = physics.entity_set.get_body_mut(body_key).unwrap();
                    body.pose = m::PhysicsPose::from(*pose);
                    let r = body.pose.rotation;
                    slots.push(body_key.0.slot());
                    poses.push([pose.translation.x, pose.translation.y, r.s as f32, r.bv.xy as f32]);
                } else if let Some(coll) = physics.entity_set.get_collider_mut(coll_key) {
                    coll.pose = m::PhysicsPose::from(*pose);         // static collider: re-uploaded with the colliders
                }
            }
            true
        });
        // the device copy: only the poses that were driven from the ECS this frame
        if !slots.is_empty() {
            check(physics.raw(), unsafe { sys::sfg_set_poses_indexed(physics.raw(), slots.len() as u32, slots.as_ptr(), poses.as_ptr() as *const f32) });
        }
    }

    /// hecs_sync.rs:185-227. Call after `GpuPhysicsWorld::tick`.
    pub fn sync_physics_to_hecs(&mut self, physics: &GpuPhysicsWorld, hecs_world: &mut hecs::World) {
        self.body_entity_map.retain(|body_key, (entity, opts)| {
            if !physics.entity_set.bodies.contains(body_key) {
                if opts.auto_delete_hecs { hecs_world.despawn(*entity).ok(); }
                return false;                                        // removed from the map even without auto-delete (:192-194)
            }
            if opts.physics_to_hecs {
                let body = physics.entity_set.get_body(BodyKey(body_key)).unwrap();
                let Ok(pose) = hecs_world.query_one_mut::<&mut m::Pose>(*entity) else { return true };
                pose.sync_from_physics(body.pose);
            }
            true
        });
        self.collider_entity_map.retain(|coll_key, (entity, opts)| {
            if !physics.entity_set.colliders.contains(coll_key) {
                if opts.auto_delete_hecs { hecs_world.despawn(*entity).ok(); }
                return false;
            }
            if opts.physics_to_hecs {
                let coll_key = ColliderKey(coll_key);
                let Ok(pose) = hecs_world.query_one_mut::<&mut m::Pose>(*entity) else { return true };
                let coll = physics.entity_set.get_collider(coll_key).unwrap();
                if let Some(body) = physics.entity_set.get_collider_body(coll_key) { pose.sync_from_physics(body.pose * coll.pose); }
                else { pose.sync_from_physics(coll.pose); }
            }
            true
        });
    }
}

/// Game::physics_tick (game.rs:297-303)
pub fn physics_tick(sync: &mut GpuHecsSync, physics: &mut GpuPhysicsWorld, world: &mut hecs::World, dt_fixed: f64, ff: &impl GpuForceField, time_scale: Option<f64>) {
    sync.sync_hecs_to_physics(physics, world);
    physics.tick(dt_fixed, time_scale, ff);
    sync.sync_physics_to_hecs(physics, world);
}

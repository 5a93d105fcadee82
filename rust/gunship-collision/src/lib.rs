//! Drop-in for Gunship's `GridCollisionSystem` (src/component/collider/grid_collision.rs:110-285) that
//! runs the collision step on a B200 through the C ABI of include/gunship_collision.h.
//!
//! What stays exactly as in the engine: `Collider`, `ColliderManager`, `BoundingVolumeManager`'s dense
//! `entities` order, `CollisionSystem::update(&mut self, &Scene, f32)`, and the public field
//! `collisions: HashSet<(Entity, Entity), FnvHashState>` that `CollisionCallbackManager::process_collisions`
//! consumes (mod.rs:602, 678-708).  What changes: `bvh_update` + `GridCollisionSystem::update` become one
//! `gsc_step`.  Errors: the engine panics on the same conditions (Cargo.toml `panic = "abort"`), so a
//! non-zero status becomes `panic!`.
//!
//! NOT compiled in the authoring image (no Rust toolchain there) -- see INTEGRATION.md.
mod ffi;

use std::collections::HashSet;
use std::ffi::CStr;
use std::hash::BuildHasherDefault;
use std::ptr;

/// `Entity` is a `u32` id (src/old/ecs.rs:7-8).
pub type Entity = u32;

/// The engine passes its own `FnvHashState` (lib/hash); any `BuildHasher` works for the shim.
pub type PairSet<S> = HashSet<(Entity, Entity), S>;

/// Mirror of `enum Collider` (mod.rs:139-183) flattened for the FFI.
#[derive(Debug, Clone, Copy, PartialEq)]
pub enum Collider {
    Sphere { offset: [f32; 3], radius: f32 },
    Box { offset: [f32; 3], widths: [f32; 3] },
    Mesh,
}

/// Derived transform of one entity (TransformManager position/rotation/scale_derived,
/// component/transform.rs:441-459); rotation is the quaternion (x, y, z, w).
#[derive(Debug, Clone, Copy)]
pub struct DerivedTransform {
    pub position: [f32; 3],
    pub rotation: [f32; 4],
    pub scale: [f32; 3],
}

pub struct GridCollisionSystem<S: std::hash::BuildHasher + Default = BuildHasherDefault<std::collections::hash_map::DefaultHasher>> {
    ctx: *mut ffi::gsc_ctx,
    n: u32,
    /// `box_slot[i]` = number of boxes before collider i: where its rotation / scale go in the packed staging arrays
    box_slot: Vec<u32>,
    is_box: Vec<bool>,
    n_box: u32,
    /// pinned staging bank the next frame is marshalled into (two banks alternate, include/gunship_collision.h)
    bank: u32,
    pairs: Vec<u32>,
    /// Same name, type shape and tuple order as grid_collision.rs:119: (later volume, earlier volume).
    pub collisions: PairSet<S>,
    pub cell_size: f32,
}

fn check(ctx: *const ffi::gsc_ctx, rc: i32) {
    if rc != ffi::GSC_OK {
        let msg = unsafe { CStr::from_ptr(ffi::gsc_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("gunship_collision: status {}: {}", rc, msg); // the reference panics here too
    }
}

impl<S: std::hash::BuildHasher + Default> GridCollisionSystem<S> {
    /// `GridCollisionSystem::new()` (grid_collision.rs:122-216); `max_colliders` replaces the engine's
    /// MAX_COMPONENTS cap (struct_component_manager.rs:9).
    pub fn new(device: i32, max_colliders: u32) -> Self {
        let cfg = ffi::gsc_config { device, max_colliders, max_pairs: 0, max_candidates: 0, max_cells: 0, flags: 0, reserved: 0 };
        let mut ctx = ptr::null_mut();
        let rc = unsafe { ffi::gsc_create(&cfg, &mut ctx) };
        check(ptr::null(), rc);
        GridCollisionSystem { ctx, n: 0, box_slot: vec![], is_box: vec![], n_box: 0, bank: 0, pairs: vec![],
                              collisions: HashSet::with_hasher(S::default()), cell_size: 0.0 }
    }

    /// Called when the collider set changes (`ColliderManager::assign` / `destroy`, mod.rs:221-223, 282-285):
    /// `entities` and `colliders` in `BoundingVolumeManager.components` order -- that order orients every tuple.
    pub fn set_colliders(&mut self, entities: &[Entity], colliders: &[Collider]) {
        assert_eq!(entities.len(), colliders.len());
        let n = entities.len();
        let (mut kind, mut off, mut size) = (Vec::with_capacity(n), Vec::with_capacity(3 * n), Vec::with_capacity(3 * n));
        for c in colliders {
            match *c {
                Collider::Sphere { offset, radius } => { kind.push(ffi::GSC_SPHERE); off.extend_from_slice(&offset); size.extend_from_slice(&[radius, 0.0, 0.0]); }
                Collider::Box { offset, widths } => { kind.push(ffi::GSC_BOX); off.extend_from_slice(&offset); size.extend_from_slice(&widths); }
                Collider::Mesh => { kind.push(ffi::GSC_MESH); off.extend_from_slice(&[0.0; 3]); size.extend_from_slice(&[0.0; 3]); } // -> GSC_ERR_KIND, as unimplemented!()
            }
        }
        let rc = unsafe { ffi::gsc_upload_colliders(self.ctx, n as u32, entities.as_ptr(), kind.as_ptr(), off.as_ptr(), size.as_ptr(), ptr::null()) };
        check(self.ctx, rc);
        self.n = n as u32;
        self.refresh_box_slots();
    }

    /// after the collider set changed: where every box finds its rotation / scale in the packed staging arrays
    fn refresh_box_slots(&mut self) {
        self.box_slot.resize(self.n as usize, 0);
        let mut n_box = 0u32;
        let rc = unsafe { ffi::gsc_box_slots(self.ctx, self.box_slot.as_mut_ptr(), self.n, &mut n_box) };
        check(self.ctx, rc);
        self.n_box = n_box;
        self.is_box.clear();
        for i in 0..self.n as usize {
            let next = if i + 1 < self.n as usize { self.box_slot[i + 1] } else { n_box };
            self.is_box.push(next > self.box_slot[i]);
        }
    }

    /// The one pass over the frame's transforms `bvh_update` makes (bounding_volume.rs:358-362), writing straight
    /// into the library's PINNED staging bank -- positions in dense order, rotation / scale packed over the boxes
    /// (spheres ignore them, mod.rs:313-317) -- then the bank is handed over (asynchronous copy).
    fn marshal(&mut self, transforms: &[DerivedTransform]) {
        assert_eq!(transforms.len(), self.n as usize);
        let (mut pos, mut quat, mut scale) = (ptr::null_mut::<f32>(), ptr::null_mut::<f32>(), ptr::null_mut::<f32>());
        unsafe {
            check(self.ctx, ffi::gsc_map_transform_staging(self.ctx, self.bank, &mut pos, &mut quat, &mut scale));
            for (i, t) in transforms.iter().enumerate() {
                ptr::copy_nonoverlapping(t.position.as_ptr(), pos.add(3 * i), 3);
                if self.is_box[i] {
                    let k = self.box_slot[i] as usize;
                    ptr::copy_nonoverlapping(t.rotation.as_ptr(), quat.add(4 * k), 4);
                    ptr::copy_nonoverlapping(t.scale.as_ptr(), scale.add(3 * k), 3);
                }
            }
            check(self.ctx, ffi::gsc_commit_transform_staging(self.ctx, self.bank, self.n, self.n_box, 1, 1));
        }
        self.bank ^= 1;
    }

    /// `bvh_update` + `GridCollisionSystem::update(&mut self, &BoundingVolumeManager)` (grid_collision.rs:218):
    /// fills `self.collisions`.  `transforms[i]` belongs to `entities[i]` of `set_colliders`.
    pub fn update(&mut self, transforms: &[DerivedTransform]) {
        self.marshal(transforms);
        let mut out: ffi::gsc_step_out = unsafe { std::mem::zeroed() };
        unsafe {
            check(self.ctx, ffi::gsc_step(self.ctx, &mut out));
        }
        self.cell_size = out.cell_size;
        self.pairs.resize(2 * out.n_pairs as usize, 0);
        let mut n = 0u64;
        let rc = unsafe { ffi::gsc_download_pairs(self.ctx, self.pairs.as_mut_ptr(), out.n_pairs, 0, &mut n) };
        check(self.ctx, rc);
        self.collisions.clear(); // grid_collision.rs:221
        for p in self.pairs.chunks(2) {
            self.collisions.insert((p[0], p[1]));
        }
    }
}

impl<S: std::hash::BuildHasher + Default> GridCollisionSystem<S> {
    /// `ColliderManager::assign` after the first `set_colliders` (struct_component_manager.rs:83-96): the new
    /// colliders go to the end of the dense arrays; the ones already on the device are not touched.
    pub fn append_colliders(&mut self, entities: &[Entity], kind: &[u8], offset3: &[f32], size3: &[f32]) {
        let rc = unsafe { ffi::gsc_append_colliders(self.ctx, entities.len() as u32, entities.as_ptr(), kind.as_ptr(), offset3.as_ptr(), size3.as_ptr()) };
        check(self.ctx, rc);
        self.n += entities.len() as u32;
        self.refresh_box_slots();
    }

    /// The cleanup loop of `CollisionSystem::update` (mod.rs:605-609): one `destroy_immediate` (swap_remove,
    /// bounding_volume.rs:82-104) per index, in order; `indices[k]` is what `self.indices.remove(&entity)` yields
    /// at that moment.
    pub fn remove_colliders(&mut self, indices: &[u32]) {
        let rc = unsafe { ffi::gsc_remove_colliders(self.ctx, indices.len() as u32, indices.as_ptr()) };
        check(self.ctx, rc);
        self.n -= indices.len() as u32;
        self.refresh_box_slots();
    }

    /// Pipelined form of `update` for a main loop that can hand over the NEXT frame's transforms while this
    /// frame's step runs: `begin_update(frame f)`, ..., `begin_update(frame f+1)` is legal before `end_update()`
    /// only through `stage_transforms`, which the C side orders behind `bvh_update` of the step in flight.
    pub fn stage_transforms(&mut self, transforms: &[DerivedTransform]) {
        self.marshal(transforms);
    }
    pub fn begin_update(&mut self) {
        unsafe { check(self.ctx, ffi::gsc_step_begin(self.ctx)); }
    }
    pub fn end_update(&mut self) {
        let mut out: ffi::gsc_step_out = unsafe { std::mem::zeroed() };
        unsafe { check(self.ctx, ffi::gsc_step_end(self.ctx, &mut out)); }
        self.cell_size = out.cell_size;
        self.pairs.resize(2 * out.n_pairs as usize, 0);
        let mut n = 0u64;
        let rc = unsafe { ffi::gsc_download_pairs(self.ctx, self.pairs.as_mut_ptr(), out.n_pairs, 0, &mut n) };
        check(self.ctx, rc);
        self.collisions.clear();
        for p in self.pairs.chunks(2) {
            self.collisions.insert((p[0], p[1]));
        }
    }
}

impl<S: std::hash::BuildHasher + Default> Drop for GridCollisionSystem<S> {
    fn drop(&mut self) {
        unsafe { ffi::gsc_destroy(self.ctx) }
    }
}

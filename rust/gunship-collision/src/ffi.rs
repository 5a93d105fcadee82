//! `extern "C"` declarations -- one to one with include/gunship_collision.h (GSC_ABI_VERSION 2).
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

pub const GSC_OK: c_int = 0;
pub const GSC_ERR_CAPACITY: c_int = 4;
pub const GSC_SPHERE: u8 = 0;
pub const GSC_BOX: u8 = 1;
pub const GSC_MESH: u8 = 2;
pub const GSC_ERR_NCCL: c_int = 3;
pub const GSC_FLAG_TIMING: u32 = 1;
pub const GSC_FLAG_GRAPH: u32 = 2;
pub const GSC_FLAG_CONTACTS: u32 = 4;
pub const GSC_MAX_PEERS: usize = 16;
pub const GSC_IPC_EXPORT_BYTES: usize = 256;
pub const GSC_NUM_STAGES: usize = 6;
pub const GSC_HALO_RECORD_BYTES: usize = 96;

pub enum gsc_ctx {}
pub enum gsc_group {}

#[repr(C)]
pub struct gsc_config {
    pub device: i32,
    pub max_colliders: u32,
    pub max_pairs: u64,
    pub max_candidates: u64,
    pub max_cells: u64,
    pub flags: u32,
    pub reserved: u32,
}

#[repr(C)]
pub struct gsc_step_out {
    pub n_pairs: u64,
    pub n_sphere_sphere: u64,
    pub n_sphere_obb: u64,
    pub n_obb_obb: u64,
    pub n_cells: u64,
    pub cell_size: f32,
    pub n_wide: u32,
    pub grid_origin: [i32; 3],
    pub grid_dims: [i32; 3],
    pub sort_passes: u32,
    pub status_flags: u32,
    pub stage_ms: [f32; GSC_NUM_STAGES],
    pub step_ms: f32,
    pub d_pairs: *const u32,
    pub d_contacts: *const f32,
    pub n_ghost: u32,
    pub reserved: u32,
}

#[repr(C)]
pub struct gsc_adjacency {
    pub n_entities: u64,
    pub n_entries: u64,
    pub d_entity: *const u32,
    pub d_offset: *const u32,
    pub d_other: *const u32,
}

#[repr(C)]
pub struct gsc_bounds {
    pub longest_axis: f32,
    pub world_min: [f32; 3],
    pub world_max: [f32; 3],
    pub status_flags: u32,
    pub home_x_max: f32,
}

extern "C" {
    pub fn gsc_create(cfg: *const gsc_config, out: *mut *mut gsc_ctx) -> c_int;
    pub fn gsc_destroy(ctx: *mut gsc_ctx);
    pub fn gsc_last_error(ctx: *const gsc_ctx) -> *const c_char;
    pub fn gsc_abi_version() -> c_int;
    pub fn gsc_set_flags(ctx: *mut gsc_ctx, flags: u32) -> c_int;
    pub fn gsc_upload_colliders(ctx: *mut gsc_ctx, n: u32, entity: *const u32, kind: *const u8, offset3: *const f32,
                                size3: *const f32, global_index: *const u32) -> c_int;
    pub fn gsc_update_transforms(ctx: *mut gsc_ctx, n: u32, pos3: *const f32, quat4: *const f32, scale3: *const f32) -> c_int;
    pub fn gsc_update_transforms_packed(ctx: *mut gsc_ctx, n: u32, pos3: *const f32, n_box: u32, box_quat4: *const f32,
                                        box_scale3: *const f32) -> c_int;
    pub fn gsc_bind_device_transforms(ctx: *mut gsc_ctx, n: u32, d_pos3: *const f32, d_quat4: *const f32,
                                      d_scale3: *const f32) -> c_int;
    pub fn gsc_step(ctx: *mut gsc_ctx, out: *mut gsc_step_out) -> c_int;
    pub fn gsc_step_begin(ctx: *mut gsc_ctx) -> c_int;
    pub fn gsc_step_end(ctx: *mut gsc_ctx, out: *mut gsc_step_out) -> c_int;
    pub fn gsc_append_colliders(ctx: *mut gsc_ctx, n_new: u32, entity: *const u32, kind: *const u8, offset3: *const f32,
                                size3: *const f32) -> c_int;
    pub fn gsc_remove_colliders(ctx: *mut gsc_ctx, n_remove: u32, index: *const u32) -> c_int;
    pub fn gsc_collider_count(ctx: *mut gsc_ctx, n_out: *mut u32) -> c_int;
    pub fn gsc_hierarchy_upload(ctx: *mut gsc_ctx, n_nodes: u32, n_rows: u32, row_begin: *const u32, parent: *const u32,
                                n_colliders: u32, collider_node: *const u32) -> c_int;
    pub fn gsc_hierarchy_update(ctx: *mut gsc_ctx, local_pos3: *const f32, local_quat4: *const f32,
                                local_scale3: *const f32) -> c_int;
    pub fn gsc_hierarchy_download(ctx: *mut gsc_ctx, pos3: *mut f32, quat4: *mut f32, scale3: *mut f32,
                                  matrix16: *mut f32) -> c_int;
    pub fn gsc_download_pairs(ctx: *mut gsc_ctx, dst: *mut u32, cap_pairs: u64, sorted: c_int, n_out: *mut u64) -> c_int;
    pub fn gsc_download_contacts(ctx: *mut gsc_ctx, dst: *mut f32, cap_pairs: u64, sorted: c_int, n_out: *mut u64) -> c_int;
    pub fn gsc_build_adjacency(ctx: *mut gsc_ctx, out: *mut gsc_adjacency) -> c_int;
    pub fn gsc_download_adjacency(ctx: *mut gsc_ctx, entity: *mut u32, offset: *mut u32, other: *mut u32, cap_entities: u64,
                                  cap_entries: u64) -> c_int;
    pub fn gsc_debug_download(ctx: *mut gsc_ctx, what: c_int, dst: *mut c_void, dst_bytes: usize) -> c_int;
    pub fn gsc_test_sphere_sphere(device: c_int, n: usize, a4: *const f32, b4: *const f32, out: *mut u8) -> c_int;
    pub fn gsc_test_sphere_obb(device: c_int, n: usize, s4: *const f32, obb16: *const f32, out: *mut u8) -> c_int;
    pub fn gsc_test_obb_obb(device: c_int, n: usize, a16: *const f32, b16: *const f32, out: *mut u8) -> c_int;
    pub fn gsc_test_aabb_aabb(device: c_int, n: usize, a6: *const f32, b6: *const f32, out: *mut u8) -> c_int;
    pub fn gsc_fnv1a_gridcell(device: c_int, n: usize, cells3: *const i16, out: *mut u64) -> c_int;
    pub fn gsc_contact_sphere_sphere(device: c_int, n: usize, hi4: *const f32, lo4: *const f32, out4: *mut f32) -> c_int;
    pub fn gsc_contact_sphere_obb(device: c_int, n: usize, s4: *const f32, obb16: *const f32, sphere_is_hi: *const u8,
                                  out4: *mut f32) -> c_int;
    pub fn gsc_contact_obb_obb(device: c_int, n: usize, hi16: *const f32, lo16: *const f32, out4: *mut f32) -> c_int;
    pub fn gsc_sort_pairs_u32(device: c_int, n: usize, keys: *mut u32, vals: *mut u32, key_bits: c_int) -> c_int;
    pub fn gsc_phase_bounds(ctx: *mut gsc_ctx, local: *mut gsc_bounds) -> c_int;
    pub fn gsc_set_global_bounds(ctx: *mut gsc_ctx, global: *const gsc_bounds) -> c_int;
    pub fn gsc_set_x_window(ctx: *mut gsc_ctx, x_lo: i32, x_hi: i32) -> c_int;
    pub fn gsc_local_x_range(ctx: *mut gsc_ctx, x_min: *mut i32, x_max: *mut i32) -> c_int;
    pub fn gsc_halo_select(ctx: *mut gsc_ctx, x_lo: i32, x_hi: i32, d_records: *mut c_void, cap_records: u32,
                           count: *mut u32) -> c_int;
    pub fn gsc_halo_append(ctx: *mut gsc_ctx, count: u32, d_records: *const c_void) -> c_int;
    pub fn gsc_ipc_export(ctx: *mut gsc_ctx, blob: *mut c_void, blob_bytes: usize) -> c_int;
    pub fn gsc_ipc_open_peer(ctx: *mut gsc_ctx, peer: u32, blob: *const c_void, blob_bytes: usize) -> c_int;
    pub fn gsc_halo_push(ctx: *mut gsc_ctx, peer: u32, x_lo: i32, x_hi: i32, peer_n_local: u32) -> c_int;
    pub fn gsc_halo_sync(ctx: *mut gsc_ctx) -> c_int;
    pub fn gsc_halo_commit(ctx: *mut gsc_ctx, n_ghost: *mut u32) -> c_int;
    pub fn gsc_phase_finish(ctx: *mut gsc_ctx, out: *mut gsc_step_out) -> c_int;
    // device-driven slab exchange
    pub fn gsc_open_peer_local(ctx: *mut gsc_ctx, peer: u32, other: *mut gsc_ctx) -> c_int;
    pub fn gsc_slab_init(ctx: *mut gsc_ctx, rank: u32, world: u32) -> c_int;
    pub fn gsc_slab_step(ctx: *mut gsc_ctx, out: *mut gsc_step_out) -> c_int;
    pub fn gsc_slab_step_begin(ctx: *mut gsc_ctx) -> c_int;
    // one handle for all GPUs of a box
    pub fn gsc_group_create(n_devices: u32, devices: *const i32, cfg: *const gsc_config, out: *mut *mut gsc_group) -> c_int;
    pub fn gsc_group_destroy(group: *mut gsc_group);
    pub fn gsc_group_last_error(group: *const gsc_group) -> *const c_char;
    pub fn gsc_group_upload_colliders(group: *mut gsc_group, n: u32, entity: *const u32, kind: *const u8, offset3: *const f32,
                                      size3: *const f32, pos3: *const f32) -> c_int;
    pub fn gsc_group_update_transforms(group: *mut gsc_group, n: u32, pos3: *const f32, quat4: *const f32,
                                       scale3: *const f32) -> c_int;
    pub fn gsc_group_step(group: *mut gsc_group, total: *mut gsc_step_out) -> c_int;
    pub fn gsc_group_download_pairs(group: *mut gsc_group, dst: *mut u32, cap_pairs: u64, sorted: c_int, n_out: *mut u64) -> c_int;
    pub fn gsc_group_rebalance(group: *mut gsc_group, n: u32, pos3: *const f32) -> c_int;
    pub fn gsc_group_reorder(group: *mut gsc_group) -> c_int;
    pub fn gsc_group_member(group: *mut gsc_group, i: u32, ctx: *mut *mut gsc_ctx, n_local: *mut u32, owned: *mut u32) -> c_int;
    // pinned transform staging
    pub fn gsc_map_transform_staging(ctx: *mut gsc_ctx, bank: u32, pos3: *mut *mut f32, box_quat4: *mut *mut f32,
                                     box_scale3: *mut *mut f32) -> c_int;
    pub fn gsc_commit_transform_staging(ctx: *mut gsc_ctx, bank: u32, n: u32, n_box: u32, with_rotation: c_int,
                                        with_scale: c_int) -> c_int;
    pub fn gsc_box_slots(ctx: *mut gsc_ctx, box_slot: *mut u32, cap: u32, n_box: *mut u32) -> c_int;
    pub fn gsc_reorder_colliders(ctx: *mut gsc_ctx, order: *mut u32, cap: u32) -> c_int;
}

/*
 * gunship_collision.h -- C ABI of the B200-native Gunship collision step.
 *
 * This is the drop-in boundary for the one data-parallel hot path of randomPoison/gunship-rs:
 * `CollisionSystem::update` = `bvh_update` + `GridCollisionSystem::update`
 * (src/component/collider/mod.rs:590-611, bounding_volume.rs:345-409, grid_collision.rs:218-285).
 * Everything above it (Collider / ColliderManager / CollisionSystem / callbacks) stays in the host
 * language; everything below it is hand-written CUDA for sm_100a.  Plain pointers and sizes only:
 * no C++ types, no torch types.  Every entry point returns an `int` status (GSC_OK == 0) and never
 * throws or aborts across the boundary; the reference panics in the same situations
 * (Cargo.toml:27-40 `panic = "abort"`), so a host wrapper turns a non-zero status into a panic.
 *
 * Threading: a context is owned by one caller thread (the reference's `update(&mut self, ..)`,
 * src/old/engine.rs:193); distinct contexts are independent.
 *
 * Ownership: input host arrays are borrowed for the duration of the call only -- with one exception: the
 * per-frame transform arrays of gsc_update_transforms / _packed are copied asynchronously (truly so when the
 * memory is pinned) and must stay valid and unmodified until the gsc_step / gsc_step_end that consumes them has
 * returned.  The library owns all device buffers; device pointers it hands out stay valid until the next
 * gsc_step / gsc_step_begin / gsc_destroy.
 */
#ifndef GUNSHIP_COLLISION_H
#define GUNSHIP_COLLISION_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSC_ABI_VERSION 2

/* status codes */
enum {
    GSC_OK           = 0,
    GSC_ERR_INVALID  = 1, /* bad argument / call order */
    GSC_ERR_CUDA     = 2, /* CUDA runtime failure (message in gsc_last_error) */
    GSC_ERR_NCCL     = 3, /* multi-GPU exchange failure: a peer GPU did not arrive within the time limit
                             (gsc_slab_step / gsc_group_step; GSC_PEER_TIMEOUT_MS, default 20000) */
    GSC_ERR_CAPACITY = 4, /* pair / candidate / cell-table capacity exceeded; needed sizes are in gsc_step_out */
    GSC_ERR_RANGE    = 5, /* a grid coordinate leaves i16 (grid_collision.rs:530-535), cell size is not
                             positive, or an input is NaN/Inf -- the reference's behaviour is undefined there */
    GSC_ERR_KIND     = 6  /* Collider::Mesh: `unimplemented!()` in the reference (mod.rs:331,343) */
};

/* collider kinds -- `enum Collider` (mod.rs:139-183) */
enum { GSC_SPHERE = 0, GSC_BOX = 1, GSC_MESH = 2 };

/* gsc_config.flags */
enum {
    GSC_FLAG_TIMING = 1u << 0, /* record per-stage CUDA-event times in gsc_step_out.stage_ms */
    GSC_FLAG_GRAPH  = 1u << 1, /* replay stages 1-5 of gsc_step from a captured CUDA graph (one graph launch instead of ~18
                                  kernel launches; matters for small scenes, where the step is launch bound) */
    GSC_FLAG_CONTACTS = 1u << 2 /* also produce a contact (unit normal, depth) per pair -- see gsc_download_contacts */
};

/* stages of one step, in launch order (names follow the reference's Stopwatch scopes) */
enum {
    GSC_STAGE_BVH_UPDATE = 0, /* "BVH Update": world shapes, AABBs, longest axis, world bounds */
    GSC_STAGE_CELL_KEYS  = 1, /* world_to_grid + cell keys + digit histograms */
    GSC_STAGE_SORT       = 2, /* radix sort of (cell, collider) */
    GSC_STAGE_CELL_TABLE = 3, /* cell start table + cell-order permutation of the volumes */
    GSC_STAGE_PAIRS      = 4, /* neighbour-cell pair enumeration + conservative AABB filter -> candidate lists per shape pair */
    GSC_STAGE_NARROW     = 5, /* exact AABB test, then sphere-sphere, sphere-OBB and OBB-OBB SAT over the candidates (+ wide volumes) */
    GSC_NUM_STAGES       = 6
};

typedef struct gsc_ctx gsc_ctx;

typedef struct gsc_config {
    int32_t  device;         /* CUDA device ordinal */
    uint32_t max_colliders;  /* capacity: local colliders + ghosts received from other slabs */
    uint64_t max_pairs;      /* capacity of the contact-pair buffer (0 = 8 * max_colliders + 65536) */
    uint64_t max_candidates; /* capacity of each of the three candidate lists (0 = 8 * max_colliders + 65536) */
    uint64_t max_cells;      /* capacity of the dense cell table (0 = 16 * max_colliders + 2^20) */
    uint32_t flags;          /* GSC_FLAG_* */
    uint32_t reserved;
} gsc_config;

typedef struct gsc_step_out {
    uint64_t n_pairs;        /* contact pairs found (may exceed max_pairs on GSC_ERR_CAPACITY) */
    uint64_t n_sphere_sphere;/* sphere/sphere candidates handed to the shape-test stage: pairs of neighbouring volumes that passed
                                the conservative (quantised) AABB filter; that stage runs the exact AABB::test_aabb first */
    uint64_t n_sphere_obb;   /* sphere/box candidates handed to the SAT stage (same definition) */
    uint64_t n_obb_obb;      /* box/box candidates handed to the SAT stage (same definition) */
    uint64_t n_cells;        /* cells of the dense grid this frame */
    float    cell_size;      /* BoundingVolumeManager::longest_axis (bounding_volume.rs:365-381) */
    uint32_t n_wide;         /* volumes whose AABB spans > 2 cells on an axis (debug_assert grid_collision.rs:403-410) */
    int32_t  grid_origin[3]; /* min cell coordinate per axis */
    int32_t  grid_dims[3];   /* cells per axis */
    uint32_t sort_passes;    /* radix passes run on the cell keys (digits of up to 9 bits) */
    uint32_t status_flags;   /* raw device error bits (diagnostics) */
    float    stage_ms[GSC_NUM_STAGES]; /* only with GSC_FLAG_TIMING */
    float    step_ms;        /* device time of the whole step (always) */
    const uint32_t *d_pairs; /* DEVICE pointer: n_pairs x (entity of higher index, entity of lower index) */
    const float *d_contacts; /* DEVICE pointer: n_pairs x (normal.xyz, depth), same order; NULL without GSC_FLAG_CONTACTS */
    uint32_t n_ghost;        /* ghost colliders other slabs pushed into this context this step (gsc_slab_step) */
    uint32_t reserved;
} gsc_step_out;

/* ---- lifetime ---- */
int  gsc_create(const gsc_config *cfg, gsc_ctx **out);
void gsc_destroy(gsc_ctx *ctx);
/* change GSC_FLAG_TIMING / GSC_FLAG_GRAPH of a live context (GSC_FLAG_CONTACTS is fixed at gsc_create) */
int  gsc_set_flags(gsc_ctx *ctx, uint32_t flags);
/* last error text of this context (never NULL); ctx == NULL gives the last gsc_create failure */
const char *gsc_last_error(const gsc_ctx *ctx);
int  gsc_abi_version(void);

/* ---- collider registration: ColliderManager::assign over the whole dense array
 *      (mod.rs:221-223; iteration order = index order, struct_component_manager.rs:112-117) ----
 * entity[n]; kind[n] (GSC_SPHERE/GSC_BOX); offset3[3n]; size3[3n] = (radius,-,-) or box widths.
 * global_index may be NULL (index i) -- it is the collider's position in the reference's
 * BoundingVolumeManager.components, which orients every output tuple (grid_collision.rs:440,497). */
int gsc_upload_colliders(gsc_ctx *ctx, uint32_t n, const uint32_t *entity, const uint8_t *kind,
                         const float *offset3, const float *size3, const uint32_t *global_index);

/* ---- per-frame derived transforms (TransformManager position/rotation/scale_derived,
 *      component/transform.rs:441-459) from HOST memory; quat is (x,y,z,w).
 *      quat4 / scale3 may be NULL to keep the previous values. ---- */
int gsc_update_transforms(gsc_ctx *ctx, uint32_t n, const float *pos3, const float *quat4, const float *scale3);

/* same frame data with the rotation / scale arrays packed over the BOXES only, in collider order (box k of the
 * dense array is the k-th collider with kind GSC_BOX): spheres ignore rotation and scale (mod.rs:313-317), so a
 * host layer that marshals transforms anyway need not send them -- 12 B per collider + 28 B per box over PCIe
 * instead of 40 B per collider.  box_quat4 / box_scale3 may be NULL to keep the previous packed values. */
int gsc_update_transforms_packed(gsc_ctx *ctx, uint32_t n, const float *pos3, uint32_t n_box, const float *box_quat4,
                                 const float *box_scale3);

/* ---- pinned transform staging: the host layer marshals the frame's derived transforms straight into library-owned
 *      PINNED buffers (bounding_volume.rs:358-362 reads every collider's transform once per frame; a host that fills
 *      pageable vectors instead pays a second, staged copy inside the driver).  Two banks alternate, so frame f+1 can be
 *      written while frame f's copy is still in flight.  pos3[3 * max_colliders] is in dense collider order;
 *      box_quat4 / box_scale3 are packed over the boxes (entry k = the k-th collider of kind GSC_BOX), exactly the
 *      arrays of gsc_update_transforms_packed.  gsc_commit_transform_staging hands bank `bank` over (asynchronous,
 *      same ordering rules as gsc_update_transforms_packed); with_rotation / with_scale == 0 keep the previous values. */
int gsc_map_transform_staging(gsc_ctx *ctx, uint32_t bank, float **pos3, float **box_quat4, float **box_scale3);
int gsc_commit_transform_staging(gsc_ctx *ctx, uint32_t bank, uint32_t n, uint32_t n_box, int with_rotation, int with_scale);
/* box_slot[i] = number of boxes before collider i (where collider i's rotation / scale go in the packed arrays) */
int gsc_box_slots(gsc_ctx *ctx, uint32_t *box_slot, uint32_t cap, uint32_t *n_box);

/* same, from DEVICE memory that the caller keeps alive until the next bind (no copy is made) */
int gsc_bind_device_transforms(gsc_ctx *ctx, uint32_t n, const float *d_pos3, const float *d_quat4, const float *d_scale3);

/* ---- the step: CollisionSystem::update (mod.rs:590-611) minus the callback dispatch ---- */
int gsc_step(gsc_ctx *ctx, gsc_step_out *out);

/* the same step in two halves, for a host loop that overlaps the NEXT frame's transform upload with this frame's
 * step (PCIe is the slowest link of a step fed from host memory): gsc_step_begin enqueues the whole step and
 * returns; gsc_step_end waits for it and reports exactly what gsc_step would have.  Between the two the only calls
 * allowed on the context are gsc_update_transforms / _packed / gsc_bind_device_transforms for the next frame -- the
 * copy is ordered behind bvh_update of the step in flight (the one reader of the transform arrays) and runs under
 * the rest of it.  The reference's update blocks (grid_collision.rs:262-272); gsc_step keeps that form. */
int gsc_step_begin(gsc_ctx *ctx);
int gsc_step_end(gsc_ctx *ctx, gsc_step_out *out);

/* ---- collider lifecycle after the first upload ----
 * ColliderManager::assign appends to the dense arrays (struct_component_manager.rs:83-96; BoundingVolumeManager
 * pushes a volume the first time bvh_update meets the entity, bounding_volume.rs:394-407): gsc_append_colliders adds
 * n_new colliders behind the uploaded ones without touching those.
 * Destruction is a swap_remove of the dense arrays (BoundingVolumeManager::destroy_immediate,
 * bounding_volume.rs:82-104, driven by the cleanup loop of CollisionSystem::update, mod.rs:605-609): the last
 * element moves into the hole, so the INDEX order -- which orients every output tuple and fixes the OBB-OBB argument
 * order -- changes exactly as in the reference.  gsc_remove_colliders applies the removals index[0], index[1], ...
 * in that order, each index meaning the dense position at the time of its removal (what the reference's
 * `indices.remove(&entity)` yields), to the device-resident static arrays; nothing is re-uploaded.
 * After either call the next frame's transforms must be provided for the new count before the next step. */
int gsc_append_colliders(gsc_ctx *ctx, uint32_t n_new, const uint32_t *entity, const uint8_t *kind,
                         const float *offset3, const float *size3);
int gsc_remove_colliders(gsc_ctx *ctx, uint32_t n_remove, const uint32_t *index);
/* colliders currently registered on the device */
int gsc_collider_count(gsc_ctx *ctx, uint32_t *n_out);

/* ---- temporal coherence (second half) ----
 * The reference rebuilds its grid from scratch every frame (grid_collision.rs:484-485) and keeps its volumes in
 * assign order.  Colliders move little from frame to frame, so the CELL ORDER of one frame is a good memory order for
 * the next ones: gsc_reorder_colliders permutes the context's dense arrays into the cell order of the step that
 * just completed.  order[slot] receives the collider's previous dense index: from now on the host hands over
 * transforms in SLOT order (transform of collider order[slot] at position slot -- the reference's bvh_update looks
 * every transform up by entity anyway, bounding_volume.rs:358-362).  Nothing observable changes: every collider keeps
 * its global index (its position in the reference's dense array), which is what orients the output tuples and fixes
 * the OBB-OBB argument order.  Afterwards the cell-order gather of the volume records and the record gathers of the
 * SAT stage run over nearly sequential memory instead of random indices.  Call it again whenever the scene has
 * drifted (every few hundred frames).  Like a context with explicit global indices, a reordered context does not
 * take gsc_append_colliders / gsc_remove_colliders (upload the new set instead). */
int gsc_reorder_colliders(gsc_ctx *ctx, uint32_t *order, uint32_t cap);

/* ---- the step BEFORE the path: transform derivation ----
 * TransformManager::update_transforms walks the rows of the hierarchy (row = depth; "the transforms in a row can be
 * processed independently so they should be done in parallel", component/transform.rs:243-251) and
 * TransformData::update (:588-600) derives, per node,
 *     matrix_derived   = parent.matrix_derived * (T(position) * (R(rotation) * S(scale)))      (:593, :602-608)
 *     position_derived = matrix_derived.translation_part()                                     (:595)
 *     rotation_derived = parent.rotation_derived * rotation                                    (:596, quaternion.rs:164-169)
 *     scale_derived    = scale * parent.scale_derived                                          (:597)
 * with the identity dummy as the parent of a root (:47-61).  Here: one kernel launch per row, one thread per node,
 * f32 operation order as written in lib/polygon_math (Matrix4 * Matrix4 matrix.rs:207-243, all four products of
 * every element kept); the derived values stay on the device and feed bvh_update directly, so a frame needs no
 * separate gsc_update_transforms.
 * Nodes are numbered row by row: row r holds nodes row_begin[r] .. row_begin[r+1]-1; parent[i] is a node of the
 * previous row (GSC_NO_PARENT for the nodes of row 0).  collider_node[k] names the node whose derived transform
 * collider k of the dense array reads (mod.rs:311-333). */
#define GSC_NO_PARENT 0xffffffffu
int gsc_hierarchy_upload(gsc_ctx *ctx, uint32_t n_nodes, uint32_t n_rows, const uint32_t *row_begin,
                         const uint32_t *parent, uint32_t n_colliders, const uint32_t *collider_node);
/* one frame: local position / rotation (x,y,z,w) / scale of every node from HOST memory (NULL keeps the previous
 * values of that array), derive, and bind the colliders' transforms for the next step */
int gsc_hierarchy_update(gsc_ctx *ctx, const float *local_pos3, const float *local_quat4, const float *local_scale3);
/* derived values of every node, to HOST memory (any pointer may be NULL): pos3[3n], quat4[4n], scale3[3n],
 * matrix16[16n] row-major */
int gsc_hierarchy_download(gsc_ctx *ctx, float *pos3, float *quat4, float *scale3, float *matrix16);

/* copy the contact pairs of the last step to host memory: dst[2*i] = entity of the later (higher
 * index) volume, dst[2*i+1] = entity of the earlier one -- the tuple order of
 * GridCollisionSystem::collisions (grid_collision.rs:119,497).  sorted != 0 orders them ascending by
 * (first << 32 | second) on the device first.  *n_out receives the pair count; dst == NULL with
 * cap_pairs == 0 only queries it. */
int gsc_download_pairs(gsc_ctx *ctx, uint32_t *dst, uint64_t cap_pairs, int sorted, uint64_t *n_out);

/* Contacts (GSC_FLAG_CONTACTS).  The reference computes boolean collisions only -- penetration depth and contact
 * normal are planned there (mod.rs:110-113) but do not exist, so there is nothing to be bit-exact WITH: this is the
 * definition this library adds (parity unpinned; checked against the same definition in oracle/ at 1e-5 relative).
 * For the pair (hi, lo) = (later volume, earlier volume): `normal` is a unit vector pointing from lo to hi, `depth`
 * the overlap along it (moving hi by depth * normal separates the two shapes along that axis).
 *   sphere-sphere : normal = (c_hi - c_lo) / |c_hi - c_lo| ((1,0,0) if the centres coincide); depth = r_hi + r_lo - dist
 *   sphere-box    : q = OBB::closest_point(centre) (mod.rs:549-568); outside: normal = +-(centre - q)/|centre - q|,
 *                   depth = r - |centre - q|; centre inside the box (no projection clamped): the face of least
 *                   penetration, depth = r + that
 *   box-box       : the axis of least overlap among the 15 SAT axes of OBB::test_obb (mod.rs:443-542), cross axes
 *                   normalised by |A_i x B_j| and skipped when nearly parallel; depth = that overlap
 * dst[4*i .. 4*i+3] = (nx, ny, nz, depth) of pair i in the order of gsc_download_pairs(sorted). */
int gsc_download_contacts(gsc_ctx *ctx, float *dst, uint64_t cap_pairs, int sorted, uint64_t *n_out);

/* ---- the consumer of the pair set: CollisionCallbackManager::process_collisions (mod.rs:678-708) builds, for
 *      every entity, the list of entities it collides with (each pair contributes both directions, :687-690) and
 *      invokes the callbacks with (entity, &[others]).  gsc_build_adjacency produces those lists on the device as
 *      CSR: entity i of d_entity (ascending) collides with d_other[d_offset[i] .. d_offset[i+1]) (ascending; the
 *      reference's order is a HashMap iteration order, i.e. unspecified).  Valid until the next step. ---- */
typedef struct gsc_adjacency {
    uint64_t n_entities;      /* entities with at least one collision */
    uint64_t n_entries;       /* 2 * n_pairs */
    const uint32_t *d_entity; /* DEVICE [n_entities] */
    const uint32_t *d_offset; /* DEVICE [n_entities + 1] */
    const uint32_t *d_other;  /* DEVICE [n_entries] */
} gsc_adjacency;
int gsc_build_adjacency(gsc_ctx *ctx, gsc_adjacency *out);
/* same, copied to host arrays (entity[cap_entities], offset[cap_entities + 1], other[cap_entries]) */
int gsc_download_adjacency(gsc_ctx *ctx, uint32_t *entity, uint32_t *offset, uint32_t *other, uint64_t cap_entities,
                           uint64_t cap_entries);

/* ---- stage-level read-back for parity tests (host buffers) ---- */
enum {
    GSC_DBG_AABB        = 0, /* float[6n]: min xyz, max xyz per collider (AABB::from_collider) */
    GSC_DBG_SORTED_KEYS = 1, /* uint32[n]: dense cell ids of world_to_grid(aabb.min) after the radix sort
                                (x-major; wide volumes carry n_cells) */
    GSC_DBG_SORTED_IDX  = 2, /* uint32[n]: collider indices after the radix sort */
    GSC_DBG_CELL_START  = 3, /* uint32[n_cells+1]: first sorted position of each cell */
    GSC_DBG_OBB         = 4, /* float[16n]: centre xyz, orientation row-major 9, half widths 3, entity bits */
    GSC_DBG_ENTITY      = 5, /* uint32[n]: entity of every collider in dense-array order */
    GSC_DBG_KIND        = 6, /* uint8[n]: kind, dense-array order */
    GSC_DBG_OFFSET      = 7, /* float[3n]: collider offset, dense-array order */
    GSC_DBG_SIZE        = 8  /* float[3n]: radius,-,- / box widths, dense-array order */
};
int gsc_debug_download(gsc_ctx *ctx, int what, void *dst, size_t dst_bytes);

/* ---- micro-bench parity: the functions benches/collision.rs and benches/grid_collision.rs call,
 *      batched over n independent inputs (HOST arrays; results in out[n]) ---- */
/* Sphere::test_sphere (mod.rs:383-389); spheres as (cx,cy,cz,r) */
int gsc_test_sphere_sphere(int device, size_t n, const float *a4, const float *b4, uint8_t *out);
/* Sphere::test_obb (mod.rs:391-394); OBB as 16 floats (centre 3, orientation row-major 9, half widths 3, pad) */
int gsc_test_sphere_obb(int device, size_t n, const float *s4, const float *obb16, uint8_t *out);
/* OBB::test_obb (mod.rs:413-546): a.test_obb(&b) */
int gsc_test_obb_obb(int device, size_t n, const float *a16, const float *b16, uint8_t *out);
/* AABB::test_aabb (bounding_volume.rs:338-343); AABBs as (min xyz, max xyz) */
int gsc_test_aabb_aabb(int device, size_t n, const float *a6, const float *b6, uint8_t *out);
/* FnvHasher over GridCell (lib/hash/src/fnv.rs:21-38, grid_collision.rs:523-528); cells as i16 xyz */
int gsc_fnv1a_gridcell(int device, size_t n, const int16_t *cells3, uint64_t *out);
/* contact definition above, batched: out4[4*i..] = (normal.xyz, depth); first argument is the later volume `hi` */
int gsc_contact_sphere_sphere(int device, size_t n, const float *hi4, const float *lo4, float *out4);
int gsc_contact_sphere_obb(int device, size_t n, const float *s4, const float *obb16, const uint8_t *sphere_is_hi, float *out4);
int gsc_contact_obb_obb(int device, size_t n, const float *hi16, const float *lo16, float *out4);
/* the radix sort used on (cell, collider) keys, exposed for testing: stable LSD over `key_bits` bits */
int gsc_sort_pairs_u32(int device, size_t n, uint32_t *keys, uint32_t *vals, int key_bits);

/* ---- multi-GPU slab support (x-slab partition, one context per GPU; the host layer does the exchange,
 *      gunship_rs_b200/slabs.py over torch.distributed/NCCL) ----
 * The reference already decomposes space (8 octant work units that each see every volume and dedupe at the
 * merge, grid_collision.rs:60-86,161-193,262-272).  Here each rank owns a subset of the colliders, ideally an
 * x-slab.  A pair is reported by exactly one rank: the owner of its "taker" -- the volume that comes later in
 * (home cell key, global index) order -- which only ever looks at home x-cells {x-1, x}.  So rank r needs as
 * ghosts the volumes of other ranks whose home x-cell lies in [xmin_r - 1, xmax_r].  One step on rank r:
 *   gsc_phase_bounds -> [all-gather of the gsc_bounds: global cell size / world box, every rank's home x-range]
 *   -> gsc_set_global_bounds -> gsc_set_x_window(xmin_r - 1, xmax_r) -> for every rank s that needs some of mine:
 *   gsc_halo_select(xmin_s - 1, xmax_s) -> [send/recv] -> gsc_halo_append -> gsc_phase_finish.
 * The union of the ranks' pair lists is the single-GPU pair set; no cross-rank dedupe is needed.
 * Volumes that span 3 cells through rounding (n_wide, expected 0) are only matched within their own rank.
 * Single-GPU callers use gsc_step, which runs the same stages back to back. */
typedef struct gsc_bounds {
    float longest_axis;  /* max AABB extent over the local colliders */
    float world_min[3];  /* min over local AABB mins */
    float world_max[3];  /* max over local AABB maxs */
    uint32_t status_flags;
    float home_x_max;    /* max over local AABB min.x: floor(world_min[0] / cell) .. floor(home_x_max / cell) is the
                            home x-cell range of the local colliders once the global cell size is known */
} gsc_bounds;

#define GSC_HALO_RECORD_BYTES 96 /* 32 B volume record + 64 B OBB record */

/* stage 0 on the local colliders; returns the local bounds */
int gsc_phase_bounds(gsc_ctx *ctx, gsc_bounds *local);
/* install the globally reduced bounds (max of longest_axis, min of world_min, max of world_max) */
int gsc_set_global_bounds(gsc_ctx *ctx, const gsc_bounds *global);
/* restrict this rank's grid to the home x-cells [x_lo, x_hi] (absolute cells; normally xmin_r - 1 .. xmax_r): the
 * cell table then covers the slab, not the world.  Must contain every local collider's home cell (GSC_ERR_INVALID
 * from gsc_phase_finish otherwise); ghosts outside it are ignored.  Valid for the current step only. */
int gsc_set_x_window(gsc_ctx *ctx, int32_t x_lo, int32_t x_hi);
/* min/max home x-cell (floor(aabb.min.x / cell), absolute) of the local colliders; empty => *x_min > *x_max */
int gsc_local_x_range(gsc_ctx *ctx, int32_t *x_min, int32_t *x_max);
/* pack into the caller's DEVICE buffer (16-byte aligned, cap_records * GSC_HALO_RECORD_BYTES) the local
 * colliders whose home x-cell lies in [x_lo, x_hi]; *count receives how many (GSC_ERR_CAPACITY if > cap) */
int gsc_halo_select(gsc_ctx *ctx, int32_t x_lo, int32_t x_hi, void *d_records, uint32_t cap_records, uint32_t *count);
/* append `count` ghost colliders (records produced by gsc_halo_select on another rank, DEVICE memory);
 * ghosts take part as the earlier volume of a pair only */
int gsc_halo_append(gsc_ctx *ctx, uint32_t count, const void *d_records);
/* -- the same exchange fused with the transfer, for GPUs of one box (NVLink / NVSwitch peer memory) --
 * Setup (once): every rank calls gsc_ipc_export, the blobs are all-gathered, every rank opens every other
 * rank's blob with gsc_ipc_open_peer(peer = rank).  Per step, instead of select / send / recv / append:
 *   gsc_halo_push(peer s, xmin_s - 1, xmax_s, n_local of s)   select + write straight into rank s's arrays
 *   gsc_halo_sync -> [barrier across ranks] -> gsc_halo_commit   adopt what the peers pushed into me
 * The receiving context must not start its next gsc_phase_bounds before all pushes of this step are done
 * (the barrier), and pushes must not start before every rank's gsc_phase_bounds of this step has returned
 * (the bounds all-gather orders that). */
#define GSC_MAX_PEERS 16
#define GSC_IPC_EXPORT_BYTES 256
int gsc_ipc_export(gsc_ctx *ctx, void *blob, size_t blob_bytes);
int gsc_ipc_open_peer(gsc_ctx *ctx, uint32_t peer, const void *blob, size_t blob_bytes);
int gsc_halo_push(gsc_ctx *ctx, uint32_t peer, int32_t x_lo, int32_t x_hi, uint32_t peer_n_local);
int gsc_halo_sync(gsc_ctx *ctx);
int gsc_halo_commit(gsc_ctx *ctx, uint32_t *n_ghost);

/* stages 1..5 on local + ghost colliders */
int gsc_phase_finish(gsc_ctx *ctx, gsc_step_out *out);

/* -- the same exchange driven from the DEVICE (round 2): no host round trip inside a step --
 * After the one-time peer setup (gsc_ipc_open_peer for ranks in other processes, gsc_open_peer_local for contexts of
 * this process) and gsc_slab_init, ONE call enqueues the whole multi-GPU step on the rank's own stream:
 *   bvh_update -> k_slab_publish (my bounds into every rank's mailbox, in peer memory) -> k_slab_gather (wait for all
 *   bounds; global cell size / world box; every rank's home x-range; what each peer needs of mine; my windowed grid)
 *   -> k_slab_push (select + write my boundary volumes into the peers' arrays) -> k_slab_done (flag) -> k_slab_wait
 *   (wait for the peers' flags, adopt the ghosts) -> stages 1..5 over local + ghost volumes
 * and the host waits once, at the end.  Every rank of the group must call it once per step (they step together);
 * a rank that does not arrive makes the others fail with GSC_ERR_NCCL after the time limit instead of hanging.
 * Ranks must run on DIFFERENT GPUs unless their steps are ordered by stream events (gsc_group_step does that):
 * two kernels of one GPU that wait for each other are not guaranteed to run at the same time. */
int gsc_open_peer_local(gsc_ctx *ctx, uint32_t peer, gsc_ctx *other);
int gsc_slab_init(gsc_ctx *ctx, uint32_t rank, uint32_t world);
int gsc_slab_step(gsc_ctx *ctx, gsc_step_out *out);
/* in two halves, as gsc_step_begin / gsc_step_end: finish with gsc_step_end */
int gsc_slab_step_begin(gsc_ctx *ctx);

/* ---- a whole box behind ONE handle: one process, one context per GPU (what a Rust / C++ engine links against) ----
 * The group owns the x-slab partition.  Colliders are registered and transforms handed over in the engine's dense
 * order; the group scatters them to the slabs (pinned staging per member), steps all members with the device-driven
 * exchange above (ordered by stream events, so members may even share a GPU) and returns the union of the members'
 * pair lists, which IS the single-GPU pair set (no cross-rank dedupe).  gsc_group_rebalance re-cuts the slabs from
 * the current positions and the per-slab work of the last step (colliders + candidate pairs; migrants change owner;
 * SURVEY 8e collectives 2 and 3). */
typedef struct gsc_group gsc_group;
int  gsc_group_create(uint32_t n_devices, const int32_t *devices, const gsc_config *cfg, gsc_group **out);
void gsc_group_destroy(gsc_group *group);
const char *gsc_group_last_error(const gsc_group *group);
/* pos3: positions used to cut the slabs (normally the first frame's); cfg.max_colliders is the capacity PER MEMBER */
int gsc_group_upload_colliders(gsc_group *group, uint32_t n, const uint32_t *entity, const uint8_t *kind,
                               const float *offset3, const float *size3, const float *pos3);
int gsc_group_update_transforms(gsc_group *group, uint32_t n, const float *pos3, const float *quat4, const float *scale3);
/* total: sums over the members (stage_ms / step_ms: max over members; d_pairs / d_contacts NULL) */
int gsc_group_step(gsc_group *group, gsc_step_out *total);
int gsc_group_download_pairs(gsc_group *group, uint32_t *dst, uint64_t cap_pairs, int sorted, uint64_t *n_out);
/* re-cut the slabs by the work of the last step; pos3 = the CURRENT positions of all n colliders (dense order), which
 * decide the new owners.  After gsc_group_rebalance and after gsc_group_reorder the next gsc_group_update_transforms must
 * carry rotation and scale again (the group keeps no copy of the engine's transform arrays). */
int gsc_group_rebalance(gsc_group *group, uint32_t n, const float *pos3);
/* temporal coherence behind the group: every member keeps its colliders in the cell order of the step that just
 * completed (gsc_reorder_colliders); the engine keeps handing over dense arrays in ITS order */
int gsc_group_reorder(gsc_group *group);
/* inspection: member i's context, its collider count and the global indices it owns (owned[n_local], may be NULL) */
int gsc_group_member(gsc_group *group, uint32_t i, gsc_ctx **ctx, uint32_t *n_local, uint32_t *owned);

#ifdef __cplusplus
}
#endif
#endif /* GUNSHIP_COLLISION_H */

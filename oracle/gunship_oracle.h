/*
 * gunship_oracle.h -- CPU restatement of Gunship's collision hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and there only as the checker / the timed
 * CPU baseline.  The product path (gunship_rs_b200/) never calls into it.
 *
 * The reference (randomPoison/gunship-rs, Rust, 2016 nightly) cannot be built
 * in this image (no rustc/cargo, unresolvable deps, collider module not part of
 * the crate -- SURVEY.md F2-F5), so this is a plain-C restatement, op for op in
 * IEEE binary32 with no FMA contraction (-ffp-contract=off), of:
 *
 *   src/component/collider/mod.rs            (Collider, CachedCollider, Sphere, OBB, tests)
 *   src/component/collider/bounding_volume.rs (AABB, BoundVolume, bvh_update)
 *   src/component/collider/grid_collision.rs  (GridCollisionSystem, WorkUnit, Worker, GridCell)
 *   lib/polygon_math/src/{lib,vector,point,matrix}.rs
 *   lib/hash/src/fnv.rs
 *
 * Parity pinning: the oracle is checked against every known-answer test the
 * reference holds for this path (mod.rs:711-782 sphere/OBB KATs,
 * quaternion_test.rs:18-33 quaternion->matrix KATs) in tests/test_oracle_kat.py.
 * The reference has NO test of the grid broadphase, AABB construction,
 * sphere-OBB or the pair-set output; for those the oracle is pinned only by the
 * restated algorithm plus the brute-force equivalence (orc_brute_force).
 * Penetration depth / contact normal: the reference has none -- parity unpinned.
 */
#ifndef GUNSHIP_ORACLE_H
#define GUNSHIP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lib/polygon_math/src/point.rs:13-20 -- Point is x,y,z,w (16 B). */
typedef struct { float x, y, z, w; } orc_point;
/* lib/polygon_math/src/vector.rs:9-15 */
typedef struct { float x, y, z; } orc_vec3;
/* lib/polygon_math/src/matrix.rs:293-295 -- row-major [[f32;3];3] */
typedef struct { float m[3][3]; } orc_mat3;
/* bounding_volume.rs:140-144 */
typedef struct { orc_point min, max; } orc_aabb;
/* mod.rs:368-372 */
typedef struct { orc_point center; float radius; } orc_sphere;
/* mod.rs:397-402 */
typedef struct { orc_point center; orc_mat3 orientation; orc_vec3 half_widths; } orc_obb;

enum { ORC_SPHERE = 0, ORC_BOX = 1, ORC_MESH = 2 };

/* mod.rs:303-308 -- tagged union, 68 B */
typedef struct {
    uint32_t tag;
    union { orc_sphere sphere; orc_obb obb; } u;
} orc_cached_collider;

/* bounding_volume.rs:115-120 -- 104 B */
typedef struct {
    uint32_t entity;
    orc_aabb aabb;
    orc_cached_collider collider;
} orc_bound_volume;

/* grid_collision.rs:523-535 */
typedef struct { int16_t x, y, z; } orc_grid_cell;

/* ---- arithmetic primitives (each cites what it follows) ---- */
void  orc_quat_to_mat3(const float q_xyzw[4], orc_mat3 *out);          /* matrix.rs:392-402 */
void  orc_mat3_rotation(float x, float y, float z, orc_mat3 *out);     /* matrix.rs:320-333 */
void  orc_cache_collider(uint8_t kind, const float offset[3], const float size[3],
                         const float pos[3], const float quat_xyzw[4], const float scale[3],
                         orc_cached_collider *out);                     /* mod.rs:311-333 */
void  orc_aabb_from_collider(const orc_cached_collider *c, orc_aabb *out); /* bounding_volume.rs:148-336 */
int   orc_test_aabb(const orc_aabb *a, const orc_aabb *b);              /* bounding_volume.rs:338-343,411-419 */
int   orc_test_sphere_sphere(const orc_sphere *a, const orc_sphere *b); /* mod.rs:383-389 */
int   orc_test_sphere_obb(const orc_sphere *s, const orc_obb *o);       /* mod.rs:391-394,549-574 */
int   orc_test_obb_obb(const orc_obb *a, const orc_obb *b);             /* mod.rs:413-546 */
int   orc_test_collider(const orc_cached_collider *a, const orc_cached_collider *b); /* mod.rs:335-345,375-381,405-411 */
int   orc_test_volume(const orc_bound_volume *a, const orc_bound_volume *b); /* bounding_volume.rs:124-132 */
uint64_t orc_fnv1a(const uint8_t *bytes, size_t n, uint64_t state);     /* fnv.rs:21-38 */
uint64_t orc_fnv1a_gridcell(orc_grid_cell c);                           /* grid_collision.rs:523 derive(Hash) + fnv.rs */
orc_grid_cell orc_world_to_grid(orc_point p, float cell_size);          /* grid_collision.rs:329-335 */

/* batch helpers for ctypes-driven parity tests (n independent pairs) */
void orc_batch_sphere_sphere(size_t n, const float *a_cxyz_r, const float *b_cxyz_r, uint8_t *out);
void orc_batch_sphere_obb(size_t n, const float *s_cxyz_r, const float *obb16, uint8_t *out);
void orc_batch_obb_obb(size_t n, const float *a_obb16, const float *b_obb16, uint8_t *out);
void orc_batch_aabb_aabb(size_t n, const float *a_minmax6, const float *b_minmax6, uint8_t *out);
void orc_batch_fnv1a_gridcell(size_t n, const int16_t *cells_xyz, uint64_t *out);

/* ---- contacts (normal.xyz, depth): NOT in the reference -- PARITY UNPINNED (see gunship_oracle.c) ---- */
void orc_contact_sphere_sphere(const orc_sphere *hi, const orc_sphere *lo, float out[4]);
void orc_contact_sphere_obb(const orc_sphere *s, const orc_obb *o, int sphere_is_hi, float out[4]);
void orc_contact_obb_obb(const orc_obb *hi, const orc_obb *lo, float out[4]);
void orc_batch_contact_sphere_sphere(size_t n, const float *hi_cxyz_r, const float *lo_cxyz_r, float *out4);
void orc_batch_contact_sphere_obb(size_t n, const float *s_cxyz_r, const float *obb16, const uint8_t *sphere_is_hi, float *out4);
void orc_batch_contact_obb_obb(size_t n, const float *hi_obb16, const float *lo_obb16, float *out4);

/* ---- the collision system (CollisionSystem::update = bvh_update + GridCollisionSystem::update) ---- */
typedef struct orc_system orc_system;

/* num_work_units in {1,2,4,8} (grid_collision.rs:130-196), num_workers >= 1 (grid_collision.rs:104). */
orc_system *orc_system_new(int num_work_units, int num_workers);
void        orc_system_free(orc_system *s);

/* bvh_update (bounding_volume.rs:345-409) over SoA inputs.
 * kind[i]: ORC_SPHERE / ORC_BOX.  offset3, size3 (sphere: size3[0]=radius; box: widths),
 * pos3, quat4 (x,y,z,w), scale3 are the *derived* transform of a root entity.
 * Returns 0, or -1 if a Mesh collider is met (reference: unimplemented!()). */
int  orc_bvh_update(orc_system *s, uint32_t n, const uint32_t *entity, const uint8_t *kind,
                    const float *offset3, const float *size3,
                    const float *pos3, const float *quat4, const float *scale3);
float orc_longest_axis(const orc_system *s);
uint32_t orc_num_volumes(const orc_system *s);
const orc_bound_volume *orc_components(const orc_system *s);

/* GridCollisionSystem::update (grid_collision.rs:218-285).  Returns number of
 * distinct collision pairs.  Pairs are (entity of later volume, entity of earlier volume). */
uint64_t orc_grid_update(orc_system *s);
/* copy the collision set out as packed u64 (first<<32 | second), ascending order */
uint64_t orc_collisions_sorted(orc_system *s, uint64_t *out, uint64_t cap);
/* statistics of the last grid update (summed over work units) */
uint64_t orc_last_candidates(const orc_system *s);   /* candidate_collisions pushed, incl. duplicates */
uint64_t orc_last_cell_entries(const orc_system *s); /* (cell, volume) insertions */
uint32_t orc_last_wide_volumes(const orc_system *s); /* volumes whose AABB spanned >2 cells on an axis (debug_assert :403-410) */
uint32_t orc_last_out_of_range(const orc_system *s); /* world_to_grid results that saturated i16 */

/* CollisionCallbackManager::process_collisions (mod.rs:678-708): per-entity lists of colliding entities as CSR,
 * entities and lists ascending.  Returns the entity count (call with cap_entities = 0 and `other` sized
 * 2 * n_pairs to query). */
uint64_t orc_process_collisions(orc_system *s, uint32_t *entity, uint32_t *offset, uint32_t *other, uint64_t cap_entities);

/* second oracle: O(N^2) over the volumes built by orc_bvh_update:
 * { (E[i],E[j]) : j<i, test_volume(V[i],V[j]) }, sorted ascending packed u64. */
uint64_t orc_brute_force_sorted(orc_system *s, uint64_t *out, uint64_t cap);

/* ---- the step before the path: TransformData::update row by row (component/transform.rs:243-251, 588-608);
 *      PARITY UNPINNED (the reference has no test of it) ---- */
void orc_transform_update(uint32_t n_rows, const uint32_t *row_begin, const uint32_t *parent, const float *lpos3,
                          const float *lquat4, const float *lscale3, float *dpos3, float *dquat4, float *dscale3,
                          float *dmat16);
/* ---- lifecycle: a sequence of destroy_immediate calls = swap_removes of the dense arrays (bounding_volume.rs:82-104) ---- */
uint32_t orc_swap_remove_sequence(uint32_t *order, uint32_t n, const uint32_t *index, uint32_t n_remove);

#ifdef __cplusplus
}
#endif
#endif

// gunship_collision.hpp -- the host side of the collision path above the C ABI, in C++17 (header only).
//
// The reference is Rust and there is no Rust toolchain in this image, so next to the Rust shim that is
// provided as source (rust/gunship-collision/) this header is the COMPILED host-side mirror of the reference's
// interface for the path: same type and method names, same argument meaning, same error behaviour -- where the
// reference panics (Cargo.toml:27-40 `panic = "abort"`) this throws gunship::Panic.  Nothing here computes
// collisions: every call forwards to libgunship_collision.so (include/gunship_collision.h).
//
//   Collider, ColliderManager, CollisionCallbackManager   src/component/collider/mod.rs:139-286, 613-708
//   BoundingVolumeManager (dense order, swap_remove)      src/component/collider/bounding_volume.rs:17-113
//   GridCollisionSystem { collisions }                    src/component/collider/grid_collision.rs:119-285
//   CollisionSystem::update                               src/component/collider/mod.rs:577-611
//   Sphere / OBB / AABB :: test_*                         mod.rs:368-575, bounding_volume.rs:338-343 (micro-benches)
//   TransformManager (what bvh_update reads of it)        src/component/transform.rs:441-459, bounding_volume.rs:359
#pragma once
#include <cmath>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "gunship_collision.h"

namespace gunship {

// the reference aborts on these; a C++ host gets an exception it can turn into std::abort()
struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

using Entity = uint32_t;  // src/old/ecs.rs:7-8

struct Vector3 { float x = 0, y = 0, z = 0; };                 // lib/polygon_math/src/vector.rs:9-15
struct Point { float x = 0, y = 0, z = 0, w = 1; };            // point.rs:13-20
struct Quaternion { float x = 0, y = 0, z = 0, w = 1; };       // quaternion.rs:22-25 as (v.x, v.y, v.z, w)
struct Matrix3 {                                               // matrix.rs:293-295, row-major
    float m[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
    static Matrix3 identity() { return Matrix3(); }
    // Matrix3::rotation(x, y, z) (matrix.rs:320-333): Euler angles applied x -> y -> z
    static Matrix3 rotation(float x, float y, float z)
    {
        const float s1 = std::sin(x), c1 = std::cos(x), s2 = std::sin(y), c2 = std::cos(y), s3 = std::sin(z), c3 = std::cos(z);
        Matrix3 r;
        r.m[0][0] = c2 * c3;                r.m[0][1] = -c2 * s3;               r.m[0][2] = s2;
        r.m[1][0] = c1 * s3 + c3 * s1 * s2; r.m[1][1] = c1 * c3 - s1 * s2 * s3; r.m[1][2] = -c2 * s1;
        r.m[2][0] = s1 * s3 - c1 * c3 * s2; r.m[2][1] = c3 * s1 + c1 * s2 * s3; r.m[2][2] = c1 * c2;
        return r;
    }
};

namespace detail {
inline void check(gsc_ctx *ctx, int rc)
{
    if (rc != GSC_OK) throw Panic(std::string("gunship_collision: status ") + std::to_string(rc) + ": " + gsc_last_error(ctx));
}
inline void check_batch(int rc)
{
    if (rc != GSC_OK) throw Panic(std::string("gunship_collision: status ") + std::to_string(rc) + ": " + gsc_last_error(nullptr));
}
struct PairHash {
    size_t operator()(const std::pair<Entity, Entity> &p) const { return std::hash<uint64_t>()(((uint64_t)p.first << 32) | p.second); }
};
}  // namespace detail

// ---- world-space shapes and their tests: the functions benches/collision.rs calls -------------------------------
struct OBB;
struct Sphere {  // mod.rs:368-372
    Point center;
    float radius = 0;
    bool test_sphere(const Sphere &other) const;  // mod.rs:383-389
    bool test_obb(const OBB &obb) const;          // mod.rs:391-394
};
struct OBB {  // mod.rs:397-402; the box axes are the COLUMNS of `orientation`
    Point center;
    Matrix3 orientation;
    Vector3 half_widths;
    bool test_obb(const OBB &other) const;  // mod.rs:413-546: self.test_obb(&other)
    void pack(float out[16]) const
    {
        out[0] = center.x; out[1] = center.y; out[2] = center.z;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) out[3 + 3 * i + j] = orientation.m[i][j];
        out[12] = half_widths.x; out[13] = half_widths.y; out[14] = half_widths.z; out[15] = 0.0f;
    }
};
struct AABB {  // bounding_volume.rs:140-144
    Point min, max;
    bool test_aabb(const AABB &other) const  // bounding_volume.rs:338-343
    {
        const float a[6] = { min.x, min.y, min.z, max.x, max.y, max.z };
        const float b[6] = { other.min.x, other.min.y, other.min.z, other.max.x, other.max.y, other.max.z };
        uint8_t out = 0;
        detail::check_batch(gsc_test_aabb_aabb(0, 1, a, b, &out));
        return out != 0;
    }
};
inline bool Sphere::test_sphere(const Sphere &other) const
{
    const float a[4] = { center.x, center.y, center.z, radius }, b[4] = { other.center.x, other.center.y, other.center.z, other.radius };
    uint8_t out = 0;
    detail::check_batch(gsc_test_sphere_sphere(0, 1, a, b, &out));
    return out != 0;
}
inline bool Sphere::test_obb(const OBB &obb) const
{
    const float s[4] = { center.x, center.y, center.z, radius };
    float o[16];
    obb.pack(o);
    uint8_t out = 0;
    detail::check_batch(gsc_test_sphere_obb(0, 1, s, o, &out));
    return out != 0;
}
inline bool OBB::test_obb(const OBB &other) const
{
    float a[16], b[16];
    pack(a);
    other.pack(b);
    uint8_t out = 0;
    detail::check_batch(gsc_test_obb_obb(0, 1, a, b, &out));
    return out != 0;
}

// ---- enum Collider (mod.rs:139-183) -----------------------------------------------------------------------------
struct Collider {
    enum Kind : uint8_t { kSphere = GSC_SPHERE, kBox = GSC_BOX, kMesh = GSC_MESH } kind = kSphere;
    Vector3 offset;
    float radius = 0;  // Sphere: NOT scaled by the transform (mod.rs:146-152)
    Vector3 widths;    // Box: multiplied by the derived scale (mod.rs:167-175)
    static Collider Sphere(Vector3 offset, float radius) { Collider c; c.kind = kSphere; c.offset = offset; c.radius = radius; return c; }
    static Collider Box(Vector3 offset, Vector3 widths) { Collider c; c.kind = kBox; c.offset = offset; c.widths = widths; return c; }
    static Collider Mesh() { Collider c; c.kind = kMesh; return c; }
};

// ---- the part of TransformManager the path reads: derived position / rotation / scale per entity ------------------
struct Transform {  // component/transform.rs:441-459 (roots: derived == local)
    Point position;
    Quaternion rotation;
    Vector3 scale{ 1, 1, 1 };
    Point position_derived() const { return position; }
    Quaternion rotation_derived() const { return rotation; }
    Vector3 scale_derived() const { return scale; }
};
class TransformManager {
public:
    Transform &assign(Entity entity) { return transforms_[entity]; }
    const Transform *get(Entity entity) const
    {
        auto it = transforms_.find(entity);
        return it == transforms_.end() ? nullptr : &it->second;
    }
    Transform *get_mut(Entity entity)
    {
        auto it = transforms_.find(entity);
        return it == transforms_.end() ? nullptr : &it->second;
    }

private:
    std::unordered_map<Entity, Transform> transforms_;
};

// ---- CollisionCallbackManager (mod.rs:627-708) ------------------------------------------------------------------
// CollisionCallback::invoke(&mut self, scene, first, others) (mod.rs:613-621); the scene argument is the caller's to capture
using CollisionCallback = std::function<void(Entity first, const Entity *others, size_t n_others)>;
using CallbackId = size_t;

class CollisionCallbackManager {
public:
    CallbackId register_callback(CollisionCallback cb)  // :639-642
    {
        callbacks_.push_back(std::move(cb));
        return callbacks_.size() - 1;
    }
    void assign(Entity entity, CallbackId id)  // :644-666
    {
        if (id >= callbacks_.size()) throw Panic("Cannot assign collision callback that has not been registered");
        entity_callbacks_[entity].push_back(id);
    }
    void unregister_all(Entity entity) { entity_callbacks_.erase(entity); }  // :668-670
    // process_collisions (:674-708) over the per-entity lists the device built (gsc_build_adjacency): entity i of
    // `entity` collides with other[offset[i] .. offset[i+1])
    void process_collisions(const std::vector<uint32_t> &entity, const std::vector<uint32_t> &offset, const std::vector<uint32_t> &other)
    {
        for (size_t i = 0; i < entity.size(); ++i) {
            auto it = entity_callbacks_.find(entity[i]);
            if (it == entity_callbacks_.end()) continue;
            for (CallbackId id : it->second) callbacks_[id](entity[i], other.data() + offset[i], offset[i + 1] - offset[i]);
        }
    }

private:
    std::vector<CollisionCallback> callbacks_;
    std::unordered_map<Entity, std::vector<CallbackId>> entity_callbacks_;
};

// ---- ColliderManager (mod.rs:196-286) + the dense order of BoundingVolumeManager (bounding_volume.rs:17-113) ------
class ColliderManager {
public:
    // mod.rs:221-223; panics on a duplicate (struct_component_manager.rs:84-87)
    void assign(Entity entity, Collider collider)
    {
        if (indices_.count(entity)) throw Panic("Component already assign to entity " + std::to_string(entity));
        indices_[entity] = entities_.size();
        entities_.push_back(entity);
        colliders_.push_back(collider);
    }
    const Collider *get(Entity entity) const  // mod.rs:278-280
    {
        auto it = indices_.find(entity);
        return it == indices_.end() ? nullptr : &colliders_[it->second];
    }
    void destroy(Entity entity)  // mod.rs:282-285: marks only; CollisionSystem::update removes at its end
    {
        if (indices_.count(entity) && marked_.insert(entity).second) marked_order_.push_back(entity);
    }
    CallbackId register_callback(CollisionCallback cb) { return callback_manager.register_callback(std::move(cb)); }  // mod.rs:237-239
    void assign_callback(Entity entity, CallbackId id) { callback_manager.assign(entity, id); }                        // mod.rs:252-254
    size_t len() const { return entities_.size(); }
    // iteration order = dense order = index order (struct_component_manager.rs:112-117); it orients every output tuple
    const std::vector<Entity> &entities() const { return entities_; }
    const std::vector<Collider> &components() const { return colliders_; }

    CollisionCallbackManager callback_manager;

private:
    friend class CollisionSystem;
    // BoundingVolumeManager::destroy_immediate (bounding_volume.rs:82-104): swap_remove; returns the index removed
    uint32_t destroy_immediate(Entity entity)
    {
        auto it = indices_.find(entity);
        if (it == indices_.end()) throw Panic("destroy_immediate: no collider for entity " + std::to_string(entity));
        const size_t index = it->second;
        indices_.erase(it);
        entities_[index] = entities_.back();
        colliders_[index] = colliders_.back();
        entities_.pop_back();
        colliders_.pop_back();
        if (index != entities_.size()) indices_[entities_[index]] = index;
        return (uint32_t)index;
    }
    std::vector<Entity> entities_;
    std::vector<Collider> colliders_;
    std::unordered_map<Entity, size_t> indices_;
    std::unordered_set<Entity> marked_;
    std::vector<Entity> marked_order_;
    size_t on_device_ = 0;  // how many of the dense elements the device arrays already hold
};

// ---- GridCollisionSystem (grid_collision.rs:119-285) --------------------------------------------------------------
class GridCollisionSystem {
public:
    using CollisionSet = std::unordered_set<std::pair<Entity, Entity>, detail::PairHash>;
    // same name and tuple order as grid_collision.rs:119: (entity of the later volume, entity of the earlier one)
    CollisionSet collisions;
    float cell_size = 0;

    // GridCollisionSystem::new() (:122-216); max_colliders replaces the engine's MAX_COMPONENTS cap
    explicit GridCollisionSystem(uint32_t max_colliders = 1000, int device = 0, uint32_t flags = 0)
    {
        gsc_config cfg{};
        cfg.device = device;
        cfg.max_colliders = max_colliders;
        cfg.flags = flags;
        detail::check(nullptr, gsc_create(&cfg, &ctx_));
    }
    ~GridCollisionSystem() { gsc_destroy(ctx_); }
    GridCollisionSystem(const GridCollisionSystem &) = delete;
    GridCollisionSystem &operator=(const GridCollisionSystem &) = delete;
    gsc_ctx *ctx() const { return ctx_; }

    // update(&mut self, &BoundingVolumeManager) (:218-285) on the frame's transforms, already on their way to the device
    void update()
    {
        gsc_step_out out{};
        detail::check(ctx_, gsc_step(ctx_, &out));
        cell_size = out.cell_size;
        pairs_.resize(2 * (size_t)out.n_pairs);
        uint64_t n = 0;
        detail::check(ctx_, gsc_download_pairs(ctx_, pairs_.data(), out.n_pairs, 0, &n));
        collisions.clear();  // :221
        for (size_t i = 0; i < (size_t)n; ++i) collisions.insert({ pairs_[2 * i], pairs_[2 * i + 1] });
    }

private:
    gsc_ctx *ctx_ = nullptr;
    std::vector<uint32_t> pairs_;
};

// ---- the same GridCollisionSystem over ALL GPUs of a box (gsc_group_*): x-slab partition, one context per device ----
// The reference's own decomposition is 8 octant work units on 8 threads merged into one set (grid_collision.rs:60-86,
// 161-193, 262-272); here the units are slabs on GPUs and `collisions` is the union of their lists.  Dense-array
// interface: colliders and derived transforms in BoundingVolumeManager.components order.
class GroupGridCollisionSystem {
public:
    GridCollisionSystem::CollisionSet collisions;
    float cell_size = 0;
    uint32_t ghosts = 0;  // halo colliders exchanged in the last update

    GroupGridCollisionSystem(const std::vector<int32_t> &devices, uint32_t max_colliders_per_device, uint32_t flags = 0)
    {
        gsc_config cfg{};
        cfg.max_colliders = max_colliders_per_device;
        cfg.flags = flags;
        const int rc = gsc_group_create((uint32_t)devices.size(), devices.data(), &cfg, &group_);
        if (rc != GSC_OK) throw Panic(std::string("gsc_group_create: ") + gsc_last_error(nullptr));
    }
    ~GroupGridCollisionSystem() { gsc_group_destroy(group_); }
    GroupGridCollisionSystem(const GroupGridCollisionSystem &) = delete;
    GroupGridCollisionSystem &operator=(const GroupGridCollisionSystem &) = delete;

    // the dense collider arrays (ColliderManager order) and the positions the slabs are cut from
    void set_colliders(const std::vector<Entity> &entities, const std::vector<Collider> &colliders, const std::vector<float> &pos3)
    {
        std::vector<uint8_t> kind;
        std::vector<float> off, size;
        for (const Collider &c : colliders) {
            kind.push_back((uint8_t)c.kind);
            off.insert(off.end(), { c.offset.x, c.offset.y, c.offset.z });
            if (c.kind == Collider::kSphere) size.insert(size.end(), { c.radius, 0.0f, 0.0f });
            else size.insert(size.end(), { c.widths.x, c.widths.y, c.widths.z });
        }
        n_ = (uint32_t)entities.size();
        check(gsc_group_upload_colliders(group_, n_, entities.data(), kind.data(), off.data(), size.data(), pos3.data()));
    }
    // bvh_update + update over the frame's derived transforms (dense order)
    void update(const std::vector<float> &pos3, const std::vector<float> &quat4, const std::vector<float> &scale3)
    {
        check(gsc_group_update_transforms(group_, n_, pos3.data(), quat4.data(), scale3.data()));
        gsc_step_out out{};
        check(gsc_group_step(group_, &out));
        cell_size = out.cell_size;
        ghosts = out.n_ghost;
        pairs_.resize(2 * (size_t)out.n_pairs);
        uint64_t n = 0;
        check(gsc_group_download_pairs(group_, pairs_.data(), out.n_pairs, 0, &n));
        collisions.clear();
        for (size_t i = 0; i < (size_t)n; ++i) collisions.insert({ pairs_[2 * i], pairs_[2 * i + 1] });
    }
    // migrants: re-cut the slabs from the current positions (dense order) and the last step's work
    void rebalance(const std::vector<float> &pos3) { check(gsc_group_rebalance(group_, n_, pos3.data())); }
    // temporal coherence: the slabs keep their colliders in the cell order of the last update
    void reorder() { check(gsc_group_reorder(group_)); }

private:
    void check(int rc) const
    {
        if (rc != GSC_OK) throw Panic(std::string("gunship_collision group: ") + gsc_group_last_error(group_));
    }
    gsc_group *group_ = nullptr;
    uint32_t n_ = 0;
    std::vector<uint32_t> pairs_;
};

// ---- CollisionSystem (mod.rs:577-611) -----------------------------------------------------------------------------
class CollisionSystem {
public:
    explicit CollisionSystem(uint32_t max_colliders = 1000, int device = 0) : grid_system(max_colliders, device) {}
    GridCollisionSystem grid_system;

    // System::update(&mut self, scene, delta) (mod.rs:590-611)
    void update(ColliderManager &cm, const TransformManager &tm)
    {
        gsc_ctx *ctx = grid_system.ctx();
        // bvh_update meets entities it has not seen yet and pushes their volumes (bounding_volume.rs:394-407)
        push_assigned(ctx, cm);
        // bvh_update reads every collider's transform (bounding_volume.rs:358-362); a missing one is an unwrap() panic
        // The ONE pass over the frame's transforms goes straight into the library's pinned staging bank (positions in
        // dense order; rotation / scale packed over the boxes -- spheres ignore them, mod.rs:313-317): no pageable
        // intermediate vectors, no second (staged) copy inside the driver.
        const size_t n = cm.entities_.size();
        float *pos = nullptr, *quat = nullptr, *scale = nullptr;
        detail::check(ctx, gsc_map_transform_staging(ctx, bank_, &pos, &quat, &scale));
        size_t k = 0;  // box slot
        for (size_t i = 0; i < n; ++i) {
            const Transform *t = tm.get(cm.entities_[i]);
            if (!t) throw Panic("called `Option::unwrap()` on a `None` value: entity " + std::to_string(cm.entities_[i]) + " has a collider but no transform");
            const Point p = t->position_derived();
            pos[3 * i] = p.x; pos[3 * i + 1] = p.y; pos[3 * i + 2] = p.z;
            if (cm.colliders_[i].kind == Collider::kBox) {
                const Quaternion q = t->rotation_derived();
                const Vector3 s = t->scale_derived();
                quat[4 * k] = q.x; quat[4 * k + 1] = q.y; quat[4 * k + 2] = q.z; quat[4 * k + 3] = q.w;
                scale[3 * k] = s.x; scale[3 * k + 1] = s.y; scale[3 * k + 2] = s.z;
                ++k;
            }
        }
        detail::check(ctx, gsc_commit_transform_staging(ctx, bank_, (uint32_t)n, (uint32_t)k, 1, 1));
        bank_ ^= 1u;
        grid_system.update();  // :596-599

        // callback_manager.process_collisions(scene, &self.grid_system.collisions) (:601): the per-entity lists
        // (:687-690) come from the device as CSR
        gsc_adjacency adj{};
        detail::check(ctx, gsc_build_adjacency(ctx, &adj));
        adj_entity_.resize(adj.n_entities); adj_offset_.resize(adj.n_entities + 1); adj_other_.resize(adj.n_entries);
        detail::check(ctx, gsc_download_adjacency(ctx, adj_entity_.data(), adj_offset_.data(), adj_other_.data(), adj.n_entities, adj.n_entries));
        cm.callback_manager.process_collisions(adj_entity_, adj_offset_, adj_other_);

        // Run cleanup of marked components (:603-609)
        std::vector<uint32_t> removed;
        for (Entity entity : cm.marked_order_) {
            cm.callback_manager.unregister_all(entity);
            removed.push_back(cm.destroy_immediate(entity));
        }
        cm.marked_.clear();
        cm.marked_order_.clear();
        if (!removed.empty()) {
            detail::check(ctx, gsc_remove_colliders(ctx, (uint32_t)removed.size(), removed.data()));
            cm.on_device_ = cm.entities_.size();
        }
    }

private:
    static void push_assigned(gsc_ctx *ctx, ColliderManager &cm)
    {
        const size_t lo = cm.on_device_, hi = cm.entities_.size();
        if (hi == lo && hi != 0) return;
        std::vector<uint8_t> kind;
        std::vector<float> off, size;
        for (size_t i = lo; i < hi; ++i) {
            const Collider &c = cm.colliders_[i];
            kind.push_back((uint8_t)c.kind);  // Mesh -> GSC_ERR_KIND, as unimplemented!() (mod.rs:331)
            off.insert(off.end(), { c.offset.x, c.offset.y, c.offset.z });
            if (c.kind == Collider::kSphere) size.insert(size.end(), { c.radius, 0.0f, 0.0f });
            else size.insert(size.end(), { c.widths.x, c.widths.y, c.widths.z });
        }
        if (lo == 0)
            detail::check(ctx, gsc_upload_colliders(ctx, (uint32_t)hi, cm.entities_.data(), kind.data(), off.data(), size.data(), nullptr));
        else
            detail::check(ctx, gsc_append_colliders(ctx, (uint32_t)(hi - lo), cm.entities_.data() + lo, kind.data(), off.data(), size.data()));
        cm.on_device_ = hi;
    }
    uint32_t bank_ = 0;  // pinned staging bank of the next frame (two alternate)
    std::vector<uint32_t> adj_entity_, adj_offset_, adj_other_;
};

}  // namespace gunship

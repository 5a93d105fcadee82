// Drives the C++ host mirror (ColliderManager / TransformManager / CollisionSystem) over a scene file written by
// tests/test_cpp_host.py and writes what the reference's game code would observe: the collision set of every frame
// and what the collision callbacks were told.  Frame script: frame 0 plain; before frame 1 every `destroy_stride`-th
// entity is destroyed (marked: it still collides in frame 1, the cleanup runs at the end of that update,
// mod.rs:603-609) and `n_late` more colliders are assigned; frame 2 runs on what is left.
//
// scene file (little endian): u32 n, n_late, n_frames, destroy_stride; then for each of the n + n_late colliders
// u32 entity, u32 kind, f32 offset[3], f32 size[3]; then per frame and collider f32 pos[3], quat[4], scale[3].
// output: per frame u64 n_pairs, then n_pairs x u64 (first << 32 | second) ascending; then u64 callback_calls,
// u64 callback_others (sum of list lengths), u64 callback_checksum.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "gunship_collision.hpp"

using namespace gunship;

template <class T>
static T rd(FILE *f)
{
    T v;
    if (std::fread(&v, sizeof v, 1, f) != 1) throw Panic("short scene file");
    return v;
}

int main(int argc, char **argv)
{
    if (argc != 3) { std::printf("usage: scene_runner scene.bin out.bin\n"); return 2; }
    try {
        FILE *f = std::fopen(argv[1], "rb");
        if (!f) throw Panic("cannot open scene file");
        const uint32_t n = rd<uint32_t>(f), n_late = rd<uint32_t>(f), n_frames = rd<uint32_t>(f), stride = rd<uint32_t>(f);
        struct Def { Entity entity; Collider collider; };
        std::vector<Def> defs(n + n_late);
        for (auto &d : defs) {
            d.entity = rd<uint32_t>(f);
            const uint32_t kind = rd<uint32_t>(f);
            float o[3], s[3];
            for (float &x : o) x = rd<float>(f);
            for (float &x : s) x = rd<float>(f);
            d.collider = kind == GSC_SPHERE ? Collider::Sphere(Vector3{ o[0], o[1], o[2] }, s[0]) : Collider::Box(Vector3{ o[0], o[1], o[2] }, Vector3{ s[0], s[1], s[2] });
        }
        std::vector<std::vector<Transform>> frames(n_frames, std::vector<Transform>(n + n_late));
        for (auto &fr : frames)
            for (auto &t : fr) {
                t.position.x = rd<float>(f); t.position.y = rd<float>(f); t.position.z = rd<float>(f);
                t.rotation.x = rd<float>(f); t.rotation.y = rd<float>(f); t.rotation.z = rd<float>(f); t.rotation.w = rd<float>(f);
                t.scale.x = rd<float>(f); t.scale.y = rd<float>(f); t.scale.z = rd<float>(f);
            }
        std::fclose(f);

        ColliderManager colliders;
        TransformManager transforms;
        CollisionSystem system(n + n_late);
        uint64_t calls = 0, others_total = 0, checksum = 0;
        const CallbackId cb = colliders.register_callback([&](Entity first, const Entity *others, size_t k) {
            ++calls;
            others_total += k;
            for (size_t i = 0; i < k; ++i) checksum += (uint64_t)first * 1000003ull + others[i];
        });
        for (uint32_t i = 0; i < n; ++i) {
            colliders.assign(defs[i].entity, defs[i].collider);
            transforms.assign(defs[i].entity);
            if (i % 3 == 0) colliders.assign_callback(defs[i].entity, cb);
        }
        bool dup_panicked = false;
        try { colliders.assign(defs[0].entity, defs[0].collider); } catch (const Panic &) { dup_panicked = true; }  // struct_component_manager.rs:84-87
        if (!dup_panicked) throw Panic("duplicate assign did not panic");

        FILE *o = std::fopen(argv[2], "wb");
        if (!o) throw Panic("cannot open output file");
        for (uint32_t fr = 0; fr < n_frames; ++fr) {
            if (fr == 1) {
                for (uint32_t i = 0; i < n; i += stride) colliders.destroy(defs[i].entity);
                for (uint32_t i = n; i < n + n_late; ++i) {
                    colliders.assign(defs[i].entity, defs[i].collider);
                    transforms.assign(defs[i].entity);
                }
            }
            for (uint32_t i = 0; i < n + n_late; ++i)
                if (Transform *t = transforms.get_mut(defs[i].entity)) *t = frames[fr][i];
            system.update(colliders, transforms);
            std::vector<uint64_t> packed;
            for (const auto &p : system.grid_system.collisions) packed.push_back(((uint64_t)p.first << 32) | p.second);
            std::sort(packed.begin(), packed.end());
            const uint64_t np = packed.size();
            std::fwrite(&np, 8, 1, o);
            std::fwrite(packed.data(), 8, packed.size(), o);
        }
        std::fwrite(&calls, 8, 1, o);
        std::fwrite(&others_total, 8, 1, o);
        std::fwrite(&checksum, 8, 1, o);
        std::fclose(o);
        std::printf("scene_runner: ok, %zu colliders left\n", colliders.len());
    } catch (const Panic &p) {
        std::printf("PANIC: %s\n", p.what());
        return 3;
    }
    return 0;
}

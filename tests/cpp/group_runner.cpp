// Drives gunship::GroupGridCollisionSystem (gsc_group_* behind the C++ host mirror) over a scene file written by
// tests/test_cpp_host.py: `world` slabs on the devices given on the command line (devices may repeat: members of a
// group may share a GPU), a re-balance after the middle frame, the sorted collision set of every frame written out.
//
// scene file (little endian): u32 n, n_frames; per collider u32 entity, u32 kind, f32 offset[3], f32 size[3]; per
// frame and collider f32 pos[3], quat[4], scale[3].   output: per frame u64 n_pairs, u64 ghosts, n_pairs x u64.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

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
    if (argc < 4) { std::printf("usage: group_runner scene.bin out.bin device [device ...]\n"); return 2; }
    try {
        std::vector<int32_t> devices;
        for (int i = 3; i < argc; ++i) devices.push_back(std::atoi(argv[i]));
        FILE *f = std::fopen(argv[1], "rb");
        if (!f) throw Panic("cannot open scene file");
        const uint32_t n = rd<uint32_t>(f), n_frames = rd<uint32_t>(f);
        std::vector<Entity> entities(n);
        std::vector<Collider> colliders(n);
        for (uint32_t i = 0; i < n; ++i) {
            entities[i] = rd<uint32_t>(f);
            const uint32_t kind = rd<uint32_t>(f);
            float o[3], s[3];
            for (float &x : o) x = rd<float>(f);
            for (float &x : s) x = rd<float>(f);
            colliders[i] = kind == GSC_SPHERE ? Collider::Sphere(Vector3{ o[0], o[1], o[2] }, s[0]) : Collider::Box(Vector3{ o[0], o[1], o[2] }, Vector3{ s[0], s[1], s[2] });
        }
        std::vector<std::vector<float>> pos(n_frames, std::vector<float>(3 * (size_t)n)), quat(n_frames, std::vector<float>(4 * (size_t)n)),
            scale(n_frames, std::vector<float>(3 * (size_t)n));
        for (uint32_t fr = 0; fr < n_frames; ++fr)
            for (uint32_t i = 0; i < n; ++i) {
                for (int a = 0; a < 3; ++a) pos[fr][3 * (size_t)i + a] = rd<float>(f);
                for (int a = 0; a < 4; ++a) quat[fr][4 * (size_t)i + a] = rd<float>(f);
                for (int a = 0; a < 3; ++a) scale[fr][3 * (size_t)i + a] = rd<float>(f);
            }
        std::fclose(f);

        GroupGridCollisionSystem system(devices, (uint32_t)(n / devices.size() * 2 + 65536));
        system.set_colliders(entities, colliders, pos[0]);
        FILE *o = std::fopen(argv[2], "wb");
        if (!o) throw Panic("cannot open output file");
        for (uint32_t fr = 0; fr < n_frames; ++fr) {
            system.update(pos[fr], quat[fr], scale[fr]);
            std::vector<uint64_t> packed;
            for (const auto &p : system.collisions) packed.push_back(((uint64_t)p.first << 32) | p.second);
            std::sort(packed.begin(), packed.end());
            const uint64_t np = packed.size(), gh = system.ghosts;
            std::fwrite(&np, 8, 1, o);
            std::fwrite(&gh, 8, 1, o);
            std::fwrite(packed.data(), 8, packed.size(), o);
            if (fr == n_frames / 2) system.rebalance(pos[fr]);
        }
        std::fclose(o);
        std::printf("group_runner: ok, %zu devices, cell %.7f\n", devices.size(), system.cell_size);
    } catch (const Panic &p) {
        std::printf("PANIC: %s\n", p.what());
        return 3;
    }
    return 0;
}

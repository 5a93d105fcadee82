// C++ counterpart of the reference's benches/grid_collision.rs `collision_bench` (:56-87) and of the shape-test
// benches of benches/collision.rs, over the C++ host mirror (host/gunship_collision.hpp): 1000 sphere colliders of
// radius 0.5 at x, y = u * 10 - 5, z = 0, a callback on every entity, then `do_collision_update` in a loop.
//
//   g++ -std=c++17 -O2 -I include -I host benches/grid_collision.cpp -L gunship_rs_b200 -lgunship_collision
//       -Wl,-rpath,$PWD/gunship_rs_b200 -o grid_collision_bench ; ./grid_collision_bench [colliders] [iterations]
//
// At this size the step is launch-latency bound (19 kernels on ~1000 colliders): it is the reference's CPU-sized
// workload (BASELINE.json configs[0]); the throughput workloads live in bench.py.
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "gunship_collision.hpp"

using namespace gunship;

// SplitMix64: the unseeded rand::random::<f32>() of the reference, made reproducible
static uint64_t g_state = 1;
static float random01()
{
    uint64_t z = (g_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}
static float random_range(float lo, float hi) { return random01() * (hi - lo) + lo; }

int main(int argc, char **argv)
{
    const uint32_t n = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 1000;
    const int iters = argc > 2 ? std::atoi(argv[2]) : 200;
    try {
        ColliderManager colliders;
        TransformManager transforms;
        CollisionSystem system(n);
        uint64_t callback_calls = 0;
        const CallbackId callback = colliders.register_callback([&](Entity, const Entity *, size_t) { ++callback_calls; });
        for (Entity entity = 1; entity <= n; ++entity) {                 // grid_collision.rs:69-81
            Transform &t = transforms.assign(entity);
            t.position.x = random01() * 10.0f - 5.0f;
            t.position.y = random01() * 10.0f - 5.0f;
            t.position.z = 0.0f;
            colliders.assign(entity, Collider::Sphere(Vector3{}, 0.5f));
            colliders.assign_callback(entity, callback);
        }
        system.update(colliders, transforms);                             // warm-up (uploads the colliders)
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < iters; ++i) system.update(colliders, transforms);  // bencher.iter(|| do_collision_update(..)) :84-86
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / iters;
        std::printf("collision_bench: %u colliders, %zu collisions, cell %.7f, %.1f us/iter, %llu callbacks\n", n,
                    system.grid_system.collisions.size(), system.grid_system.cell_size, us, (unsigned long long)callback_calls);

        // benches/collision.rs: 10 x 10 shape tests per iteration
        std::vector<Sphere> spheres(10);
        std::vector<OBB> obbs(10);
        for (auto &s : spheres) {
            s.center.x = random_range(-5, 5); s.center.y = random_range(-5, 5); s.center.z = random_range(-5, 5);
            s.radius = random_range(0.5f, 2.0f);
        }
        for (auto &o : obbs) {
            o.center.x = random_range(-5, 5); o.center.y = random_range(-5, 5); o.center.z = random_range(-5, 5);
            o.orientation = Matrix3::rotation(random_range(0, 3.14159265f), random_range(0, 3.14159265f), random_range(0, 3.14159265f));
            o.half_widths = Vector3{ random_range(0.5f, 2.0f), random_range(0.5f, 2.0f), random_range(0.5f, 2.0f) };
        }
        int hits = 0;
        for (const auto &l : spheres) for (const auto &r : spheres) hits += l.test_sphere(r);   // sphere_sphere_x100
        for (const auto &l : obbs) for (const auto &r : obbs) hits += l.test_obb(r);            // obb_obb_x100
        for (const auto &l : spheres) for (const auto &r : obbs) hits += l.test_obb(r);         // sphere_obb_x100
        std::printf("shape tests: %d of 300 hit\n", hits);
    } catch (const Panic &p) {
        std::printf("PANIC: %s\n", p.what());
        return 3;
    }
    return 0;
}

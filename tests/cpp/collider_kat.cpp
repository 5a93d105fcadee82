// The reference's own unit tests of the collider module (src/component/collider/mod.rs:711-782), written against
// the C++ host mirror the way the Rust tests are written against the Rust types.  Needs a CUDA device: the shape
// tests run on the GPU (there is no CPU fallback); without one the first call throws gunship::Panic.
#include <cstdio>

#include "gunship_collision.hpp"

using namespace gunship;

static int failures = 0;
#define CHECK(cond)                                                      \
    do {                                                                 \
        if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

// mod.rs:713-718: every assertion in both argument orders
#define sphere_test(lhs, rhs, expected)           \
    do {                                          \
        CHECK((rhs).test_sphere(lhs) == (expected)); \
        CHECK((lhs).test_sphere(rhs) == (expected)); \
    } while (0)
#define obb_test(lhs, rhs, expected)           \
    do {                                       \
        CHECK((rhs).test_obb(lhs) == (expected)); \
        CHECK((lhs).test_obb(rhs) == (expected)); \
    } while (0)

static Point pt(float x, float y, float z) { Point p; p.x = x; p.y = y; p.z = z; return p; }

static void sphere_sphere_tests()  // mod.rs:711-755
{
    Sphere unit{ pt(0, 0, 0), 1.0f };
    Sphere far_unit{ pt(2, 2, 2), 1.0f };
    Sphere odd_radius{ pt(1.7f, 1.7f, 1.7f), 0.3f };
    Sphere unit_edge{ pt(2, 0, 0), 1.0f };
    sphere_test(unit, unit, true);
    sphere_test(far_unit, far_unit, true);
    sphere_test(odd_radius, odd_radius, true);
    sphere_test(unit, far_unit, false);
    sphere_test(unit, odd_radius, false);
    sphere_test(far_unit, odd_radius, true);
    sphere_test(unit, unit_edge, true);  // exact touch counts (mod.rs:388, :750)
}

static void obb_obb_tests()  // mod.rs:757-782
{
    OBB unit{ pt(0, 0, 0), Matrix3::identity(), Vector3{ 0.5f, 0.5f, 0.5f } };
    OBB rot_z{ pt(1, 0, 0), Matrix3::rotation(0.0f, 0.0f, 3.14159265358979323846f * 0.25f), Vector3{ 0.5f, 0.5f, 0.5f } };
    obb_test(unit, unit, true);
    obb_test(unit, rot_z, true);
    // extra negatives (not from the reference)
    OBB far = rot_z;
    far.center = pt(3, 0, 0);
    obb_test(unit, far, false);
}

static void sphere_obb_and_aabb()
{
    OBB unit{ pt(0, 0, 0), Matrix3::identity(), Vector3{ 0.5f, 0.5f, 0.5f } };
    Sphere touching{ pt(1.5f, 0, 0), 1.0f };   // |p - closest|^2 == r*r exactly: strict, no epsilon (mod.rs:393)
    Sphere inside{ pt(0.25f, 0, 0), 0.1f };
    Sphere away{ pt(3, 3, 3), 1.0f };
    CHECK(touching.test_obb(unit) == false);
    CHECK(inside.test_obb(unit) == true);
    CHECK(away.test_obb(unit) == false);
    AABB a{ pt(0, 0, 0), pt(1, 1, 1) }, b{ pt(1, 1, 1), pt(2, 2, 2) }, c{ pt(1.5f, 0, 0), pt(2, 1, 1) };
    CHECK(a.test_aabb(b) == true);   // inclusive (bounding_volume.rs:415-418)
    CHECK(a.test_aabb(c) == false);
}

int main()
{
    try {
        sphere_sphere_tests();
        obb_obb_tests();
        sphere_obb_and_aabb();
    } catch (const Panic &p) {
        std::printf("PANIC: %s\n", p.what());
        return 3;
    }
    std::printf("%s\n", failures ? "collider_kat: FAILED" : "collider_kat: ok");
    return failures ? 1 : 0;
}

// gsc_math.cuh -- device-side f32 arithmetic of the collision path, bit-exact with the reference.
//
// Every expression below is written with the explicit round-to-nearest intrinsics
// (__fmul_rn / __fadd_rn / __fsub_rn / __fdiv_rn), which the compiler never contracts into FMA,
// so the result equals the Rust source's separate-multiply-and-add evaluation regardless of
// -fmad.  Association is left-to-right exactly as the reference writes it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gsc {

#define GSC_DEV __device__ __forceinline__

constexpr float kEpsilon = 1e-6f;  // lib/polygon_math/src/lib.rs:24

GSC_DEV float mul(float a, float b) { return __fmul_rn(a, b); }
GSC_DEV float add(float a, float b) { return __fadd_rn(a, b); }
GSC_DEV float sub(float a, float b) { return __fsub_rn(a, b); }

// Random-access 16 B load for gathers of 32 B / 64 B records: read-only path with the smallest L2 prefetch
// size PTX offers.  By default a missing sector makes L2 fetch its whole 128 B line from HBM (ncu: 3.8
// sectors per request on the cell-order gather), which quadruples the DRAM bytes of a 32 B gather.
GSC_DEV float4 ldg_gather(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.nc.L2::64B.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

GSC_DEV uint32_t ldg_gather_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// 32 B in ONE access.  sm_100 has 256-bit global loads and stores (PTX ld/st.global.v8.b32 -> LDG.E.256 / STG.E.256):
// a 32 B volume record is one instruction and one L1 request instead of two, a 64 B box record two instead of four.
// The record arrays are where the load / store unit's wavefronts go in bvh_update (strided 16 B stores) and in the
// SAT stage (random 32 / 64 B gathers), so halving them is worth more than any re-ordering of the same 16 B accesses.
// The address must be 32 B aligned: rec + 2 * i and obb + 4 * i of arrays that come straight from cudaMalloc.
// -DGSC_NO_LDST256 keeps the 128-bit pairs (A/B measurements).
struct F8 {
    float4 lo, hi;
};
#ifndef GSC_NO_LDST256
GSC_DEV F8 ld256_gather(const float4 *p)  // random access: read-only path, 64 B L2 prefetch granule (see ldg_gather)
{
    F8 v;
    asm volatile("ld.global.nc.L2::64B.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v.lo.x), "=f"(v.lo.y), "=f"(v.lo.z), "=f"(v.lo.w), "=f"(v.hi.x), "=f"(v.hi.y), "=f"(v.hi.z), "=f"(v.hi.w)
                 : "l"(p));
    return v;
}
GSC_DEV F8 ld256(const float4 *p)  // streaming access, read-only path
{
    F8 v;
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v.lo.x), "=f"(v.lo.y), "=f"(v.lo.z), "=f"(v.lo.w), "=f"(v.hi.x), "=f"(v.hi.y), "=f"(v.hi.z), "=f"(v.hi.w)
                 : "l"(p));
    return v;
}
GSC_DEV void st256(float4 *p, float4 a, float4 b)
{
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
                 "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
                 : "memory");
}
#else
GSC_DEV F8 ld256_gather(const float4 *p) { return F8{ ldg_gather(p), ldg_gather(p + 1) }; }
GSC_DEV F8 ld256(const float4 *p) { return F8{ __ldg(p), __ldg(p + 1) }; }
GSC_DEV void st256(float4 *p, float4 a, float4 b)
{
    p[0] = a;
    p[1] = b;
}
#endif

// vector.rs:190-196  Dot: x*x' + y*y' + z*z'
GSC_DEV float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    return add(add(mul(ax, bx), mul(ay, by)), mul(az, bz));
}

// World-space oriented box: OBB (mod.rs:397-402).  m is the row-major Matrix3 (matrix.rs:293-295);
// the box axes are its COLUMNS (mod.rs:419,555).
struct Obb {
    float cx, cy, cz;
    float m[3][3];
    float hx, hy, hz;
};

// 64-byte storage record of an Obb: 4 x float4.
//   q0 = (cx, cy, cz, m00)  q1 = (m01, m02, m10, m11)  q2 = (m12, m20, m21, m22)  q3 = (hx, hy, hz, entity bits)
GSC_DEV void obb_store(float4 *dst, const Obb &o, uint32_t entity)
{
    st256(dst, make_float4(o.cx, o.cy, o.cz, o.m[0][0]), make_float4(o.m[0][1], o.m[0][2], o.m[1][0], o.m[1][1]));
    st256(dst + 2, make_float4(o.m[1][2], o.m[2][0], o.m[2][1], o.m[2][2]), make_float4(o.hx, o.hy, o.hz, __uint_as_float(entity)));
}

GSC_DEV Obb obb_load(const float4 *src, uint32_t *entity)
{
    const F8 h0 = ld256_gather(src), h1 = ld256_gather(src + 2);
    const float4 q0 = h0.lo, q1 = h0.hi, q2 = h1.lo, q3 = h1.hi;
    Obb o;
    o.cx = q0.x; o.cy = q0.y; o.cz = q0.z;
    o.m[0][0] = q0.w; o.m[0][1] = q1.x; o.m[0][2] = q1.y;
    o.m[1][0] = q1.z; o.m[1][1] = q1.w; o.m[1][2] = q2.x;
    o.m[2][0] = q2.y; o.m[2][1] = q2.z; o.m[2][2] = q2.w;
    o.hx = q3.x; o.hy = q3.y; o.hz = q3.z;
    if (entity) *entity = __float_as_uint(q3.w);
    return o;
}

// matrix.rs:392-402  From<Orientation> for Matrix3, q = (x,y,z,w)
GSC_DEV void quat_to_mat3(float x, float y, float z, float w, float m[3][3])
{
    const float ww = mul(w, w), xx = mul(x, x), yy = mul(y, y), zz = mul(z, z);
    const float x2 = mul(2.0f, x), y2 = mul(2.0f, y), w2 = mul(2.0f, w);
    m[0][0] = sub(sub(add(ww, xx), yy), zz);
    m[0][1] = sub(mul(x2, y), mul(w2, z));
    m[0][2] = add(mul(x2, z), mul(w2, y));
    m[1][0] = add(mul(x2, y), mul(w2, z));
    m[1][1] = sub(add(sub(ww, xx), yy), zz);
    m[1][2] = sub(mul(y2, z), mul(w2, x));
    m[2][0] = sub(mul(x2, z), mul(w2, y));
    m[2][1] = add(mul(y2, z), mul(w2, x));
    m[2][2] = add(sub(sub(ww, xx), yy), zz);
}

// bounding_volume.rs:161-327  AABB of an OBB: 8 vertices centre + M.(half*sign) in the reference's
// sign order, running min/max with `if v < min {..} else if v > max {..}`.
GSC_DEV void obb_aabb(const Obb &o, float mn[3], float mx[3])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float sx = (k & 4) ? -1.0f : 1.0f, sy = (k & 2) ? -1.0f : 1.0f, sz = (k & 1) ? -1.0f : 1.0f;
        const float vx = mul(o.hx, sx), vy = mul(o.hy, sy), vz = mul(o.hz, sz);
        // matrix.rs:453-463  M.v row by row
        const float px = add(o.cx, add(add(mul(o.m[0][0], vx), mul(o.m[0][1], vy)), mul(o.m[0][2], vz)));
        const float py = add(o.cy, add(add(mul(o.m[1][0], vx), mul(o.m[1][1], vy)), mul(o.m[1][2], vz)));
        const float pz = add(o.cz, add(add(mul(o.m[2][0], vx), mul(o.m[2][1], vy)), mul(o.m[2][2], vz)));
        if (k == 0) {
            mn[0] = mx[0] = px; mn[1] = mx[1] = py; mn[2] = mx[2] = pz;
        } else {
            if (px < mn[0]) mn[0] = px; else if (px > mx[0]) mx[0] = px;
            if (py < mn[1]) mn[1] = py; else if (py > mx[1]) mx[1] = py;
            if (pz < mn[2]) mn[2] = pz; else if (pz > mx[2]) mx[2] = pz;
        }
    }
}

// The same AABB without walking the vertices.  With a = fl(m0*hx), b = fl(m1*hy), d = fl(m2*hz) a vertex coordinate
// is fl(c + fl(fl(+-a +- b) +- d)) (multiplying by the sign -1 is exact and rounding is symmetric).  Rounding is
// monotone, so over the 8 sign choices the largest inner sum is S = fl(fl(|a| + |b|) + |d|), the smallest is -S, and
// the running min / max of the reference ends at fl(c - S) / fl(c + S) -- bit for bit, including the sign of a zero
// result, as long as S != 0 (then a zero bound can only come from c + s = 0 with s != 0, which rounds to +0 either
// way; `if v < min .. else if v > max` cannot skip an update because min <= max).  A degenerate box (S == 0 on an
// axis: zero half width or a zero matrix row, where the sign of a zero bound would depend on the vertex order) takes
// the literal path.  12 multiply-adds instead of 8 x 18 plus the compare chain.
GSC_DEV void obb_aabb_fast(const Obb &o, float mn[3], float mx[3])
{
    float S[3];
    const float c[3] = { o.cx, o.cy, o.cz };
#pragma unroll
    for (int r = 0; r < 3; ++r)
        S[r] = add(add(fabsf(mul(o.m[r][0], o.hx)), fabsf(mul(o.m[r][1], o.hy))), fabsf(mul(o.m[r][2], o.hz)));
    if (S[0] == 0.0f || S[1] == 0.0f || S[2] == 0.0f) {
        obb_aabb(o, mn, mx);
        return;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        mn[r] = sub(c[r], S[r]);
        mx[r] = add(c[r], S[r]);
    }
}

// bounding_volume.rs:411-419
GSC_DEV bool test_ranges(float min_a, float max_a, float min_b, float max_b)
{
    return !(min_a > max_b || min_b > max_a || max_a < min_b || max_b < min_a);
}

// bounding_volume.rs:338-343 (inclusive: touching boxes overlap)
GSC_DEV bool test_aabb(const float amn[3], const float amx[3], const float bmn[3], const float bmx[3])
{
    return test_ranges(amn[0], amx[0], bmn[0], bmx[0]) && test_ranges(amn[1], amx[1], bmn[1], bmx[1]) &&
           test_ranges(amn[2], amx[2], bmn[2], bmx[2]);
}

// mod.rs:383-389  Sphere::test_sphere
GSC_DEV bool test_sphere_sphere(float ax, float ay, float az, float ar, float bx, float by, float bz, float br)
{
    const float dx = sub(ax, bx), dy = sub(ay, by), dz = sub(az, bz);
    const float dist_sqr = add(add(mul(dx, dx), mul(dy, dy)), mul(dz, dz));
    const float rs = add(ar, br);
    const float diff = sub(dist_sqr, mul(rs, rs));
    return diff < 0.0f || fabsf(diff) < kEpsilon;
}

// mod.rs:391-394 + 549-574  Sphere::test_obb via OBB::closest_point (strict <)
GSC_DEV bool test_sphere_obb(float sx, float sy, float sz, float sr, const Obb &o)
{
    const float dx = sub(sx, o.cx), dy = sub(sy, o.cy), dz = sub(sz, o.cz);
    float rx = o.cx, ry = o.cy, rz = o.cz;
    const float h[3] = { o.hx, o.hy, o.hz };
#pragma unroll
    for (int axis = 0; axis < 3; ++axis) {
        const float ax = o.m[0][axis], ay = o.m[1][axis], az = o.m[2][axis];  // column `axis`
        float dist = dot3(dx, dy, dz, ax, ay, az);
        dist = fminf(fmaxf(dist, -h[axis]), h[axis]);  // lib.rs:41-45 clamp
        rx = add(rx, mul(ax, dist));
        ry = add(ry, mul(ay, dist));
        rz = add(rz, mul(az, dist));
    }
    const float ex = sub(sx, rx), ey = sub(sy, ry), ez = sub(sz, rz);
    const float dist_sqr = add(add(mul(ex, ex), mul(ey, ey)), mul(ez, ez));
    return dist_sqr < mul(sr, sr);
}

// mod.rs:413-546  OBB::test_obb -- `a` is self, `b` the argument.  15-axis SAT.
GSC_DEV bool test_obb_obb(const Obb &a, const Obb &b)
{
    float r[3][3], ar[3][3];
#pragma unroll
    for (int row = 0; row < 3; ++row)
#pragma unroll
        for (int col = 0; col < 3; ++col) {
            // a.col(row) . b.col(col)
            r[row][col] = dot3(a.m[0][row], a.m[1][row], a.m[2][row], b.m[0][col], b.m[1][col], b.m[2][col]);
            ar[row][col] = add(fabsf(r[row][col]), kEpsilon);
        }
    const float tx = sub(b.cx, a.cx), ty = sub(b.cy, a.cy), tz = sub(b.cz, a.cz);
    float t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)  // t * a.orientation.transpose(): row i of A^T is column i of A
        t[i] = add(add(mul(a.m[0][i], tx), mul(a.m[1][i], ty)), mul(a.m[2][i], tz));
    const float ah[3] = { a.hx, a.hy, a.hz };
    const float bh[3] = { b.hx, b.hy, b.hz };
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // L = A_i
        const float rb = add(add(mul(bh[0], ar[i][0]), mul(bh[1], ar[i][1])), mul(bh[2], ar[i][2]));
        if (fabsf(t[i]) > add(ah[i], rb)) return false;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // L = B_i
        const float ra = add(add(mul(ah[0], ar[0][i]), mul(ah[1], ar[1][i])), mul(ah[2], ar[2][i]));
        const float proj = add(add(mul(t[0], r[0][i]), mul(t[1], r[1][i])), mul(t[2], r[2][i]));
        if (fabsf(proj) > add(ra, bh[i])) return false;
    }
    float ra, rb;
    // A0 x B0
    ra = add(mul(ah[1], ar[2][0]), mul(ah[2], ar[1][0]));
    rb = add(mul(bh[1], ar[0][2]), mul(bh[2], ar[0][1]));
    if (fabsf(sub(mul(t[2], r[1][0]), mul(t[1], r[2][0]))) > add(ra, rb)) return false;
    // A0 x B1
    ra = add(mul(ah[1], ar[2][1]), mul(ah[2], ar[1][1]));
    rb = add(mul(bh[0], ar[0][2]), mul(bh[2], ar[0][0]));
    if (fabsf(sub(mul(t[2], r[1][1]), mul(t[1], r[2][1]))) > add(ra, rb)) return false;
    // A0 x B2
    ra = add(mul(ah[1], ar[2][2]), mul(ah[2], ar[1][2]));
    rb = add(mul(bh[0], ar[0][1]), mul(bh[1], ar[0][0]));
    if (fabsf(sub(mul(t[2], r[1][2]), mul(t[1], r[2][2]))) > add(ra, rb)) return false;
    // A1 x B0
    ra = add(mul(ah[0], ar[2][0]), mul(ah[2], ar[0][0]));
    rb = add(mul(bh[1], ar[1][2]), mul(bh[2], ar[1][1]));
    if (fabsf(sub(mul(t[0], r[2][0]), mul(t[2], r[0][0]))) > add(ra, rb)) return false;
    // A1 x B1
    ra = add(mul(ah[0], ar[2][1]), mul(ah[2], ar[0][1]));
    rb = add(mul(bh[0], ar[1][2]), mul(bh[2], ar[1][0]));
    if (fabsf(sub(mul(t[0], r[2][1]), mul(t[2], r[0][1]))) > add(ra, rb)) return false;
    // A1 x B2
    ra = add(mul(ah[0], ar[2][2]), mul(ah[2], ar[0][2]));
    rb = add(mul(bh[0], ar[1][1]), mul(bh[1], ar[1][0]));
    if (fabsf(sub(mul(t[0], r[2][2]), mul(t[2], r[0][2]))) > add(ra, rb)) return false;
    // A2 x B0
    ra = add(mul(ah[0], ar[1][0]), mul(ah[1], ar[0][0]));
    rb = add(mul(bh[1], ar[2][2]), mul(bh[2], ar[2][1]));
    if (fabsf(sub(mul(t[1], r[0][0]), mul(t[0], r[1][0]))) > add(ra, rb)) return false;
    // A2 x B1
    ra = add(mul(ah[0], ar[1][1]), mul(ah[1], ar[0][1]));
    rb = add(mul(bh[0], ar[2][2]), mul(bh[2], ar[2][0]));
    if (fabsf(sub(mul(t[1], r[0][1]), mul(t[0], r[1][1]))) > add(ra, rb)) return false;
    // A2 x B2
    ra = add(mul(ah[0], ar[1][2]), mul(ah[1], ar[0][2]));
    rb = add(mul(bh[0], ar[2][1]), mul(bh[1], ar[2][0]));
    if (fabsf(sub(mul(t[1], r[0][2]), mul(t[0], r[1][2]))) > add(ra, rb)) return false;
    return true;
}

// ---- contacts (normal.xyz, depth) -------------------------------------------------------------------
// NOT in the reference (it only plans them, mod.rs:110-113): this is the definition BASELINE.json's
// north_star asks for, documented in include/gunship_collision.h.  `hi` is the later volume of the pair, `lo`
// the earlier one; the unit normal points from lo to hi, depth is the overlap along it.  Same f32 operation
// order as the checker in oracle/ (separate mul/add, IEEE sqrt and divide).
GSC_DEV float4 contact_sphere_sphere(float hx, float hy, float hz, float hr, float lx, float ly, float lz, float lr)
{
    const float dx = sub(hx, lx), dy = sub(hy, ly), dz = sub(hz, lz);
    const float dist = __fsqrt_rn(add(add(mul(dx, dx), mul(dy, dy)), mul(dz, dz)));
    float4 c;
    if (dist > 0.0f) { c.x = __fdiv_rn(dx, dist); c.y = __fdiv_rn(dy, dist); c.z = __fdiv_rn(dz, dist); }
    else { c.x = 1.0f; c.y = 0.0f; c.z = 0.0f; }
    c.w = sub(add(hr, lr), dist);
    return c;
}

// from OBB::closest_point (mod.rs:549-568): outside, the normal is (centre - closest point) normalised; a centre
// inside the box leaves through the face of least penetration
GSC_DEV float4 contact_sphere_obb(float sx, float sy, float sz, float sr, const Obb &o, bool sphere_is_hi)
{
    const float dx = sub(sx, o.cx), dy = sub(sy, o.cy), dz = sub(sz, o.cz);
    float rx = o.cx, ry = o.cy, rz = o.cz;
    const float h[3] = { o.hx, o.hy, o.hz };
    float proj[3];
#pragma unroll
    for (int axis = 0; axis < 3; ++axis) {
        const float ax = o.m[0][axis], ay = o.m[1][axis], az = o.m[2][axis];
        const float dist = dot3(dx, dy, dz, ax, ay, az);
        proj[axis] = dist;
        const float cl = fminf(fmaxf(dist, -h[axis]), h[axis]);
        rx = add(rx, mul(ax, cl));
        ry = add(ry, mul(ay, cl));
        rz = add(rz, mul(az, cl));
    }
    const float vx = sub(sx, rx), vy = sub(sy, ry), vz = sub(sz, rz);
    const float d2 = add(add(mul(vx, vx), mul(vy, vy)), mul(vz, vz));
    // inside <=> no projection was clamped (d2 alone is not robust: the rebuilt point differs by rounding)
    const bool inside = fabsf(proj[0]) <= h[0] && fabsf(proj[1]) <= h[1] && fabsf(proj[2]) <= h[2];
    float4 c;
    if (!inside && d2 > 0.0f) {
        const float dd = __fsqrt_rn(d2);
        c.x = __fdiv_rn(vx, dd); c.y = __fdiv_rn(vy, dd); c.z = __fdiv_rn(vz, dd);
        c.w = sub(sr, dd);
    } else {
        int best = 0;
        float pen = sub(h[0], fabsf(proj[0]));
#pragma unroll
        for (int i = 1; i < 3; ++i) {
            const float p = sub(h[i], fabsf(proj[i]));
            if (p < pen) { pen = p; best = i; }
        }
        const float pb = best == 0 ? proj[0] : (best == 1 ? proj[1] : proj[2]);
        const float sgn = pb < 0.0f ? -1.0f : 1.0f;
        c.x = mul(best == 0 ? o.m[0][0] : (best == 1 ? o.m[0][1] : o.m[0][2]), sgn);
        c.y = mul(best == 0 ? o.m[1][0] : (best == 1 ? o.m[1][1] : o.m[1][2]), sgn);
        c.z = mul(best == 0 ? o.m[2][0] : (best == 1 ? o.m[2][1] : o.m[2][2]), sgn);
        c.w = add(sr, pen);
    }
    if (!sphere_is_hi) { c.x = -c.x; c.y = -c.y; c.z = -c.z; }
    return c;
}

// least-penetration axis of the 15 SAT axes of OBB::test_obb (mod.rs:443-542), a = hi, b = lo.  Cross axes are
// normalised by |A_i x B_j| = sqrt(1 - r_ij^2) and skipped when nearly parallel.
GSC_DEV float4 contact_obb_obb(const Obb &a, const Obb &b)
{
    float r[3][3], ar[3][3];
#pragma unroll
    for (int row = 0; row < 3; ++row)
#pragma unroll
        for (int col = 0; col < 3; ++col) {
            r[row][col] = dot3(a.m[0][row], a.m[1][row], a.m[2][row], b.m[0][col], b.m[1][col], b.m[2][col]);
            ar[row][col] = add(fabsf(r[row][col]), kEpsilon);
        }
    const float tx = sub(b.cx, a.cx), ty = sub(b.cy, a.cy), tz = sub(b.cz, a.cz);
    float t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = add(add(mul(a.m[0][i], tx), mul(a.m[1][i], ty)), mul(a.m[2][i], tz));
    const float ah[3] = { a.hx, a.hy, a.hz };
    const float bh[3] = { b.hx, b.hy, b.hz };
    float best_pen = INFINITY, best_s = 0.0f, best_len = 1.0f;
    int best_axis = 0;  // 0..2 A_i, 3..5 B_i, 6 + 3 i + j cross
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float rb = add(add(mul(bh[0], ar[i][0]), mul(bh[1], ar[i][1])), mul(bh[2], ar[i][2]));
        const float pen = sub(add(ah[i], rb), fabsf(t[i]));
        if (pen < best_pen) { best_pen = pen; best_s = t[i]; best_axis = i; }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float ra = add(add(mul(ah[0], ar[0][i]), mul(ah[1], ar[1][i])), mul(ah[2], ar[2][i]));
        const float proj = add(add(mul(t[0], r[0][i]), mul(t[1], r[1][i])), mul(t[2], r[2][i]));
        const float pen = sub(add(ra, bh[i]), fabsf(proj));
        if (pen < best_pen) { best_pen = pen; best_s = proj; best_axis = 3 + i; }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        constexpr int nx[3] = { 1, 2, 0 }, pv[3] = { 2, 0, 1 };
        const int i1 = nx[i], i2 = pv[i];
        const int ilo = i1 < i2 ? i1 : i2, ihi = i1 < i2 ? i2 : i1;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int j1 = nx[j], j2 = pv[j];
            const int jlo = j1 < j2 ? j1 : j2, jhi = j1 < j2 ? j2 : j1;
            const float ra = add(mul(ah[ilo], ar[ihi][j]), mul(ah[ihi], ar[ilo][j]));
            const float rb = add(mul(bh[jlo], ar[i][jhi]), mul(bh[jhi], ar[i][jlo]));
            const float sv = sub(mul(t[i2], r[i1][j]), mul(t[i1], r[i2][j]));
            const float len2 = sub(1.0f, mul(r[i][j], r[i][j]));
            if (len2 > kEpsilon) {
                const float len = __fsqrt_rn(len2);
                const float pen = __fdiv_rn(sub(add(ra, rb), fabsf(sv)), len);
                if (pen < best_pen) { best_pen = pen; best_s = sv; best_axis = 6 + 3 * i + j; best_len = len; }
            }
        }
    }
    float lx, ly, lz;
    if (best_axis < 3) {
        lx = best_axis == 0 ? a.m[0][0] : (best_axis == 1 ? a.m[0][1] : a.m[0][2]);
        ly = best_axis == 0 ? a.m[1][0] : (best_axis == 1 ? a.m[1][1] : a.m[1][2]);
        lz = best_axis == 0 ? a.m[2][0] : (best_axis == 1 ? a.m[2][1] : a.m[2][2]);
    } else if (best_axis < 6) {
        const int k = best_axis - 3;
        lx = k == 0 ? b.m[0][0] : (k == 1 ? b.m[0][1] : b.m[0][2]);
        ly = k == 0 ? b.m[1][0] : (k == 1 ? b.m[1][1] : b.m[1][2]);
        lz = k == 0 ? b.m[2][0] : (k == 1 ? b.m[2][1] : b.m[2][2]);
    } else {
        const int i = (best_axis - 6) / 3, j = (best_axis - 6) % 3;
        const float ux = i == 0 ? a.m[0][0] : (i == 1 ? a.m[0][1] : a.m[0][2]);
        const float uy = i == 0 ? a.m[1][0] : (i == 1 ? a.m[1][1] : a.m[1][2]);
        const float uz = i == 0 ? a.m[2][0] : (i == 1 ? a.m[2][1] : a.m[2][2]);
        const float wx = j == 0 ? b.m[0][0] : (j == 1 ? b.m[0][1] : b.m[0][2]);
        const float wy = j == 0 ? b.m[1][0] : (j == 1 ? b.m[1][1] : b.m[1][2]);
        const float wz = j == 0 ? b.m[2][0] : (j == 1 ? b.m[2][1] : b.m[2][2]);
        lx = __fdiv_rn(sub(mul(uy, wz), mul(uz, wy)), best_len);
        ly = __fdiv_rn(sub(mul(uz, wx), mul(ux, wz)), best_len);
        lz = __fdiv_rn(sub(mul(ux, wy), mul(uy, wx)), best_len);
    }
    if (best_s > 0.0f) { lx = -lx; ly = -ly; lz = -lz; }  // L pointed from a to b: the normal goes lo -> hi
    return make_float4(lx, ly, lz, best_pen);
}

// grid_collision.rs:329-335  (p / cell_size).floor(), IEEE divide
GSC_DEV float cell_coord_f(float p, float cell_size) { return floorf(__fdiv_rn(p, cell_size)); }

// lib/hash/src/fnv.rs:21-38 over the six little-endian bytes of GridCell{x,y,z: i16}
GSC_DEV uint64_t fnv1a_gridcell(int16_t x, int16_t y, int16_t z)
{
    uint64_t h = 0xcbf29ce484222325ull;
    const uint16_t v[3] = { (uint16_t)x, (uint16_t)y, (uint16_t)z };
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        h = (h ^ (uint64_t)(v[i] & 0xff)) * 0x100000001b3ull;
        h = (h ^ (uint64_t)(v[i] >> 8)) * 0x100000001b3ull;
    }
    return h;
}

// monotone float <-> uint mapping for atomicMin/atomicMax on floats of either sign
GSC_DEV uint32_t f2ord(float f)
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float ord2f(uint32_t u)
{
    const uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

}  // namespace gsc

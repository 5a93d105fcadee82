/*
 * gunship_oracle.c -- CPU restatement of Gunship's collision hot path (plain C11).
 *
 * TEST INFRASTRUCTURE ONLY (see gunship_oracle.h).  Build with
 *     gcc -O2 -ffp-contract=off -fno-fast-math -msse2 -mfpmath=sse -fPIC -shared -pthread
 * so every f32 expression is evaluated exactly as the Rust source writes it:
 * IEEE binary32, round-to-nearest, separate multiply and add (rustc never
 * contracts to FMA), left-to-right association.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * the reference checkout).  Nothing here is copied: the reference is Rust with
 * HashMap/Vec/threads; this is C with hand-rolled open-addressing tables and
 * pthreads, restating the same algorithm.
 */
#define _GNU_SOURCE
#include "gunship_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* ------------------------------------------------------------------------ */
/* polygon_math primitives                                                   */
/* ------------------------------------------------------------------------ */

#define ORC_EPSILON 1e-6f /* lib/polygon_math/src/lib.rs:24 */

/* lib.rs:31-35  IsZero for f32 */
static inline int f_is_zero(float x) { return fabsf(x) < ORC_EPSILON; }

/* lib.rs:41-45  Clamp for f32: f32::min(f32::max(self, min), max) */
static inline float f_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

/* vector.rs:190-196  Dot for Vector3: x*x' + y*y' + z*z' (left to right) */
static inline float v_dot(orc_vec3 a, orc_vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

/* vector.rs:127-129 */
static inline float v_mag2(orc_vec3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }

/* matrix.rs:335-341 */
static inline orc_vec3 m_col(const orc_mat3 *m, int c)
{
    orc_vec3 v = { m->m[0][c], m->m[1][c], m->m[2][c] };
    return v;
}

/* matrix.rs:453-463  `Vector3 * Matrix3` is M.v, row by row */
static inline orc_vec3 v_mul_m(orc_vec3 v, const orc_mat3 *r)
{
    orc_vec3 o;
    o.x = r->m[0][0] * v.x + r->m[0][1] * v.y + r->m[0][2] * v.z;
    o.y = r->m[1][0] * v.x + r->m[1][1] * v.y + r->m[1][2] * v.z;
    o.z = r->m[2][0] * v.x + r->m[2][1] * v.y + r->m[2][2] * v.z;
    return o;
}

/* point.rs:110-116  Point - Point -> Vector3 */
static inline orc_vec3 p_sub(orc_point a, orc_point b)
{
    orc_vec3 v = { a.x - b.x, a.y - b.y, a.z - b.z };
    return v;
}

/* point.rs:118-134  Point + Vector3 (w forced to 1) */
static inline orc_point p_add_v(orc_point p, orc_vec3 v)
{
    orc_point o = { p.x + v.x, p.y + v.y, p.z + v.z, 1.0f };
    return o;
}

/* point.rs:136-147  Point - Vector3 */
static inline orc_point p_sub_v(orc_point p, orc_vec3 v)
{
    orc_point o = { p.x - v.x, p.y - v.y, p.z - v.z, 1.0f };
    return o;
}

/* matrix.rs:392-402  From<Orientation> for Matrix3 (the only in-tree quat->Matrix3) */
void orc_quat_to_mat3(const float q[4], orc_mat3 *o)
{
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    o->m[0][0] = w * w + x * x - y * y - z * z;
    o->m[0][1] = 2.0f * x * y - 2.0f * w * z;
    o->m[0][2] = 2.0f * x * z + 2.0f * w * y;
    o->m[1][0] = 2.0f * x * y + 2.0f * w * z;
    o->m[1][1] = w * w - x * x + y * y - z * z;
    o->m[1][2] = 2.0f * y * z - 2.0f * w * x;
    o->m[2][0] = 2.0f * x * z - 2.0f * w * y;
    o->m[2][1] = 2.0f * y * z + 2.0f * w * x;
    o->m[2][2] = w * w - x * x - y * y + z * z;
}

/* matrix.rs:320-333  Matrix3::rotation (Euler x -> y -> z) */
void orc_mat3_rotation(float x, float y, float z, orc_mat3 *o)
{
    const float s1 = sinf(x), c1 = cosf(x);
    const float s2 = sinf(y), c2 = cosf(y);
    const float s3 = sinf(z), c3 = cosf(z);
    o->m[0][0] = c2 * c3;
    o->m[0][1] = -c2 * s3;
    o->m[0][2] = s2;
    o->m[1][0] = c1 * s3 + c3 * s1 * s2;
    o->m[1][1] = c1 * c3 - s1 * s2 * s3;
    o->m[1][2] = -c2 * s1;
    o->m[2][0] = s1 * s3 - c1 * c3 * s2;
    o->m[2][1] = c3 * s1 + c1 * s2 * s3;
    o->m[2][2] = c1 * c2;
}

/* ------------------------------------------------------------------------ */
/* collider cache + AABB                                                     */
/* ------------------------------------------------------------------------ */

/* mod.rs:311-333  CachedCollider::from_collider_transform */
void orc_cache_collider(uint8_t kind, const float offset[3], const float size[3],
                        const float pos[3], const float quat[4], const float scale[3],
                        orc_cached_collider *out)
{
    memset(out, 0, sizeof *out);
    orc_point position = { pos[0], pos[1], pos[2], 1.0f };
    orc_vec3 off = { offset[0], offset[1], offset[2] };
    if (kind == ORC_SPHERE) {
        out->tag = ORC_SPHERE;
        out->u.sphere.center = p_add_v(position, off); /* :315 */
        out->u.sphere.radius = size[0];                /* :316 -- radius is NOT scaled */
    } else if (kind == ORC_BOX) {
        out->tag = ORC_BOX;
        /* :320  widths * scale * 0.5  == (widths*scale)*0.5, component-wise then scalar */
        orc_vec3 hw;
        hw.x = size[0] * scale[0];
        hw.y = size[1] * scale[1];
        hw.z = size[2] * scale[2];
        hw.x *= 0.5f;
        hw.y *= 0.5f;
        hw.z *= 0.5f;
        out->u.obb.half_widths = hw;
        out->u.obb.center = p_add_v(position, off);     /* :321 */
        orc_quat_to_mat3(quat, &out->u.obb.orientation); /* :322 */
    } else {
        out->tag = ORC_MESH; /* :331 unimplemented!() */
    }
}

/* bounding_volume.rs:148-336  AABB::from_collider */
void orc_aabb_from_collider(const orc_cached_collider *c, orc_aabb *out)
{
    if (c->tag == ORC_SPHERE) {
        /* :150-159 */
        const float r = c->u.sphere.radius;
        orc_vec3 hw = { r, r, r };
        out->min = p_sub_v(c->u.sphere.center, hw);
        out->max = p_add_v(c->u.sphere.center, hw);
        return;
    }
    /* :161-327  eight vertices in the order (+,+,+) (+,+,-) (+,-,+) (+,-,-) (-,+,+) (-,+,-) (-,-,+) (-,-,-);
     * running min/max with `if v < min {..} else if v > max {..}` per axis. */
    static const float sign[8][3] = {
        { 1.0f, 1.0f, 1.0f },  { 1.0f, 1.0f, -1.0f },  { 1.0f, -1.0f, 1.0f },  { 1.0f, -1.0f, -1.0f },
        { -1.0f, 1.0f, 1.0f }, { -1.0f, 1.0f, -1.0f }, { -1.0f, -1.0f, 1.0f }, { -1.0f, -1.0f, -1.0f },
    };
    const orc_obb *o = &c->u.obb;
    orc_point mn, mx;
    for (int k = 0; k < 8; ++k) {
        orc_vec3 hv = { o->half_widths.x * sign[k][0], o->half_widths.y * sign[k][1], o->half_widths.z * sign[k][2] };
        orc_point vert = p_add_v(o->center, v_mul_m(hv, &o->orientation));
        if (k == 0) {
            mn = vert;
            mx = vert;
            continue;
        }
        if (vert.x < mn.x) mn.x = vert.x; else if (vert.x > mx.x) mx.x = vert.x;
        if (vert.y < mn.y) mn.y = vert.y; else if (vert.y > mx.y) mx.y = vert.y;
        if (vert.z < mn.z) mn.z = vert.z; else if (vert.z > mx.z) mx.z = vert.z;
    }
    out->min = mn;
    out->max = mx;
}

/* bounding_volume.rs:411-419 */
static inline int test_ranges(float min_a, float max_a, float min_b, float max_b)
{
    return !(min_a > max_b || min_b > max_a || max_a < min_b || max_b < min_a);
}

/* bounding_volume.rs:338-343 */
int orc_test_aabb(const orc_aabb *a, const orc_aabb *b)
{
    return test_ranges(a->min.x, a->max.x, b->min.x, b->max.x)
        && test_ranges(a->min.y, a->max.y, b->min.y, b->max.y)
        && test_ranges(a->min.z, a->max.z, b->min.z, b->max.z);
}

/* ------------------------------------------------------------------------ */
/* narrowphase                                                               */
/* ------------------------------------------------------------------------ */

/* mod.rs:383-389 */
int orc_test_sphere_sphere(const orc_sphere *a, const orc_sphere *b)
{
    const float dist_sqr = v_mag2(p_sub(a->center, b->center));
    const float max_dist_sqr = (a->radius + b->radius) * (a->radius + b->radius);
    const float diff = dist_sqr - max_dist_sqr;
    return diff < 0.0f || f_is_zero(diff);
}

/* mod.rs:549-568  OBB::closest_point */
static orc_point obb_closest_point(const orc_obb *o, orc_point p)
{
    const orc_vec3 d = p_sub(p, o->center);
    orc_point result = o->center;
    const float hw[3] = { o->half_widths.x, o->half_widths.y, o->half_widths.z };
    for (int axis = 0; axis < 3; ++axis) {
        const orc_vec3 local_axis = m_col(&o->orientation, axis);
        float dist = v_dot(d, local_axis);
        dist = f_clamp(dist, -hw[axis], hw[axis]);
        /* `dist * local_axis` is Vector3*f32 component-wise (vector.rs:272-296), then Point += Vector3 */
        orc_vec3 step = { local_axis.x * dist, local_axis.y * dist, local_axis.z * dist };
        result = p_add_v(result, step);
    }
    return result;
}

/* mod.rs:391-394 + :571-574  (strict <, no epsilon) */
int orc_test_sphere_obb(const orc_sphere *s, const orc_obb *o)
{
    const orc_point closest = obb_closest_point(o, s->center);
    const float dist_sqr = v_mag2(p_sub(s->center, closest));
    return dist_sqr < s->radius * s->radius;
}

/* mod.rs:413-546  15-axis SAT; `a` is self, `b` the argument */
int orc_test_obb_obb(const orc_obb *a, const orc_obb *b)
{
    float r[3][3], abs_r[3][3];
    /* :415-423 */
    for (int row = 0; row < 3; ++row)
        for (int col = 0; col < 3; ++col)
            r[row][col] = v_dot(m_col(&a->orientation, row), m_col(&b->orientation, col));

    /* :426-429  t = (b.center - a.center) * a.orientation.transpose()  (= A^T . t) */
    orc_vec3 tv = p_sub(b->center, a->center);
    orc_mat3 at;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            at.m[i][j] = a->orientation.m[j][i];
    tv = v_mul_m(tv, &at);
    const float t[3] = { tv.x, tv.y, tv.z };

    /* :433-441 */
    for (int row = 0; row < 3; ++row)
        for (int col = 0; col < 3; ++col)
            abs_r[row][col] = fabsf(r[row][col]) + ORC_EPSILON;

    const float ah[3] = { a->half_widths.x, a->half_widths.y, a->half_widths.z };
    const float bh[3] = { b->half_widths.x, b->half_widths.y, b->half_widths.z };

    /* :444-451  L = A0, A1, A2 */
    for (int i = 0; i < 3; ++i) {
        const float ra = ah[i];
        const float rb = bh[0] * abs_r[i][0] + bh[1] * abs_r[i][1] + bh[2] * abs_r[i][2];
        if (fabsf(t[i]) > ra + rb) return 0;
    }
    /* :454-461  L = B0, B1, B2 */
    for (int i = 0; i < 3; ++i) {
        const float ra = ah[0] * abs_r[0][i] + ah[1] * abs_r[1][i] + ah[2] * abs_r[2][i];
        const float rb = bh[i];
        const float proj = t[0] * r[0][i] + t[1] * r[1][i] + t[2] * r[2][i];
        if (fabsf(proj) > ra + rb) return 0;
    }
    float ra, rb;
    /* :464-470  A0 x B0 */
    ra = ah[1] * abs_r[2][0] + ah[2] * abs_r[1][0];
    rb = bh[1] * abs_r[0][2] + bh[2] * abs_r[0][1];
    if (fabsf(t[2] * r[1][0] - t[1] * r[2][0]) > ra + rb) return 0;
    /* :473-479  A0 x B1 */
    ra = ah[1] * abs_r[2][1] + ah[2] * abs_r[1][1];
    rb = bh[0] * abs_r[0][2] + bh[2] * abs_r[0][0];
    if (fabsf(t[2] * r[1][1] - t[1] * r[2][1]) > ra + rb) return 0;
    /* :482-488  A0 x B2 */
    ra = ah[1] * abs_r[2][2] + ah[2] * abs_r[1][2];
    rb = bh[0] * abs_r[0][1] + bh[1] * abs_r[0][0];
    if (fabsf(t[2] * r[1][2] - t[1] * r[2][2]) > ra + rb) return 0;
    /* :491-497  A1 x B0 */
    ra = ah[0] * abs_r[2][0] + ah[2] * abs_r[0][0];
    rb = bh[1] * abs_r[1][2] + bh[2] * abs_r[1][1];
    if (fabsf(t[0] * r[2][0] - t[2] * r[0][0]) > ra + rb) return 0;
    /* :500-506  A1 x B1 */
    ra = ah[0] * abs_r[2][1] + ah[2] * abs_r[0][1];
    rb = bh[0] * abs_r[1][2] + bh[2] * abs_r[1][0];
    if (fabsf(t[0] * r[2][1] - t[2] * r[0][1]) > ra + rb) return 0;
    /* :509-515  A1 x B2 */
    ra = ah[0] * abs_r[2][2] + ah[2] * abs_r[0][2];
    rb = bh[0] * abs_r[1][1] + bh[1] * abs_r[1][0];
    if (fabsf(t[0] * r[2][2] - t[2] * r[0][2]) > ra + rb) return 0;
    /* :518-524  A2 x B0 */
    ra = ah[0] * abs_r[1][0] + ah[1] * abs_r[0][0];
    rb = bh[1] * abs_r[2][2] + bh[2] * abs_r[2][1];
    if (fabsf(t[1] * r[0][0] - t[0] * r[1][0]) > ra + rb) return 0;
    /* :527-533  A2 x B1 */
    ra = ah[0] * abs_r[1][1] + ah[1] * abs_r[0][1];
    rb = bh[0] * abs_r[2][2] + bh[2] * abs_r[2][0];
    if (fabsf(t[1] * r[0][1] - t[0] * r[1][1]) > ra + rb) return 0;
    /* :536-542  A2 x B2 */
    ra = ah[0] * abs_r[1][2] + ah[1] * abs_r[0][2];
    rb = bh[0] * abs_r[2][1] + bh[1] * abs_r[2][0];
    if (fabsf(t[1] * r[0][2] - t[0] * r[1][2]) > ra + rb) return 0;
    return 1; /* :545 */
}

/* mod.rs:335-345 (CachedCollider::test), :375-381 (Sphere::test_collider), :405-411 (OBB::test_collider) */
int orc_test_collider(const orc_cached_collider *a, const orc_cached_collider *b)
{
    if (a->tag == ORC_SPHERE) {
        if (b->tag == ORC_SPHERE) return orc_test_sphere_sphere(&a->u.sphere, &b->u.sphere);
        return orc_test_sphere_obb(&a->u.sphere, &b->u.obb);
    }
    if (b->tag == ORC_SPHERE) return orc_test_sphere_obb(&b->u.sphere, &a->u.obb);
    return orc_test_obb_obb(&a->u.obb, &b->u.obb);
}

/* bounding_volume.rs:124-132 */
int orc_test_volume(const orc_bound_volume *a, const orc_bound_volume *b)
{
    if (orc_test_aabb(&a->aabb, &b->aabb))
        if (orc_test_collider(&a->collider, &b->collider))
            return 1;
    return 0;
}

/* ------------------------------------------------------------------------ */
/* contacts (depth / normal) -- NOT in the reference (it only        */
/* plans them, mod.rs:110-113).  PARITY UNPINNED: these functions define the    */
/* outputs BASELINE.json's north_star asks for and are the checker of the CUDA  */
/* implementation of the same definition (include/gunship_collision.h).         */
/* Convention: `hi` is the later volume of the pair, `lo` the earlier one; the  */
/* unit normal points from lo to hi, depth is the overlap along it.             */
/* ------------------------------------------------------------------------ */

void orc_contact_sphere_sphere(const orc_sphere *hi, const orc_sphere *lo, float out[4])
{
    const orc_vec3 d = p_sub(hi->center, lo->center);
    const float dist = sqrtf(v_mag2(d));
    if (dist > 0.0f) { out[0] = d.x / dist; out[1] = d.y / dist; out[2] = d.z / dist; }
    else { out[0] = 1.0f; out[1] = 0.0f; out[2] = 0.0f; }
    out[3] = (hi->radius + lo->radius) - dist;
}

/* from OBB::closest_point (mod.rs:549-568): outside, the normal is (centre - closest point) normalised; a
 * centre inside the box leaves through the face of least penetration */
void orc_contact_sphere_obb(const orc_sphere *s, const orc_obb *o, int sphere_is_hi, float out[4])
{
    const orc_vec3 d = p_sub(s->center, o->center);
    orc_point res = o->center;
    const float hw[3] = { o->half_widths.x, o->half_widths.y, o->half_widths.z };
    float proj[3];
    for (int axis = 0; axis < 3; ++axis) {
        const orc_vec3 ax = m_col(&o->orientation, axis);
        const float dist = v_dot(d, ax);
        proj[axis] = dist;
        const float c = f_clamp(dist, -hw[axis], hw[axis]);
        orc_vec3 step = { ax.x * c, ax.y * c, ax.z * c };
        res = p_add_v(res, step);
    }
    const orc_vec3 v = p_sub(s->center, res);
    const float d2 = v_mag2(v);
    /* inside <=> no projection was clamped (d2 alone is not robust: the rebuilt point differs by rounding) */
    const int inside = fabsf(proj[0]) <= hw[0] && fabsf(proj[1]) <= hw[1] && fabsf(proj[2]) <= hw[2];
    float n[3], depth;
    if (!inside && d2 > 0.0f) {
        const float dd = sqrtf(d2);
        n[0] = v.x / dd; n[1] = v.y / dd; n[2] = v.z / dd;
        depth = s->radius - dd;
    } else {
        int best = 0;
        float pen = hw[0] - fabsf(proj[0]);
        for (int i = 1; i < 3; ++i) {
            const float p = hw[i] - fabsf(proj[i]);
            if (p < pen) { pen = p; best = i; }
        }
        const float sgn = proj[best] < 0.0f ? -1.0f : 1.0f;
        const orc_vec3 ax = m_col(&o->orientation, best);
        n[0] = ax.x * sgn; n[1] = ax.y * sgn; n[2] = ax.z * sgn;
        depth = s->radius + pen;
    }
    if (!sphere_is_hi) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
    out[0] = n[0]; out[1] = n[1]; out[2] = n[2]; out[3] = depth;
}

/* least-penetration axis of the 15 SAT axes of OBB::test_obb (mod.rs:443-542), `a` = hi, `b` = lo.
 * Cross axes are normalised by |A_i x B_j| = sqrt(1 - r_ij^2) and skipped when nearly parallel. */
void orc_contact_obb_obb(const orc_obb *a, const orc_obb *b, float out[4])
{
    float r[3][3], abs_r[3][3];
    for (int row = 0; row < 3; ++row)
        for (int col = 0; col < 3; ++col)
            r[row][col] = v_dot(m_col(&a->orientation, row), m_col(&b->orientation, col));
    orc_vec3 tv = p_sub(b->center, a->center);
    orc_mat3 at;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            at.m[i][j] = a->orientation.m[j][i];
    tv = v_mul_m(tv, &at);
    const float t[3] = { tv.x, tv.y, tv.z };
    for (int row = 0; row < 3; ++row)
        for (int col = 0; col < 3; ++col)
            abs_r[row][col] = fabsf(r[row][col]) + ORC_EPSILON;
    const float ah[3] = { a->half_widths.x, a->half_widths.y, a->half_widths.z };
    const float bh[3] = { b->half_widths.x, b->half_widths.y, b->half_widths.z };

    float best_pen = INFINITY, best_s = 0.0f, best_len = 1.0f;
    int best_kind = 0, best_i = 0, best_j = 0;
    for (int i = 0; i < 3; ++i) { /* L = A_i */
        const float rb = bh[0] * abs_r[i][0] + bh[1] * abs_r[i][1] + bh[2] * abs_r[i][2];
        const float pen = (ah[i] + rb) - fabsf(t[i]);
        if (pen < best_pen) { best_pen = pen; best_s = t[i]; best_kind = 0; best_i = i; }
    }
    for (int i = 0; i < 3; ++i) { /* L = B_i */
        const float ra = ah[0] * abs_r[0][i] + ah[1] * abs_r[1][i] + ah[2] * abs_r[2][i];
        const float proj = t[0] * r[0][i] + t[1] * r[1][i] + t[2] * r[2][i];
        const float pen = (ra + bh[i]) - fabsf(proj);
        if (pen < best_pen) { best_pen = pen; best_s = proj; best_kind = 1; best_i = i; }
    }
    for (int i = 0; i < 3; ++i) { /* L = A_i x B_j, same expressions as mod.rs:463-542 */
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        for (int j = 0; j < 3; ++j) {
            const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            const float ra = ah[i1 < i2 ? i1 : i2] * abs_r[i1 < i2 ? i2 : i1][j] + ah[i1 < i2 ? i2 : i1] * abs_r[i1 < i2 ? i1 : i2][j];
            const float rb = bh[j1 < j2 ? j1 : j2] * abs_r[i][j1 < j2 ? j2 : j1] + bh[j1 < j2 ? j2 : j1] * abs_r[i][j1 < j2 ? j1 : j2];
            const float sv = t[i2] * r[i1][j] - t[i1] * r[i2][j];
            const float len2 = 1.0f - r[i][j] * r[i][j];
            if (len2 > ORC_EPSILON) {
                const float len = sqrtf(len2);
                const float pen = ((ra + rb) - fabsf(sv)) / len;
                if (pen < best_pen) { best_pen = pen; best_s = sv; best_kind = 2; best_i = i; best_j = j; best_len = len; }
            }
        }
    }
    orc_vec3 L;
    if (best_kind == 0) L = m_col(&a->orientation, best_i);
    else if (best_kind == 1) L = m_col(&b->orientation, best_i);
    else {
        const orc_vec3 u = m_col(&a->orientation, best_i), w = m_col(&b->orientation, best_j);
        L.x = (u.y * w.z - u.z * w.y) / best_len;
        L.y = (u.z * w.x - u.x * w.z) / best_len;
        L.z = (u.x * w.y - u.y * w.x) / best_len;
    }
    if (best_s > 0.0f) { L.x = -L.x; L.y = -L.y; L.z = -L.z; } /* L pointed from a to b: the normal goes lo -> hi */
    out[0] = L.x; out[1] = L.y; out[2] = L.z; out[3] = best_pen;
}

/* ------------------------------------------------------------------------ */
/* FNV-1a + grid cells                                                       */
/* ------------------------------------------------------------------------ */

#define FNV_BASIS 0xcbf29ce484222325ull /* fnv.rs:23 */
#define FNV_PRIME 0x100000001b3ull      /* fnv.rs:34 */

/* fnv.rs:31-38 */
uint64_t orc_fnv1a(const uint8_t *bytes, size_t n, uint64_t h)
{
    for (size_t i = 0; i < n; ++i) {
        h ^= (uint64_t)bytes[i];
        h *= FNV_PRIME;
    }
    return h;
}

/* grid_collision.rs:523 derive(Hash): write_i16(x), write_i16(y), write_i16(z), native (little) endian */
uint64_t orc_fnv1a_gridcell(orc_grid_cell c)
{
    uint8_t b[6];
    memcpy(b + 0, &c.x, 2);
    memcpy(b + 2, &c.y, 2);
    memcpy(b + 4, &c.z, 2);
    return orc_fnv1a(b, 6, FNV_BASIS);
}

/* (Entity, Entity) derive(Hash): write_u32 twice (src/old/ecs.rs:7-8) */
static inline uint64_t fnv_pair(uint32_t a, uint32_t b)
{
    uint8_t buf[8];
    memcpy(buf, &a, 4);
    memcpy(buf + 4, &b, 4);
    return orc_fnv1a(buf, 8, FNV_BASIS);
}

static inline int16_t to_grid_coord(float v, uint32_t *out_of_range)
{
    /* `as i16` on an out-of-range float was undefined in 2016 Rust; the C-ABI rejects such scenes
     * (SURVEY §7 "hard parts").  Saturate here and count, so tests can assert the count is 0. */
    if (!(v >= -32768.0f)) { if (out_of_range) ++*out_of_range; return INT16_MIN; }
    if (!(v <= 32767.0f))  { if (out_of_range) ++*out_of_range; return INT16_MAX; }
    return (int16_t)v;
}

/* grid_collision.rs:329-335 */
static inline orc_grid_cell world_to_grid(orc_point p, float cell_size, uint32_t *oor)
{
    orc_grid_cell c;
    c.x = to_grid_coord(floorf(p.x / cell_size), oor);
    c.y = to_grid_coord(floorf(p.y / cell_size), oor);
    c.z = to_grid_coord(floorf(p.z / cell_size), oor);
    return c;
}

orc_grid_cell orc_world_to_grid(orc_point p, float cell_size) { return world_to_grid(p, cell_size, NULL); }

/* ---- batch helpers ---- */
static void load_obb16(const float *f, orc_obb *o)
{
    /* 16 floats: center xyz, 9 orientation (row-major), half xyz, pad */
    o->center.x = f[0]; o->center.y = f[1]; o->center.z = f[2]; o->center.w = 1.0f;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o->orientation.m[i][j] = f[3 + 3 * i + j];
    o->half_widths.x = f[12]; o->half_widths.y = f[13]; o->half_widths.z = f[14];
}
static void load_sphere4(const float *f, orc_sphere *s)
{
    s->center.x = f[0]; s->center.y = f[1]; s->center.z = f[2]; s->center.w = 1.0f; s->radius = f[3];
}
void orc_batch_sphere_sphere(size_t n, const float *a, const float *b, uint8_t *out)
{
    for (size_t i = 0; i < n; ++i) {
        orc_sphere sa, sb;
        load_sphere4(a + 4 * i, &sa);
        load_sphere4(b + 4 * i, &sb);
        out[i] = (uint8_t)orc_test_sphere_sphere(&sa, &sb);
    }
}
void orc_batch_sphere_obb(size_t n, const float *s, const float *o16, uint8_t *out)
{
    for (size_t i = 0; i < n; ++i) {
        orc_sphere sa;
        orc_obb ob;
        load_sphere4(s + 4 * i, &sa);
        load_obb16(o16 + 16 * i, &ob);
        out[i] = (uint8_t)orc_test_sphere_obb(&sa, &ob);
    }
}
void orc_batch_obb_obb(size_t n, const float *a16, const float *b16, uint8_t *out)
{
    for (size_t i = 0; i < n; ++i) {
        orc_obb oa, ob;
        load_obb16(a16 + 16 * i, &oa);
        load_obb16(b16 + 16 * i, &ob);
        out[i] = (uint8_t)orc_test_obb_obb(&oa, &ob);
    }
}
void orc_batch_contact_sphere_sphere(size_t n, const float *hi, const float *lo, float *out4)
{
    for (size_t i = 0; i < n; ++i) {
        orc_sphere a, b;
        load_sphere4(hi + 4 * i, &a);
        load_sphere4(lo + 4 * i, &b);
        orc_contact_sphere_sphere(&a, &b, out4 + 4 * i);
    }
}

void orc_batch_contact_sphere_obb(size_t n, const float *s, const float *o16, const uint8_t *sphere_is_hi, float *out4)
{
    for (size_t i = 0; i < n; ++i) {
        orc_sphere a;
        orc_obb o;
        load_sphere4(s + 4 * i, &a);
        load_obb16(o16 + 16 * i, &o);
        orc_contact_sphere_obb(&a, &o, sphere_is_hi[i], out4 + 4 * i);
    }
}

void orc_batch_contact_obb_obb(size_t n, const float *a16, const float *b16, float *out4)
{
    for (size_t i = 0; i < n; ++i) {
        orc_obb a, b;
        load_obb16(a16 + 16 * i, &a);
        load_obb16(b16 + 16 * i, &b);
        orc_contact_obb_obb(&a, &b, out4 + 4 * i);
    }
}

void orc_batch_aabb_aabb(size_t n, const float *a6, const float *b6, uint8_t *out)
{
    for (size_t i = 0; i < n; ++i) {
        orc_aabb a = { { a6[6 * i], a6[6 * i + 1], a6[6 * i + 2], 1.0f }, { a6[6 * i + 3], a6[6 * i + 4], a6[6 * i + 5], 1.0f } };
        orc_aabb b = { { b6[6 * i], b6[6 * i + 1], b6[6 * i + 2], 1.0f }, { b6[6 * i + 3], b6[6 * i + 4], b6[6 * i + 5], 1.0f } };
        out[i] = (uint8_t)orc_test_aabb(&a, &b);
    }
}
void orc_batch_fnv1a_gridcell(size_t n, const int16_t *cells, uint64_t *out)
{
    for (size_t i = 0; i < n; ++i) {
        orc_grid_cell c = { cells[3 * i], cells[3 * i + 1], cells[3 * i + 2] };
        out[i] = orc_fnv1a_gridcell(c);
    }
}

/* ------------------------------------------------------------------------ */
/* open-addressing containers standing in for HashMap/HashSet<_, FnvHashState> */
/* ------------------------------------------------------------------------ */

/* set of packed (first<<32 | second) pairs; 0 is never a valid pair here because it would mean
 * entity 0 colliding with entity 0 -- but be safe and keep an explicit occupancy flag via key+1. */
typedef struct {
    uint64_t *slot; /* stores key+1; 0 == empty */
    uint64_t cap;   /* power of two */
    uint64_t len;
} pair_set;

static void pair_set_init(pair_set *s) { s->cap = 1024; s->len = 0; s->slot = calloc(s->cap, sizeof(uint64_t)); }
static void pair_set_free(pair_set *s) { free(s->slot); s->slot = NULL; }
static void pair_set_clear(pair_set *s) { memset(s->slot, 0, s->cap * sizeof(uint64_t)); s->len = 0; }
static int pair_set_insert_raw(pair_set *s, uint64_t key);
static void pair_set_grow(pair_set *s)
{
    pair_set n = { calloc(s->cap * 2, sizeof(uint64_t)), s->cap * 2, 0 };
    for (uint64_t i = 0; i < s->cap; ++i)
        if (s->slot[i]) pair_set_insert_raw(&n, s->slot[i] - 1);
    free(s->slot);
    *s = n;
}
/* returns 1 if newly inserted */
static int pair_set_insert_raw(pair_set *s, uint64_t key)
{
    if ((s->len + 1) * 10 > s->cap * 7) pair_set_grow(s);
    uint64_t h = fnv_pair((uint32_t)(key >> 32), (uint32_t)key) & (s->cap - 1);
    for (;;) {
        if (!s->slot[h]) { s->slot[h] = key + 1; ++s->len; return 1; }
        if (s->slot[h] == key + 1) return 0;
        h = (h + 1) & (s->cap - 1);
    }
}
static int pair_set_contains(const pair_set *s, uint64_t key)
{
    uint64_t h = fnv_pair((uint32_t)(key >> 32), (uint32_t)key) & (s->cap - 1);
    for (;;) {
        if (!s->slot[h]) return 0;
        if (s->slot[h] == key + 1) return 1;
        h = (h + 1) & (s->cap - 1);
    }
}

/* Vec<*const BoundVolume> */
typedef struct { const orc_bound_volume **p; uint32_t len, cap; } vol_vec;

static inline void vol_vec_push(vol_vec *v, const orc_bound_volume *b)
{
    if (v->len == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 4;
        v->p = realloc(v->p, (size_t)v->cap * sizeof *v->p);
    }
    v->p[v->len++] = b;
}

/* CollisionGrid = HashMap<GridCell, Vec<*const BoundVolume>, FnvHashState>  (grid_collision.rs:107) */
typedef struct {
    uint64_t key; /* 1 + packed 48-bit cell; 0 == empty */
    vol_vec  vec;
} grid_slot;

typedef struct {
    grid_slot *slot;
    uint64_t cap, len;
} cell_grid;

static inline uint64_t pack_cell(orc_grid_cell c)
{
    return ((uint64_t)(uint16_t)c.x) | ((uint64_t)(uint16_t)c.y << 16) | ((uint64_t)(uint16_t)c.z << 32);
}

typedef struct { const orc_bound_volume *a, *b; } candidate;

/* WorkUnit (grid_collision.rs:297-336) */
typedef struct {
    pair_set collisions;
    orc_aabb bounds;
    cell_grid grid;
    float cell_size;
    /* stats */
    uint64_t n_candidates, n_cell_entries;
    uint32_t n_wide, n_oor;
} work_unit;

/* Worker (grid_collision.rs:345-351): candidate list + cell_cache are per worker thread */
typedef struct {
    candidate *cand;
    uint64_t n_cand, cap_cand;
    vol_vec *cell_cache;
    uint64_t n_cache, cap_cache;
} worker_state;

struct orc_system {
    /* BoundingVolumeManager (bounding_volume.rs:17-113) */
    orc_bound_volume *components;
    uint32_t n, cap;
    float longest_axis;
    /* ThreadData.volumes: the per-frame clone (grid_collision.rs:243-244, 340-343) */
    orc_bound_volume *volumes;
    uint32_t vol_cap;
    /* work units + pool */
    int num_units, num_workers;
    work_unit *units;
    worker_state *workers;
    pthread_t *threads;
    pthread_mutex_t mu;
    pthread_cond_t cv_pending, cv_done;
    int *pending;     /* stack of unit ids (Mutex<Vec<WorkUnit>>, popped from the back :374) */
    int n_pending;
    int *done;        /* the mpsc channel */
    int n_done;
    int shutdown;
    /* GridCollisionSystem.collisions (grid_collision.rs:119) */
    pair_set collisions;
    uint64_t last_candidates, last_entries;
    uint32_t last_wide, last_oor;
};

static void grid_init(cell_grid *g) { g->cap = 1024; g->len = 0; g->slot = calloc(g->cap, sizeof(grid_slot)); }

static vol_vec *grid_entry(cell_grid *g, worker_state *w, orc_grid_cell c);

static void grid_grow(cell_grid *g)
{
    grid_slot *old = g->slot;
    uint64_t oc = g->cap;
    g->cap *= 2;
    g->slot = calloc(g->cap, sizeof(grid_slot));
    for (uint64_t i = 0; i < oc; ++i) {
        if (!old[i].key) continue;
        uint64_t k = old[i].key - 1;
        orc_grid_cell c = { (int16_t)(k & 0xffff), (int16_t)((k >> 16) & 0xffff), (int16_t)((k >> 32) & 0xffff) };
        uint64_t h = orc_fnv1a_gridcell(c) & (g->cap - 1);
        while (g->slot[h].key) h = (h + 1) & (g->cap - 1);
        g->slot[h] = old[i];
    }
    free(old);
}

/* work.grid.entry(cell).or_insert_with(|| cell_cache.pop().unwrap_or(Vec::new()))  (grid_collision.rs:434-436) */
static vol_vec *grid_entry(cell_grid *g, worker_state *w, orc_grid_cell c)
{
    if ((g->len + 1) * 10 > g->cap * 7) grid_grow(g);
    const uint64_t key = pack_cell(c) + 1;
    uint64_t h = orc_fnv1a_gridcell(c) & (g->cap - 1);
    for (;;) {
        grid_slot *s = &g->slot[h];
        if (s->key == key) return &s->vec;
        if (!s->key) {
            s->key = key;
            if (w->n_cache) s->vec = w->cell_cache[--w->n_cache];
            else memset(&s->vec, 0, sizeof s->vec);
            ++g->len;
            return &s->vec;
        }
        h = (h + 1) & (g->cap - 1);
    }
}

/* the `test_cell` closure (grid_collision.rs:419-445) */
static inline void test_cell(work_unit *wu, worker_state *w, const orc_bound_volume *bvh, orc_grid_cell cell)
{
    vol_vec *v = grid_entry(&wu->grid, w, cell);
    if (w->n_cand + v->len > w->cap_cand) {
        while (w->n_cand + v->len > w->cap_cand) w->cap_cand = w->cap_cand ? w->cap_cand * 2 : 4096;
        w->cand = realloc(w->cand, w->cap_cand * sizeof *w->cand);
    }
    for (uint32_t i = 0; i < v->len; ++i) { /* :439-441 */
        w->cand[w->n_cand].a = bvh;
        w->cand[w->n_cand].b = v->p[i];
        ++w->n_cand;
    }
    vol_vec_push(v, bvh); /* :444 */
    ++wu->n_cell_entries;
}

static inline orc_grid_cell mk_cell(int16_t x, int16_t y, int16_t z) { orc_grid_cell c = { x, y, z }; return c; }

/* Worker::do_broadphase (grid_collision.rs:389-490) */
static void do_broadphase(orc_system *s, work_unit *wu, worker_state *w)
{
    const orc_bound_volume *vols = s->volumes;
    for (uint32_t i = 0; i < s->n; ++i) {
        const orc_bound_volume *bvh = &vols[i];
        if (!orc_test_aabb(&bvh->aabb, &wu->bounds)) continue; /* :397 */
        const orc_grid_cell mn = world_to_grid(bvh->aabb.min, wu->cell_size, &wu->n_oor); /* :401 */
        const orc_grid_cell mx = world_to_grid(bvh->aabb.max, wu->cell_size, &wu->n_oor); /* :402 */
        if (mx.x - mn.x > 1 || mx.y - mn.y > 1 || mx.z - mn.z > 1) ++wu->n_wide; /* debug_assert :403-410 */

        test_cell(wu, w, bvh, mn); /* :447 */
        const int ox = mn.x < mx.x, oy = mn.y < mx.y, oz = mn.z < mx.z; /* :449-451 */
        if (ox) { /* :454-480 */
            test_cell(wu, w, bvh, mk_cell(mx.x, mn.y, mn.z));
            if (oy) {
                test_cell(wu, w, bvh, mk_cell(mn.x, mx.y, mn.z));
                test_cell(wu, w, bvh, mk_cell(mx.x, mx.y, mn.z));
                if (oz) {
                    test_cell(wu, w, bvh, mk_cell(mn.x, mn.y, mx.z));
                    test_cell(wu, w, bvh, mk_cell(mn.x, mx.y, mx.z));
                    test_cell(wu, w, bvh, mk_cell(mx.x, mn.y, mx.z));
                    test_cell(wu, w, bvh, mk_cell(mx.x, mx.y, mx.z));
                }
            } else if (oz) {
                test_cell(wu, w, bvh, mk_cell(mn.x, mn.y, mx.z));
                test_cell(wu, w, bvh, mk_cell(mx.x, mn.y, mx.z));
            }
        } else if (oy) {
            test_cell(wu, w, bvh, mk_cell(mn.x, mx.y, mn.z));
            if (oz) {
                test_cell(wu, w, bvh, mk_cell(mn.x, mn.y, mx.z));
                test_cell(wu, w, bvh, mk_cell(mn.x, mx.y, mx.z));
            }
        } else if (oz) {
            test_cell(wu, w, bvh, mk_cell(mn.x, mn.y, mx.z));
        }
    }
    /* :486-489  drain the grid, park the cleared Vecs in the worker's cell_cache */
    cell_grid *g = &wu->grid;
    for (uint64_t i = 0; i < g->cap; ++i) {
        if (!g->slot[i].key) continue;
        g->slot[i].vec.len = 0;
        if (w->n_cache == w->cap_cache) {
            w->cap_cache = w->cap_cache ? w->cap_cache * 2 : 1024;
            w->cell_cache = realloc(w->cell_cache, w->cap_cache * sizeof *w->cell_cache);
        }
        w->cell_cache[w->n_cache++] = g->slot[i].vec;
        g->slot[i].key = 0;
    }
    g->len = 0;
}

/* Worker::do_narrowphase (grid_collision.rs:492-513) */
static void do_narrowphase(work_unit *wu, worker_state *w)
{
    wu->n_candidates += w->n_cand;
    for (uint64_t i = 0; i < w->n_cand; ++i) {
        const orc_bound_volume *bvh = w->cand[i].a, *other = w->cand[i].b;
        const uint64_t key = ((uint64_t)bvh->entity << 32) | other->entity; /* :497 */
        if (pair_set_contains(&wu->collisions, key)) continue;            /* :502 Entry::Occupied */
        if (orc_test_volume(bvh, other))                                   /* :505 */
            pair_set_insert_raw(&wu->collisions, key);                     /* :507 */
    }
    w->n_cand = 0; /* drain(0..) */
}

typedef struct { orc_system *s; int id; } worker_arg;

/* Worker::start (grid_collision.rs:363-387) */
static void *worker_main(void *p)
{
    worker_arg *arg = p;
    orc_system *s = arg->s;
    worker_state *w = &s->workers[arg->id];
    free(arg);
    for (;;) {
        pthread_mutex_lock(&s->mu);
        while (s->n_pending == 0 && !s->shutdown) pthread_cond_wait(&s->cv_pending, &s->mu);
        if (s->shutdown) { pthread_mutex_unlock(&s->mu); return NULL; }
        const int uid = s->pending[--s->n_pending]; /* pending.pop() :374 */
        pthread_mutex_unlock(&s->mu);

        work_unit *wu = &s->units[uid];
        do_broadphase(s, wu, w);
        do_narrowphase(wu, w);

        pthread_mutex_lock(&s->mu); /* channel.send(work) :385 */
        s->done[s->n_done++] = uid;
        pthread_cond_signal(&s->cv_done);
        pthread_mutex_unlock(&s->mu);
    }
}

static orc_aabb mk_bounds(float x0, float y0, float z0, float x1, float y1, float z1)
{
    orc_aabb b = { { x0, y0, z0, 1.0f }, { x1, y1, z1, 1.0f } };
    return b;
}

/* GridCollisionSystem::new (grid_collision.rs:122-216) */
orc_system *orc_system_new(int num_units, int num_workers)
{
    if (!(num_units == 1 || num_units == 2 || num_units == 4 || num_units == 8) || num_workers < 1) return NULL;
    orc_system *s = calloc(1, sizeof *s);
    s->num_units = num_units;
    s->num_workers = num_workers;
    s->units = calloc((size_t)num_units, sizeof(work_unit));
    const float MIN = -FLT_MAX, MAX = FLT_MAX; /* std::f32::{MIN,MAX} */
    orc_aabb b[8];
    if (num_units == 1) {
        /* :130-134 -- NOTE the reference's single work unit is the all-negative octant only; restated verbatim. */
        b[0] = mk_bounds(MIN, MIN, MIN, 0.0f, 0.0f, 0.0f);
    } else if (num_units == 2) { /* :135-143 */
        b[0] = mk_bounds(MIN, MIN, MIN, 0.0f, MAX, MAX);
        b[1] = mk_bounds(0.0f, MIN, MIN, MAX, MAX, MAX);
    } else if (num_units == 4) { /* :144-160 */
        b[0] = mk_bounds(MIN, MIN, MIN, 0.0f, 0.0f, MAX);
        b[1] = mk_bounds(MIN, 0.0f, MIN, 0.0f, MAX, MAX);
        b[2] = mk_bounds(0.0f, MIN, MIN, MAX, 0.0f, MAX);
        b[3] = mk_bounds(0.0f, 0.0f, MIN, MAX, MAX, MAX);
    } else { /* :161-193 */
        b[0] = mk_bounds(MIN, MIN, MIN, 0.0f, 0.0f, 0.0f);
        b[1] = mk_bounds(MIN, MIN, 0.0f, 0.0f, 0.0f, MAX);
        b[2] = mk_bounds(MIN, 0.0f, MIN, 0.0f, MAX, 0.0f);
        b[3] = mk_bounds(MIN, 0.0f, 0.0f, 0.0f, MAX, MAX);
        b[4] = mk_bounds(0.0f, MIN, MIN, MAX, 0.0f, 0.0f);
        b[5] = mk_bounds(0.0f, MIN, 0.0f, MAX, 0.0f, MAX);
        b[6] = mk_bounds(0.0f, 0.0f, MIN, MAX, MAX, 0.0f);
        b[7] = mk_bounds(0.0f, 0.0f, 0.0f, MAX, MAX, MAX);
    }
    for (int i = 0; i < num_units; ++i) {
        s->units[i].bounds = b[i];
        s->units[i].cell_size = 1.0f; /* :319 */
        pair_set_init(&s->units[i].collisions);
        grid_init(&s->units[i].grid);
    }
    pair_set_init(&s->collisions);
    s->workers = calloc((size_t)num_workers, sizeof(worker_state));
    s->pending = calloc((size_t)num_units, sizeof(int));
    s->done = calloc((size_t)num_units, sizeof(int));
    pthread_mutex_init(&s->mu, NULL);
    pthread_cond_init(&s->cv_pending, NULL);
    pthread_cond_init(&s->cv_done, NULL);
    s->threads = calloc((size_t)num_workers, sizeof(pthread_t));
    for (int i = 0; i < num_workers; ++i) { /* :198-207 persistent workers */
        worker_arg *a = malloc(sizeof *a);
        a->s = s;
        a->id = i;
        pthread_create(&s->threads[i], NULL, worker_main, a);
    }
    return s;
}

void orc_system_free(orc_system *s)
{
    if (!s) return;
    pthread_mutex_lock(&s->mu);
    s->shutdown = 1;
    pthread_cond_broadcast(&s->cv_pending);
    pthread_mutex_unlock(&s->mu);
    for (int i = 0; i < s->num_workers; ++i) pthread_join(s->threads[i], NULL);
    for (int i = 0; i < s->num_units; ++i) {
        pair_set_free(&s->units[i].collisions);
        cell_grid *g = &s->units[i].grid;
        for (uint64_t k = 0; k < g->cap; ++k) if (g->slot[k].key) free(g->slot[k].vec.p);
        free(g->slot);
    }
    for (int i = 0; i < s->num_workers; ++i) {
        for (uint64_t k = 0; k < s->workers[i].n_cache; ++k) free(s->workers[i].cell_cache[k].p);
        free(s->workers[i].cell_cache);
        free(s->workers[i].cand);
    }
    pair_set_free(&s->collisions);
    free(s->units); free(s->workers); free(s->threads); free(s->pending); free(s->done);
    free(s->components); free(s->volumes);
    pthread_mutex_destroy(&s->mu);
    pthread_cond_destroy(&s->cv_pending);
    pthread_cond_destroy(&s->cv_done);
    free(s);
}

/* bvh_update (bounding_volume.rs:345-409).  The entity->index HashMap lookups (:359, :394) are
 * replaced by the dense iteration order itself (collider i <-> volume i), which is what they resolve
 * to when colliders are assigned once and never destroyed. collision_region (:384-390) is a dead
 * output (never read by the grid) and is not restated. */
int orc_bvh_update(orc_system *s, uint32_t n, const uint32_t *entity, const uint8_t *kind,
                   const float *offset3, const float *size3,
                   const float *pos3, const float *quat4, const float *scale3)
{
    if (n > s->cap) {
        s->cap = n;
        s->components = realloc(s->components, (size_t)n * sizeof *s->components);
    }
    s->n = n;
    s->longest_axis = 0.0f; /* :352 */
    for (uint32_t i = 0; i < n; ++i) {
        if (kind[i] != ORC_SPHERE && kind[i] != ORC_BOX) return -1;
        orc_bound_volume *bv = &s->components[i];
        bv->entity = entity[i];
        orc_cache_collider(kind[i], offset3 + 3 * i, size3 + 3 * i, pos3 + 3 * i, quat4 + 4 * i, scale3 + 3 * i, &bv->collider);
        orc_aabb_from_collider(&bv->collider, &bv->aabb);
        const float dx = bv->aabb.max.x - bv->aabb.min.x; /* :365-381 */
        const float dy = bv->aabb.max.y - bv->aabb.min.y;
        const float dz = bv->aabb.max.z - bv->aabb.min.z;
        if (dx > s->longest_axis) s->longest_axis = dx;
        if (dy > s->longest_axis) s->longest_axis = dy;
        if (dz > s->longest_axis) s->longest_axis = dz;
    }
    return 0;
}

float orc_longest_axis(const orc_system *s) { return s->longest_axis; }
uint32_t orc_num_volumes(const orc_system *s) { return s->n; }
const orc_bound_volume *orc_components(const orc_system *s) { return s->components; }

/* GridCollisionSystem::update (grid_collision.rs:218-285) */
uint64_t orc_grid_update(orc_system *s)
{
    pair_set_clear(&s->collisions); /* :221 */
    for (int i = 0; i < s->num_units; ++i) {
        s->units[i].cell_size = s->longest_axis; /* :238-240 */
        s->units[i].n_candidates = s->units[i].n_cell_entries = 0;
        s->units[i].n_wide = s->units[i].n_oor = 0;
    }
    /* :243-244  volumes.clone_from(bvh_manager.components()) */
    if (s->n > s->vol_cap) {
        s->vol_cap = s->n;
        s->volumes = realloc(s->volumes, (size_t)s->n * sizeof *s->volumes);
    }
    if (s->n) memcpy(s->volumes, s->components, (size_t)s->n * sizeof *s->volumes);

    pthread_mutex_lock(&s->mu); /* :246-258 */
    for (int i = 0; i < s->num_units; ++i) s->pending[i] = i;
    s->n_pending = s->num_units;
    s->n_done = 0;
    pthread_cond_broadcast(&s->cv_pending);
    pthread_mutex_unlock(&s->mu);

    int processed = 0; /* :262-272  recv + serial merge */
    while (processed < s->num_units) {
        pthread_mutex_lock(&s->mu);
        while (s->n_done == 0) pthread_cond_wait(&s->cv_done, &s->mu);
        const int uid = s->done[--s->n_done];
        pthread_mutex_unlock(&s->mu);
        work_unit *wu = &s->units[uid];
        pair_set *ps = &wu->collisions;
        for (uint64_t k = 0; k < ps->cap; ++k) { /* drain() + insert */
            if (!ps->slot[k]) continue;
            pair_set_insert_raw(&s->collisions, ps->slot[k] - 1);
            ps->slot[k] = 0;
        }
        ps->len = 0;
        ++processed;
    }
    s->last_candidates = s->last_entries = 0;
    s->last_wide = s->last_oor = 0;
    for (int i = 0; i < s->num_units; ++i) {
        s->last_candidates += s->units[i].n_candidates;
        s->last_entries += s->units[i].n_cell_entries;
        s->last_wide += s->units[i].n_wide;
        s->last_oor += s->units[i].n_oor;
    }
    return s->collisions.len;
}

static int cmp_u64(const void *a, const void *b)
{
    const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

uint64_t orc_collisions_sorted(orc_system *s, uint64_t *out, uint64_t cap)
{
    uint64_t n = 0;
    for (uint64_t k = 0; k < s->collisions.cap; ++k) {
        if (!s->collisions.slot[k]) continue;
        if (n < cap) out[n] = s->collisions.slot[k] - 1;
        ++n;
    }
    if (n <= cap) qsort(out, n, sizeof(uint64_t), cmp_u64);
    return n;
}

uint64_t orc_last_candidates(const orc_system *s) { return s->last_candidates; }
uint64_t orc_last_cell_entries(const orc_system *s) { return s->last_entries; }
uint32_t orc_last_wide_volumes(const orc_system *s) { return s->last_wide; }
uint32_t orc_last_out_of_range(const orc_system *s) { return s->last_oor; }

/* CollisionCallbackManager::process_collisions, the consumer of the pair set (mod.rs:678-708): every pair
 * (entity, other) appends `other` to entity's list and `entity` to other's list (:687-690); callbacks then get
 * (entity, &[others]) per entity with a non-empty list (:694-705).  HashMap iteration order is unspecified in the
 * reference, so the lists are reported sorted: entities ascending, each list ascending.
 * Returns the number of entities; entity[cap_e], offset[cap_e + 1], other[2 * n_pairs]. */
static int cmp_u64_pair(const void *a, const void *b)
{
    const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}
uint64_t orc_process_collisions(orc_system *s, uint32_t *entity, uint32_t *offset, uint32_t *other, uint64_t cap_entities)
{
    const uint64_t h = orc_collisions_sorted(s, NULL, 0);
    uint64_t *pairs = malloc((h ? h : 1) * sizeof(uint64_t));
    uint64_t *ent = malloc((h ? 2 * h : 1) * sizeof(uint64_t));
    orc_collisions_sorted(s, pairs, h);
    for (uint64_t i = 0; i < h; ++i) {
        const uint64_t a = pairs[i] >> 32, b = pairs[i] & 0xffffffffu;
        ent[2 * i] = (a << 32) | b;      /* a's list gets b */
        ent[2 * i + 1] = (b << 32) | a;  /* b's list gets a */
    }
    qsort(ent, 2 * h, sizeof(uint64_t), cmp_u64_pair);
    uint64_t n = 0;
    for (uint64_t i = 0; i < 2 * h; ++i) {
        if (i == 0 || (ent[i] >> 32) != (ent[i - 1] >> 32)) {
            if (n < cap_entities) { entity[n] = (uint32_t)(ent[i] >> 32); offset[n] = (uint32_t)i; }
            ++n;
        }
        other[i] = (uint32_t)(ent[i] & 0xffffffffu);
    }
    if (n <= cap_entities) offset[n] = (uint32_t)(2 * h);
    free(pairs);
    free(ent);
    return n;
}

/* SURVEY F9 second oracle */
uint64_t orc_brute_force_sorted(orc_system *s, uint64_t *out, uint64_t cap)
{
    uint64_t n = 0;
    for (uint32_t i = 1; i < s->n; ++i)
        for (uint32_t j = 0; j < i; ++j)
            if (orc_test_volume(&s->components[i], &s->components[j])) {
                if (n < cap) out[n] = ((uint64_t)s->components[i].entity << 32) | s->components[j].entity;
                ++n;
            }
    if (n <= cap) qsort(out, n, sizeof(uint64_t), cmp_u64);
    return n;
}

/* ======================================================================================================
 * Transform derivation -- the step BEFORE the collision path.
 * TransformManager::update_transforms (component/transform.rs:243-251) walks the rows (row = depth) and calls
 * TransformData::update (:588-600) on every node.  The reference holds no test for it: PARITY UNPINNED beyond
 * the restated arithmetic and the identity-parent property.
 * `Matrix4::from_quaternion` (:604) no longer exists in lib/polygon_math; the only in-tree quaternion -> Matrix4
 * definition is `From<Orientation> for Matrix4` (matrix.rs:160-172), which is what is restated here.
 * ====================================================================================================== */
typedef struct { float m[4][4]; } orc_mat4;

/* matrix.rs:207-243: every element is (a0*b0) + (a1*b1) + (a2*b2) + (a3*b3), summed left to right */
static orc_mat4 m4_mul(const orc_mat4 *a, const orc_mat4 *b)
{
    orc_mat4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.m[i][j] = (a->m[i][0] * b->m[0][j]) + (a->m[i][1] * b->m[1][j]) + (a->m[i][2] * b->m[2][j]) + (a->m[i][3] * b->m[3][j]);
    return r;
}

static orc_mat4 m4_identity(void)
{
    orc_mat4 r;
    memset(&r, 0, sizeof r);
    r.m[0][0] = r.m[1][1] = r.m[2][2] = r.m[3][3] = 1.0f;
    return r;
}

/* TransformData::local_matrix (transform.rs:602-608): position * (rotation * scale) */
static orc_mat4 transform_local_matrix(const float p[3], const float q[4], const float s[3])
{
    orc_mat4 position = m4_identity(), rotation = m4_identity(), scale = m4_identity();
    position.m[0][3] = p[0]; position.m[1][3] = p[1]; position.m[2][3] = p[2]; /* matrix.rs:44-56 */
    orc_mat3 r3;
    orc_quat_to_mat3(q, &r3);                                                   /* matrix.rs:160-172, same formulas as :392-402 */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) rotation.m[i][j] = r3.m[i][j];
    scale.m[0][0] = s[0]; scale.m[1][1] = s[1]; scale.m[2][2] = s[2];           /* matrix.rs:103-110 */
    const orc_mat4 rs = m4_mul(&rotation, &scale);
    return m4_mul(&position, &rs);
}

/* row_begin[n_rows + 1]; parent[i] = node of the previous row or 0xffffffff (the identity dummy, transform.rs:47-61).
 * Outputs per node: pos3, quat4 (x,y,z,w), scale3, matrix16 (row-major).  Rows in order, as update_transforms. */
void orc_transform_update(uint32_t n_rows, const uint32_t *row_begin, const uint32_t *parent, const float *lpos3,
                          const float *lquat4, const float *lscale3, float *dpos3, float *dquat4, float *dscale3,
                          float *dmat16)
{
    for (uint32_t r = 0; r < n_rows; ++r)
        for (uint32_t i = row_begin[r]; i < row_begin[r + 1]; ++i) {
            orc_mat4 pm = m4_identity();
            float pq[4] = { 0.0f, 0.0f, 0.0f, 1.0f }, ps[3] = { 1.0f, 1.0f, 1.0f };
            if (parent[i] != 0xffffffffu) {
                const uint32_t p = parent[i];
                memcpy(&pm, dmat16 + 16 * (size_t)p, sizeof pm);
                memcpy(pq, dquat4 + 4 * (size_t)p, sizeof pq);
                memcpy(ps, dscale3 + 3 * (size_t)p, sizeof ps);
            }
            const float *q = lquat4 + 4 * (size_t)i, *s = lscale3 + 3 * (size_t)i;
            const orc_mat4 local = transform_local_matrix(lpos3 + 3 * (size_t)i, q, s);
            const orc_mat4 derived = m4_mul(&pm, &local);                       /* :593 */
            memcpy(dmat16 + 16 * (size_t)i, &derived, sizeof derived);
            dpos3[3 * (size_t)i] = derived.m[0][3];                             /* :595, matrix.rs:137-139 */
            dpos3[3 * (size_t)i + 1] = derived.m[1][3];
            dpos3[3 * (size_t)i + 2] = derived.m[2][3];
            /* :596 parent.rotation_derived * self.rotation; quaternion.rs:164-169:
             *   v = cross(a.v, b.v) + b.w * a.v + a.w * b.v ; w = a.w * b.w - dot(a.v, b.v)   (vector.rs:66-72, 190-196) */
            const float ax = pq[0], ay = pq[1], az = pq[2], aw = pq[3], bx = q[0], by = q[1], bz = q[2], bw = q[3];
            const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
            dquat4[4 * (size_t)i] = (cx + ax * bw) + bx * aw;
            dquat4[4 * (size_t)i + 1] = (cy + ay * bw) + by * aw;
            dquat4[4 * (size_t)i + 2] = (cz + az * bw) + bz * aw;
            dquat4[4 * (size_t)i + 3] = aw * bw - (ax * bx + ay * by + az * bz);
            /* :597 self.scale * parent.scale_derived */
            dscale3[3 * (size_t)i] = s[0] * ps[0];
            dscale3[3 * (size_t)i + 1] = s[1] * ps[1];
            dscale3[3 * (size_t)i + 2] = s[2] * ps[2];
        }
}

/* ======================================================================================================
 * Collider lifecycle: BoundingVolumeManager::destroy_immediate (bounding_volume.rs:82-104) is a
 * swap_remove of the dense `entities` / `components` arrays: the last element moves into the hole (nothing moves
 * when the hole IS the last element).  `order[n]` stands for the dense arrays (one id per element); removals are
 * applied in the given order, each index addressing the array as it is at that moment.  Returns the new length,
 * or (uint32_t)-1 on an out-of-range index (the reference's `indices.remove` would have returned None).
 * ====================================================================================================== */
uint32_t orc_swap_remove_sequence(uint32_t *order, uint32_t n, const uint32_t *index, uint32_t n_remove)
{
    for (uint32_t k = 0; k < n_remove; ++k) {
        const uint32_t i = index[k];
        if (i >= n) return (uint32_t)-1;
        order[i] = order[n - 1]; /* Vec::swap_remove */
        --n;
    }
    return n;
}

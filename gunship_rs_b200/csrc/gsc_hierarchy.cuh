// gsc_hierarchy.cuh -- transform derivation, the step BEFORE the collision path.
//
// TransformManager::update_transforms (component/transform.rs:243-251) walks the hierarchy row by row (row =
// depth) and the reference itself notes that "the transforms in a row can be processed independently so they
// should be done in parallel": here a row is one kernel launch, a node one thread.  TransformData::update
// (:588-600) per node:
//     matrix_derived   = parent.matrix_derived * (T(position) * (R(rotation) * S(scale)))
//     position_derived = matrix_derived.translation_part()
//     rotation_derived = parent.rotation_derived * rotation
//     scale_derived    = scale * parent.scale_derived
// The Matrix4 products keep all four multiply-adds of every element in the order matrix.rs:207-243 writes them
// (multiplications by the literal 0.0 / 1.0 entries of T, R, S included), so the result is bit-identical to the
// reference's f32 evaluation, signs of zero and all.  The kernels are HBM streams: 40 B in (+ 92 B of the parent
// through L2) and 104 B out per node.
#pragma once
#include "gsc_math.cuh"

namespace gsc {

constexpr uint32_t kNoParent = 0xffffffffu;
constexpr int kHierThreads = 128;

struct Mat4 {
    float m[4][4];
};

// matrix.rs:207-243
GSC_DEV Mat4 mat4_mul(const Mat4 &a, const Mat4 &b)
{
    Mat4 r;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            r.m[i][j] = add(add(add(mul(a.m[i][0], b.m[0][j]), mul(a.m[i][1], b.m[1][j])), mul(a.m[i][2], b.m[2][j])),
                            mul(a.m[i][3], b.m[3][j]));
    return r;
}

GSC_DEV Mat4 mat4_identity()
{
    Mat4 r;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) r.m[i][j] = i == j ? 1.0f : 0.0f;
    return r;
}

// Per node, device resident:  lpos3 / lquat / lscale3 (local, as uploaded)  ->  dpos3 / dquat / dscale3 / dmat (derived;
// dmat = 4 x float4 rows).  Nodes [begin, end) are one row; their parents lie in the previous row, already derived.
__global__ void __launch_bounds__(kHierThreads)
k_transform_row(uint32_t begin, uint32_t end, const uint32_t *__restrict__ parent, const float *__restrict__ lpos3,
                const float4 *__restrict__ lquat, const float *__restrict__ lscale3, float *__restrict__ dpos3,
                float4 *__restrict__ dquat, float *__restrict__ dscale3, float4 *__restrict__ dmat)
{
    const uint32_t i = begin + blockIdx.x * kHierThreads + threadIdx.x;
    if (i >= end) return;
    const float px = lpos3[3 * (size_t)i], py = lpos3[3 * (size_t)i + 1], pz = lpos3[3 * (size_t)i + 2];
    const float4 q = lquat[i];
    const float sx = lscale3[3 * (size_t)i], sy = lscale3[3 * (size_t)i + 1], sz = lscale3[3 * (size_t)i + 2];

    // the dummy parent of a root: identity everything (transform.rs:47-61)
    Mat4 pm = mat4_identity();
    float4 pq = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
    float psx = 1.0f, psy = 1.0f, psz = 1.0f;
    const uint32_t p = parent[i];
    if (p != kNoParent) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float4 row = dmat[4 * (size_t)p + r];
            pm.m[r][0] = row.x; pm.m[r][1] = row.y; pm.m[r][2] = row.z; pm.m[r][3] = row.w;
        }
        pq = dquat[p];
        psx = dscale3[3 * (size_t)p]; psy = dscale3[3 * (size_t)p + 1]; psz = dscale3[3 * (size_t)p + 2];
    }

    // local_matrix (transform.rs:602-608) = position * (rotation * scale)
    Mat4 T = mat4_identity(), R = mat4_identity(), S = mat4_identity();
    T.m[0][3] = px; T.m[1][3] = py; T.m[2][3] = pz;  // matrix.rs:44-56
    float r3[3][3];
    quat_to_mat3(q.x, q.y, q.z, q.w, r3);            // matrix.rs:160-172 (same formulas as :392-402)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) R.m[a][b] = r3[a][b];
    S.m[0][0] = sx; S.m[1][1] = sy; S.m[2][2] = sz;  // matrix.rs:103-110
    const Mat4 local = mat4_mul(T, mat4_mul(R, S));
    const Mat4 d = mat4_mul(pm, local);               // transform.rs:593
#pragma unroll
    for (int r = 0; r < 4; ++r) dmat[4 * (size_t)i + r] = make_float4(d.m[r][0], d.m[r][1], d.m[r][2], d.m[r][3]);
    dpos3[3 * (size_t)i] = d.m[0][3];                  // :595 translation_part (matrix.rs:137-139)
    dpos3[3 * (size_t)i + 1] = d.m[1][3];
    dpos3[3 * (size_t)i + 2] = d.m[2][3];
    // :596 parent.rotation_derived * rotation; quaternion.rs:164-169: v = cross(a.v, b.v) + b.w * a.v + a.w * b.v,
    // w = a.w * b.w - dot(a.v, b.v)
    const float cx = sub(mul(pq.y, q.z), mul(pq.z, q.y));
    const float cy = sub(mul(pq.z, q.x), mul(pq.x, q.z));
    const float cz = sub(mul(pq.x, q.y), mul(pq.y, q.x));
    float4 dq;
    dq.x = add(add(cx, mul(pq.x, q.w)), mul(q.x, pq.w));
    dq.y = add(add(cy, mul(pq.y, q.w)), mul(q.y, pq.w));
    dq.z = add(add(cz, mul(pq.z, q.w)), mul(q.z, pq.w));
    dq.w = sub(mul(pq.w, q.w), dot3(pq.x, pq.y, pq.z, q.x, q.y, q.z));
    dquat[i] = dq;
    // :597 scale * parent.scale_derived
    dscale3[3 * (size_t)i] = mul(sx, psx);
    dscale3[3 * (size_t)i + 1] = mul(sy, psy);
    dscale3[3 * (size_t)i + 2] = mul(sz, psz);
}

// collider k reads the derived transform of node collider_node[k] (mod.rs:311-333): gather into the context's
// per-collider transform arrays, the layout bvh_update streams
__global__ void __launch_bounds__(256)
k_hier_bind(uint32_t n, const uint32_t *__restrict__ collider_node, const float *__restrict__ dpos3,
            const float4 *__restrict__ dquat, const float *__restrict__ dscale3, float *__restrict__ pos3,
            float4 *__restrict__ quat, float *__restrict__ scale3)
{
    const uint32_t k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    const size_t j = collider_node[k];
    pos3[3 * (size_t)k] = dpos3[3 * j];
    pos3[3 * (size_t)k + 1] = dpos3[3 * j + 1];
    pos3[3 * (size_t)k + 2] = dpos3[3 * j + 2];
    quat[k] = dquat[j];
    scale3[3 * (size_t)k] = dscale3[3 * j];
    scale3[3 * (size_t)k + 1] = dscale3[3 * j + 1];
    scale3[3 * (size_t)k + 2] = dscale3[3 * j + 2];
}

// lifecycle: the net effect of a sequence of swap_removes on the dense static arrays is a set of
// moves dst <- src with every src >= new length > every dst (gsc_remove_colliders), so one gather pass is hazard free
__global__ void __launch_bounds__(256)
k_apply_moves(uint32_t n_moves, const uint2 *__restrict__ moves, uint8_t *__restrict__ kind, float *__restrict__ offset3,
              float *__restrict__ size3, uint32_t *__restrict__ entity)
{
    const uint32_t m = blockIdx.x * 256 + threadIdx.x;
    if (m >= n_moves) return;
    const size_t d = moves[m].x, s = moves[m].y;
    kind[d] = kind[s];
    entity[d] = entity[s];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        offset3[3 * d + a] = offset3[3 * s + a];
        size3[3 * d + a] = size3[3 * s + a];
    }
}

}  // namespace gsc

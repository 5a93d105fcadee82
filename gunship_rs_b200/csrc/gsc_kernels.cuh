// gsc_kernels.cuh -- the collision step as hand-written sm_100a kernels.
//
// Stage map (reference -> kernel):
//   bvh_update (bounding_volume.rs:345-409)                      -> k_bvh_update
//   longest_axis / cell size, grid extent                        -> k_grid_setup
//   WorkUnit::world_to_grid (grid_collision.rs:329-335)          -> k_cell_keys (+ digit histograms)
//   HashMap<GridCell, Vec<..>> grouping (grid_collision.rs:434)  -> k_radix_upsweep/scan/downsweep (gsc_sort.cuh)
//   cell lookup                                                  -> k_cell_table_permute
//   test_cell candidate generation + HashMap dedupe (:439-441,
//   :497-511) + BoundVolume::test (bounding_volume.rs:124-132)   -> k_pairs, k_narrow, k_wide_pairs
//
// Why neighbour search instead of the reference's 8-corner insertion: with cell = longest AABB
// extent every AABB spans <= 2 cells per axis, so two volumes share a corner cell iff their home
// cells (cell of aabb.min) differ by at most one per axis AND their AABBs allow it; the AABB test
// itself is the exact filter.  Each volume is therefore entered ONCE, under its home
// cell, and looks at the 13 "earlier" neighbour cells plus its own.  Volumes that rounding makes
// span 3 cells ("wide", grid_collision.rs:403-410) leave that argument and are handled exactly by
// k_wide_pairs, which restates the corner-cell rule literally.
#pragma once
#include <limits.h>

#include "gsc_math.cuh"
#include "gsc_sort.cuh"

namespace gsc {

// ---- status bits (FrameBlock::status) ----
enum : uint32_t {
    kStNonFinite = 1u << 0,
    kStMesh = 1u << 1,
    kStRange = 1u << 2,
    kStCellSize = 1u << 3,
    kStCells = 1u << 4,
    kStPairs = 1u << 5,
    kStCands = 1u << 6,
    kStWindow = 1u << 7,
    kStPeerTimeout = 1u << 8,  // a peer GPU did not publish its bounds / finish its halo push in time
    kStGhostCap = 1u << 9,     // local + pushed ghost volumes exceed the context's capacity
    kStWideBorder = 1u << 10,  // a 3-cell-wide volume touches a slab border: its cross-slab pairs would be lost
    kStCheck = 1u << 11,       // -DGSC_CHECKED builds only: an internal bounds check failed (see GSC_CHECK)
};

// compute-sanitizer is closed on the pool this was developed on, so the kernels carry their own bounds checks for the
// places that do raw address arithmetic (shared-memory slot addresses, staged indices, scatter destinations, peer
// writes).  A -DGSC_CHECKED build (tools/checked_build.sh) evaluates them and reports a failure through the status word
// -- the step then fails with GSC_ERR_CUDA -- instead of trapping; the production build compiles them away.
#ifdef GSC_CHECKED
#define GSC_CHECK(fb_, cond) do { if (!(cond)) atomicOr(&(fb_)->status, kStCheck); } while (0)
#else
#define GSC_CHECK(fb_, cond) do { } while (0)
#endif

// ---- volume record: 32 B = 2 x float4, the unit the broadphase streams ----
//   sphere: r0 = (cx, cy, cz, radius)          r1 = (entity bits, 0, global index bits, meta)
//   box:    r0 = (min.x, min.y, min.z, max.x)  r1 = (max.y, max.z, global index bits, meta)
//   meta = local index << 4 | flags
constexpr uint32_t kMetaBox = 1u, kMetaGhost = 2u;

struct Volume {
    float mn[3], mx[3];
    float cx, cy, cz, r;  // sphere only
    uint32_t entity;      // sphere only
    uint32_t gidx, meta;
};

GSC_DEV Volume volume_decode(const float4 r0, const float4 r1)
{
    Volume v;
    v.gidx = __float_as_uint(r1.z);
    v.meta = __float_as_uint(r1.w);
    if (v.meta & kMetaBox) {
        v.mn[0] = r0.x; v.mn[1] = r0.y; v.mn[2] = r0.z;
        v.mx[0] = r0.w; v.mx[1] = r1.x; v.mx[2] = r1.y;
        v.cx = v.cy = v.cz = v.r = 0.0f;
        v.entity = 0;
    } else {
        // bounding_volume.rs:150-159: min = centre - (r,r,r), max = centre + (r,r,r)
        v.cx = r0.x; v.cy = r0.y; v.cz = r0.z; v.r = r0.w;
        v.mn[0] = sub(r0.x, r0.w); v.mn[1] = sub(r0.y, r0.w); v.mn[2] = sub(r0.z, r0.w);
        v.mx[0] = add(r0.x, r0.w); v.mx[1] = add(r0.y, r0.w); v.mx[2] = add(r0.z, r0.w);
        v.entity = __float_as_uint(r1.x);
    }
    return v;
}

struct GridParams {
    float cell_size;
    int32_t origin[3];
    int32_t dims[3];
    uint32_t num_cells;  // dims product; also the key given to wide volumes
    int32_t valid;
    int32_t windowed;  // the grid covers one x-slab of the world (multi-GPU step)
    float qorg[3], qscale[3], qmargin;  // conservative box quantisation of k_pairs: t = (v - qorg) * qscale, qscale = 128 / cell (x, y), 64 / cell (z)
};

// ---- multi-GPU slab exchange without the host (k_slab_*) ----
constexpr int kMaxPeers = 16;  // == GSC_MAX_PEERS

// what k_slab_gather works out for this rank from all ranks' bounds (absolute home x-cells)
struct SlabState {
    int32_t push_lo[kMaxPeers], push_hi[kMaxPeers];  // my volumes with home x-cell in [lo, hi] go to peer s (lo > hi: none)
    uint32_t peer_n_local[kMaxPeers];                // where peer s's ghost region starts
    int32_t x_lo[kMaxPeers], x_hi[kMaxPeers];        // every rank's home x-range (lo > hi: no colliders)
    uint32_t n_ghost;                                // ghosts adopted this step (k_slab_wait)
    uint32_t pad;
};

// Mailbox of a context, written by the PEERS over NVLink / NVSwitch peer memory (one slot per sender), never zeroed:
// `seq` / `done_seq` carry the step number, so a slot is valid exactly when its sequence equals the current step.
struct SlabMailSlot {
    uint32_t bounds[8];  // the sender's FrameBlock::bounds after its bvh_update
    uint32_t n_local;
    uint32_t status;
    uint32_t pad[5];
    uint32_t seq;        // written last (release)
};
struct SlabMail {
    SlabMailSlot from[kMaxPeers];
    uint32_t done_seq[kMaxPeers];  // sender s has finished pushing its halo into this context for step done_seq[s]
};
struct FrameBlock;
struct SlabPeers {  // kernel parameter: the mapped arrays of every rank (own entry included)
    float4 *rec[kMaxPeers];
    float4 *obb[kMaxPeers];
    FrameBlock *fb[kMaxPeers];
    SlabMail *mail[kMaxPeers];
    uint32_t cap[kMaxPeers];
};

// Per-step device control block.  Zeroed at the start of every step.
struct FrameBlock {
    uint32_t bounds[8];  // [0] ord(longest axis); [1..3] ~ord(world min xyz); [4..6] ord(world max xyz); [7] ord(max aabb.min.x)
    uint32_t status;
    uint32_t n_wide;
    unsigned long long n_pairs, n_ss, n_so, n_oo;
    uint32_t xrange[2];  // ~(min home x + 32768), (max home x + 32768) of local volumes
    uint32_t n_gaps;      // long empty key gaps queued for k_fill_gaps
    uint32_t n_ghost_in;  // ghosts written into this context by peer GPUs this step (k_halo_push, system-scope atomic)
    uint32_t narrow_ticket;  // work queue of k_narrow (chunks of the three candidate lists)
    uint32_t n_total;     // local + ghost volumes of this step (k_grid_setup); kernels launched for a capacity clamp to it
    uint32_t n_overflow;  // tiles k_pairs_tiled could not stage (queued for k_pairs)
    uint32_t pad1;
    unsigned long long reserved0;
    GridParams grid;
    SortControl sort;
    SlabState slab;
};

// ---- block-aggregated stream compaction ------------------------------------------------------
// Hits are ballot-compacted per warp into a warp-private shared-memory buffer; when the CTA is
// done ONE atomicAdd per stream reserves the CTA's slice of the global output and every warp
// copies its buffer out.  (A warp whose buffer fills mid-kernel spills with a warp-level atomic.)
template <int WARPS, int CAP>
struct EmitBuf {
    uint2 items[WARPS][CAP];
    float4 extra[WARPS][CAP];  // contact (normal.xyz, depth) travelling with the pair (GSC_FLAG_CONTACTS)
    uint32_t count[WARPS];
    unsigned long long base;
};

template <int WARPS, int CAP>
GSC_DEV void emit_init(EmitBuf<WARPS, CAP> &b)
{
    if (threadIdx.x < WARPS) b.count[threadIdx.x] = 0;
}

// `out_extra` == nullptr: pairs only
template <int WARPS, int CAP>
GSC_DEV void emit(EmitBuf<WARPS, CAP> &b, bool pred, uint2 item, float4 extra, uint2 *__restrict__ out,
                  float4 *__restrict__ out_extra, unsigned long long *counter, unsigned long long cap)
{
    const uint32_t mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cnt = __popc(mask);
    uint32_t have = b.count[warp];
    if (have + cnt > CAP) {  // spill this warp's buffer
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)have);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (uint32_t i = lane; i < have; i += 32)
            if (base + i < cap) {
                out[base + i] = b.items[warp][i];
                if (out_extra) out_extra[base + i] = b.extra[warp][i];
            }
        have = 0;
        __syncwarp();
    }
    if (pred) {
        const uint32_t at = have + __popc(mask & ((1u << lane) - 1u));
        b.items[warp][at] = item;
        if (out_extra) b.extra[warp][at] = extra;
    }
    __syncwarp();
    if (lane == 0) b.count[warp] = have + cnt;
    __syncwarp();
}

template <int WARPS, int CAP>
GSC_DEV void emit_finish(EmitBuf<WARPS, CAP> &b, uint2 *__restrict__ out, float4 *__restrict__ out_extra,
                         unsigned long long *counter, unsigned long long cap)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (int w = 0; w < WARPS; ++w) total += b.count[w];
        b.base = total ? atomicAdd(counter, (unsigned long long)total) : 0ull;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long base = b.base;
    for (int w = 0; w < warp; ++w) base += b.count[w];
    const uint32_t have = b.count[warp];
    for (uint32_t i = lane; i < have; i += 32)
        if (base + i < cap) {
            out[base + i] = b.items[warp][i];
            if (out_extra) out_extra[base + i] = b.extra[warp][i];
        }
}

// ---- stage 0: bvh_update -----------------------------------------------------------------------
// One thread per collider: CachedCollider::from_collider_transform (mod.rs:311-333) +
// AABB::from_collider (bounding_volume.rs:148-336) + the longest-axis fold (:365-381) and the world
// bounds (needed to place the dense grid), reduced warp -> CTA -> one atomic per CTA.
constexpr int kBvhThreads = 256;

#ifndef GSC_BVH_MIN_CTAS
#define GSC_BVH_MIN_CTAS 8  // 6 (40 registers) / 8 (32 registers, 56 B of spills): 0.398 / 0.394 ms at 16M
#endif
__global__ void __launch_bounds__(kBvhThreads, GSC_BVH_MIN_CTAS)
k_bvh_update(uint32_t first, uint32_t n, const uint8_t *__restrict__ kind, const float *__restrict__ offset3,
             const float *__restrict__ size3, const float *__restrict__ pos3, const float *__restrict__ quat4,
             const float *__restrict__ scale3, const uint32_t *__restrict__ entity,
             const uint32_t *__restrict__ gidx, const uint32_t *__restrict__ box_slot, float4 *__restrict__ rec,
             float4 *__restrict__ obb, FrameBlock *__restrict__ fb)
{
    // box_slot != nullptr: quat4 / scale3 are packed over the boxes only (gsc_update_transforms_packed);
    // box_slot[i] = number of boxes before collider i
    // colliders [first, n): the host layer launches one grid per arriving chunk of a host-to-device transform copy
    const uint32_t i = first + blockIdx.x * kBvhThreads + threadIdx.x;
    float ext = 0.0f, mnx_max = -INFINITY;
    float mn[3] = { INFINITY, INFINITY, INFINITY }, mx[3] = { -INFINITY, -INFINITY, -INFINITY };
    uint32_t st = 0;
    if (i < n) {
        const uint32_t k = kind[i];
        const float ox = offset3[3 * i], oy = offset3[3 * i + 1], oz = offset3[3 * i + 2];
        const float px = pos3[3 * i], py = pos3[3 * i + 1], pz = pos3[3 * i + 2];
        const uint32_t e = entity[i];
        const uint32_t g = gidx ? gidx[i] : i;
        // mod.rs:315 / :321  centre = position_derived + offset
        const float cx = add(px, ox), cy = add(py, oy), cz = add(pz, oz);
        if (k == 0) {
            const float r = size3[3 * i];  // mod.rs:316 radius is not scaled
            mn[0] = sub(cx, r); mn[1] = sub(cy, r); mn[2] = sub(cz, r);
            mx[0] = add(cx, r); mx[1] = add(cy, r); mx[2] = add(cz, r);
            st256(rec + 2 * (size_t)i, make_float4(cx, cy, cz, r),
                  make_float4(__uint_as_float(e), 0.0f, __uint_as_float(g), __uint_as_float(i << 4)));
        } else if (k == 1) {
            Obb o;
            // mod.rs:320  half_widths = widths * scale * 0.5
            const size_t t = box_slot ? box_slot[i] : i;  // where this box's rotation / scale live
            o.hx = mul(mul(size3[3 * i], scale3[3 * t]), 0.5f);
            o.hy = mul(mul(size3[3 * i + 1], scale3[3 * t + 1]), 0.5f);
            o.hz = mul(mul(size3[3 * i + 2], scale3[3 * t + 2]), 0.5f);
            o.cx = cx; o.cy = cy; o.cz = cz;
            const float4 q = reinterpret_cast<const float4 *>(quat4)[t];
            quat_to_mat3(q.x, q.y, q.z, q.w, o.m);  // mod.rs:322
            obb_aabb_fast(o, mn, mx);  // == the 8-vertex walk of bounding_volume.rs:161-327, bit for bit
            obb_store(obb + 4 * (size_t)i, o, e);
            st256(rec + 2 * (size_t)i, make_float4(mn[0], mn[1], mn[2], mx[0]),
                  make_float4(mx[1], mx[2], __uint_as_float(g), __uint_as_float((i << 4) | kMetaBox)));
        } else {
            st |= kStMesh;  // mod.rs:331 unimplemented!()
            // keep the record decodable: a zero-size sphere far away is never produced; mark and bail
            st256(rec + 2 * (size_t)i, make_float4(0.f, 0.f, 0.f, 0.f),
                  make_float4(__uint_as_float(e), 0.0f, __uint_as_float(g), __uint_as_float(i << 4)));
            mn[0] = mn[1] = mn[2] = mx[0] = mx[1] = mx[2] = 0.0f;
        }
        // bounding_volume.rs:365-381: longest_axis = max(0, every diff), strict > from 0.0
        const float dx = sub(mx[0], mn[0]), dy = sub(mx[1], mn[1]), dz = sub(mx[2], mn[2]);
        ext = fmaxf(fmaxf(fmaxf(0.0f, dx), dy), dz);
        mnx_max = mn[0];
        const float chk = mn[0] + mn[1] + mn[2] + mx[0] + mx[1] + mx[2];
        if (!(fabsf(chk) <= 3.0e38f)) st |= kStNonFinite;  // NaN or Inf anywhere
    }
    // warp reduce
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ext = fmaxf(ext, __shfl_xor_sync(0xffffffffu, ext, o));
        mnx_max = fmaxf(mnx_max, __shfl_xor_sync(0xffffffffu, mnx_max, o));
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        st |= __shfl_xor_sync(0xffffffffu, st, o);
    }
    __shared__ float s_red[kBvhThreads / 32][8];
    __shared__ uint32_t s_st[kBvhThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_red[warp][0] = ext;
        s_red[warp][7] = mnx_max;
        for (int a = 0; a < 3; ++a) { s_red[warp][1 + a] = mn[a]; s_red[warp][4 + a] = mx[a]; }
        s_st[warp] = st;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        const int c = threadIdx.x;
        float v = s_red[0][c];
        for (int w = 1; w < kBvhThreads / 32; ++w) v = (c >= 1 && c <= 3) ? fminf(v, s_red[w][c]) : fmaxf(v, s_red[w][c]);
        if (c >= 1 && c <= 3) {
            if (v != INFINITY) atomicMax(&fb->bounds[c], ~f2ord(v));
        } else if (c == 0 || v != -INFINITY) {
            atomicMax(&fb->bounds[c], f2ord(v));
        }
    }
    if (threadIdx.x == 32) {
        uint32_t s = 0;
        for (int w = 0; w < kBvhThreads / 32; ++w) s |= s_st[w];
        if (s) atomicOr(&fb->status, s);
    }
}

// ---- grid placement: one thread -----------------------------------------------------------------
// cell size = longest axis (grid_collision.rs:238-240); dense grid covers floor(world min / cell) ..
// floor(world max / cell), which bounds every home cell because floor(fl(p / c)) is monotone in p.
// `has_window`: the caller (multi-GPU slab layer) restricts the grid to home x-cells [x_lo, x_hi] (absolute).
GSC_DEV void grid_setup(FrameBlock *fb, unsigned long long max_cells, uint32_t n_total, int has_window, int x_lo, int x_hi)
{
    GridParams g;
    g.valid = 0;
    g.windowed = has_window;
    g.num_cells = 0;
    g.cell_size = ord2f(fb->bounds[0]);
    for (int a = 0; a < 3; ++a) { g.origin[a] = 0; g.dims[a] = 0; g.qorg[a] = 0.0f; g.qscale[a] = 0.0f; }
    g.qmargin = 1.0f;
    fb->sort.num_passes = 0;
    fb->sort.key_bits = 0;
    fb->sort.digit_bits = 0;
    fb->n_total = n_total;
    uint32_t st = 0;
    if (n_total == 0 || fb->bounds[0] == 0) {  // nothing to do (no colliders): valid stays 0, not an error
        fb->grid = g;
        return;
    }
    if (!(g.cell_size > 0.0f)) st |= kStCellSize;
    if (fb->status & (kStNonFinite | kStMesh)) st |= fb->status;
    if (st == 0) {
        unsigned long long cells = 1;
        for (int a = 0; a < 3; ++a) {
            const float wmin = ord2f(~fb->bounds[1 + a]), wmax = ord2f(fb->bounds[4 + a]);
            const float ext = wmax - wmin;
            g.qorg[a] = wmin;
            g.qscale[a] = __fdiv_rn(a == 2 ? 64.0f : 128.0f, g.cell_size);
            // f32 error of t = (v - wmin) * qscale is < 3 * 2^-24 * t: one quantum of slack covers it up to t = 2^21
            // (16384 cells on an axis); beyond that widen the slack
            g.qmargin = fmaxf(g.qmargin, 1.0f + ceilf(ext * g.qscale[a] * 2.4e-7f - 0.5f));
            float lo = cell_coord_f(wmin, g.cell_size);
            float hi = cell_coord_f(wmax, g.cell_size);
            if (a == 0 && has_window) {  // only the slab's columns get cells; home cells outside are dropped by k_cell_keys
                lo = fmaxf(lo, (float)x_lo);
                hi = fminf(fmaxf(hi, lo), fmaxf((float)x_hi, lo));
            }
            // GridCoord = i16 (grid_collision.rs:535): outside it the reference's cast is undefined
            if (!(lo >= -32768.0f) || !(hi <= 32767.0f)) { st |= kStRange; break; }
            // one apron cell on every side: neighbour keys of any home cell are valid cells (k_pairs)
            g.origin[a] = (int32_t)lo - 1;
            g.dims[a] = (int32_t)hi - (int32_t)lo + 3;
            cells *= (unsigned long long)g.dims[a];
        }
        if (st == 0) {
            if (max_cells == ~0ull) {
                // probe (gsc_set_global_bounds): only the cell size / quantisation are needed yet; the real
                // table is sized when the slab window is known
                g.valid = 1;
            } else if (cells > max_cells || cells >= (1ull << 31)) st |= kStCells;
            else {
                g.num_cells = (uint32_t)cells;
                g.valid = 1;
                int bits = 0;
                while ((1ull << bits) <= cells) ++bits;  // keys take values 0..cells (cells == wide)
                fb->sort.num_passes = radix_passes(bits);
                fb->sort.key_bits = bits;
                fb->sort.digit_bits = radix_digit_bits(bits);
            }
        }
    }
    if (st) atomicOr(&fb->status, st);
    fb->grid = g;
}

__global__ void k_grid_setup(FrameBlock *fb, unsigned long long max_cells, uint32_t n_total, int has_window, int x_lo, int x_hi)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    grid_setup(fb, max_cells, n_total, has_window, x_lo, x_hi);
}

GSC_DEV bool home_cell(const Volume &v, const GridParams &g, int c_mn[3], int c_mx[3])
{
    bool wide = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        c_mn[a] = (int)cell_coord_f(v.mn[a], g.cell_size) - g.origin[a];
        c_mx[a] = (int)cell_coord_f(v.mx[a], g.cell_size) - g.origin[a];
        wide |= (c_mx[a] - c_mn[a]) > 1;
    }
    return wide;
}

// ---- stage 1: cell keys -----------------------------------------------------------------------------
// key = (x * dims.y + y) * dims.z + z of the home cell: x is the major axis, so x-slabs of the
// multi-GPU partition are contiguous key ranges and cells of one (x,y) row are contiguous in z.
constexpr int kKeyThreads = 256;

// One CTA per sort tile (kSortTile volumes), so that the digit counts of the FIRST radix pass come for free:
// tile_hist[d * num_tiles + tile] is what k_radix_upsweep would compute for pass 0.
#ifndef GSC_KEYS_MIN_CTAS
#define GSC_KEYS_MIN_CTAS 8  // 34 registers without a bound (7 CTAs) / 32 with (8 CTAs): 0.126 / 0.122 ms at 16M
#endif
__global__ void __launch_bounds__(kKeyThreads, GSC_KEYS_MIN_CTAS)
k_cell_keys(uint32_t n_cap, const uint32_t *__restrict__ n_dev, const float4 *__restrict__ rec, uint32_t *__restrict__ keys,
            uint32_t *__restrict__ wide_list, uint32_t *__restrict__ tile_hist, FrameBlock *__restrict__ fb)
{
    constexpr int WARPS = kKeyThreads / 32;
    __shared__ __align__(16) uint16_t s_hist[WARPS][kRadix];
    const GridParams g = fb->grid;
    if (!g.valid) return;
    // (slab steps launch for a capacity: the step's volume count -- local + ghosts -- then lives on the device;
    // otherwise the host knows it and no CTA waits for a load before it can start)
    const uint32_t n_total = n_dev ? min(*n_dev, n_cap) : n_cap;
    const uint32_t num_tiles = (n_total + kSortTile - 1) / kSortTile;
    if (blockIdx.x >= num_tiles) return;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t mask = (1u << fb->sort.digit_bits) - 1u;
#pragma unroll
    for (int i = 0; i < WARPS * kRadix / (8 * kKeyThreads); ++i)
        reinterpret_cast<uint4 *>(&s_hist[0][0])[tid + i * kKeyThreads] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    const uint32_t tile = blockIdx.x, tile_base = tile * (uint32_t)kSortTile;
#pragma unroll 4
    for (int j = 0; j < kSortTile / kKeyThreads; ++j) {
        const uint32_t i = tile_base + j * kKeyThreads + tid;
        if (i >= n_total) break;
        const F8 rv = ld256(rec + 2 * (size_t)i);
        const Volume v = volume_decode(rv.lo, rv.hi);
        int c0[3], c1[3];
        const bool wide = home_cell(v, g, c0, c1);
        uint32_t key;
        if (c0[0] < 1 || c0[0] > g.dims[0] - 2) {
            // outside the slab's x window: a ghost there cannot meet any local taker -- park it behind the
            // cells; a LOCAL volume there means the caller's window is wrong
            key = g.num_cells;
            if (!(v.meta & kMetaGhost)) atomicOr(&fb->status, kStWindow);
        } else if (wide) {
            key = g.num_cells;
            // Wide volumes are matched by k_wide_pairs against the LOCAL volumes only.  In a slab step that loses the
            // pairs a wide volume has across a slab border: refuse instead of dropping them silently (it needs an
            // AABB that spans three cells through rounding AND sits on a cut: never seen in practice).
            if (g.windowed && ((v.meta & kMetaGhost) || c0[0] <= 2 || c0[0] >= g.dims[0] - 3)) atomicOr(&fb->status, kStWideBorder);
            if (!(v.meta & kMetaGhost)) wide_list[atomicAdd(&fb->n_wide, 1u)] = i;
        } else {
            key = ((uint32_t)c0[0] * (uint32_t)g.dims[1] + (uint32_t)c0[1]) * (uint32_t)g.dims[2] + (uint32_t)c0[2];
        }
        keys[i] = key;
        hist_add(s_hist[warp], key & mask);
    }
    __syncthreads();
    uint32_t c0 = 0, c1 = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        const uint32_t v = reinterpret_cast<const uint32_t *>(s_hist[w])[tid];
        c0 += v & 0xffffu;
        c1 += v >> 16;
    }
    tile_hist[(size_t)(2 * tid) * num_tiles + tile] = c0;
    tile_hist[(size_t)(2 * tid + 1) * num_tiles + tile] = c1;
}

// What the pairs stage streams in CELL ORDER is 12 B per volume, not the 32 B record: the reach flags say whether the
// AABB's max corner lies in the next cell along y / z (cell(max) > cell(min)), which is what lets the pairs stage prune
// rows exactly.  (Rounds 1 and 2a also kept a 32 B cell-ordered copy of the record for the exact AABB test of the pairs
// tail; the test now runs in the SAT stage on the records that stage loads anyway, so the copy -- 512 MB written per
// 16M step -- is gone.)
constexpr uint32_t kMetaReachY = 4u, kMetaReachZ = 8u;

// and, for the filter loop of k_pairs, an 8 B conservatively quantised box per volume plus its ordering word:
//   q[p]  = (Wmin, Wmax | G)      w3[p] = gidx << 4 | meta flags
// Each W packs the three axes of the quantised min (max) corner into one word -- x: bits 0-9, y: bits 11-20, z: bits
// 22-30 -- with a guard bit G above every field (bits 10, 21, 31).  The quantum is cell/128 on x, y and cell/64 on z,
// qmin = floor(t) - margin, qmax = ceil(t) + margin with t = (v - world_min) * quanta_per_cell / cell, and only the
// LOW bits of the integer are kept: coordinates are compared modulo 8 cells.  That is enough because k_pairs only
// ever compares volumes whose home cells differ by at most one per axis: A.max - B.min and B.max - A.min then lie
// in (-2, 3) cells = (-256, 384) quanta (x, y) resp. (-128, 192) (z), inside the signed range of the field.  So
//   t1 = (A.Wmax | G) - B.Wmin,   t2 = (B.Wmax | G) - A.Wmin      (the guard bits absorb the borrows)
// hold the three differences each, and the boxes overlap on every axis  <=>  no field of t1 or t2 is negative
//   <=>  ((t1 | t2) & kQSign) == 0:
// two subtractions and one logic op per candidate instead of six compares.  Exact overlap always implies quantised
// overlap (never a false negative; the f32 error of t is below the margin); the exact test runs on the survivors only
// (in the SAT stage, k_narrow).
constexpr uint32_t kQGuard = (1u << 10) | (1u << 21) | (1u << 31);
constexpr uint32_t kQSign = (1u << 9) | (1u << 20) | (1u << 30);

struct SortedVolumes {
    uint2 *q;
    uint32_t *w3;
};

GSC_DEV uint32_t quant_lo(float v, float org, float scale, float margin)
{
    return (uint32_t)fmaxf(floorf((v - org) * scale) - margin, 0.0f);
}
GSC_DEV uint32_t quant_hi(float v, float org, float scale, float margin)
{
    return (uint32_t)(ceilf((v - org) * scale) + margin);
}
GSC_DEV uint32_t quant_pack(uint32_t x, uint32_t y, uint32_t z)
{
    return (x & 1023u) | ((y & 1023u) << 11) | ((z & 511u) << 22);
}

// ---- stage 3: cell start table + the cell-ordered 12 B pair records (quantised box, ordering word) ----
// start[c] = first sorted position whose key >= c, for c in [0, num_cells]; thread p fills the gap
// (key[p-1], key[p]].  Long gaps (empty space) are filled by the whole warp.
constexpr int kTableThreads = 256;
#ifndef GSC_TABLE_IPT
#define GSC_TABLE_IPT 2
#endif
#ifndef GSC_TABLE_MIN_CTAS
#define GSC_TABLE_MIN_CTAS 7  // measured at 16M (slot order): 1 position per thread 0.290 ms; 2 positions 0.284 (48 registers,
#endif                        // 5 CTAs), 0.249 with a bound of 6 CTAs (40 registers), 0.233 with 7 (32 registers: 8 CTAs
                              // resident); 4 positions 0.275 (72 registers)
constexpr int kTableIpt = GSC_TABLE_IPT;  // sorted positions per thread: their record gathers are in flight together
constexpr int kTableSpan = kTableThreads * kTableIpt;  // positions per CTA
constexpr int kGapLong = 1 << 14;  // empty-cell runs longer than this are filled by the whole grid
constexpr int kGapCap = 4096;
constexpr int kWarpFillMax = 4096;  // cells a warp fills transposed; wider spans (rare) go gap by gap

// start[] for the cells the warp's 32 consecutive positions own: lane's position p owns the key gap (klo, kcur]
GSC_DEV void table_fill_warp(uint32_t p, uint32_t klo, uint32_t kcur, int lane, uint32_t C, uint32_t n_total,
                             uint32_t *__restrict__ cell_start, uint4 *__restrict__ gaps, FrameBlock *__restrict__ fb)
{
    // The warp's 32 positions own the contiguous cells [first, last].  Normally the warp fills them TRANSPOSED: lane l
    // takes cells first + l, first + 32 + l, ... and finds each cell's start -- position of the first key >= the
    // cell -- by a 5-step binary search over the warp's sorted keys (shuffles).  32 cells per trip whatever the gap
    // sizes, instead of every lane walking its own gap (sparse space: 20-cell gaps in every lane, 32 serial loops).
    const uint32_t first = __shfl_sync(0xffffffffu, klo, 0), last = __shfl_sync(0xffffffffu, kcur, 31);
    const uint32_t span = last + 1u - first;  // >= 0: klo of lane 0 <= its key + 1 <= last + 1
    if (span <= (uint32_t)kWarpFillMax) {
        const uint32_t p0 = p - lane;
        for (uint32_t t = 0; t < span; t += 32) {
            const uint32_t c = first + t + lane;
            uint32_t j = 0;  // number of keys of this warp below c (keys ascend; < 32 for every c <= last)
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint32_t k = __shfl_sync(0xffffffffu, kcur, j + step - 1);
                if (k < c) j += step;
            }
            GSC_CHECK(fb, c > last || (c <= C && p0 + j <= n_total));
            if (c <= last) cell_start[c] = p0 + j;
        }
        return;
    }
    // a warp that spans a large run of empty cells (blob boundary, space between slabs): every lane its own gap
    int64_t lo = klo, hi = kcur;
    const int64_t len = hi - lo + 1;
    if (len > kGapLong) {  // a huge run of empty cells: hand it to k_fill_gaps
        const uint32_t slot = atomicAdd(&fb->n_gaps, 1u);
        if (slot < (uint32_t)kGapCap) {
            gaps[slot] = make_uint4((uint32_t)lo, (uint32_t)hi, p, 0u);
            hi = lo - 1;
        }
    }
    const bool is_long = hi - lo + 1 > 8;
    if (!is_long)
        for (int64_t c = lo; c <= hi; ++c) cell_start[c] = p;
    uint32_t long_mask = __ballot_sync(0xffffffffu, is_long);
    while (long_mask) {
        const int src = __ffs(long_mask) - 1;
        long_mask &= long_mask - 1;
        const int64_t slo = __shfl_sync(0xffffffffu, lo, src);
        const int64_t shi = __shfl_sync(0xffffffffu, hi, src);
        const uint32_t sp = __shfl_sync(0xffffffffu, p, src);
        for (int64_t c = slo + lane; c <= shi; c += 32) cell_start[c] = sp;
    }
}

__global__ void __launch_bounds__(kTableThreads, GSC_TABLE_MIN_CTAS)
k_cell_table_permute(uint32_t n_cap, const uint32_t *__restrict__ n_dev, const uint32_t *__restrict__ keys_a, const uint32_t *__restrict__ vals_a,
                     const uint32_t *__restrict__ keys_b, const uint32_t *__restrict__ vals_b,
                     const float4 *__restrict__ rec, SortedVolumes sv,
                     uint32_t *__restrict__ cell_start, uint4 *__restrict__ gaps, FrameBlock *__restrict__ fb)
{
    const GridParams g = fb->grid;
    if (!g.valid) return;
    const uint32_t n_total = n_dev ? min(*n_dev, n_cap) : n_cap;
    if (blockIdx.x * (uint32_t)kTableSpan > n_total) return;  // (position n_total itself closes the table)
    const bool in_b = fb->sort.num_passes & 1;  // ping-pong parity of the sort
    const uint32_t *keys = in_b ? keys_b : keys_a;
    const uint32_t *vals = in_b ? vals_b : vals_a;
    const uint32_t C = g.num_cells;
    // thread positions p in [0, n_total]; position p fills start[] for the cells of the key gap (key[p-1], key[p]]
    // (keys clamped to C; positions behind n_total count as key C, so their gaps are empty and the kernel needs no
    // tail special case).  A thread takes kTableIpt positions kTableThreads apart and has their record gathers -- the
    // one latency-bound access of the kernel -- in flight together.
    uint32_t p[kTableIpt], kcur[kTableIpt], klo[kTableIpt];
    F8 rv[kTableIpt];
#pragma unroll
    for (int u = 0; u < kTableIpt; ++u) {
        p[u] = blockIdx.x * (uint32_t)kTableSpan + u * kTableThreads + threadIdx.x;
        kcur[u] = p[u] < n_total ? min(keys[p[u]], C) : C;
        klo[u] = p[u] == 0 ? 0u : (p[u] - 1 < n_total ? min(keys[p[u] - 1], C) : C) + 1u;
        if (p[u] < n_total) rv[u] = ld256_gather(rec + 2 * (size_t)vals[p[u]]);
    }
#pragma unroll
    for (int u = 0; u < kTableIpt; ++u) {
        if (p[u] < n_total) {
            const Volume v = volume_decode(rv[u].lo, rv[u].hi);
            uint32_t meta = v.meta;
            if (cell_coord_f(v.mx[1], g.cell_size) > cell_coord_f(v.mn[1], g.cell_size)) meta |= kMetaReachY;
            if (cell_coord_f(v.mx[2], g.cell_size) > cell_coord_f(v.mn[2], g.cell_size)) meta |= kMetaReachZ;
            const uint32_t wmin = quant_pack(quant_lo(v.mn[0], g.qorg[0], g.qscale[0], g.qmargin),
                                             quant_lo(v.mn[1], g.qorg[1], g.qscale[1], g.qmargin),
                                             quant_lo(v.mn[2], g.qorg[2], g.qscale[2], g.qmargin));
            const uint32_t wmax = quant_pack(quant_hi(v.mx[0], g.qorg[0], g.qscale[0], g.qmargin),
                                             quant_hi(v.mx[1], g.qorg[1], g.qscale[1], g.qmargin),
                                             quant_hi(v.mx[2], g.qorg[2], g.qscale[2], g.qmargin));
            sv.q[p[u]] = make_uint2(wmin, wmax | kQGuard);
            sv.w3[p[u]] = (v.gidx << 4) | (meta & 15u);
        }
    }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int u = 0; u < kTableIpt; ++u) table_fill_warp(p[u], klo[u], kcur[u], lane, C, n_total, cell_start, gaps, fb);
}

__global__ void __launch_bounds__(256)
k_fill_gaps(uint32_t *__restrict__ cell_start, const uint4 *__restrict__ gaps, const FrameBlock *__restrict__ fb)
{
    const uint32_t n = min(fb->n_gaps, (uint32_t)kGapCap);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t k = 0; k < n; ++k) {
        const uint4 gp = gaps[k];
        for (uint64_t c = (uint64_t)gp.x + blockIdx.x * blockDim.x + threadIdx.x; c <= gp.y; c += stride) cell_start[c] = gp.z;
    }
}

}  // namespace gsc

#include "gsc_pairs.cuh"  // stage 4: neighbour-cell pair enumeration + AABB filter (k_pairs_tiled, k_pairs)

namespace gsc {

// ---- stage 5: SAT narrowphase over the deferred candidates ------------------------------------------
constexpr int kNarrowThreads = 256;
constexpr int kNarrowWarps = kNarrowThreads / 32;
constexpr int kNarrowCap = 96;
#ifndef GSC_NARROW_CTAS
#define GSC_NARROW_CTAS 5  // 3 / 4 / 5 measured with the exact AABB test in this kernel (slot order): 0.334 / 0.286 / 0.273 ms at 16M
#endif
constexpr int kNarrowCtasPerSm = GSC_NARROW_CTAS;  // resident CTAs per SM the persistent grid is sized for

// All three shape tests in ONE persistent kernel over a work queue of candidate chunks: box-box chunks first
// (the long SAT tests), then sphere-box, then sphere-sphere.  Run as three kernels each of them was latency
// bound with the GPU to itself (issue 9-22%, DRAM 38-58% of peak: one random 128 B line per record); mixed in
// one grid, CTAs doing SAT arithmetic overlap with CTAs waiting on record gathers.
constexpr int kNarrowChunk = kNarrowThreads * 2;  // candidates per ticket

template <bool CONTACTS>
__global__ void __launch_bounds__(kNarrowThreads, kNarrowCtasPerSm)
k_narrow(const uint2 *__restrict__ cand_ss, const uint2 *__restrict__ cand_so, const uint2 *__restrict__ cand_oo,
         unsigned long long cap_cands, const float4 *__restrict__ rec, const float4 *__restrict__ obb,
         uint2 *__restrict__ pairs, float4 *__restrict__ contacts, unsigned long long cap_pairs, FrameBlock *__restrict__ fb)
{
    __shared__ EmitBuf<kNarrowWarps, kNarrowCap> s_out;
    __shared__ uint32_t s_ticket;
    emit_init(s_out);
    const unsigned long long n_oo = min(fb->n_oo, cap_cands), n_so = min(fb->n_so, cap_cands), n_ss = min(fb->n_ss, cap_cands);
    const uint32_t c_oo = (uint32_t)((n_oo + kNarrowChunk - 1) / kNarrowChunk);
    const uint32_t c_so = (uint32_t)((n_so + kNarrowChunk - 1) / kNarrowChunk);
    const uint32_t c_ss = (uint32_t)((n_ss + kNarrowChunk - 1) / kNarrowChunk);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_ticket = atomicAdd(&fb->narrow_ticket, 1u);
        __syncthreads();
        uint32_t t = s_ticket;
        if (t >= c_oo + c_so + c_ss) break;
        const int cls = t < c_oo ? 2 : (t < c_oo + c_so ? 1 : 0);
        t -= cls == 2 ? 0u : (cls == 1 ? c_oo : c_oo + c_so);
        const unsigned long long n = cls == 2 ? n_oo : (cls == 1 ? n_so : n_ss);
        const uint2 *cand = cls == 2 ? cand_oo : (cls == 1 ? cand_so : cand_ss);
#pragma unroll 1
        for (int k = 0; k < kNarrowChunk / kNarrowThreads; ++k) {
            const unsigned long long i = (unsigned long long)t * kNarrowChunk + k * kNarrowThreads + threadIdx.x;
            bool hit = false;
            uint2 item = make_uint2(0, 0);
            float4 ct = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n) {
                const uint2 c = cand[i];
                if (cls == 2) {  // OBB::test_obb (mod.rs:413-546) evaluated as later.test_obb(earlier) (grid_collision.rs:505, mod.rs:408)
                    uint32_t ea, eb;
                    const Obb a = obb_load(obb + 4 * (size_t)c.x, &ea);
                    const Obb bb = obb_load(obb + 4 * (size_t)c.y, &eb);
                    // BoundVolume::test runs AABB::test_aabb first (bounding_volume.rs:124-132).  The candidates only passed
                    // the conservative quantised filter; the exact AABBs are re-derived here, bit for bit, with the function
                    // bvh_update used on the very same OBB records (AABB::from_collider, bounding_volume.rs:161-327)
                    float amn[3], amx[3], bmn[3], bmx[3];
                    obb_aabb_fast(a, amn, amx);
                    obb_aabb_fast(bb, bmn, bmx);
                    hit = test_aabb(amn, amx, bmn, bmx) && test_obb_obb(a, bb);
                    item = make_uint2(ea, eb);
                    if (CONTACTS && hit) ct = contact_obb_obb(a, bb);
                } else if (cls == 1) {  // Sphere::test_obb (mod.rs:391-394): either order calls sphere.test_obb(obb) (mod.rs:378,407)
                    const bool hi_is_sphere = c.x >> 31;
                    const uint32_t hi = c.x & 0x7fffffffu, lo = c.y;
                    const uint32_t si = hi_is_sphere ? hi : lo, oi = hi_is_sphere ? lo : hi;
                    const F8 sr = ld256_gather(rec + 2 * (size_t)si);
                    const float4 s0 = sr.lo, s1 = sr.hi;
                    uint32_t oe;
                    const Obb o = obb_load(obb + 4 * (size_t)oi, &oe);
                    float omn[3], omx[3];
                    obb_aabb_fast(o, omn, omx);
                    // a sphere's AABB: centre -+ (r, r, r) (bounding_volume.rs:150-159)
                    const float smn[3] = { sub(s0.x, s0.w), sub(s0.y, s0.w), sub(s0.z, s0.w) };
                    const float smx[3] = { add(s0.x, s0.w), add(s0.y, s0.w), add(s0.z, s0.w) };
                    hit = test_aabb(smn, smx, omn, omx) && test_sphere_obb(s0.x, s0.y, s0.z, s0.w, o);
                    const uint32_t se = __float_as_uint(s1.x);
                    item = hi_is_sphere ? make_uint2(se, oe) : make_uint2(oe, se);
                    if (CONTACTS && hit) ct = contact_sphere_obb(s0.x, s0.y, s0.z, s0.w, o, hi_is_sphere);
                } else {  // Sphere::test_sphere (mod.rs:383-389)
#ifndef GSC_NO_LDST256
                    const F8 hr = ld256_gather(rec + 2 * (size_t)c.x), lr = ld256_gather(rec + 2 * (size_t)c.y);
                    const float4 h0 = hr.lo, l0 = lr.lo;
#else
                    const float4 h0 = ldg_gather(rec + 2 * (size_t)c.x), l0 = ldg_gather(rec + 2 * (size_t)c.y);
#endif
                    const float hmn[3] = { sub(h0.x, h0.w), sub(h0.y, h0.w), sub(h0.z, h0.w) };
                    const float hmx[3] = { add(h0.x, h0.w), add(h0.y, h0.w), add(h0.z, h0.w) };
                    const float lmn[3] = { sub(l0.x, l0.w), sub(l0.y, l0.w), sub(l0.z, l0.w) };
                    const float lmx[3] = { add(l0.x, l0.w), add(l0.y, l0.w), add(l0.z, l0.w) };
                    hit = test_aabb(hmn, hmx, lmn, lmx) && test_sphere_sphere(h0.x, h0.y, h0.z, h0.w, l0.x, l0.y, l0.z, l0.w);
                    if (hit) {
#ifndef GSC_NO_LDST256
                        const float4 h1 = hr.hi, l1 = lr.hi;  // (the entity ids came with the same 32 B access)
#else
                        const float4 h1 = ldg_gather(rec + 2 * (size_t)c.x + 1), l1 = ldg_gather(rec + 2 * (size_t)c.y + 1);
#endif
                        item = make_uint2(__float_as_uint(h1.x), __float_as_uint(l1.x));
                        if (CONTACTS) ct = contact_sphere_sphere(h0.x, h0.y, h0.z, h0.w, l0.x, l0.y, l0.z, l0.w);
                    }
                }
            }
            emit(s_out, hit, item, ct, pairs, CONTACTS ? contacts : nullptr, &fb->n_pairs, cap_pairs);
        }
    }
    emit_finish(s_out, pairs, CONTACTS ? contacts : nullptr, &fb->n_pairs, cap_pairs);
}
// (Measured and dropped: a software-pipelined variant that staged the NEXT candidate's records in thread-private
// shared-memory slots with cp.async while testing the current one ran this stage in 1.0-1.2 ms instead of 0.46 ms at
// 16M -- eight scattered 16 B LDGSTS per thread cost far more in the L1/L2 request path than the overlap wins; the
// gathers already overlap across the 32 resident warps of an SM.)

// full BoundVolume::test (bounding_volume.rs:124-132) between two arbitrary volumes, `hi` = later index
GSC_DEV bool test_volume_full(const Volume &hi, const Volume &lo, const float4 *__restrict__ obb, uint32_t *e_hi,
                              uint32_t *e_lo, bool want_contact, float4 *ct)
{
    if (!test_aabb(hi.mn, hi.mx, lo.mn, lo.mx)) return false;
    const bool hb = hi.meta & kMetaBox, lb = lo.meta & kMetaBox;
    bool hit;
    if (!hb && !lb) {
        *e_hi = hi.entity; *e_lo = lo.entity;
        hit = test_sphere_sphere(hi.cx, hi.cy, hi.cz, hi.r, lo.cx, lo.cy, lo.cz, lo.r);
        if (hit && want_contact) *ct = contact_sphere_sphere(hi.cx, hi.cy, hi.cz, hi.r, lo.cx, lo.cy, lo.cz, lo.r);
    } else if (hb && lb) {
        const Obb a = obb_load(obb + 4 * (size_t)(hi.meta >> 4), e_hi);
        const Obb b = obb_load(obb + 4 * (size_t)(lo.meta >> 4), e_lo);
        hit = test_obb_obb(a, b);
        if (hit && want_contact) *ct = contact_obb_obb(a, b);
    } else if (hb) {
        const Obb o = obb_load(obb + 4 * (size_t)(hi.meta >> 4), e_hi);
        *e_lo = lo.entity;
        hit = test_sphere_obb(lo.cx, lo.cy, lo.cz, lo.r, o);
        if (hit && want_contact) *ct = contact_sphere_obb(lo.cx, lo.cy, lo.cz, lo.r, o, false);
    } else {
        const Obb o = obb_load(obb + 4 * (size_t)(lo.meta >> 4), e_lo);
        *e_hi = hi.entity;
        hit = test_sphere_obb(hi.cx, hi.cy, hi.cz, hi.r, o);
        if (hit && want_contact) *ct = contact_sphere_obb(hi.cx, hi.cy, hi.cz, hi.r, o, true);
    }
    return hit;
}

// Volumes that span 3 cells on an axis (possible only through rounding of p / cell_size): the
// reference enters them in their CORNER cells only (grid_collision.rs:447-480), so they collide with
// another volume iff the two share a corner cell.  Restated literally, wide x everything.
__global__ void __launch_bounds__(kNarrowThreads)
k_wide_pairs(uint32_t n_local, const uint32_t *__restrict__ wide_list, const float4 *__restrict__ rec,
             const float4 *__restrict__ obb, uint2 *__restrict__ pairs, float4 *__restrict__ contacts,
             unsigned long long cap_pairs, FrameBlock *__restrict__ fb)
{
    __shared__ EmitBuf<kNarrowWarps, kNarrowCap> s_out;
    const GridParams g = fb->grid;
    const uint32_t n_wide = fb->n_wide;
    if (!g.valid || n_wide == 0) return;
    emit_init(s_out);
    __syncthreads();
    const unsigned long long total = (unsigned long long)n_wide * n_local;
    const unsigned long long stride = (unsigned long long)gridDim.x * kNarrowThreads;
    for (unsigned long long b = (unsigned long long)blockIdx.x * kNarrowThreads; b < total; b += stride) {
        const unsigned long long t = b + threadIdx.x;
        bool hit = false;
        uint2 item = make_uint2(0, 0);
        float4 ct = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < total) {
            const uint32_t wi = wide_list[t / n_local];
            const uint32_t j = (uint32_t)(t % n_local);
            if (wi != j) {
                const Volume W = volume_decode(__ldg(rec + 2 * (size_t)wi), __ldg(rec + 2 * (size_t)wi + 1));
                const Volume J = volume_decode(__ldg(rec + 2 * (size_t)j), __ldg(rec + 2 * (size_t)j + 1));
                int w0[3], w1[3], j0[3], j1[3];
                home_cell(W, g, w0, w1);
                const bool j_wide = home_cell(J, g, j0, j1);
                // a wide-wide pair is taken once, by the iteration whose J is the later volume
                bool take = !j_wide || (W.gidx < J.gidx);
#pragma unroll
                for (int a = 0; a < 3; ++a)
                    take &= (w0[a] == j0[a]) || (w0[a] == j1[a]) || (w1[a] == j0[a]) || (w1[a] == j1[a]);
                if (take) {
                    const bool w_hi = W.gidx > J.gidx;
                    uint32_t eh = 0, el = 0;
                    hit = w_hi ? test_volume_full(W, J, obb, &eh, &el, contacts != nullptr, &ct)
                               : test_volume_full(J, W, obb, &eh, &el, contacts != nullptr, &ct);
                    item = make_uint2(eh, el);
                }
            }
        }
        emit(s_out, hit, item, ct, pairs, contacts, &fb->n_pairs, cap_pairs);
    }
    emit_finish(s_out, pairs, contacts, &fb->n_pairs, cap_pairs);
}

// ---- multi-GPU slab support: home-x range, halo selection, ghost append ------------------------------
// Home x-cell range of the local volumes, in absolute cell coordinates (floor(aabb.min.x / cell)).
__global__ void __launch_bounds__(256)
k_local_x_range(uint32_t n_local, const float4 *__restrict__ rec, FrameBlock *__restrict__ fb)
{
    const GridParams g = fb->grid;
    if (!g.valid) return;
    int lo = INT_MAX, hi = INT_MIN;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += gridDim.x * blockDim.x) {
        const Volume v = volume_decode(__ldg(rec + 2 * (size_t)i), __ldg(rec + 2 * (size_t)i + 1));
        const int x = (int)cell_coord_f(v.mn[0], g.cell_size);
        lo = min(lo, x);
        hi = max(hi, x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        atomicMax(&fb->xrange[0], ~(uint32_t)(lo + 32768));
        atomicMax(&fb->xrange[1], (uint32_t)(hi + 32768));
    }
}

// Halo record = 96 B: the 32 B volume record followed by the 64 B OBB record (undefined for spheres).
// Selects the local volumes whose home x-cell lies in [x_lo, x_hi] (absolute cells) into `out`,
// warp-aggregated; *count keeps counting past `cap` so the caller can size the buffer.
__global__ void __launch_bounds__(256)
k_halo_select(uint32_t n_local, const float4 *__restrict__ rec, const float4 *__restrict__ obb, int x_lo, int x_hi,
              float4 *__restrict__ out, uint32_t cap, uint32_t *__restrict__ count, const FrameBlock *__restrict__ fb)
{
    const GridParams g = fb->grid;
    if (!g.valid) return;
    const int lane = threadIdx.x & 31;
    const uint32_t n_round = (n_local + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        bool sel = false;
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
        if (i < n_local) {
            r0 = __ldg(rec + 2 * (size_t)i);
            r1 = __ldg(rec + 2 * (size_t)i + 1);
            const Volume v = volume_decode(r0, r1);
            const int x = (int)cell_coord_f(v.mn[0], g.cell_size);
            sel = x >= x_lo && x <= x_hi;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, sel);
        if (m == 0) continue;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(count, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        const uint32_t at = base + __popc(m & ((1u << lane) - 1u));
        if (sel && at < cap) {
            float4 *d = out + 6 * (size_t)at;
            d[0] = r0;
            d[1] = r1;
            if (__float_as_uint(r1.w) & kMetaBox) {
#pragma unroll
                for (int k = 0; k < 4; ++k) d[2 + k] = __ldg(obb + 4 * (size_t)i + k);
            }
        }
    }
}

// Fused select + transfer over peer memory (NVLink / NVSwitch): the local volumes whose home x-cell lies in
// [x_lo, x_hi] are written straight into the PEER context's record arrays behind its local volumes, as ghosts
// with their new local index.  Each CTA first collects the indices it selects in shared memory, then reserves
// its slice of the peer's ghost region with ONE system-scope atomic on the peer's FrameBlock (remote atomics
// cost microseconds: one per warp made this kernel 5x slower), then streams the records across.
// No staging buffer, no count exchange, no append kernel on the receiving side.
constexpr int kPushThreads = 256;
constexpr int kPushList = 1024;  // records a CTA scans per round (= selected indices it can stage)

GSC_DEV void halo_push_cta(uint32_t n_local, const float4 *__restrict__ rec, const float4 *__restrict__ obb, int x_lo, int x_hi,
                           float4 *__restrict__ peer_rec, float4 *__restrict__ peer_obb, FrameBlock *peer_fb, uint32_t peer_first,
                           uint32_t peer_cap, const GridParams &g, uint32_t cta, uint32_t n_ctas)
{
    __shared__ uint32_t s_list[kPushList];
    __shared__ uint32_t s_count, s_base;
    const int lane = threadIdx.x & 31;
    // Blocks of kPushList records are dealt out round-robin over the CTAs.  (Round 2: with the colliders kept in cell
    // order the volumes a peer needs are ONE contiguous piece of the array -- the lowest / highest home x-cells -- and a
    // contiguous share per CTA left the whole transfer to a handful of CTAs.)
    const uint32_t end = n_local;
    for (uint32_t r0 = cta * (uint32_t)kPushList; r0 < end; r0 += n_ctas * (uint32_t)kPushList) {
        const uint32_t r1 = min(end, r0 + (uint32_t)kPushList);
        if (threadIdx.x == 0) s_count = 0;
        __syncthreads();
        for (uint32_t i = r0 + threadIdx.x; i < ((r1 - r0 + 31u) & ~31u) + r0; i += kPushThreads) {
            bool sel = false;
            if (i < r1) {
                const Volume v = volume_decode(__ldg(rec + 2 * (size_t)i), __ldg(rec + 2 * (size_t)i + 1));
                const int x = (int)cell_coord_f(v.mn[0], g.cell_size);
                sel = x >= x_lo && x <= x_hi;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, sel);
            if (m == 0) continue;
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&s_count, (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (sel) s_list[min(base + __popc(m & ((1u << lane) - 1u)), (uint32_t)kPushList - 1u)] = i;
        }
        __syncthreads();
        const uint32_t cnt = s_count;
        if (cnt == 0) continue;  // uniform
        if (threadIdx.x == 0) s_base = atomicAdd_system(&peer_fb->n_ghost_in, cnt);
        __syncthreads();
        const uint32_t first = peer_first + s_base;
        for (uint32_t k = threadIdx.x; k < cnt; k += kPushThreads) {
            const uint32_t i = s_list[k], at = first + k;
            if (at >= peer_cap) continue;
            const float4 r0v = __ldg(rec + 2 * (size_t)i);
            float4 r1v = __ldg(rec + 2 * (size_t)i + 1);
            const uint32_t box = __float_as_uint(r1v.w) & kMetaBox;
            r1v.w = __uint_as_float((at << 4) | box | kMetaGhost);
            peer_rec[2 * (size_t)at] = r0v;
            peer_rec[2 * (size_t)at + 1] = r1v;
            if (box) {
#pragma unroll
                for (int q = 0; q < 4; ++q) peer_obb[4 * (size_t)at + q] = __ldg(obb + 4 * (size_t)i + q);
            }
        }
        __syncthreads();
    }
    __threadfence_system();  // the pushed records are visible to the peer before this rank's done flag can be
}

// host-driven exchange: the caller worked out the range (gsc_halo_push)
__global__ void __launch_bounds__(kPushThreads)
k_halo_push(uint32_t n_local, const float4 *__restrict__ rec, const float4 *__restrict__ obb, int x_lo, int x_hi,
            float4 *__restrict__ peer_rec, float4 *__restrict__ peer_obb, FrameBlock *peer_fb, uint32_t peer_first,
            uint32_t peer_cap, const FrameBlock *__restrict__ fb)
{
    const GridParams g = fb->grid;
    if (!g.valid) return;
    halo_push_cta(n_local, rec, obb, x_lo, x_hi, peer_rec, peer_obb, peer_fb, peer_first, peer_cap, g, blockIdx.x, gridDim.x);
}

// device-driven exchange: blockIdx.y = peer; the range comes from k_slab_gather (most peers need nothing: those CTAs
// leave at once)
__global__ void __launch_bounds__(kPushThreads)
k_slab_push(uint32_t n_local, const float4 *__restrict__ rec, const float4 *__restrict__ obb, SlabPeers peers, uint32_t rank,
            const FrameBlock *__restrict__ fb)
{
    const uint32_t s = blockIdx.y;
    const int x_lo = fb->slab.push_lo[s], x_hi = fb->slab.push_hi[s];
    if (s == rank || x_lo > x_hi || !(fb->bounds[0] != 0)) return;
    GridParams g = fb->grid;  // (only the cell size is used)
    halo_push_cta(n_local, rec, obb, x_lo, x_hi, peers.rec[s], peers.obb[s], peers.fb[s], fb->slab.peer_n_local[s], peers.cap[s], g,
                  blockIdx.x, gridDim.x);
}

// ---- multi-GPU slab exchange driven from the device ---------------------------------------------------------------
// Round 1 ran the exchange from the host: D2H of the bounds + stream sync, a host all-gather, an H2D of the global
// bounds, the push launches, a second sync, a host barrier, a third sync to read the ghost count -- 0.26 ms per
// step whatever the number of GPUs.  Here the ranks talk through each other's mapped memory and every wait is a
// one-warp kernel on the rank's own stream:
//   k_slab_publish  my bounds -> slot `rank` of every rank's mailbox (system-scope release)
//   k_slab_gather   wait for all slots of my mailbox; reduce the global bounds; every rank's home x-range; what each
//                   peer needs of mine; place my (windowed) grid
//   k_halo_push     (one launch per peer, range read from the control block) select + write into the peer's arrays
//   k_slab_done     my pushes are complete -> done flag in every peer's mailbox
//   k_slab_wait     wait for every peer's done flag; adopt the ghosts: n_total = n_local + n_ghost_in
// The host enqueues all of it plus stages 1-5 and waits ONCE, at the end of the step.
GSC_DEV uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
GSC_DEV void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
GSC_DEV unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag == seq; false on timeout
GSC_DEV bool slab_wait_flag(const uint32_t *flag, uint32_t seq, unsigned long long timeout_ns)
{
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(flag) != seq) {
        if (global_timer_ns() - t0 > timeout_ns) return false;
        __nanosleep(64);
    }
    return true;
}

__global__ void k_slab_publish(const FrameBlock *__restrict__ fb, SlabPeers peers, uint32_t rank, uint32_t world, uint32_t n_local,
                               uint32_t seq)
{
    const uint32_t s = threadIdx.x;
    if (s >= world) return;
    SlabMailSlot *dst = &peers.mail[s]->from[rank];
#pragma unroll
    for (int i = 0; i < 8; ++i) dst->bounds[i] = fb->bounds[i];
    dst->n_local = n_local;
    dst->status = fb->status;
    __threadfence_system();
    st_release_sys(&dst->seq, seq);
}

__global__ void k_slab_gather(FrameBlock *fb, const SlabMail *mail, uint32_t rank, uint32_t world, uint32_t seq,
                              unsigned long long max_cells, uint32_t n_hint, unsigned long long timeout_ns)
{
    // one warp; lane r owns rank r's mailbox slot: the waits AND the reads of the slots run in parallel (a serial walk
    // over 8 slots of system-scope loads costs ~60 us)
    const uint32_t r = threadIdx.x;
    const bool mine = r < world;
    bool ok = true;
    if (mine) ok = slab_wait_flag(&mail->from[r].seq, seq, timeout_ns);
    if (!__all_sync(0xffffffffu, ok)) {
        if (r == 0) {
            atomicOr(&fb->status, kStPeerTimeout);
            fb->grid.valid = 0;
        }
        return;
    }
    // global bounds: bounds[] are ordered-uint encodings, every one of them reduces with max
    uint32_t b[8], st = 0, n_loc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = 0;
    if (mine) {
        const SlabMailSlot &m = mail->from[r];  // (lane r acquired this slot's sequence number above)
#pragma unroll
        for (int i = 0; i < 8; ++i) b[i] = m.bounds[i];
        st = m.status;
        n_loc = m.n_local;
    }
    uint32_t gb[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) gb[i] = __reduce_max_sync(0xffffffffu, b[i]);
    st = __reduce_or_sync(0xffffffffu, st);
    const float cell = ord2f(gb[0]);
    // home x-cell = floor(fl(aabb.min.x / cell)) (grid_collision.rs:329-335): range of rank r from its smallest / largest aabb.min.x
    const bool has = mine && n_loc != 0 && b[0] != 0 && gb[0] != 0 && cell > 0.0f && b[1] != 0 && b[7] != 0;
    float lo = 1.0f, hi = 0.0f;
    if (has) {
        lo = cell_coord_f(ord2f(~b[1]), cell);
        hi = cell_coord_f(ord2f(b[7]), cell);
        if (!(lo >= -32768.0f) || !(hi <= 32767.0f)) { lo = 1.0f; hi = 0.0f; }  // (grid_setup reports the range error)
    }
    const int32_t x_lo = (int32_t)lo, x_hi = (int32_t)hi;
    const int32_t my_lo = __shfl_sync(0xffffffffu, x_lo, rank), my_hi = __shfl_sync(0xffffffffu, x_hi, rank);
    SlabState &sl = fb->slab;
    if (mine) {
        sl.x_lo[r] = x_lo;
        sl.x_hi[r] = x_hi;
        sl.peer_n_local[r] = n_loc;
        // rank r's takers look at home x-cells {x - 1, x}: it needs the foreign volumes in [x_lo_r - 1, x_hi_r]
        int32_t plo = 1, phi = 0;
        if (r != rank && my_lo <= my_hi && x_lo <= x_hi) {
            plo = max(x_lo - 1, my_lo);
            phi = min(x_hi, my_hi);
        }
        sl.push_lo[r] = plo;
        sl.push_hi[r] = phi;
    }
    if (r != 0) return;
    for (int i = 0; i < 7; ++i) fb->bounds[i] = gb[i];
    if (st) atomicOr(&fb->status, st);
    // my grid covers my slab (+ the one column of ghosts below it), not the world
    grid_setup(fb, max_cells, n_hint, my_lo <= my_hi ? 1 : 0, my_lo - 1, my_hi);
    if (my_lo > my_hi) fb->grid.valid = 0;  // no local colliders: nothing to do on this rank (ghosts never take)
}

__global__ void k_slab_done(SlabPeers peers, uint32_t rank, uint32_t world, uint32_t seq)
{
    const uint32_t s = threadIdx.x;
    if (s >= world || s == rank) return;
    __threadfence_system();
    st_release_sys(&peers.mail[s]->done_seq[rank], seq);
}

__global__ void k_slab_wait(FrameBlock *fb, const SlabMail *mail, uint32_t rank, uint32_t world, uint32_t seq, uint32_t n_local,
                            uint32_t cap, unsigned long long timeout_ns)
{
    const uint32_t s = threadIdx.x;
    bool ok = true;
    if (s < world && s != rank) ok = slab_wait_flag(&mail->done_seq[s], seq, timeout_ns);
    const bool all_ok = __all_sync(0xffffffffu, ok);
    if (threadIdx.x != 0) return;
    if (!all_ok) {
        atomicOr(&fb->status, kStPeerTimeout);
        fb->grid.valid = 0;
        return;
    }
    uint32_t ghosts = atomicAdd_system(&fb->n_ghost_in, 0u);  // (the peers' reservations were system-scope atomics on this word)
    if ((unsigned long long)n_local + ghosts > cap) {
        atomicOr(&fb->status, kStGhostCap);
        fb->grid.valid = 0;  // the pushed records beyond the capacity were dropped: no partial result
        ghosts = cap - n_local;
    }
    fb->slab.n_ghost = ghosts;
    fb->n_total = n_local + ghosts;
}

// Appends ghosts received from another slab behind the local volumes: new local index, ghost flag.
__global__ void __launch_bounds__(256)
k_halo_append(uint32_t count, const float4 *__restrict__ in, uint32_t first, float4 *__restrict__ rec, float4 *__restrict__ obb)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const float4 *s = in + 6 * (size_t)j;
    const uint32_t i = first + j;
    float4 r0 = s[0], r1 = s[1];
    const uint32_t box = __float_as_uint(r1.w) & kMetaBox;
    r1.w = __uint_as_float((i << 4) | box | kMetaGhost);
    rec[2 * (size_t)i] = r0;
    rec[2 * (size_t)i + 1] = r1;
    if (box) {
#pragma unroll
        for (int k = 0; k < 4; ++k) obb[4 * (size_t)i + k] = s[2 + k];
    }
}

// ---- micro-bench parity kernels (benches/collision.rs, benches/grid_collision.rs) -------------------
__global__ void k_batch_sphere_sphere(size_t n, const float4 *a, const float4 *b, uint8_t *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = a[i], q = b[i];
    out[i] = test_sphere_sphere(p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w);
}
GSC_DEV Obb obb_from16(const float *f)
{
    Obb o;
    o.cx = f[0]; o.cy = f[1]; o.cz = f[2];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) o.m[i][j] = f[3 + 3 * i + j];
    o.hx = f[12]; o.hy = f[13]; o.hz = f[14];
    return o;
}
__global__ void k_batch_sphere_obb(size_t n, const float4 *s, const float *o16, uint8_t *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = s[i];
    out[i] = test_sphere_obb(p.x, p.y, p.z, p.w, obb_from16(o16 + 16 * i));
}
__global__ void k_batch_obb_obb(size_t n, const float *a16, const float *b16, uint8_t *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = test_obb_obb(obb_from16(a16 + 16 * i), obb_from16(b16 + 16 * i));
}
__global__ void k_batch_contact_ss(size_t n, const float4 *hi, const float4 *lo, float4 *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 h = hi[i], l = lo[i];
    out[i] = contact_sphere_sphere(h.x, h.y, h.z, h.w, l.x, l.y, l.z, l.w);
}
__global__ void k_batch_contact_so(size_t n, const float4 *s, const float *o16, const uint8_t *sphere_is_hi, float4 *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = s[i];
    out[i] = contact_sphere_obb(p.x, p.y, p.z, p.w, obb_from16(o16 + 16 * i), sphere_is_hi[i] != 0);
}
__global__ void k_batch_contact_oo(size_t n, const float *a16, const float *b16, float4 *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = contact_obb_obb(obb_from16(a16 + 16 * i), obb_from16(b16 + 16 * i));
}
__global__ void k_batch_aabb_aabb(size_t n, const float *a6, const float *b6, uint8_t *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *a = a6 + 6 * i, *b = b6 + 6 * i;
    const float amn[3] = { a[0], a[1], a[2] }, amx[3] = { a[3], a[4], a[5] };
    const float bmn[3] = { b[0], b[1], b[2] }, bmx[3] = { b[3], b[4], b[5] };
    out[i] = test_aabb(amn, amx, bmn, bmx);
}
__global__ void k_batch_fnv1a(size_t n, const int16_t *cells, uint64_t *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = fnv1a_gridcell(cells[3 * i], cells[3 * i + 1], cells[3 * i + 2]);
}

}  // namespace gsc

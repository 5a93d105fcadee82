// gsc_pairs.cuh -- stage 4 of the collision step: neighbour-cell pair enumeration + AABB filter.
// (included by gsc_kernels.cuh after the cell-table stage; needs FrameBlock, SortedVolumes, kMeta*, kQ*)
//
// Replaces test_cell's candidate generation and the HashMap dedupe of do_narrowphase
// (grid_collision.rs:439-441, 497-511) + AABB::test_aabb (bounding_volume.rs:338-343, 411-419).
//
// Candidates of a cell-sorted volume A are the volumes of the 13 cells that precede A's home cell in key order
// among its 26 neighbours plus A's own cell; a pair inside one cell is taken by the member with the larger global
// index.  Every unordered pair is therefore visited exactly once -- the sort order itself is the dedupe.
// Because keys are x-major / z-minor and start[] is monotone, the candidates of A are FIVE contiguous runs of the
// cell-sorted arrays, one per (dx, dy) row: (-1,-1) (-1,0) (-1,+1) (0,-1) (0,0), each covering cells z-1 .. z+1
// (z-1 .. z for the own row).  The grid carries a one-cell apron of empty cells, so neighbour keys are plain offsets
// of A's key.  Rows / cells that A's own AABB cannot reach are pruned exactly: B in row y+1 (or cell z+1) means
// cell(B.min) > cell(A.max) there, hence B.min > A.max by monotonicity of floor(fl(p / c)).
//
// Two kernels:
//   k_pairs        the production path: one thread per cell-sorted volume A walks A's five runs through L1 (lanes of one
//                  cell walk the same run together, so their candidate loads coalesce), four candidates per trip, doing
//                  only the quantised overlap test; survivors go into private shared-memory slots with a predicated
//                  store.  k_pairs<false>: one CTA per kPairThreads volumes of the whole sorted array.
//   k_pairs_tiled  the `north_star` form, built and measured in round 2 (opt-in: GSC_PAIRS_TILED=1).  One CTA = a TILE
//                  of kTileA consecutive cell-sorted volumes.  Their runs are monotone in the sorted position, so per
//                  row the union of the tile's runs is ONE contiguous segment of the sorted array: the five segments
//                  are staged in shared memory once, coalesced (cp.async), every candidate byte is read from L2/HBM
//                  once per tile.  The (A, run) work is then FLATTENED: the tile's non-empty runs are listed with their
//                  prefix sums and every lane takes an equal, contiguous share of the tile's tests.  Bit-exact, and
//                  1.7-2x SLOWER than k_pairs at 16M (profiles/r02_pairs_tiled_v1_ncu.md: 690 instructions of tile
//                  set-up per 32 volumes against 120, and a divergent run switch in every iteration of the flattened
//                  loop because runs are only 5.6 candidates long).  Tiles whose segments do not fit the staging
//                  buffers (hundreds of volumes per cell) are queued and walked by k_pairs<true>, persistent over the
//                  queue (normally empty: it exits at once).
// Both end in the same tail: the survivors of the quantised filter are classified by shape pair and oriented by global
// index from their 4 B ordering words alone and appended to the three candidate lists of the SAT stage, which runs the
// EXACT inclusive AABB test (on AABBs it re-derives bit for bit from the records it loads anyway) before the shape test.
// (Round 2 measured two alternatives and dropped them: the exact test on the 32 B cell-ordered records in this tail --
// 4 dependent 16 B gathers per survivor, 25% of the kernel's instructions, 31% of its stall samples -- and deciding
// sphere-sphere pairs right here: +0.13 ms in this kernel for -0.01 ms in k_narrow.)
#pragma once

namespace gsc {

struct PairOut {
    uint2 *cand_ss;   // sphere-sphere candidates (pairs that passed the conservative quantised AABB filter)
    uint2 *cand_so;
    uint2 *cand_oo;
    unsigned long long cap_cands;
};
// what the tail needs to classify a filtered pair WITHOUT touching the 32 B volume records: the 4 B ordering words
// (gidx << 4 | flags: box bit, global index) and the sort's index stream (local index of a sorted position)
struct PairRefs {
    const uint32_t *w3;
    const uint32_t *vals;
};

// Classify + orient a filtered (A at sorted position p, B at q) pair from the ordering words alone.  The EXACT
// AABB::test_aabb (bounding_volume.rs:338-343) is not run here: the SAT stage re-derives both AABBs bit for bit from the
// records it loads anyway (a sphere's from centre -+ radius, a box's with obb_aabb_fast from its OBB record -- the very
// function and inputs bvh_update used) and tests them before the shape test.  Round 2: this took four dependent 16 B
// gathers per survivor and their decode out of the pairs kernel (its tail was 25% of the instructions and 31% of the
// stall samples).
// Both members are addressed by their LOCAL index (where the 32 B volume record and the 64 B OBB record live); once the
// colliders are kept in a previous frame's cell order (gsc_reorder_colliders) those are neighbours in memory too.
// item = (later volume [| sphere flag], earlier volume); returns the class 0 sphere-sphere, 1 sphere-box, 2 box-box.
GSC_DEV int pair_classify(uint32_t p, uint32_t q, PairRefs refs, uint2 *item)
{
    const uint32_t aw = __ldg(refs.w3 + p), bw = __ldg(refs.w3 + q);
    const bool a_box = aw & kMetaBox, b_box = bw & kMetaBox;
    const bool a_hi = (aw >> 4) > (bw >> 4);  // tuple = (later index, earlier index), grid_collision.rs:440,497
    const uint32_t a_ref = __ldg(refs.vals + p), b_ref = __ldg(refs.vals + q);
    uint32_t hi = a_hi ? a_ref : b_ref;
    const uint32_t lo = a_hi ? b_ref : a_ref;
    const int cls = (int)a_box + (int)b_box;
    if (cls == 1 && (a_hi ? !a_box : !b_box)) hi |= 0x80000000u;  // the later volume is the sphere
    *item = make_uint2(hi, lo);
    return cls;
}

// slow path of a full CTA list: classify and append one pair straight to the global lists
__device__ __noinline__ void pair_spill(uint32_t p, uint32_t q, PairRefs refs, PairOut out, FrameBlock *fb)
{
    uint2 item;
    const int cls = pair_classify(p, q, refs, &item);
    unsigned long long *ctr = cls == 0 ? &fb->n_ss : (cls == 1 ? &fb->n_so : &fb->n_oo);
    uint2 *dst = cls == 0 ? out.cand_ss : (cls == 1 ? out.cand_so : out.cand_oo);
    const unsigned long long at = atomicAdd(ctr, 1ull);
    if (at < out.cap_cands) dst[at] = item;
}

// ---- the common tail: classify + append a CTA's filtered pairs -------------------------------------------------
template <int WARPS>
struct TailShared {
    uint32_t warp_cls[WARPS][3];
    unsigned long long base[3];
};

// `fetch(i)` yields the i-th listed pair as (p, q) sorted positions.  All threads of the CTA call.
template <int THREADS, class Fetch>
GSC_DEV void pairs_tail(uint32_t have, Fetch fetch, PairRefs refs, const PairOut &out, FrameBlock *__restrict__ fb,
                        TailShared<THREADS / 32> &ts)
{
    constexpr int WARPS = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t base = 0; base < have; base += THREADS) {
        const uint32_t i = base + tid;
        int cls = 3;
        uint2 item = make_uint2(0, 0);
        if (i < have) {
            const uint2 e = fetch(i);
            cls = pair_classify(e.x, e.y, refs, &item);
        }
        uint32_t m[3], rank = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            m[c] = __ballot_sync(0xffffffffu, cls == c);
            if (cls == c) rank = __popc(m[c] & ((1u << lane) - 1u));
        }
        if (lane < 3) ts.warp_cls[warp][lane] = __popc(m[lane]);
        __syncthreads();
        if (tid < 3) {
            uint32_t tot = 0;
            for (int w = 0; w < WARPS; ++w) {
                const uint32_t c = ts.warp_cls[w][tid];
                ts.warp_cls[w][tid] = tot;
                tot += c;
            }
            unsigned long long *ctr = tid == 0 ? &fb->n_ss : (tid == 1 ? &fb->n_so : &fb->n_oo);
            ts.base[tid] = tot ? atomicAdd(ctr, (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        if (cls < 3) {
            uint2 *dst = cls == 0 ? out.cand_ss : (cls == 1 ? out.cand_so : out.cand_oo);
            const unsigned long long at = ts.base[cls] + ts.warp_cls[warp][cls] + rank;
            if (at < out.cap_cands) dst[at] = item;
        }
        __syncthreads();
    }
}

// row r of the half shell: base key of the row's middle cell, relative to A's key
GSC_DEV uint32_t pair_row_key(uint32_t key, int r, uint32_t dz, uint32_t dydz)
{
    return key - (r < 3 ? dydz : 0u) + (r == 2 ? dz : 0u) - ((r == 0 || r == 3) ? dz : 0u);
}

constexpr int kPairRows = 5;

// =================================================================================================================
// k_pairs_tiled
// =================================================================================================================
#ifndef GSC_TILE_A
#define GSC_TILE_A 256
#endif
#ifndef GSC_TILE_PRIV
#define GSC_TILE_PRIV 8
#endif
#ifndef GSC_TILE_CTAS
#define GSC_TILE_CTAS 4
#endif
constexpr int kTileA = GSC_TILE_A;            // volumes (= threads) per CTA
constexpr int kTileWarps = kTileA / 32;
constexpr int kTileCand = 8 * kTileA;         // staged candidate boxes per tile (typical: 5.1 * kTileA)
constexpr int kTileOwn = 3 * kTileA;          // staged ordering words of the own-row segment (typical: 1.1 * kTileA)
constexpr int kTileRuns = 5 * kTileA;         // run records (typical 3.6 * kTileA of the 6 * kTileA possible)
constexpr int kTilePriv = GSC_TILE_PRIV;      // private survivor slots per lane
constexpr int kTileList = 4 * kTileA;         // survivors listed per CTA; denser tiles spill one by one
constexpr int kTileBlock = 4;                 // tests between two drain checks (kTileBlock <= kTilePriv / 2)
// (the tile scan packs runs << 20 | tests into one word: kTileA * kTileCand < 2^20, 6 * kTileA < 2^12)
static_assert(kTileA % 32 == 0 && kTileA <= 256 && kTileCand <= 65536, "tile geometry");
constexpr uint32_t tile_search_step(uint32_t n) { uint32_t s = 1; while (2 * s <= n) s *= 2; return s; }
constexpr uint32_t kTileSearchStep = tile_search_step(kTileRuns);  // highest power of two <= kTileRuns
static_assert(2 * kTileBlock <= kTilePriv, "a block of tests must fit the free private slots");

struct TileShared {
    uint4 run[kTileRuns + 1];      // {first test, first staged candidate, length, A | own-cell << 31}; [R] = sentinel
    uint2 cand[kTileCand];         // staged quantised boxes of the five segments, back to back
    uint32_t cand_w3[kTileOwn];    // ordering words of the own-row segment
    uint2 a_box[kTileA];           // the tile's own quantised boxes
    uint32_t a_w3[kTileA];
    uint32_t priv[kTilePriv * kTileA];  // slot-major: lane t owns priv[j * kTileA + t]
    uint32_t list[kTileList];      // survivors, packed A << 16 | staged index
    uint32_t seg_lo[kPairRows], seg_hi[kPairRows];
    uint32_t count;
    uint32_t scan_warp[kTileWarps];
    TailShared<kTileWarps> tail;
};

// exclusive scan of one value per thread over the CTA (all threads call); *total = sum
template <int WARPS>
GSC_DEV uint32_t block_excl_scan(uint32_t v, uint32_t *s_warp, uint32_t *total)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        const uint32_t s = s_warp[w];
        if (w < warp) base += s;
        tot += s;
    }
    *total = tot;
    __syncthreads();
    return base + inc - v;
}

GSC_DEV void cp_async8(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
GSC_DEV void cp_async4(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
GSC_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__global__ void __launch_bounds__(kTileA, GSC_TILE_CTAS)
k_pairs_tiled(uint32_t n_cap, const uint32_t *__restrict__ n_dev, const uint32_t *__restrict__ keys_a, const uint32_t *__restrict__ keys_b,
              const uint32_t *__restrict__ vals_a, const uint32_t *__restrict__ vals_b, SortedVolumes sv,
              const uint32_t *__restrict__ cell_start, PairOut out, uint32_t *__restrict__ overflow_tiles,
              FrameBlock *__restrict__ fb)
{
    extern __shared__ __align__(16) unsigned char tile_smem[];
    TileShared &sh = *reinterpret_cast<TileShared *>(tile_smem);
    const GridParams g = fb->grid;
    if (!g.valid) return;
    const uint32_t n_total = n_dev ? min(*n_dev, n_cap) : n_cap;  // (device-side count: slab steps only)
    const uint32_t p0 = blockIdx.x * (uint32_t)kTileA;
    if (p0 >= n_total) return;
    const uint32_t *keys = (fb->sort.num_passes & 1) ? keys_b : keys_a;
    const PairRefs refs{ sv.w3, (fb->sort.num_passes & 1) ? vals_b : vals_a };
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < kPairRows) {
        sh.seg_lo[tid] = 0xffffffffu;
        sh.seg_hi[tid] = 0u;
    }
    if (tid == 0) sh.count = 0;
    __syncthreads();

    // ---- phase 1: this thread's volume A and its five runs -------------------------------------------------------
    const uint32_t p = p0 + tid;
    uint32_t lo[kPairRows], hi[kPairRows], qsame = 0;
#pragma unroll
    for (int r = 0; r < kPairRows; ++r) lo[r] = hi[r] = 0;
    {
        uint2 aq = make_uint2(0u, 0u);
        uint32_t a_w3 = 0;
        if (p < n_total) {
            aq = __ldg(sv.q + p);
            const uint32_t key = keys[p];
            a_w3 = __ldg(sv.w3 + p);
            // ghosts only serve as the earlier volume; wide volumes (key == num_cells) go to k_wide_pairs
            const bool active = !(a_w3 & kMetaGhost) && key < g.num_cells;
            const uint32_t dz = (uint32_t)g.dims[2], dydz = (uint32_t)g.dims[1] * dz;
            const uint32_t zhi = (a_w3 & kMetaReachZ) ? 2u : 1u;  // end = start[kk + zhi]
            const bool ry = a_w3 & kMetaReachY;
#pragma unroll
            for (int r = 0; r < kPairRows; ++r) {
                const uint32_t kk = pair_row_key(key, r, dz, dydz);  // (apron => always a valid cell)
                const bool use = active && (r != 2 || ry);
                lo[r] = use ? ldg_gather_u32(cell_start + kk - 1) : 0u;
                hi[r] = use ? ldg_gather_u32(cell_start + kk + (r == 4 ? 1u : zhi)) : 0u;
            }
            qsame = active ? ldg_gather_u32(cell_start + key) : 0u;
        }
        sh.a_box[tid] = aq;
        sh.a_w3[tid] = a_w3;
    }

    // ---- phase 2: per row, the union of the tile's runs is one contiguous segment ---------------------------------
#pragma unroll
    for (int r = 0; r < kPairRows; ++r) {
        const bool has = hi[r] > lo[r];
        const uint32_t mn = __reduce_min_sync(0xffffffffu, has ? lo[r] : 0xffffffffu);
        const uint32_t mx = __reduce_max_sync(0xffffffffu, has ? hi[r] : 0u);
        if (lane == 0 && mx) {
            atomicMin(&sh.seg_lo[r], mn);
            atomicMax(&sh.seg_hi[r], mx);
        }
    }
    __syncthreads();
    uint32_t seg_lo[kPairRows], seg_off[kPairRows + 1];
    seg_off[0] = 0;
#pragma unroll
    for (int r = 0; r < kPairRows; ++r) {
        const uint32_t l = sh.seg_lo[r], h = sh.seg_hi[r];
        seg_lo[r] = h ? l : 0u;
        seg_off[r + 1] = seg_off[r] + (h ? h - l : 0u);
    }
    const uint32_t own_len = seg_off[kPairRows] - seg_off[kPairRows - 1];
    // staged index -> global sorted position (for the few survivors): s + (seg_lo - seg_off) of the segment that holds s
    auto global_pos = [&](uint32_t s) {
        uint32_t d = seg_lo[0] - seg_off[0];
#pragma unroll
        for (int rr = 1; rr < kPairRows; ++rr)
            if (s >= seg_off[rr]) d = seg_lo[rr] - seg_off[rr];
        return s + d;
    };

    // ---- phase 3: count this thread's non-empty runs and tests; scan over the tile --------------------------------
    // the own row is split at A's own cell: [lo, qsame) is taken whole, [qsame, hi) only the earlier volumes (ordering rule)
    uint32_t rl[kPairRows + 1], rb[kPairRows + 1];  // run lengths / first global candidate
#pragma unroll
    for (int r = 0; r < kPairRows - 1; ++r) { rl[r] = hi[r] - lo[r]; rb[r] = lo[r]; }
    rl[4] = qsame > lo[4] ? qsame - lo[4] : 0u;   rb[4] = lo[4];
    rl[5] = hi[4] > qsame ? hi[4] - qsame : 0u;   rb[5] = qsame;
    if (hi[4] == 0) rl[4] = rl[5] = 0;
    uint32_t my_runs = 0, my_tests = 0;
#pragma unroll
    for (int r = 0; r <= kPairRows; ++r) {
        my_runs += rl[r] ? 1u : 0u;
        my_tests += rl[r];
    }
    // tiles whose segments do not fit the staging buffers go to k_pairs (CTA-uniform decisions).  A run lies inside
    // its row's segment, so below this check no thread has more than kTileCand tests and the packed scan cannot carry
    bool overflow = seg_off[kPairRows] > (uint32_t)kTileCand || own_len > (uint32_t)kTileOwn;
    uint32_t total = 0, excl = 0;
    if (!overflow) {
        excl = block_excl_scan<kTileWarps>((my_runs << 20) | my_tests, sh.scan_warp, &total);
        overflow = (total >> 20) > (uint32_t)kTileRuns;
    }
    if (overflow) {
        if (tid == 0) overflow_tiles[atomicAdd(&fb->n_overflow, 1u)] = blockIdx.x;
        return;
    }
    const uint32_t R = total >> 20, W = total & 0xfffffu;
    if (W == 0) return;

    // ---- phase 4: stage the segments (asynchronously) and list the runs --------------------------------------------
#pragma unroll
    for (int r = 0; r < kPairRows; ++r) {
        const uint32_t len = seg_off[r + 1] - seg_off[r];
        for (uint32_t i = tid; i < len; i += kTileA) cp_async8(&sh.cand[seg_off[r] + i], sv.q + seg_lo[r] + i);
    }
    for (uint32_t i = tid; i < own_len; i += kTileA) cp_async4(&sh.cand_w3[i], sv.w3 + seg_lo[kPairRows - 1] + i);
    {
        uint32_t ri = excl >> 20, P = excl & 0xfffffu;
#pragma unroll
        for (int r = 0; r <= kPairRows; ++r) {
            if (rl[r]) {
                const int row = r < kPairRows ? r : kPairRows - 1;
                GSC_CHECK(fb, ri < (uint32_t)kTileRuns && rb[r] >= seg_lo[row] && seg_off[row] + (rb[r] - seg_lo[row]) + rl[r] <= seg_off[row + 1]);
                sh.run[ri] = make_uint4(P, seg_off[row] + (rb[r] - seg_lo[row]), rl[r], (uint32_t)tid | (r == kPairRows ? 0x80000000u : 0u));
                ++ri;
                P += rl[r];
            }
        }
        if (tid == 0) sh.run[R] = make_uint4(W, 0u, 0u, 0u);
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- phase 5: every lane takes an equal, contiguous share of the tile's W tests --------------------------------
    const uint32_t K = (W + kTileA - 1) / kTileA;
    const uint32_t t0 = tid * K;
    uint32_t cnt = t0 < W ? min(K, W - t0) : 0u;
    const uint32_t pa0 = (uint32_t)__cvta_generic_to_shared(&sh.priv[tid]);
    constexpr uint32_t kSlotStride = kTileA * 4u;
    uint32_t pa = pa0;
    if (cnt) {
        uint32_t i = 0;  // the run that holds test t0: largest i with run[i].first <= t0
#pragma unroll
        for (uint32_t step = kTileSearchStep; step; step >>= 1) {
            const uint32_t j = i + step;
            if (j < R && sh.run[j].x <= t0) i = j;
        }
        uint4 rec = sh.run[i];
        uint32_t sidx = rec.y + (t0 - rec.x), rem = rec.z - (t0 - rec.x), tag = rec.w;
        uint2 aq = sh.a_box[tag & 0xffffu];
        uint32_t aw3 = sh.a_w3[tag & 0xffffu];
        const uint32_t own_off = seg_off[kPairRows - 1];
        while (cnt) {
            const uint32_t nb = min(cnt, (uint32_t)kTileBlock);
#pragma unroll
            for (int k = 0; k < kTileBlock; ++k) {
                if ((uint32_t)k < nb) {
                    if (rem == 0) {  // next run
                        rec = sh.run[++i];
                        sidx = rec.y;
                        rem = rec.z;
                        tag = rec.w;
                        aq = sh.a_box[tag & 0xffffu];
                        aw3 = sh.a_w3[tag & 0xffffu];
                    }
                    // quantised AABB::test_aabb (bounding_volume.rs:338-343, 411-419; inclusive): three packed
                    // differences per subtraction, overlap on every axis iff no field went negative
                    GSC_CHECK(fb, sidx < seg_off[kPairRows] && i < R && (tag & 0xffffu) < (uint32_t)kTileA);
                    const uint2 b = sh.cand[sidx];
                    const uint32_t t1 = aq.y - b.x;  // A.qmax - B.qmin per axis
                    const uint32_t t2 = b.y - aq.x;  // B.qmax - A.qmin per axis
                    bool pass = ((t1 | t2) & kQSign) == 0u;
                    if (tag >> 31) pass = pass && sh.cand_w3[sidx - own_off] < aw3;  // same cell: the later global index takes the pair
                    if (pass) {
                        GSC_CHECK(fb, pa - pa0 < kTilePriv * kSlotStride && (!(tag >> 31) || sidx - own_off < own_len));
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(pa), "r"((tag << 16) | sidx) : "memory");
                        pa += kSlotStride;
                    }
                    ++sidx;
                    --rem;
                }
            }
            cnt -= nb;
            if (pa - pa0 > (kTilePriv - kTileBlock) * kSlotStride) {  // fewer than one block of free slots left
                const uint32_t have = (pa - pa0) / kSlotStride;
                const uint32_t at = atomicAdd(&sh.count, have);
                for (uint32_t j = 0; j < have; ++j) {
                    const uint32_t w = sh.priv[j * kTileA + tid];
                    if (at + j < (uint32_t)kTileList) sh.list[at + j] = w;
                    else pair_spill(p0 + (w >> 16), global_pos(w & 0xffffu), refs, out, fb);
                }
                pa = pa0;
            }
        }
    }

    // ---- compact the private slots into the CTA list: warp scan + one shared atomic per warp ----
    {
        const uint32_t have = (pa - pa0) / kSlotStride;
        uint32_t incl = have;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        uint32_t wbase = 0;
        if (lane == 31 && incl) wbase = atomicAdd(&sh.count, incl);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        const uint32_t at = wbase + incl - have;
        for (uint32_t j = 0; j < have; ++j) {
            const uint32_t w = sh.priv[j * kTileA + tid];
            if (at + j < (uint32_t)kTileList) sh.list[at + j] = w;
            else pair_spill(p0 + (w >> 16), global_pos(w & 0xffffu), refs, out, fb);
        }
    }
    __syncthreads();

    // ---- decide + append the tile's survivors ----
    const uint32_t listed = min(sh.count, (uint32_t)kTileList);
    auto fetch = [&](uint32_t i) {
        const uint32_t w = sh.list[i];
        return make_uint2(p0 + (w >> 16), global_pos(w & 0xffffu));
    };
    pairs_tail<kTileA>(listed, fetch, refs, out, fb, sh.tail);
}

// =================================================================================================================
// k_pairs: thread-per-volume walk through L1 (round 1), persistent over the tiles k_pairs_tiled could not stage
// =================================================================================================================
// One thread per cell-sorted volume A; each lane walks its five runs doing only the quantised overlap test; a
// survivor's position is stored with a predicated write into a few PRIVATE shared-memory slots of the lane (no
// atomics, no divergent branch).  At the end of a slice the slots are compacted into one CTA list and handed to the
// common tail.
#ifndef GSC_PAIR_THREADS
#define GSC_PAIR_THREADS 64
#endif
#ifndef GSC_PAIR_MIN_CTAS
#define GSC_PAIR_MIN_CTAS 24  // 16 / 20 / 24 resident CTAs measured with the round-2 tail: 0.926 / 0.852 / 0.842 ms at 16M
#endif
#ifndef GSC_PAIR_PRIV
#define GSC_PAIR_PRIV 12  // 8 / 12 / 16 / 24 measured: 0.942 / 0.903 / 0.913 / 0.936 ms at 16M
#endif
constexpr int kPairThreads = GSC_PAIR_THREADS;
constexpr int kPairWarps = kPairThreads / 32;
constexpr int kPairTrip = 4;    // candidates per loop trip
constexpr int kPairPriv = GSC_PAIR_PRIV;    // private survivor slots per lane; a lane that fills them drains early
constexpr int kPairCap = 4 * kPairThreads;   // survivors staged per CTA (4 per volume); denser CTAs spill one by one
static_assert(kTileA % kPairThreads == 0, "a tile is cut into whole slices");

struct PairShared {
    uint32_t priv[kPairPriv * kPairThreads];  // slot-major: lane t owns priv[j * kPairThreads + t]
    uint2 list[kPairCap];
    uint32_t count;
    TailShared<kPairWarps> tail;
};

// a lane whose private slots are (nearly) full moves them to the CTA list (rare: >= kPairPriv - 1 survivors)
__device__ __noinline__ void pair_drain_lane(PairShared &sh, uint32_t p, uint32_t cnt, int tid, PairRefs refs, PairOut out, FrameBlock *fb)
{
    const uint32_t at = atomicAdd(&sh.count, cnt);
    for (uint32_t j = 0; j < cnt; ++j) {
        const uint32_t q = sh.priv[j * kPairThreads + tid];
        if (at + j < (uint32_t)kPairCap) sh.list[at + j] = make_uint2(p, q);
        else pair_spill(p, q, refs, out, fb);
    }
}

// QUEUE == false: one CTA per kPairThreads volumes over the whole sorted array (the round-1 form of the stage);
// QUEUE == true: persistent CTAs over the tiles queued in out.overflow_tiles by k_pairs_tiled
template <bool QUEUE>
__global__ void __launch_bounds__(kPairThreads, QUEUE ? 16 : GSC_PAIR_MIN_CTAS)
k_pairs(uint32_t n_cap, const uint32_t *__restrict__ n_dev, const uint32_t *__restrict__ keys_a, const uint32_t *__restrict__ keys_b,
        const uint32_t *__restrict__ vals_a, const uint32_t *__restrict__ vals_b, SortedVolumes sv,
        const uint32_t *__restrict__ cell_start, PairOut out, const uint32_t *__restrict__ overflow_tiles, FrameBlock *__restrict__ fb)
{
    __shared__ PairShared sh;
    const GridParams g = fb->grid;
    if (!g.valid) return;
    const uint32_t n_total = n_dev ? min(*n_dev, n_cap) : n_cap;  // (device-side count: slab steps only)
    const uint32_t *keys = (fb->sort.num_passes & 1) ? keys_b : keys_a;
    const PairRefs refs{ sv.w3, (fb->sort.num_passes & 1) ? vals_b : vals_a };
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr uint32_t kSlices = kTileA / kPairThreads;  // slices of kPairThreads volumes per tile
    const uint32_t n_slices = QUEUE ? fb->n_overflow * kSlices : gridDim.x;
    for (uint32_t e = blockIdx.x; e < n_slices; e += gridDim.x) {
        const uint32_t p = QUEUE ? overflow_tiles[e / kSlices] * (uint32_t)kTileA + (e % kSlices) * kPairThreads + tid
                                 : e * (uint32_t)kPairThreads + tid;
        if (!QUEUE && e * (uint32_t)kPairThreads >= n_total) return;
        if (tid == 0) sh.count = 0;
        __syncthreads();

        // A's quantised box: packed min corner and guarded packed max corner (see SortedVolumes)
        uint32_t a_min = 0, a_maxg = 0, a_w3 = 0, qsame = 0;
        uint32_t lo[kPairRows], hi[kPairRows];
#pragma unroll
        for (int r = 0; r < kPairRows; ++r) lo[r] = hi[r] = 0;
        if (p < n_total) {
            const uint2 aq = __ldg(sv.q + p);
            const uint32_t key = keys[p];
            a_min = aq.x;
            a_maxg = aq.y;
            a_w3 = __ldg(sv.w3 + p);
            const bool active = !(a_w3 & kMetaGhost) && key < g.num_cells;
            const uint32_t dz = (uint32_t)g.dims[2], dydz = (uint32_t)g.dims[1] * dz;
            const uint32_t zhi = (a_w3 & kMetaReachZ) ? 2u : 1u;
            const bool ry = a_w3 & kMetaReachY;
#pragma unroll
            for (int r = 0; r < kPairRows; ++r) {
                const uint32_t kk = pair_row_key(key, r, dz, dydz);
                const bool use = active && (r != 2 || ry);
                lo[r] = use ? ldg_gather_u32(cell_start + kk - 1) : 0u;
                hi[r] = use ? ldg_gather_u32(cell_start + kk + (r == 4 ? 1u : zhi)) : 0u;
            }
            qsame = active ? ldg_gather_u32(cell_start + key) : 0u;
        }

        // The five runs are walked one after the other (row-outer); next free private slot as a shared-window byte
        // address (st.shared with a plain register address)
        const uint32_t pa0 = (uint32_t)__cvta_generic_to_shared(&sh.priv[tid]);
        constexpr uint32_t kSlotStride = kPairThreads * 4u;
        uint32_t pa = pa0;
        auto test = [&](uint32_t q, uint2 b, bool take) {
            const uint32_t t1 = a_maxg - b.x;  // A.qmax - B.qmin per axis
            const uint32_t t2 = b.y - a_min;   // B.qmax - A.qmin per axis
            if ((((t1 | t2) & kQSign) == 0u) && take) {
                GSC_CHECK(fb, pa - pa0 < kPairPriv * kSlotStride && q < n_total);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(pa), "r"(q) : "memory");
                pa += kSlotStride;
            }
        };
        auto drain_if_full = [&]() {
            if (pa - pa0 > (kPairPriv - kPairTrip) * kSlotStride) {  // fewer than one trip of free slots left
                pair_drain_lane(sh, p, (pa - pa0) / kSlotStride, tid, refs, out, fb);
                pa = pa0;
            }
        };
#pragma unroll
        for (int r = 0; r < kPairRows; ++r) {
            const bool own = r == kPairRows - 1;  // own row (cells z-1, z): same-cell pairs go to the larger global index
            const uint32_t qe = hi[r];
#pragma unroll 1
            for (uint32_t q = lo[r]; q < qe; q += kPairTrip) {
                // (reads up to kPairTrip - 1 records past the run: always inside the array, which is padded, and
                // masked out of the test below)
                const uint2 *bp = sv.q + q;
                uint2 b[kPairTrip];
                uint32_t w[kPairTrip];
#pragma unroll
                for (int k = 0; k < kPairTrip; ++k) b[k] = __ldg(bp + k);
                if (own) {
#pragma unroll
                    for (int k = 0; k < kPairTrip; ++k) w[k] = __ldg(sv.w3 + q + k);
                }
#pragma unroll
                for (int k = 0; k < kPairTrip; ++k) {
                    const uint32_t qk = q + k;
                    test(qk, b[k], qk < qe && (!own || qk < qsame || w[k] < a_w3));
                }
                drain_if_full();
            }
            __syncwarp();  // reconverge before the next run: lanes must walk each run together
        }

        // ---- compact the private slots into the CTA list: warp scan + one shared atomic per warp ----
        {
            const uint32_t cnt = (pa - pa0) / kSlotStride;
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            uint32_t wbase = 0;
            if (lane == 31 && incl) wbase = atomicAdd(&sh.count, incl);
            wbase = __shfl_sync(0xffffffffu, wbase, 31);
            const uint32_t at = wbase + incl - cnt;
            for (uint32_t j = 0; j < cnt; ++j) {
                const uint32_t q = sh.priv[j * kPairThreads + tid];
                if (at + j < (uint32_t)kPairCap) sh.list[at + j] = make_uint2(p, q);
                else pair_spill(p, q, refs, out, fb);
            }
        }
        __syncthreads();
        pairs_tail<kPairThreads>(min(sh.count, (uint32_t)kPairCap), [&](uint32_t i) { return sh.list[i]; }, refs, out, fb, sh.tail);
        if (!QUEUE) return;
        __syncthreads();
    }
}

}  // namespace gsc

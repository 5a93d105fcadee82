// gsc_sort.cuh -- hand-written LSD radix sort of (cell key, collider index) pairs for sm_100a.
//
// Replaces the reference's per-thread `HashMap<GridCell, Vec<*const BoundVolume>>` grouping
// (grid_collision.rs:107, 434-444): colliders are grouped by cell by SORTING (key, index) pairs.
//
// One-sweep design, 8-bit digits: every pass reads each (key, value) once and writes it once.
//   * digit histograms of ALL passes are accumulated by the producer of the keys (k_cell_keys) or by
//     k_radix_histogram, so a pass needs no counting sweep of its own;
//   * a pass kernel takes a tile of THREADS*IPT items per CTA, ranks them with warp-level
//     __match_any_sync multisplit (stable), publishes its per-digit tile counts and resolves the
//     exclusive prefix over earlier tiles with decoupled look-back (flag bits in the same word as the
//     count, so no fences), stages the tile digit-sorted in shared memory and writes each digit run
//     contiguously;
//   * tile ids come from an atomic ticket, so a tile's predecessors are always resident or done.
// HBM bytes per pass: 16 B per pair (8 read + 8 written); the first pass synthesises the index and
// reads only 4 B.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gsc {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kMaxPasses = 4;  // 32-bit keys
constexpr int kSortThreads = 256;
constexpr int kSortIpt = 16;
constexpr int kSortTile = kSortThreads * kSortIpt;

constexpr uint32_t kLbAgg = 1u << 30;   // tile aggregate available
constexpr uint32_t kLbIncl = 1u << 31;  // inclusive prefix available
constexpr uint32_t kLbMask = (1u << 30) - 1;

// number of 8-bit passes for keys < 2^key_bits
__host__ __device__ inline int radix_passes(int key_bits) { return key_bits <= 0 ? 0 : (key_bits + kRadixBits - 1) / kRadixBits; }

// Per-step control block living in device memory: written by k_grid_setup, read by the sort passes,
// so that the number of passes (a function of the data-dependent grid size) needs no host round trip.
struct SortControl {
    int num_passes;
};

// exclusive scan of one value per thread over the first 256 threads of a CTA (all threads call)
__device__ __forceinline__ uint32_t block_excl_scan256(uint32_t v, uint32_t *s_warp /*[8]*/, uint32_t *total)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (warp < 8 && lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t s = s_warp[w];
        if (w < warp) base += s;
        tot += s;
    }
    if (total) *total = tot;
    __syncthreads();
    return base + inc - v;
}

// Standalone histogram of all passes (used when the keys were not produced by k_cell_keys).
__global__ void __launch_bounds__(256) k_radix_histogram(const uint32_t *__restrict__ keys, uint32_t n, int passes,
                                                         uint32_t *__restrict__ hist /*[kMaxPasses][256]*/)
{
    __shared__ uint32_t s_hist[kMaxPasses][kRadix];
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t k = keys[i];
        for (int p = 0; p < passes; ++p) atomicAdd(&s_hist[p][(k >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) {
        const uint32_t c = (&s_hist[0][0])[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// One radix pass.  grid = ceil(n / kSortTile) CTAs of kSortThreads threads.
//   keys_in/vals_in -> keys_out/vals_out; vals_in == nullptr means "value = position" (first pass).
//   ctl->num_passes is data dependent: passes beyond it exit at once and the ping-pong simply stops.
// Register budget: 64 (4 CTAs = 32 warps per SM).  Values are loaded only after the keys have been
// ranked and staged, so keys and values are never live together.
constexpr int kLookBatch = 8;  // predecessors inspected per look-back round trip (independent loads)

__global__ void __launch_bounds__(kSortThreads, 4)
k_onesweep_pass(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t n, int pass,
                const SortControl *__restrict__ ctl, const uint32_t *__restrict__ hist /*[kMaxPasses][256]*/,
                uint32_t *lookback /*[kMaxPasses][tiles][256]*/, uint32_t *tile_counter /*[kMaxPasses]*/,
                uint32_t num_tiles)
{
    constexpr int WARPS = kSortThreads / 32;
    constexpr int WTILE = 32 * kSortIpt;
    __shared__ uint32_t s_hist[WARPS][kRadix];
    __shared__ uint32_t s_keys[kSortTile];
    __shared__ uint32_t s_vals[kSortTile];
    __shared__ uint32_t s_goff[kRadix];
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_tile;

    if (pass >= ctl->num_passes) return;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int shift = pass * kRadixBits;
    if (tid == 0) s_tile = atomicAdd(&tile_counter[pass], 1u);
    for (int i = tid; i < WARPS * kRadix; i += kSortThreads) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * (uint32_t)kSortTile;
    const uint32_t warp_base = tile_base + warp * WTILE + lane;

    uint32_t key[kSortIpt];
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j) {
        const uint32_t idx = warp_base + j * 32;
        key[j] = idx < n ? keys_in[idx] : 0xffffffffu;
    }

    // ---- stable ranking inside the warp: order is (j, lane) == original order ----
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint16_t rank[kSortIpt];
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j) {
        const uint32_t d = (key[j] >> shift) & 255u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t before = s_hist[warp][d];
        const uint32_t r = __popc(peers & lt_mask);
        __syncwarp();
        if (r == 0) s_hist[warp][d] = before + __popc(peers);
        __syncwarp();
        rank[j] = (uint16_t)(before + r);
    }
    __syncthreads();

    // ---- per digit (thread t owns digit t): warp counts -> warp offsets, tile count ----
    uint32_t tile_count = 0;
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t c = s_hist[w][tid];
            s_hist[w][tid] = run;
            run += c;
        }
        tile_count = run;
    }
    volatile uint32_t *lb = lookback + (size_t)pass * num_tiles * kRadix;
    // publish the aggregate as early as possible so later tiles can look back through us
    lb[(size_t)tile * kRadix + tid] = tile_count | (tile == 0 ? kLbIncl : kLbAgg);

    const uint32_t digit_base = block_excl_scan256(hist[pass * kRadix + tid], s_warp, nullptr);
    const uint32_t bin_start = block_excl_scan256(tile_count, s_warp, nullptr);

    // ---- decoupled look-back, kLookBatch predecessors per round trip ----
    uint32_t excl = 0;
    if (tile != 0) {
        int prev = (int)tile - 1;
        bool done = false;
        while (!done) {
            uint32_t v[kLookBatch];
#pragma unroll
            for (int j = 0; j < kLookBatch; ++j)
                v[j] = (prev - j >= 0) ? (uint32_t)lb[(size_t)(prev - j) * kRadix + tid] : (uint32_t)(1u << 31);  // before tile 0: empty inclusive prefix
            int used = 0;
#pragma unroll
            for (int j = 0; j < kLookBatch; ++j) {
                if (done || used != j) continue;
                if ((v[j] & ~kLbMask) == 0) continue;  // not published yet: retry from this one
                excl += v[j] & kLbMask;
                ++used;
                if (v[j] & kLbIncl) done = true;
            }
            prev -= used;
        }
        lb[(size_t)tile * kRadix + tid] = (excl + tile_count) | kLbIncl;
    }
    // position of digit run d in the output minus its position in the tile
    s_goff[tid] = digit_base + excl - bin_start;
    // per-warp offsets become absolute tile positions
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s_hist[w][tid] += bin_start;
    __syncthreads();

    // ---- stage the tile digit-sorted in shared memory, then write digit runs contiguously ----
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j) {
        const uint32_t d = (key[j] >> shift) & 255u;
        const uint32_t pos = (uint32_t)rank[j] + s_hist[warp][d];
        rank[j] = (uint16_t)pos;
        s_keys[pos] = key[j];
    }
    if (vals_in) {
#pragma unroll
        for (int j = 0; j < kSortIpt; ++j) {
            const uint32_t idx = warp_base + j * 32;
            key[j] = idx < n ? vals_in[idx] : 0u;  // key[] now holds the values
        }
    } else {
#pragma unroll
        for (int j = 0; j < kSortIpt; ++j) key[j] = warp_base + j * 32;
    }
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j) s_vals[rank[j]] = key[j];
    __syncthreads();
    const uint32_t remaining = n - tile_base;
    const uint32_t valid = remaining < (uint32_t)kSortTile ? remaining : (uint32_t)kSortTile;
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j) {
        const uint32_t i = tid + j * kSortThreads;
        if (i < valid) {
            const uint32_t k = s_keys[i];
            const uint32_t dst = s_goff[(k >> shift) & 255u] + i;
            keys_out[dst] = k;
            vals_out[dst] = s_vals[i];
        }
    }
}

}  // namespace gsc

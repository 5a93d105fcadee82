// gsc_sort.cuh -- hand-written LSD radix sort of (cell key, collider index) pairs for sm_100a.
//
// Replaces the reference's per-thread `HashMap<GridCell, Vec<*const BoundVolume>>` grouping
// (grid_collision.rs:107, 434-444): colliders are grouped by cell by SORTING (key, index) pairs.
//
// LSD, digits of up to 9 bits (the width is chosen per sort so that the passes share the key bits evenly), three kernels per pass (upsweep / scan / downsweep, see below):
//   * the downsweep takes a tile of THREADS*IPT items per CTA, ranks them with warp-level
//     __match_any_sync multisplit (stable), adds the scanned tile/digit offsets, stages the tile
//     digit-sorted in shared memory and writes each digit run contiguously.
// HBM bytes per pass: 20 B per pair (4 upsweep + 8 read + 8 written); the first pass synthesises the
// index and reads only 4 + 4 B.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gsc {

constexpr int kRadixBits = 9;            // widest digit: 512 bins
constexpr int kRadix = 1 << kRadixBits;
constexpr int kMaxPasses = 4;            // 32-bit keys
constexpr int kSortThreads = 256;        // each thread owns the two digits 2t and 2t+1
constexpr int kSortIpt = 16;
constexpr int kSortTile = kSortThreads * kSortIpt;

// passes for keys < 2^key_bits, and the digit width the passes share: 27-bit cell keys (16M colliders at 0.25 per cell)
// take 3 passes of 9 bits instead of 4 of 8; 24 bits take 3 of 8; 32 bits 4 of 8
__host__ __device__ inline int radix_passes(int key_bits) { return key_bits <= 0 ? 0 : (key_bits + kRadixBits - 1) / kRadixBits; }
__host__ __device__ inline int radix_digit_bits(int key_bits)
{
    const int p = radix_passes(key_bits);
    return p ? (key_bits + p - 1) / p : 0;
}

// Per-step control block living in device memory: written by k_grid_setup, read by the sort passes,
// so that the number of passes (a function of the data-dependent grid size) needs no host round trip.
struct SortControl {
    int num_passes;
    int key_bits;    // significant key bits
    int digit_bits;  // radix_digit_bits(key_bits): every pass ranks on this many bits (the last one possibly fewer)
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

// ---- one radix pass = upsweep (per-tile digit counts) -> scan (over tiles, per digit) -> downsweep ----
// A decoupled look-back ("onesweep") pass was measured first: on B200 the sort is instruction-issue bound, not
// bandwidth bound (44 GB/s of HBM per SM), and the look-back's polling loops were half of all issued
// instructions.  The extra 4 B/item read of the upsweep costs ~20 us per pass at 16M; the polling cost 70.
//   tile_hist[d * num_tiles + t] : count of digit d in tile t, then its exclusive prefix over tiles
//   digit_total[d]               : count of digit d in the whole array
//   ctl->num_passes is data dependent: passes beyond it exit at once and the ping-pong simply stops.

// digit counters of one warp as u16 pairs: word t holds the digits 2t (low half) and 2t+1 (high half)
__device__ __forceinline__ void hist_add(uint16_t *h, uint32_t d) { atomicAdd(reinterpret_cast<uint32_t *>(h) + (d >> 1), 1u << ((d & 1u) * 16)); }

__global__ void __launch_bounds__(kSortThreads)
k_radix_upsweep(const uint32_t *__restrict__ keys_in, uint32_t n, int pass, const SortControl *__restrict__ ctl,
                uint32_t *__restrict__ tile_hist, uint32_t num_tiles, const uint32_t *__restrict__ n_dev)
{
    constexpr int WARPS = kSortThreads / 32;
    __shared__ __align__(16) uint16_t s_hist[WARPS][kRadix];  // a warp counts at most 512 items: u16 is plenty
    if (pass >= ctl->num_passes) return;
    if (n_dev) {  // launched for a capacity: the real count lives on the device (multi-GPU steps: local + ghosts)
        n = min(n, *n_dev);
        num_tiles = (n + kSortTile - 1) / kSortTile;
        if (blockIdx.x >= num_tiles) return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int shift = pass * ctl->digit_bits;
    const uint32_t mask = (1u << ctl->digit_bits) - 1u;
#pragma unroll
    for (int i = 0; i < WARPS * kRadix / (8 * kSortThreads); ++i)
        reinterpret_cast<uint4 *>(&s_hist[0][0])[tid + i * kSortThreads] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    const uint32_t tile = blockIdx.x, tile_base = tile * (uint32_t)kSortTile;
    const uint32_t warp_base = tile_base + warp * (32 * kSortIpt) + lane;
    uint32_t key[kSortIpt];
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j) {
        const uint32_t idx = warp_base + j * 32;
        key[j] = idx < n ? keys_in[idx] : 0xffffffffu;
    }
    // plain shared atomics on the warp's private counters (MATCH.ANY-based aggregation was measured 5x slower:
    // the instruction serialises over the distinct values in the warp)
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j) hist_add(s_hist[warp], (key[j] >> shift) & mask);
    __syncthreads();
    uint32_t c0 = 0, c1 = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        const uint32_t v = reinterpret_cast<const uint32_t *>(s_hist[w])[tid];
        c0 += v & 0xffffu;
        c1 += v >> 16;
    }
    tile_hist[(size_t)(2 * tid) * num_tiles + tile] = c0;  // (padding of the last tile lands in the top digit of that tile: harmless)
    tile_hist[(size_t)(2 * tid + 1) * num_tiles + tile] = c1;
}

// grid = kRadix CTAs: CTA d turns the counts of digit d into exclusive prefixes over the tiles.
// The row is pulled through shared memory so that global accesses are coalesced.
constexpr int kScanChunk = 4096;  // tiles per round (16 KB of shared memory)

__global__ void __launch_bounds__(256)
k_radix_scan(uint32_t *__restrict__ tile_hist, uint32_t num_tiles, int pass, const SortControl *__restrict__ ctl,
             uint32_t *__restrict__ digit_total, const uint32_t *__restrict__ n_dev)
{
    constexpr int PER = kScanChunk / 256;                 // entries per thread
    __shared__ uint32_t s_row[kScanChunk + kScanChunk / PER];  // padded: entry i lives at i + i / PER (no bank conflicts)
    __shared__ uint32_t s_warp[8];
    if (pass >= ctl->num_passes) return;
    if (n_dev) num_tiles = min(num_tiles, (*n_dev + kSortTile - 1) / kSortTile);
    if (blockIdx.x >= (1u << ctl->digit_bits)) {  // digits this sort does not use
        if (threadIdx.x == 0) digit_total[blockIdx.x] = 0;
        return;
    }
    uint32_t *row = tile_hist + (size_t)blockIdx.x * num_tiles;
    const int tid = threadIdx.x;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < num_tiles; base += kScanChunk) {
        const uint32_t cnt = min((uint32_t)kScanChunk, num_tiles - base);
        for (uint32_t i = tid; i < kScanChunk; i += 256) s_row[i + i / PER] = i < cnt ? row[base + i] : 0u;
        __syncthreads();
        uint32_t v[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            v[k] = s_row[tid * (PER + 1) + k];
            sum += v[k];
        }
        uint32_t total = 0;
        uint32_t run = carry + block_excl_scan256(sum, s_warp, &total);
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            s_row[tid * (PER + 1) + k] = run;
            run += v[k];
        }
        __syncthreads();
        for (uint32_t i = tid; i < cnt; i += 256) row[base + i] = s_row[i + i / PER];
        carry += total;
        __syncthreads();
    }
    if (tid == 0) digit_total[blockIdx.x] = carry;
}

// Stable ranking of a warp's kSortIpt x 32 keys by their digit: order is (j, lane) == original order.  Peers = the
// lanes holding the same digit, from one ballot per digit bit (MATCH.ANY serialises over the distinct values of the
// warp and was measured ~50 cycles per instruction and SM).  A digit narrower than BITS only adds ballots of zeros.
template <int BITS>
__device__ __forceinline__ void radix_rank(const uint32_t (&key)[kSortIpt], int shift, uint32_t mask, uint32_t *hist /* this warp's */,
                                           int lane, uint16_t (&rank)[kSortIpt])
{
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j) {
        const uint32_t d = (key[j] >> shift) & mask;
        uint32_t peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < BITS; ++b) {
            // peers &= (my bit b is set) ? ballot : ~ballot -- as four instructions (LOP3 to a predicate, VOTE, SEL, LOP3);
            // written in C++ the compiler derives the ballot predicate and the select mask separately (six)
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t, bal, m;\n\t"
                "and.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\t"
                "vote.sync.ballot.b32 bal, p, 0xffffffff;\n\t"
                "selp.b32 m, 0, 0xffffffff, p;\n\t"
                "lop3.b32 %0, %0, bal, m, 0x60;\n\t}"
                : "+r"(peers)
                : "r"(d), "r"(1u << b));
        }
        const uint32_t before = hist[d];
        const uint32_t r = __popc(peers & lt_mask);
        __syncwarp();
        if (r == 0) hist[d] = before + __popc(peers);
        __syncwarp();
        rank[j] = (uint16_t)(before + r);
    }
}

// grid = num_tiles CTAs of kSortThreads threads, 64 registers (4 CTAs = 32 warps per SM).
//   keys_in/vals_in -> keys_out/vals_out; vals_in == nullptr means "value = position" (first pass).
// ncu (round 1): this kernel is bound by the shared-memory / MIO pipe (mio_throttle + short_scoreboard), not by
// HBM: every shared-memory instruction below is there because it has to be.
__global__ void __launch_bounds__(kSortThreads, 4)
k_radix_downsweep(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                  uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t n, int pass,
                  const SortControl *__restrict__ ctl, const uint32_t *__restrict__ tile_hist,
                  const uint32_t *__restrict__ digit_total, uint32_t num_tiles, const uint32_t *__restrict__ n_dev)
{
    constexpr int WARPS = kSortThreads / 32;
    constexpr int WTILE = 32 * kSortIpt;
    // per-warp digit counts, then offsets within the digit; dead once the keys are staged, so the staged values
    // reuse the same 16 KB (keeps the kernel under the 48 KB of static shared memory)
    static_assert(sizeof(uint32_t) * WARPS * kRadix == sizeof(uint32_t) * kSortTile, "hist/vals alias");
    __shared__ __align__(16) uint32_t s_alias[kSortTile];
    uint32_t(*s_hist)[kRadix] = reinterpret_cast<uint32_t(*)[kRadix]>(s_alias);
    uint32_t *s_vals = s_alias;
    __shared__ uint32_t s_keys[kSortTile];
    __shared__ uint32_t s_bin[kRadix];   // first tile position of each digit
    __shared__ uint32_t s_goff[kRadix];  // output position of digit run d minus its tile position
    __shared__ uint32_t s_warp[8];

    if (pass >= ctl->num_passes) return;
    if (n_dev) {
        n = min(n, *n_dev);
        num_tiles = (n + kSortTile - 1) / kSortTile;
        if (blockIdx.x >= num_tiles) return;
    }

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int shift = pass * ctl->digit_bits;
    const uint32_t mask = (1u << ctl->digit_bits) - 1u;
#pragma unroll
    for (int i = 0; i < WARPS * kRadix / (4 * kSortThreads); ++i)
        reinterpret_cast<uint4 *>(s_alias)[tid + i * kSortThreads] = make_uint4(0u, 0u, 0u, 0u);
    const uint32_t tile = blockIdx.x;
    const uint32_t tile_base = tile * (uint32_t)kSortTile;
    const uint32_t warp_base = tile_base + warp * WTILE + lane;
    // this tile's output offsets for the thread's two digits: issued now, consumed after the ranking
    const uint32_t tile_excl0 = tile_hist[(size_t)(2 * tid) * num_tiles + tile];
    const uint32_t tile_excl1 = tile_hist[(size_t)(2 * tid + 1) * num_tiles + tile];
    const uint2 dtotal = reinterpret_cast<const uint2 *>(digit_total)[tid];

    uint32_t key[kSortIpt];
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j) {
        const uint32_t idx = warp_base + j * 32;
        key[j] = idx < n ? keys_in[idx] : 0xffffffffu;
    }
    __syncthreads();

    // ---- stable ranking inside the warp: order is (j, lane) == original order ----
    // peers = lanes holding the same digit, from one ballot per significant digit bit (MATCH.ANY serialises over
    // the distinct values of the warp and was measured ~50 cycles per instruction and SM)
    const int digit_bits = min(ctl->digit_bits, ctl->key_bits - shift);
    uint16_t rank[kSortIpt];
#ifndef GSC_SORT_RANK_GENERIC
    // the ballot loop is unrolled for the digit width at hand (grid-uniform): nine run-time "is this bit in use"
    // predicates kept the predicate file full and made the compiler re-create every ballot predicate from a register
#ifdef GSC_SORT_RANK9_ONLY
    radix_rank<9>(key, shift, mask, s_hist[warp], lane, rank);
#else
    if (digit_bits == 9) radix_rank<9>(key, shift, mask, s_hist[warp], lane, rank);
    else if (digit_bits == 8) radix_rank<8>(key, shift, mask, s_hist[warp], lane, rank);
    else radix_rank<7>(key, shift, mask, s_hist[warp], lane, rank);  // (<= 7 bits: the extra ballots see zeros)
#endif
#else
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j) {
        const uint32_t d = (key[j] >> shift) & mask;
        uint32_t peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < kRadixBits; ++b) {
            if (b < digit_bits) {  // warp-uniform
                const bool bit = (d >> b) & 1u;
                const uint32_t bal = __ballot_sync(0xffffffffu, bit);
                peers &= bit ? bal : ~bal;
            }
        }
        const uint32_t before = s_hist[warp][d];
        const uint32_t r = __popc(peers & lt_mask);
        __syncwarp();
        if (r == 0) s_hist[warp][d] = before + __popc(peers);
        __syncwarp();
        rank[j] = (uint16_t)(before + r);
    }
#endif
    __syncthreads();

    // ---- per digit pair (thread t owns digits 2t, 2t+1): warp counts -> offsets within the digit, tile counts ----
    uint32_t cnt0 = 0, cnt1 = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        uint2 *pair = reinterpret_cast<uint2 *>(s_hist[w]) + tid;
        const uint2 v = *pair;
        *pair = make_uint2(cnt0, cnt1);
        cnt0 += v.x;
        cnt1 += v.y;
    }
    // exclusive scans over the 512 digits in digit order: scan the pair sums, the odd digit adds its even sibling
    const uint32_t digit_base = block_excl_scan256(dtotal.x + dtotal.y, s_warp, nullptr);
    const uint32_t bin_start = block_excl_scan256(cnt0 + cnt1, s_warp, nullptr);
    s_bin[2 * tid] = bin_start;
    s_bin[2 * tid + 1] = bin_start + cnt0;
    s_goff[2 * tid] = digit_base + tile_excl0 - bin_start;
    s_goff[2 * tid + 1] = digit_base + dtotal.x + tile_excl1 - (bin_start + cnt0);
    __syncthreads();

    // ---- stage the tile digit-sorted in shared memory ----
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j) {
        const uint32_t d = (key[j] >> shift) & mask;
        const uint32_t pos = (uint32_t)rank[j] + s_hist[warp][d] + s_bin[d];
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
    __syncthreads();  // every warp is done reading the counters that s_vals overwrites
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j) s_vals[rank[j]] = key[j];
    __syncthreads();

    // ---- write digit runs contiguously ----
    const uint32_t remaining = n - tile_base;
    const uint32_t valid = remaining < (uint32_t)kSortTile ? remaining : (uint32_t)kSortTile;
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j) {
        const uint32_t i = tid + j * kSortThreads;
        if (i < valid) {
            const uint32_t k = s_keys[i];
            const uint32_t dst = s_goff[(k >> shift) & mask] + i;
#ifdef GSC_CHECKED
            if (dst >= n) { asm volatile("trap;"); }  // (the sort has no status word: a checked build traps)
#endif
            keys_out[dst] = k;
            vals_out[dst] = s_vals[i];
        }
    }
}

}  // namespace gsc

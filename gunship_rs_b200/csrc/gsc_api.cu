// gsc_api.cu -- the C ABI of include/gunship_collision.h: context, buffers, stage orchestration.
//
// A step is a fixed sequence of kernel launches on one stream with NO host round trip in the middle:
// data-dependent quantities (cell size, grid extent, number of radix passes, candidate counts) stay in
// a device-side control block (FrameBlock) that later kernels read, and the host reads it back once
// at the end together with the pair count.
#include "../../include/gunship_collision.h"

#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "gsc_hierarchy.cuh"
#include "gsc_kernels.cuh"

using namespace gsc;

namespace {

thread_local std::string g_create_error;

constexpr int kGridStrideBlocks = 148 * 8;  // 148 SMs x resident CTAs for the grid-stride kernels

// device allocations freed when the scope ends
struct Scratch {
    void *p[4] = { nullptr, nullptr, nullptr, nullptr };
    ~Scratch() { for (void *q : p) if (q) cudaFree(q); }
};

}  // namespace

static_assert(kMaxPeers == GSC_MAX_PEERS, "kMaxPeers");
constexpr size_t kMailOffset = (sizeof(FrameBlock) + 255) & ~size_t(255);  // the mailbox lives behind the control block

struct gsc_ctx {
    gsc_config cfg{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    std::string err;

    uint32_t cap = 0;       // max colliders (local + ghost)
    uint32_t n_local = 0;   // uploaded colliders
    uint32_t n_ghost = 0;   // ghosts appended this step
    bool have_colliders = false, have_transforms = false, have_gidx = false;

    // static collider data + transforms (SoA, as uploaded)
    uint32_t *d_entity = nullptr, *d_gidx = nullptr, *d_box_slot = nullptr;
    uint32_t n_box = 0;
    bool packed = false;  // d_quat / d_scale hold boxes only (gsc_update_transforms_packed)
    // d_quat / d_scale match the CURRENT collider set and layout: cleared by upload / append / remove (the arrays then
    // hold another set's values, packed or in pre-swap_remove order), so "NULL keeps the previous values" is refused
    // until rotation and scale were sent again (a set without boxes never needs them)
    bool rot_scale_valid = false;
    bool sorted_valid = false;  // d_vals of the last step is the cell order of the CURRENT local collider set (gsc_reorder_colliders)
    // pinned transform staging (gsc_map_transform_staging), allocated on first use
    float *h_stage[2][3] = { { nullptr, nullptr, nullptr }, { nullptr, nullptr, nullptr } };
    // GSC_FLAG_GRAPH: stages 1-5 captured once per launch geometry
    cudaGraphExec_t graph_exec = nullptr;
    uint32_t graph_n = 0, graph_local = 0;
    bool graph_window = false, graph_slab = false;
    int32_t graph_win[2] = { 0, 0 };
    // host -> device transform copies are cut into chunks; bvh_update starts on a chunk as soon as it has landed
    static constexpr int kChunks = 8;
    uint32_t chunk_begin[kChunks + 1] = {}, chunk_box[kChunks + 1] = {};
    cudaEvent_t ev_chunk[kChunks] = {};
    bool chunked = false;
    uint8_t *d_kind = nullptr;
    float *d_offset = nullptr, *d_size = nullptr;
    float *d_pos = nullptr, *d_quat = nullptr, *d_scale = nullptr;          // owned
    const float *b_pos = nullptr, *b_quat = nullptr, *b_scale = nullptr;    // in use (owned or bound)

    // per-step buffers
    float4 *d_rec = nullptr, *d_obb = nullptr;
    uint2 *d_qrec_sorted = nullptr;
    uint32_t *d_w3_sorted = nullptr;
    uint32_t *d_keys[2] = { nullptr, nullptr }, *d_vals[2] = { nullptr, nullptr };
    uint32_t *d_cell_start = nullptr, *d_wide = nullptr;
    uint32_t *d_overflow = nullptr;  // tiles of the pairs stage that did not fit its staging buffers
    bool pairs_tiled = false;        // GSC_PAIRS_TILED=1: the tile-staged, work-flattened pairs kernel (measured slower, see DESIGN)
    uint4 *d_gaps = nullptr;
    bool have_window = false;
    int32_t win_lo = 0, win_hi = 0;
    uint2 *d_pairs = nullptr, *d_cand_ss = nullptr, *d_cand_so = nullptr, *d_cand_oo = nullptr;
    uint2 *d_pairs_sorted = nullptr;
    float4 *d_contacts = nullptr, *d_contacts_sorted = nullptr;  // GSC_FLAG_CONTACTS
    // adjacency (CSR) of the last step's pair set, built on demand
    uint32_t *d_adj_entity = nullptr, *d_adj_offset = nullptr, *d_adj_other = nullptr;
    uint64_t adj_cap = 0, adj_entities = 0, adj_entries = 0;
    bool adj_valid = false;
    uint32_t *d_perm = nullptr;  // permutation that sorts the pairs of the last step
    bool perm_valid = false;
    FrameBlock *d_fb = nullptr;
    FrameBlock *h_fb = nullptr;  // pinned
    uint32_t *h_enc = nullptr;   // pinned staging of gsc_set_global_bounds
    uint32_t *d_sort_scratch = nullptr;
    // halo staging
    float4 *d_halo_rec = nullptr, *d_halo_obb = nullptr;
    uint32_t *d_halo_count = nullptr;
    uint32_t halo_cap = 0;

    // peer contexts opened over CUDA IPC (multi-GPU halo push)
    struct Peer {
        float4 *rec = nullptr, *obb = nullptr;
        FrameBlock *fb = nullptr;
        uint32_t cap = 0;
        bool self = false;   // this context itself
        bool local = false;  // another context of this process (gsc_open_peer_local): nothing to unmap
        SlabMail *mail() const { return fb ? reinterpret_cast<SlabMail *>(reinterpret_cast<char *>(fb) + kMailOffset) : nullptr; }
    } peers[GSC_MAX_PEERS];

    // device-driven slab exchange (gsc_slab_*): mailbox behind the FrameBlock in the same allocation
    SlabMail *d_mail = nullptr;
    uint32_t slab_rank = 0, slab_world = 0, slab_seq = 0;
    bool slab_ready = false, slab_step = false;  // slab_step: the step in flight / last collected ran the device-driven exchange
    unsigned long long peer_timeout_ns = 20000000000ull;  // GSC_PEER_TIMEOUT_MS
    cudaEvent_t ev_slab[2] = { nullptr, nullptr };  // behind the publish / behind the done flags (in-process groups order on them)

    uint64_t max_pairs = 0, max_cands = 0, max_cells = 0;
    uint64_t last_pairs = 0;
    bool bounds_done = false, global_bounds_set = false;
    std::vector<uint8_t> h_kind;  // host mirror of d_kind: box slots / chunk cuts are recomputed from it on append / remove

    // transform hierarchy (gsc_hierarchy_*): local and derived values per node, device resident
    uint32_t hier_nodes = 0, hier_cap = 0, hier_colliders = 0;
    std::vector<uint32_t> hier_row_begin;
    uint32_t *d_hier_parent = nullptr, *d_hier_node = nullptr;
    float *d_lpos = nullptr, *d_lquat = nullptr, *d_lscale = nullptr;
    float *d_dpos = nullptr, *d_dquat = nullptr, *d_dscale = nullptr, *d_dmat = nullptr;
    bool hier_local[3] = { false, false, false }, hier_derived = false;
    bool in_flight = false;  // gsc_step_begin enqueued a step that gsc_step_end has not collected yet

    cudaEvent_t ev[GSC_NUM_STAGES + 1] = {};
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_fence = nullptr;
};

namespace {

int fail(gsc_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_error = buf;
    return code;
}

#define CU(c, expr)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail((c), GSC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define NOT_IN_FLIGHT(c, what)                                                                                \
    do {                                                                                                      \
        if ((c)->in_flight) return fail((c), GSC_ERR_INVALID, what ": a step is in flight (call gsc_step_end first)"); \
    } while (0)

template <class T>
cudaError_t dmalloc(T **p, size_t count)
{
    return cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T));
}

__global__ void k_fill_default_transforms(uint32_t n, float *quat4, float *scale3)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    quat4[4 * (size_t)i] = 0.f; quat4[4 * (size_t)i + 1] = 0.f; quat4[4 * (size_t)i + 2] = 0.f; quat4[4 * (size_t)i + 3] = 1.f;
    scale3[3 * (size_t)i] = 1.f; scale3[3 * (size_t)i + 1] = 1.f; scale3[3 * (size_t)i + 2] = 1.f;
}

// AABB::from_collider as the records encode it, expanded to (min xyz, max xyz) for parity read-back
__global__ void k_decode_aabb(uint32_t n, const float4 *rec, float *out6)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Volume v = volume_decode(rec[2 * (size_t)i], rec[2 * (size_t)i + 1]);
    for (int a = 0; a < 3; ++a) { out6[6 * (size_t)i + a] = v.mn[a]; out6[6 * (size_t)i + 3 + a] = v.mx[a]; }
}

inline uint32_t tiles_for(uint32_t n) { return (n + kSortTile - 1) / kSortTile; }

// Runs the radix passes on `n` (key, value) pairs.  Passes beyond ctl->num_passes exit on the device.
// scratch = [kRadix * tiles] tile histogram + [kRadix] digit totals
// `have_pass0_hist`: the producer of the keys (k_cell_keys) already filled the tile histogram of pass 0
// `n_dev`: optional device-side element count (<= n); the launches are then sized for n and clamp to *n_dev
void launch_sort(cudaStream_t s, uint32_t n, uint32_t *keys[2], uint32_t *vals[2], bool first_vals_identity,
                 int max_passes, const SortControl *ctl, uint32_t *scratch, bool have_pass0_hist = false,
                 const uint32_t *n_dev = nullptr)
{
    const uint32_t tiles = tiles_for(n);
    if (tiles == 0) return;
    uint32_t *tile_hist = scratch, *digit_total = scratch + (size_t)kRadix * tiles;
    for (int pass = 0; pass < max_passes; ++pass) {
        const int in = pass & 1, out = in ^ 1;
        const uint32_t *vin = (pass == 0 && first_vals_identity) ? nullptr : vals[in];
        if (pass > 0 || !have_pass0_hist) k_radix_upsweep<<<tiles, kSortThreads, 0, s>>>(keys[in], n, pass, ctl, tile_hist, tiles, n_dev);
        k_radix_scan<<<kRadix, 256, 0, s>>>(tile_hist, tiles, pass, ctl, digit_total, n_dev);
        k_radix_downsweep<<<tiles, kSortThreads, 0, s>>>(keys[in], vin, keys[out], vals[out], n, pass, ctl, tile_hist,
                                                        digit_total, tiles, n_dev);
    }
}

inline size_t sort_scratch_words(uint32_t n) { return (size_t)kRadix * tiles_for(n) + kRadix; }

int check_ctx(gsc_ctx *c)
{
    if (!c) return GSC_ERR_INVALID;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(c, GSC_ERR_CUDA, "cudaSetDevice(%d): %s", c->device, cudaGetErrorString(e));
    return GSC_OK;
}

int status_to_error(gsc_ctx *c, const FrameBlock &fb)
{
    const uint32_t st = fb.status;
    if (st & kStMesh) return fail(c, GSC_ERR_KIND, "Collider::Mesh is not supported (unimplemented!() in the reference, mod.rs:331)");
    if (st & kStNonFinite) return fail(c, GSC_ERR_RANGE, "a collider produced a NaN/Inf bounding box");
    if (st & kStCellSize) return fail(c, GSC_ERR_RANGE, "cell size (longest AABB axis) is not positive: %g", (double)fb.grid.cell_size);
    if (st & kStRange) return fail(c, GSC_ERR_RANGE, "a grid coordinate leaves the i16 range of GridCoord (grid_collision.rs:535)");
    if (st & kStWindow) return fail(c, GSC_ERR_INVALID, "gsc_set_x_window does not cover the home cell of a local collider");
    if (st & kStCheck) return fail(c, GSC_ERR_CUDA, "internal bounds check failed (GSC_CHECKED build)");
    if (st & kStWideBorder)
        return fail(c, GSC_ERR_RANGE, "a volume whose AABB spans 3 cells (rounding; gsc_step_out.n_wide) lies on a slab border: "
                                      "multi-GPU steps match such volumes inside their own slab only");
    if (st & kStPeerTimeout)
        return fail(c, GSC_ERR_NCCL, "slab exchange: a peer GPU did not publish its bounds / finish its halo push within %llu ms "
                                     "(a rank failed, or the ranks are not stepping together)", c->peer_timeout_ns / 1000000ull);
    if (st & kStGhostCap)
        return fail(c, GSC_ERR_CAPACITY, "slab exchange: %u local + %u pushed ghost colliders > max_colliders %u", c->n_local,
                    fb.n_ghost_in, c->cap);
    if (st & kStCells) return fail(c, GSC_ERR_CAPACITY, "dense cell table too small: raise gsc_config.max_cells (have %llu)", (unsigned long long)c->max_cells);
    return GSC_OK;
}

// box_slot[i] = number of boxes before collider i (where a box finds its rotation / scale in the packed arrays of
// gsc_update_transforms_packed), and the chunk cuts of the host-to-device transform copies; from the host mirror of kind
int rebuild_box_slots(gsc_ctx *c)
{
    const uint32_t n = c->n_local;
    std::vector<uint32_t> slot(n);
    uint32_t n_box = 0;
    for (uint32_t i = 0; i < n; ++i) {
        slot[i] = n_box;
        n_box += c->h_kind[i] == GSC_BOX;
    }
    c->n_box = n_box;
    for (int k = 0; k <= gsc_ctx::kChunks; ++k) {
        const uint32_t b = k == gsc_ctx::kChunks ? n : (uint32_t)(((uint64_t)n * k / gsc_ctx::kChunks) & ~4095ull);
        c->chunk_begin[k] = b;
        c->chunk_box[k] = b < n ? slot[b] : n_box;
    }
    if (n) {
        CU(c, cudaMemcpyAsync(c->d_box_slot, slot.data(), (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));  // `slot` dies with this scope
    }
    return GSC_OK;
}

}  // namespace

extern "C" {

int gsc_abi_version(void) { return GSC_ABI_VERSION; }

const char *gsc_last_error(const gsc_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int gsc_create(const gsc_config *cfg, gsc_ctx **out)
{
    if (!cfg || !out) return fail(nullptr, GSC_ERR_INVALID, "gsc_create: null argument");
    *out = nullptr;
    if (cfg->max_colliders >= (1u << 28)) return fail(nullptr, GSC_ERR_INVALID, "max_colliders must be < 2^28");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, GSC_ERR_CUDA, "no CUDA device: %s -- this library has no CPU fallback", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, GSC_ERR_INVALID, "device %d out of range", cfg->device);
    gsc_ctx *c = new gsc_ctx();
    c->cfg = *cfg;
    c->device = cfg->device;
    c->cap = std::max<uint32_t>(cfg->max_colliders, 1);
    c->max_pairs = cfg->max_pairs ? cfg->max_pairs : 8ull * c->cap + 65536;
    c->max_cands = cfg->max_candidates ? cfg->max_candidates : 8ull * c->cap + 65536;
    c->max_cells = cfg->max_cells ? cfg->max_cells : 16ull * c->cap + (1ull << 20);
    c->max_cells = std::min<uint64_t>(c->max_cells, (1ull << 31) - 2);
    const size_t M = c->cap;
#define CR(expr)                                                                                     \
    do {                                                                                             \
        cudaError_t e2_ = (expr);                                                                    \
        if (e2_ != cudaSuccess) {                                                                    \
            fail(nullptr, GSC_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e2_));            \
            gsc_destroy(c);                                                                          \
            return GSC_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)
    CR(cudaSetDevice(c->device));
    CR(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CR(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CR(dmalloc(&c->d_entity, M));
    CR(dmalloc(&c->d_gidx, M));
    CR(dmalloc(&c->d_box_slot, M));
    CR(dmalloc(&c->d_kind, M));
    CR(dmalloc(&c->d_offset, 3 * M));
    CR(dmalloc(&c->d_size, 3 * M));
    CR(dmalloc(&c->d_pos, 3 * M));
    CR(dmalloc(&c->d_quat, 4 * M));
    CR(dmalloc(&c->d_scale, 3 * M));
    CR(dmalloc(&c->d_rec, 2 * M));
    CR(dmalloc(&c->d_qrec_sorted, M + kPairTrip));  // k_pairs reads whole trips
    CR(dmalloc(&c->d_w3_sorted, M + kPairTrip));
    CR(dmalloc(&c->d_obb, 4 * M));
    for (int i = 0; i < 2; ++i) {
        CR(dmalloc(&c->d_keys[i], M));
        CR(dmalloc(&c->d_vals[i], M));
    }
    CR(dmalloc(&c->d_cell_start, c->max_cells + 2));
    CR(dmalloc(&c->d_wide, M));
    CR(dmalloc(&c->d_overflow, (M + kTileA - 1) / kTileA + 1));
    CR(cudaFuncSetAttribute(k_pairs_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileShared)));
    {
        const char *e = getenv("GSC_PAIRS_TILED");
        c->pairs_tiled = e && e[0] == '1';
    }
    CR(dmalloc(&c->d_gaps, (size_t)kGapCap));
    CR(dmalloc(&c->d_pairs, c->max_pairs));
    if (cfg->flags & GSC_FLAG_CONTACTS) CR(dmalloc(&c->d_contacts, c->max_pairs));
    CR(dmalloc(&c->d_cand_ss, c->max_cands));
    CR(dmalloc(&c->d_cand_so, c->max_cands));
    CR(dmalloc(&c->d_cand_oo, c->max_cands));
    {   // FrameBlock (zeroed every step) + SlabMail (never zeroed after this) in ONE allocation: one IPC handle maps both
        void *blk = nullptr;
        CR(cudaMalloc(&blk, kMailOffset + sizeof(SlabMail)));
        CR(cudaMemset(blk, 0, kMailOffset + sizeof(SlabMail)));
        c->d_fb = reinterpret_cast<FrameBlock *>(blk);
        c->d_mail = reinterpret_cast<SlabMail *>(reinterpret_cast<char *>(blk) + kMailOffset);
    }
    for (auto &ev : c->ev_slab) CR(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    if (const char *e = getenv("GSC_PEER_TIMEOUT_MS")) c->peer_timeout_ns = strtoull(e, nullptr, 10) * 1000000ull;
    CR(dmalloc(&c->d_sort_scratch, sort_scratch_words(c->cap)));
    CR(dmalloc(&c->d_halo_count, 1));
    CR(cudaHostAlloc((void **)&c->h_fb, sizeof(FrameBlock), cudaHostAllocDefault));
    CR(cudaHostAlloc((void **)&c->h_enc, 8 * sizeof(uint32_t), cudaHostAllocDefault));
    for (auto &ev : c->ev) CR(cudaEventCreate(&ev));
    for (auto &ev : c->ev_chunk) CR(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CR(cudaEventCreate(&c->ev_begin));
    CR(cudaEventCreate(&c->ev_end));
    CR(cudaEventCreateWithFlags(&c->ev_fence, cudaEventDisableTiming));
    // identity quaternion / unit scale so a sphere-only caller may pass NULL for them
    k_fill_default_transforms<<<(unsigned)((M + 255) / 256), 256, 0, c->stream>>>((uint32_t)M, c->d_quat, c->d_scale);
    CR(cudaGetLastError());
    CR(cudaStreamSynchronize(c->stream));
#undef CR
    c->b_pos = c->d_pos;
    c->b_quat = c->d_quat;
    c->b_scale = c->d_scale;
    *out = c;
    return GSC_OK;
}

void gsc_destroy(gsc_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto &ev : c->ev_slab)
        if (ev) cudaEventDestroy(ev);
    for (auto &p : c->peers)
        if (p.rec && !p.self && !p.local) {
            cudaIpcCloseMemHandle(p.rec);
            cudaIpcCloseMemHandle(p.obb);
            cudaIpcCloseMemHandle(p.fb);
        }
    void *bufs[] = { c->d_entity, c->d_gidx, c->d_box_slot, c->d_kind, c->d_offset, c->d_size, c->d_pos, c->d_quat, c->d_scale, c->d_rec,
                     c->d_qrec_sorted, c->d_w3_sorted, c->d_obb, c->d_keys[0], c->d_keys[1], c->d_vals[0], c->d_vals[1], c->d_cell_start,
                     c->d_wide, c->d_overflow, c->d_gaps, c->d_pairs, c->d_cand_ss, c->d_cand_so, c->d_cand_oo, c->d_pairs_sorted, c->d_contacts, c->d_contacts_sorted, c->d_perm, c->d_adj_entity, c->d_adj_offset, c->d_adj_other, c->d_fb, c->d_sort_scratch,
                     c->d_halo_rec, c->d_halo_obb, c->d_halo_count, c->d_hier_parent, c->d_hier_node, c->d_lpos, c->d_lquat,
                     c->d_lscale, c->d_dpos, c->d_dquat, c->d_dscale, c->d_dmat };
    for (void *b : bufs)
        if (b) cudaFree(b);
    for (auto &bank : c->h_stage)
        for (float *&b : bank)
            if (b) cudaFreeHost(b);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->h_fb) cudaFreeHost(c->h_fb);
    if (c->h_enc) cudaFreeHost(c->h_enc);
    for (auto &ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto &ev : c->ev_chunk)
        if (ev) cudaEventDestroy(ev);
    if (c->ev_begin) cudaEventDestroy(c->ev_begin);
    if (c->ev_end) cudaEventDestroy(c->ev_end);
    if (c->ev_fence) cudaEventDestroy(c->ev_fence);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

int gsc_upload_colliders(gsc_ctx *c, uint32_t n, const uint32_t *entity, const uint8_t *kind, const float *offset3,
                         const float *size3, const uint32_t *global_index)
{
    if (int rc = check_ctx(c)) return rc;
    NOT_IN_FLIGHT(c, "gsc_upload_colliders");
    if (n > c->cap) return fail(c, GSC_ERR_CAPACITY, "gsc_upload_colliders: %u colliders > max_colliders %u", n, c->cap);
    if (n && (!entity || !kind || !offset3 || !size3)) return fail(c, GSC_ERR_INVALID, "gsc_upload_colliders: null array");
    for (uint32_t i = 0; i < n; ++i)
        if (kind[i] > GSC_BOX)
            return fail(c, GSC_ERR_KIND, "collider %u is a Mesh: unimplemented!() in the reference (mod.rs:331)", i);
    c->h_kind.assign(kind, kind + n);
    c->n_local = n;
    if (int rc = rebuild_box_slots(c)) return rc;
    if (n) {
        CU(c, cudaMemcpyAsync(c->d_entity, entity, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(c->d_kind, kind, (size_t)n, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(c->d_offset, offset3, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(c->d_size, size3, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
        if (global_index) CU(c, cudaMemcpyAsync(c->d_gidx, global_index, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    }
    CU(c, cudaStreamSynchronize(c->stream));
    c->have_gidx = global_index != nullptr;
    c->have_colliders = true;
    c->have_transforms = false;
    c->rot_scale_valid = false;
    c->sorted_valid = false;
    c->packed = false;
    c->chunked = false;
    c->hier_colliders = 0;
    return GSC_OK;
}

int gsc_set_flags(gsc_ctx *c, uint32_t flags)
{
    if (int rc = check_ctx(c)) return rc;
    NOT_IN_FLIGHT(c, "gsc_set_flags");
    const uint32_t fixed = GSC_FLAG_CONTACTS;  // decides what was allocated
    if ((flags & fixed) != (c->cfg.flags & fixed)) return fail(c, GSC_ERR_INVALID, "gsc_set_flags: GSC_FLAG_CONTACTS is fixed at gsc_create");
    c->cfg.flags = flags;
    return GSC_OK;
}

int gsc_collider_count(gsc_ctx *c, uint32_t *n_out)
{
    if (int rc = check_ctx(c)) return rc;
    if (!n_out) return fail(c, GSC_ERR_INVALID, "gsc_collider_count: null output");
    *n_out = c->n_local;
    return GSC_OK;
}

// ColliderManager::assign after the first upload: the dense arrays grow at the end (struct_component_manager.rs:83-96,
// bounding_volume.rs:394-407); the colliders already on the device are not touched.
int gsc_append_colliders(gsc_ctx *c, uint32_t n_new, const uint32_t *entity, const uint8_t *kind, const float *offset3,
                         const float *size3)
{
    if (int rc = check_ctx(c)) return rc;
    NOT_IN_FLIGHT(c, "gsc_append_colliders");
    if (!c->have_colliders) return fail(c, GSC_ERR_INVALID, "gsc_append_colliders before gsc_upload_colliders");
    if (c->have_gidx) return fail(c, GSC_ERR_INVALID, "gsc_append_colliders: contexts with explicit global indices (slabs) re-upload instead");
    if (n_new && (!entity || !kind || !offset3 || !size3)) return fail(c, GSC_ERR_INVALID, "gsc_append_colliders: null array");
    const uint32_t n0 = c->n_local;
    if ((uint64_t)n0 + n_new > c->cap)
        return fail(c, GSC_ERR_CAPACITY, "gsc_append_colliders: %u + %u colliders > max_colliders %u", n0, n_new, c->cap);
    for (uint32_t i = 0; i < n_new; ++i)
        if (kind[i] > GSC_BOX)
            return fail(c, GSC_ERR_KIND, "collider %u is a Mesh: unimplemented!() in the reference (mod.rs:331)", n0 + i);
    if (n_new) {
        CU(c, cudaMemcpyAsync(c->d_entity + n0, entity, (size_t)n_new * 4, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(c->d_kind + n0, kind, (size_t)n_new, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(c->d_offset + 3 * (size_t)n0, offset3, (size_t)n_new * 12, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(c->d_size + 3 * (size_t)n0, size3, (size_t)n_new * 12, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    c->h_kind.insert(c->h_kind.end(), kind, kind + n_new);
    c->n_local = n0 + n_new;
    if (int rc = rebuild_box_slots(c)) return rc;
    c->have_transforms = false;
    c->rot_scale_valid = false;
    c->sorted_valid = false;
    c->packed = false;
    c->chunked = false;
    c->hier_colliders = 0;
    return GSC_OK;
}

// BoundingVolumeManager::destroy_immediate (bounding_volume.rs:82-104) for a sequence of dense indices: each one is a
// swap_remove (the last element moves into the hole).  The host works out the NET moves of the whole sequence --
// position -> original position of the element that ends up there -- and one kernel applies them to the static
// arrays on the device.  Every source lies at or behind the new length and every destination before it, so the
// gather has no read-after-write hazard.
int gsc_remove_colliders(gsc_ctx *c, uint32_t n_remove, const uint32_t *index)
{
    if (int rc = check_ctx(c)) return rc;
    NOT_IN_FLIGHT(c, "gsc_remove_colliders");
    if (!c->have_colliders) return fail(c, GSC_ERR_INVALID, "gsc_remove_colliders before gsc_upload_colliders");
    if (c->have_gidx) return fail(c, GSC_ERR_INVALID, "gsc_remove_colliders: contexts with explicit global indices (slabs) re-upload instead");
    if (n_remove && !index) return fail(c, GSC_ERR_INVALID, "gsc_remove_colliders: null index array");
    if (n_remove > c->n_local) return fail(c, GSC_ERR_INVALID, "gsc_remove_colliders: %u removals from %u colliders", n_remove, c->n_local);
    for (uint32_t k = 0; k < n_remove; ++k)  // validate before anything is touched
        if (index[k] >= c->n_local - k)
            return fail(c, GSC_ERR_INVALID, "gsc_remove_colliders: index[%u] = %u but only %u colliders are left", k, index[k], c->n_local - k);
    uint32_t len = c->n_local;
    std::unordered_map<uint32_t, uint32_t> origin;  // position -> original position of its current element (absent: itself)
    origin.reserve(2 * (size_t)n_remove + 8);
    auto at = [&](uint32_t pos) { auto it = origin.find(pos); return it == origin.end() ? pos : it->second; };
    for (uint32_t k = 0; k < n_remove; ++k) {
        const uint32_t i = index[k];
        const uint32_t last = len - 1;
        if (i != last) origin[i] = at(last);
        origin.erase(last);
        c->h_kind[i] = c->h_kind[last];  // the same swap_remove on the host mirror
        c->h_kind.pop_back();
        len = last;
    }
    std::vector<uint2> moves;
    moves.reserve(origin.size());
    for (const auto &kv : origin)
        if (kv.first < len && kv.second != kv.first) moves.push_back(make_uint2(kv.first, kv.second));
    if (!moves.empty()) {
        Scratch tmp;
        CU(c, cudaMalloc(&tmp.p[0], moves.size() * sizeof(uint2)));
        CU(c, cudaMemcpyAsync(tmp.p[0], moves.data(), moves.size() * sizeof(uint2), cudaMemcpyHostToDevice, c->stream));
        k_apply_moves<<<(unsigned)((moves.size() + 255) / 256), 256, 0, c->stream>>>((uint32_t)moves.size(), (const uint2 *)tmp.p[0],
                                                                                     c->d_kind, c->d_offset, c->d_size, c->d_entity);
        CU(c, cudaGetLastError());
        CU(c, cudaStreamSynchronize(c->stream));
    }
    c->n_local = len;
    if (int rc = rebuild_box_slots(c)) return rc;
    c->have_transforms = false;
    c->rot_scale_valid = false;
    c->sorted_valid = false;
    c->packed = false;
    c->chunked = false;
    c->hier_colliders = 0;
    return GSC_OK;
}

// Host -> device copy of one frame's transforms, cut into chunks on the copy stream with an event per chunk, so
// that the next step's bvh_update overlaps the tail of the transfer (PCIe is the slowest link of the e2e step).
// `box_packed`: quat / scale arrays are indexed by box slot instead of collider index.
static int copy_transforms(gsc_ctx *c, uint32_t n, const float *pos3, const float *quat4, const float *scale3, bool box_packed)
{
    cudaStream_t cs = c->copy_stream;
    // The copy must not overtake the one reader of the transform arrays: bvh_update of the last enqueued step
    // (ev_fence is recorded right behind it).  The REST of a step in flight (gsc_step_begin) keeps running under
    // the transfer -- that overlap is the point of the begin/end form.
    CU(c, cudaStreamWaitEvent(cs, c->ev_fence, 0));
    for (int k = 0; k < gsc_ctx::kChunks; ++k) {
        const size_t b = c->chunk_begin[k], e = c->chunk_begin[k + 1];
        const size_t qb = box_packed ? c->chunk_box[k] : b, qe = box_packed ? c->chunk_box[k + 1] : e;
        if (e > b) CU(c, cudaMemcpyAsync(c->d_pos + 3 * b, pos3 + 3 * b, (e - b) * 12, cudaMemcpyHostToDevice, cs));
        if (quat4 && qe > qb) CU(c, cudaMemcpyAsync(c->d_quat + 4 * qb, quat4 + 4 * qb, (qe - qb) * 16, cudaMemcpyHostToDevice, cs));
        if (scale3 && qe > qb) CU(c, cudaMemcpyAsync(c->d_scale + 3 * qb, scale3 + 3 * qb, (qe - qb) * 12, cudaMemcpyHostToDevice, cs));
        CU(c, cudaEventRecord(c->ev_chunk[k], cs));
    }
    c->chunked = true;
    return GSC_OK;
}

int gsc_update_transforms(gsc_ctx *c, uint32_t n, const float *pos3, const float *quat4, const float *scale3)
{
    if (int rc = check_ctx(c)) return rc;
    if (!c->have_colliders || n != c->n_local)
        return fail(c, GSC_ERR_INVALID, "gsc_update_transforms: n=%u does not match the %u uploaded colliders", n, c->n_local);
    if (n && !pos3) return fail(c, GSC_ERR_INVALID, "gsc_update_transforms: null positions");
    if ((c->packed || !c->rot_scale_valid) && (!quat4 || !scale3) && c->n_box)
        return fail(c, GSC_ERR_INVALID, c->packed ? "gsc_update_transforms: after a packed update, pass rotation and scale again"
                                                  : "gsc_update_transforms: the collider set changed (upload / append / remove): pass rotation and scale again");
    if (int rc = copy_transforms(c, n, pos3, quat4, scale3, false)) return rc;
    c->rot_scale_valid = true;
    c->packed = false;
    c->b_pos = c->d_pos;
    c->b_quat = c->d_quat;
    c->b_scale = c->d_scale;
    c->have_transforms = true;
    return GSC_OK;
}

int gsc_update_transforms_packed(gsc_ctx *c, uint32_t n, const float *pos3, uint32_t n_box, const float *box_quat4,
                                 const float *box_scale3)
{
    if (int rc = check_ctx(c)) return rc;
    if (!c->have_colliders || n != c->n_local || n_box != c->n_box)
        return fail(c, GSC_ERR_INVALID, "gsc_update_transforms_packed: n=%u / n_box=%u do not match the uploaded %u colliders / %u boxes",
                    n, n_box, c->n_local, c->n_box);
    if (n && !pos3) return fail(c, GSC_ERR_INVALID, "gsc_update_transforms_packed: null positions");
    if (n_box && (!c->packed || !c->rot_scale_valid) && (!box_quat4 || !box_scale3))
        return fail(c, GSC_ERR_INVALID, "gsc_update_transforms_packed: the first packed update (also after upload / append / remove) needs rotation and scale");
    if (int rc = copy_transforms(c, n, pos3, box_quat4, box_scale3, true)) return rc;
    c->rot_scale_valid = true;
    c->packed = true;
    c->b_pos = c->d_pos;
    c->b_quat = c->d_quat;
    c->b_scale = c->d_scale;
    c->have_transforms = true;
    return GSC_OK;
}

// ---- pinned transform staging -----------------------------------------------------------------------------------
int gsc_map_transform_staging(gsc_ctx *c, uint32_t bank, float **pos3, float **box_quat4, float **box_scale3)
{
    if (int rc = check_ctx(c)) return rc;
    if (bank > 1 || !pos3 || !box_quat4 || !box_scale3) return fail(c, GSC_ERR_INVALID, "gsc_map_transform_staging: bad argument");
    const size_t per[3] = { 12, 16, 12 };
    for (int k = 0; k < 3; ++k)
        if (!c->h_stage[bank][k]) CU(c, cudaHostAlloc((void **)&c->h_stage[bank][k], std::max<size_t>(c->cap, 1) * per[k], cudaHostAllocDefault));
    *pos3 = c->h_stage[bank][0];
    *box_quat4 = c->h_stage[bank][1];
    *box_scale3 = c->h_stage[bank][2];
    return GSC_OK;
}

int gsc_commit_transform_staging(gsc_ctx *c, uint32_t bank, uint32_t n, uint32_t n_box, int with_rotation, int with_scale)
{
    if (int rc = check_ctx(c)) return rc;
    if (bank > 1 || !c->h_stage[bank][0]) return fail(c, GSC_ERR_INVALID, "gsc_commit_transform_staging: bank %u was not mapped", bank);
    return gsc_update_transforms_packed(c, n, c->h_stage[bank][0], n_box, with_rotation ? c->h_stage[bank][1] : nullptr,
                                        with_scale ? c->h_stage[bank][2] : nullptr);
}

int gsc_box_slots(gsc_ctx *c, uint32_t *box_slot, uint32_t cap, uint32_t *n_box)
{
    if (int rc = check_ctx(c)) return rc;
    if (!c->have_colliders) return fail(c, GSC_ERR_INVALID, "gsc_box_slots before gsc_upload_colliders");
    if (n_box) *n_box = c->n_box;
    if (!box_slot) return GSC_OK;
    if (cap < c->n_local) return fail(c, GSC_ERR_CAPACITY, "gsc_box_slots: need room for %u entries", c->n_local);
    uint32_t k = 0;
    for (uint32_t i = 0; i < c->n_local; ++i) {
        box_slot[i] = k;
        k += c->h_kind[i] == GSC_BOX;
    }
    return GSC_OK;
}

int gsc_bind_device_transforms(gsc_ctx *c, uint32_t n, const float *d_pos3, const float *d_quat4, const float *d_scale3)
{
    if (int rc = check_ctx(c)) return rc;
    if (!c->have_colliders || n != c->n_local)
        return fail(c, GSC_ERR_INVALID, "gsc_bind_device_transforms: n=%u does not match the %u uploaded colliders", n, c->n_local);
    if (n && !d_pos3) return fail(c, GSC_ERR_INVALID, "gsc_bind_device_transforms: null positions");
    if (d_quat4 && (reinterpret_cast<uintptr_t>(d_quat4) & 15))
        return fail(c, GSC_ERR_INVALID, "gsc_bind_device_transforms: quaternions must be 16-byte aligned");
    if (!c->rot_scale_valid && (!d_quat4 || !d_scale3) && c->n_box)
        return fail(c, GSC_ERR_INVALID, "gsc_bind_device_transforms: no rotation / scale on the device for the current collider set: pass them");
    // (`packed` describes the layout of the OWNED rotation / scale arrays; a NULL here keeps using them as they are,
    // caller arrays are always dense -- so the two cannot be mixed while the owned ones are packed)
    if (c->packed && c->n_box && ((d_quat4 == nullptr) != (d_scale3 == nullptr)))
        return fail(c, GSC_ERR_INVALID, "gsc_bind_device_transforms: after a packed update pass rotation AND scale, or neither");
    c->chunked = false;
    c->b_pos = d_pos3;
    c->b_quat = d_quat4 ? d_quat4 : c->d_quat;
    c->b_scale = d_scale3 ? d_scale3 : c->d_scale;
    c->have_transforms = true;
    return GSC_OK;
}

// ---- stages ------------------------------------------------------------------------------------------

static int stage_bounds(gsc_ctx *c)
{
    if (!c->have_colliders || !c->have_transforms)
        return fail(c, GSC_ERR_INVALID, "step before colliders and transforms were provided");
    cudaStream_t s = c->stream;
    CU(c, cudaEventRecord(c->ev_begin, s));
    CU(c, cudaMemsetAsync(c->d_fb, 0, sizeof(FrameBlock), s));
    c->n_ghost = 0;
    const bool timing = c->cfg.flags & GSC_FLAG_TIMING;
    if (timing) CU(c, cudaEventRecord(c->ev[0], s));
    for (int k = 0; k < gsc_ctx::kChunks && c->n_local; ++k) {
        // owned transforms arrive chunk by chunk on the copy stream; caller-bound device arrays are one chunk
        const uint32_t b = c->chunked ? c->chunk_begin[k] : 0, e = c->chunked ? c->chunk_begin[k + 1] : c->n_local;
        if (c->chunked) CU(c, cudaStreamWaitEvent(s, c->ev_chunk[k], 0));
        if (e > b)
            k_bvh_update<<<(e - b + kBvhThreads - 1) / kBvhThreads, kBvhThreads, 0, s>>>(
                b, e, c->d_kind, c->d_offset, c->d_size, c->b_pos, c->b_quat, c->b_scale, c->d_entity,
                c->have_gidx ? c->d_gidx : nullptr, (c->packed && c->b_quat == c->d_quat) ? c->d_box_slot : nullptr, c->d_rec, c->d_obb, c->d_fb);
        if (!c->chunked) break;
    }
    CU(c, cudaGetLastError());
    CU(c, cudaEventRecord(c->ev_fence, s));  // the frame's transforms are consumed: the next frame's copy may land
    if (timing) CU(c, cudaEventRecord(c->ev[1], s));
    c->bounds_done = true;
    c->global_bounds_set = false;
    c->have_window = false;
    return GSC_OK;
}

// The launches of stages 1-5 and the read-back of the control block.
// `slab`: the step's volume count (local + pushed ghosts) and the windowed grid were produced on the device by
// k_slab_gather / k_slab_wait: launches are sized for the capacity and clamp to FrameBlock::n_total
static int stage_launches(gsc_ctx *c, bool slab, bool timing)
{
    cudaStream_t s = c->stream;
    const uint32_t n_total = slab ? c->cap : c->n_local + c->n_ghost;
    const uint32_t n_launch = n_total;
    const uint32_t *n_dev = slab ? &c->d_fb->n_total : nullptr;
    FrameBlock *fb = c->d_fb;
    if (!slab) k_grid_setup<<<1, 32, 0, s>>>(fb, c->max_cells, n_total, c->have_window ? 1 : 0, c->win_lo, c->win_hi);
    if (n_total) {
        k_cell_keys<<<tiles_for(n_total), kKeyThreads, 0, s>>>(n_total, n_dev, c->d_rec, c->d_keys[0], c->d_wide, c->d_sort_scratch, fb);
        if (timing) CU(c, cudaEventRecord(c->ev[2], s));
        // the keys are cell ids <= max_cells: no more radix passes can ever be needed than that many bits take (passes
        // beyond the step's real count exit at once on the device, but each of their launches still costs ~4 us)
        int key_bits_max = 1;
        while (key_bits_max < 32 && (1ull << key_bits_max) <= c->max_cells) ++key_bits_max;
        launch_sort(s, n_total, c->d_keys, c->d_vals, true, std::min(kMaxPasses, radix_passes(key_bits_max)), &fb->sort,
                    c->d_sort_scratch, true, n_dev);
        if (timing) CU(c, cudaEventRecord(c->ev[3], s));
        const uint32_t tb = (n_total + 1 + kTableSpan - 1) / kTableSpan;
        SortedVolumes sv{ c->d_qrec_sorted, c->d_w3_sorted };
        k_cell_table_permute<<<tb, kTableThreads, 0, s>>>(n_total, n_dev, c->d_keys[0], c->d_vals[0], c->d_keys[1], c->d_vals[1],
                                                         c->d_rec, sv, c->d_cell_start, c->d_gaps, fb);
        k_fill_gaps<<<148 * 2, 256, 0, s>>>(c->d_cell_start, c->d_gaps, fb);
        if (timing) CU(c, cudaEventRecord(c->ev[4], s));
        // sphere-sphere pairs are decided inside the pairs stage unless a contact has to be computed for them
        PairOut po{ c->d_cand_ss, c->d_cand_so, c->d_cand_oo, c->max_cands };
        if (c->pairs_tiled) {  // (experimental, measured slower than k_pairs: DESIGN section 4)
            k_pairs_tiled<<<(n_launch + kTileA - 1) / kTileA, kTileA, sizeof(TileShared), s>>>(n_launch, n_dev, c->d_keys[0], c->d_keys[1],
                                                                                            c->d_vals[0], c->d_vals[1], sv, c->d_cell_start,
                                                                                            po, c->d_overflow, fb);
            // tiles too dense for the staging buffers (normally none: the kernel exits at once)
            k_pairs<true><<<148 * 16, kPairThreads, 0, s>>>(n_launch, n_dev, c->d_keys[0], c->d_keys[1], c->d_vals[0], c->d_vals[1], sv,
                                                            c->d_cell_start, po, c->d_overflow, fb);
        } else {
            k_pairs<false><<<(n_launch + kPairThreads - 1) / kPairThreads, kPairThreads, 0, s>>>(n_launch, n_dev, c->d_keys[0], c->d_keys[1],
                                                                                              c->d_vals[0], c->d_vals[1], sv,
                                                                                              c->d_cell_start, po, nullptr, fb);
        }
        if (timing) CU(c, cudaEventRecord(c->ev[5], s));
        if (c->d_contacts)
            k_narrow<true><<<148 * kNarrowCtasPerSm, kNarrowThreads, 0, s>>>(c->d_cand_ss, c->d_cand_so, c->d_cand_oo, c->max_cands, c->d_rec,
                                                            c->d_obb, c->d_pairs, c->d_contacts, c->max_pairs, fb);
        else
            k_narrow<false><<<148 * kNarrowCtasPerSm, kNarrowThreads, 0, s>>>(c->d_cand_ss, c->d_cand_so, c->d_cand_oo, c->max_cands, c->d_rec,
                                                             c->d_obb, c->d_pairs, nullptr, c->max_pairs, fb);
        k_wide_pairs<<<148 * 2, kNarrowThreads, 0, s>>>(c->n_local, c->d_wide, c->d_rec, c->d_obb, c->d_pairs, c->d_contacts,
                                                       c->max_pairs, fb);  // (normally no wide volumes: exits at once)
        if (timing) CU(c, cudaEventRecord(c->ev[6], s));
    } else if (timing) {
        for (int i = 2; i <= GSC_NUM_STAGES; ++i) CU(c, cudaEventRecord(c->ev[i], s));
    }
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(c->h_fb, fb, sizeof(FrameBlock), cudaMemcpyDeviceToHost, s));
    return GSC_OK;
}

static int stage_enqueue(gsc_ctx *c, bool slab = false)
{
    cudaStream_t s = c->stream;
    const bool timing = c->cfg.flags & GSC_FLAG_TIMING;
    // GSC_FLAG_GRAPH: the launch sequence has no host round trip and its geometry only depends on the collider count
    // (and the slab window), so it is captured once and replayed with ONE graph launch per step
    if ((c->cfg.flags & GSC_FLAG_GRAPH) && !timing) {
        // (a slab step launches for the capacity and reads the real count from the device: its geometry never changes)
        const uint32_t n_total = slab ? c->cap : c->n_local + c->n_ghost;
        const bool window = !slab && c->have_window;
        const bool same = c->graph_exec && c->graph_n == n_total && c->graph_local == c->n_local && c->graph_slab == slab &&
                          c->graph_window == window &&
                          (!window || (c->graph_win[0] == c->win_lo && c->graph_win[1] == c->win_hi));
        if (!same) {
            if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
            c->graph_exec = nullptr;
            cudaGraph_t graph = nullptr;
            CU(c, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            const int rc = stage_launches(c, slab, false);
            const cudaError_t e = cudaStreamEndCapture(s, &graph);
            if (rc != GSC_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return fail(c, GSC_ERR_CUDA, "graph capture: %s", cudaGetErrorString(e));
            const cudaError_t e2 = cudaGraphInstantiate(&c->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e2 != cudaSuccess) { c->graph_exec = nullptr; return fail(c, GSC_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e2)); }
            c->graph_n = n_total;
            c->graph_local = c->n_local;  // (baked into k_wide_pairs' arguments)
            c->graph_slab = slab;
            c->graph_window = window;
            c->graph_win[0] = c->win_lo;
            c->graph_win[1] = c->win_hi;
        }
        CU(c, cudaGraphLaunch(c->graph_exec, s));
    } else if (int rc = stage_launches(c, slab, timing)) {
        return rc;
    }
    CU(c, cudaEventRecord(c->ev_end, s));
    c->bounds_done = false;
    c->perm_valid = false;
    c->adj_valid = false;
    c->last_pairs = 0;
    return GSC_OK;
}

static int stage_collect(gsc_ctx *c, gsc_step_out *out)
{
    const bool timing = c->cfg.flags & GSC_FLAG_TIMING;
    CU(c, cudaStreamSynchronize(c->stream));
    const FrameBlock &h = *c->h_fb;
    c->last_pairs = std::min<uint64_t>(h.n_pairs, c->max_pairs);
    if (c->slab_step) c->n_ghost = h.slab.n_ghost;
    c->sorted_valid = h.status == 0 && h.grid.valid;
    if (out) {
        memset(out, 0, sizeof *out);
        out->n_pairs = h.n_pairs;
        out->n_sphere_sphere = h.n_ss;
        out->n_sphere_obb = h.n_so;
        out->n_obb_obb = h.n_oo;
        out->n_cells = h.grid.num_cells;
        out->cell_size = h.grid.cell_size;
        out->n_wide = h.n_wide;
        for (int a = 0; a < 3; ++a) {
            out->grid_origin[a] = h.grid.origin[a];
            out->grid_dims[a] = h.grid.dims[a];
        }
        out->sort_passes = (uint32_t)h.sort.num_passes;
        out->status_flags = h.status;
        out->d_pairs = reinterpret_cast<const uint32_t *>(c->d_pairs);
        out->d_contacts = reinterpret_cast<const float *>(c->d_contacts);
        out->n_ghost = c->slab_step ? h.slab.n_ghost : c->n_ghost;
        cudaEventElapsedTime(&out->step_ms, c->ev_begin, c->ev_end);
        if (timing)
            for (int i = 0; i < GSC_NUM_STAGES; ++i) cudaEventElapsedTime(&out->stage_ms[i], c->ev[i], c->ev[i + 1]);
    }
    if (int rc = status_to_error(c, h)) { c->last_pairs = 0; return rc; }
    if (h.n_ss > c->max_cands || h.n_so > c->max_cands || h.n_oo > c->max_cands)
        return fail(c, GSC_ERR_CAPACITY, "candidate list overflow: need %llu / %llu / %llu, have %llu (gsc_config.max_candidates)",
                    (unsigned long long)h.n_ss, (unsigned long long)h.n_so, (unsigned long long)h.n_oo,
                    (unsigned long long)c->max_cands);
    if (h.n_pairs > c->max_pairs)
        return fail(c, GSC_ERR_CAPACITY, "pair buffer overflow: need %llu, have %llu (gsc_config.max_pairs)",
                    (unsigned long long)h.n_pairs, (unsigned long long)c->max_pairs);
    return GSC_OK;
}

static int stage_finish(gsc_ctx *c, gsc_step_out *out)
{
    if (int rc = stage_enqueue(c)) return rc;
    return stage_collect(c, out);
}

int gsc_step(gsc_ctx *c, gsc_step_out *out)
{
    if (int rc = check_ctx(c)) return rc;
    NOT_IN_FLIGHT(c, "gsc_step");
    c->slab_step = false;
    if (int rc = stage_bounds(c)) return rc;
    return stage_finish(c, out);
}

int gsc_step_begin(gsc_ctx *c)
{
    if (int rc = check_ctx(c)) return rc;
    NOT_IN_FLIGHT(c, "gsc_step_begin");
    c->slab_step = false;
    if (int rc = stage_bounds(c)) return rc;
    if (int rc = stage_enqueue(c)) {
        c->bounds_done = false;  // nothing of this step is usable: the next call starts over
        return rc;
    }
    c->in_flight = true;
    return GSC_OK;
}

int gsc_step_end(gsc_ctx *c, gsc_step_out *out)
{
    if (int rc = check_ctx(c)) return rc;
    if (!c->in_flight) return fail(c, GSC_ERR_INVALID, "gsc_step_end without gsc_step_begin");
    c->in_flight = false;
    return stage_collect(c, out);
}

int gsc_phase_bounds(gsc_ctx *c, gsc_bounds *local)
{
    if (int rc = check_ctx(c)) return rc;
    NOT_IN_FLIGHT(c, "gsc_phase_bounds");
    if (!local) return fail(c, GSC_ERR_INVALID, "gsc_phase_bounds: null output");
    c->slab_step = false;
    if (int rc = stage_bounds(c)) return rc;
    CU(c, cudaMemcpyAsync(c->h_fb, c->d_fb, sizeof(FrameBlock), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    const FrameBlock &h = *c->h_fb;
    local->status_flags = h.status;
    if (c->n_local == 0 || h.bounds[0] == 0) {
        local->longest_axis = 0.0f;
        for (int a = 0; a < 3; ++a) { local->world_min[a] = INFINITY; local->world_max[a] = -INFINITY; }
        local->home_x_max = -INFINITY;
    } else {
        local->home_x_max = ord2f(h.bounds[7]);
        local->longest_axis = ord2f(h.bounds[0]);
        for (int a = 0; a < 3; ++a) {
            local->world_min[a] = ord2f(~h.bounds[1 + a]);
            local->world_max[a] = ord2f(h.bounds[4 + a]);
        }
    }
    return GSC_OK;
}

static uint32_t host_f2ord(float f)
{
    uint32_t b;
    memcpy(&b, &f, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

int gsc_set_global_bounds(gsc_ctx *c, const gsc_bounds *g)
{
    if (int rc = check_ctx(c)) return rc;
    if (!g || !c->bounds_done) return fail(c, GSC_ERR_INVALID, "gsc_set_global_bounds: call gsc_phase_bounds first");
    uint32_t *enc = c->h_enc;  // pinned, owned by the context: no sync needed before returning
    enc[0] = host_f2ord(g->longest_axis);
    for (int a = 0; a < 3; ++a) {
        enc[1 + a] = ~host_f2ord(g->world_min[a]);
        enc[4 + a] = host_f2ord(g->world_max[a]);
    }
    enc[7] = g->status_flags;
    CU(c, cudaMemcpyAsync(c->d_fb, enc, 7 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    if (g->status_flags) CU(c, cudaMemcpyAsync(&c->d_fb->status, enc + 7, 4, cudaMemcpyHostToDevice, c->stream));
    // place the (global) grid now: gsc_local_x_range / gsc_halo_select / gsc_halo_push need the cell size
    k_grid_setup<<<1, 32, 0, c->stream>>>(c->d_fb, ~0ull, c->cap, 0, 0, 0);
    CU(c, cudaGetLastError());
    c->global_bounds_set = true;
    return GSC_OK;
}

// ---- halo push over peer memory ------------------------------------------------------------------------
struct IpcBlob {  // what gsc_ipc_export hands to the other ranks (GSC_IPC_EXPORT_BYTES)
    cudaIpcMemHandle_t rec, obb, fb;
    uint32_t cap;
    uint32_t device;
    uint64_t ctx_tag;  // address of the exporting context: lets a rank recognise its own blob
    uint32_t pad[12];
};
static_assert(sizeof(IpcBlob) == GSC_IPC_EXPORT_BYTES, "IpcBlob size");

int gsc_ipc_export(gsc_ctx *c, void *blob, size_t blob_bytes)
{
    if (int rc = check_ctx(c)) return rc;
    if (!blob || blob_bytes < sizeof(IpcBlob)) return fail(c, GSC_ERR_INVALID, "gsc_ipc_export: need %zu bytes", sizeof(IpcBlob));
    IpcBlob b;
    memset(&b, 0, sizeof b);
    CU(c, cudaIpcGetMemHandle(&b.rec, c->d_rec));
    CU(c, cudaIpcGetMemHandle(&b.obb, c->d_obb));
    CU(c, cudaIpcGetMemHandle(&b.fb, c->d_fb));
    b.cap = c->cap;
    b.device = (uint32_t)c->device;
    b.ctx_tag = (uint64_t)(uintptr_t)c ^ ((uint64_t)getpid() << 40);
    memcpy(blob, &b, sizeof b);
    return GSC_OK;
}

int gsc_ipc_open_peer(gsc_ctx *c, uint32_t peer, const void *blob, size_t blob_bytes)
{
    if (int rc = check_ctx(c)) return rc;
    if (peer >= GSC_MAX_PEERS || !blob || blob_bytes < sizeof(IpcBlob)) return fail(c, GSC_ERR_INVALID, "gsc_ipc_open_peer: bad argument");
    IpcBlob b;
    memcpy(&b, blob, sizeof b);
    gsc_ctx::Peer &p = c->peers[peer];
    if (p.rec && !p.self && !p.local) {  // re-open: drop the old mapping
        cudaIpcCloseMemHandle(p.rec);
        cudaIpcCloseMemHandle(p.obb);
        cudaIpcCloseMemHandle(p.fb);
    }
    p = gsc_ctx::Peer();
    if (b.ctx_tag == ((uint64_t)(uintptr_t)c ^ ((uint64_t)getpid() << 40))) {  // my own export
        p.rec = c->d_rec;
        p.obb = c->d_obb;
        p.fb = c->d_fb;
        p.cap = c->cap;
        p.self = true;
        return GSC_OK;
    }
    CU(c, cudaIpcOpenMemHandle((void **)&p.rec, b.rec, cudaIpcMemLazyEnablePeerAccess));
    CU(c, cudaIpcOpenMemHandle((void **)&p.obb, b.obb, cudaIpcMemLazyEnablePeerAccess));
    CU(c, cudaIpcOpenMemHandle((void **)&p.fb, b.fb, cudaIpcMemLazyEnablePeerAccess));
    p.cap = b.cap;
    return GSC_OK;
}

int gsc_halo_push(gsc_ctx *c, uint32_t peer, int32_t x_lo, int32_t x_hi, uint32_t peer_n_local)
{
    if (int rc = check_ctx(c)) return rc;
    if (peer >= GSC_MAX_PEERS || !c->peers[peer].rec || c->peers[peer].self)
        return fail(c, GSC_ERR_INVALID, "gsc_halo_push: peer %u was not opened with gsc_ipc_open_peer", peer);
    if (!c->global_bounds_set) return fail(c, GSC_ERR_INVALID, "gsc_halo_push: call gsc_set_global_bounds first");
    const gsc_ctx::Peer &p = c->peers[peer];
    if (c->n_local && x_lo <= x_hi)
        k_halo_push<<<kGridStrideBlocks, 256, 0, c->stream>>>(c->n_local, c->d_rec, c->d_obb, x_lo, x_hi, p.rec, p.obb, p.fb,
                                                            peer_n_local, p.cap, c->d_fb);
    CU(c, cudaGetLastError());
    return GSC_OK;  // asynchronous: gsc_halo_sync waits for it
}

int gsc_halo_sync(gsc_ctx *c)
{
    if (int rc = check_ctx(c)) return rc;
    CU(c, cudaStreamSynchronize(c->stream));
    return GSC_OK;
}

int gsc_halo_commit(gsc_ctx *c, uint32_t *n_ghost)
{
    if (int rc = check_ctx(c)) return rc;
    if (!c->bounds_done) return fail(c, GSC_ERR_INVALID, "gsc_halo_commit: call gsc_phase_bounds first");
    uint32_t pushed = 0;
    CU(c, cudaMemcpyAsync(&pushed, &c->d_fb->n_ghost_in, sizeof pushed, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (n_ghost) *n_ghost = pushed;
    if ((uint64_t)c->n_local + c->n_ghost + pushed > c->cap)
        return fail(c, GSC_ERR_CAPACITY, "gsc_halo_commit: %u local + %u pushed ghosts > max_colliders %u", c->n_local, pushed, c->cap);
    c->n_ghost += pushed;
    return GSC_OK;
}

// ---- slab exchange driven from the device (k_slab_*) -----------------------------------------------------------
static SlabPeers slab_peers(const gsc_ctx *c)
{
    SlabPeers sp;
    memset(&sp, 0, sizeof sp);
    for (uint32_t s = 0; s < c->slab_world; ++s) {
        const gsc_ctx::Peer &p = c->peers[s];
        sp.rec[s] = p.rec;
        sp.obb[s] = p.obb;
        sp.fb[s] = p.fb;
        sp.mail[s] = p.mail();
        sp.cap[s] = p.cap;
    }
    return sp;
}

// phase A: bvh_update over the local colliders, then my bounds go into every rank's mailbox
static int slab_phase_a(gsc_ctx *c)
{
    if (!c->slab_ready) return fail(c, GSC_ERR_INVALID, "slab step: call gsc_slab_init first");
    if (int rc = stage_bounds(c)) return rc;
    c->slab_step = true;
    ++c->slab_seq;
    k_slab_publish<<<1, 32, 0, c->stream>>>(c->d_fb, slab_peers(c), c->slab_rank, c->slab_world, c->n_local, c->slab_seq);
    CU(c, cudaGetLastError());
    CU(c, cudaEventRecord(c->ev_slab[0], c->stream));
    return GSC_OK;
}

// phase B: wait for every rank's bounds, place my grid, push my boundary volumes into the peers that need them
static int slab_phase_b(gsc_ctx *c)
{
    cudaStream_t s = c->stream;
    k_slab_gather<<<1, 32, 0, s>>>(c->d_fb, c->d_mail, c->slab_rank, c->slab_world, c->slab_seq, c->max_cells, c->n_local + 1,
                                   c->peer_timeout_ns);
    if (c->n_local && c->slab_world > 1)
        k_slab_push<<<dim3(148 * 2, c->slab_world), kPushThreads, 0, s>>>(c->n_local, c->d_rec, c->d_obb, slab_peers(c), c->slab_rank, c->d_fb);
    k_slab_done<<<1, 32, 0, s>>>(slab_peers(c), c->slab_rank, c->slab_world, c->slab_seq);
    CU(c, cudaGetLastError());
    CU(c, cudaEventRecord(c->ev_slab[1], s));
    return GSC_OK;
}

// phase C: wait for the peers' pushes, adopt the ghosts, stages 1-5
static int slab_phase_c(gsc_ctx *c)
{
    k_slab_wait<<<1, 32, 0, c->stream>>>(c->d_fb, c->d_mail, c->slab_rank, c->slab_world, c->slab_seq, c->n_local, c->cap,
                                         c->peer_timeout_ns);
    CU(c, cudaGetLastError());
    return stage_enqueue(c, true);
}

int gsc_slab_init(gsc_ctx *c, uint32_t rank, uint32_t world)
{
    if (int rc = check_ctx(c)) return rc;
    NOT_IN_FLIGHT(c, "gsc_slab_init");
    if (world == 0 || world > GSC_MAX_PEERS || rank >= world) return fail(c, GSC_ERR_INVALID, "gsc_slab_init: rank %u of %u", rank, world);
    gsc_ctx::Peer &me = c->peers[rank];
    if (!me.rec) {  // a rank is its own peer `rank`
        me = gsc_ctx::Peer();
        me.rec = c->d_rec;
        me.obb = c->d_obb;
        me.fb = c->d_fb;
        me.cap = c->cap;
        me.self = true;
    }
    for (uint32_t s = 0; s < world; ++s)
        if (!c->peers[s].rec) return fail(c, GSC_ERR_INVALID, "gsc_slab_init: peer %u was not opened (gsc_ipc_open_peer / gsc_open_peer_local)", s);
    c->slab_rank = rank;
    c->slab_world = world;
    c->slab_ready = true;
    return GSC_OK;
}

int gsc_open_peer_local(gsc_ctx *c, uint32_t peer, gsc_ctx *other)
{
    if (int rc = check_ctx(c)) return rc;
    if (peer >= GSC_MAX_PEERS || !other) return fail(c, GSC_ERR_INVALID, "gsc_open_peer_local: bad argument");
    gsc_ctx::Peer &p = c->peers[peer];
    if (p.rec && !p.self && !p.local) {
        cudaIpcCloseMemHandle(p.rec);
        cudaIpcCloseMemHandle(p.obb);
        cudaIpcCloseMemHandle(p.fb);
    }
    p = gsc_ctx::Peer();
    if (other->device != c->device) {
        int can = 0;
        CU(c, cudaDeviceCanAccessPeer(&can, c->device, other->device));
        if (!can) return fail(c, GSC_ERR_CUDA, "gsc_open_peer_local: device %d cannot access device %d", c->device, other->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(other->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) return fail(c, GSC_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", other->device, cudaGetErrorString(e));
    }
    p.rec = other->d_rec;
    p.obb = other->d_obb;
    p.fb = other->d_fb;
    p.cap = other->cap;
    p.self = other == c;
    p.local = other != c;
    return GSC_OK;
}

int gsc_slab_step_begin(gsc_ctx *c)
{
    if (int rc = check_ctx(c)) return rc;
    NOT_IN_FLIGHT(c, "gsc_slab_step_begin");
    if (int rc = slab_phase_a(c)) return rc;
    if (int rc = slab_phase_b(c)) return rc;
    if (int rc = slab_phase_c(c)) return rc;
    c->in_flight = true;
    return GSC_OK;
}

int gsc_slab_step(gsc_ctx *c, gsc_step_out *out)
{
    if (int rc = gsc_slab_step_begin(c)) return rc;
    return gsc_step_end(c, out);
}

int gsc_set_x_window(gsc_ctx *c, int32_t x_lo, int32_t x_hi)
{
    if (int rc = check_ctx(c)) return rc;
    if (!c->bounds_done) return fail(c, GSC_ERR_INVALID, "gsc_set_x_window: call gsc_phase_bounds first");
    c->have_window = true;
    c->win_lo = x_lo;
    c->win_hi = x_hi;
    return GSC_OK;
}

int gsc_phase_finish(gsc_ctx *c, gsc_step_out *out)
{
    if (int rc = check_ctx(c)) return rc;
    if (!c->bounds_done) return fail(c, GSC_ERR_INVALID, "gsc_phase_finish: call gsc_phase_bounds first");
    return stage_finish(c, out);
}

int gsc_local_x_range(gsc_ctx *c, int32_t *x_min, int32_t *x_max)
{
    if (int rc = check_ctx(c)) return rc;
    if (!x_min || !x_max) return fail(c, GSC_ERR_INVALID, "gsc_local_x_range: null output");
    if (!c->global_bounds_set) return fail(c, GSC_ERR_INVALID, "gsc_local_x_range: call gsc_set_global_bounds first");
    cudaStream_t s = c->stream;
    CU(c, cudaMemsetAsync(c->d_fb->xrange, 0, sizeof(uint32_t) * 2, s));
    if (c->n_local) k_local_x_range<<<kGridStrideBlocks, 256, 0, s>>>(c->n_local, c->d_rec, c->d_fb);
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(c->h_fb, c->d_fb, sizeof(FrameBlock), cudaMemcpyDeviceToHost, s));
    CU(c, cudaStreamSynchronize(s));
    const FrameBlock &h = *c->h_fb;
    if (int rc = status_to_error(c, h)) return rc;
    if (h.xrange[1] == 0) {  // no local volumes (or no grid): empty range
        *x_min = 1;
        *x_max = 0;
    } else {
        *x_min = (int32_t)(~h.xrange[0]) - 32768;
        *x_max = (int32_t)h.xrange[1] - 32768;
    }
    return GSC_OK;
}

int gsc_halo_select(gsc_ctx *c, int32_t x_lo, int32_t x_hi, void *d_records, uint32_t cap_records, uint32_t *count)
{
    if (int rc = check_ctx(c)) return rc;
    if (!count || (cap_records && !d_records)) return fail(c, GSC_ERR_INVALID, "gsc_halo_select: null argument");
    if (!c->global_bounds_set) return fail(c, GSC_ERR_INVALID, "gsc_halo_select: call gsc_set_global_bounds first");
    if (reinterpret_cast<uintptr_t>(d_records) & 15) return fail(c, GSC_ERR_INVALID, "gsc_halo_select: buffer must be 16-byte aligned");
    cudaStream_t s = c->stream;
    CU(c, cudaMemsetAsync(c->d_halo_count, 0, sizeof(uint32_t), s));
    if (c->n_local && x_lo <= x_hi)
        k_halo_select<<<kGridStrideBlocks, 256, 0, s>>>(c->n_local, c->d_rec, c->d_obb, x_lo, x_hi, (float4 *)d_records,
                                                      cap_records, c->d_halo_count, c->d_fb);
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(count, c->d_halo_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CU(c, cudaStreamSynchronize(s));
    if (*count > cap_records)
        return fail(c, GSC_ERR_CAPACITY, "gsc_halo_select: %u records selected, buffer holds %u", *count, cap_records);
    return GSC_OK;
}

int gsc_halo_append(gsc_ctx *c, uint32_t count, const void *d_records)
{
    if (int rc = check_ctx(c)) return rc;
    if (!c->bounds_done) return fail(c, GSC_ERR_INVALID, "gsc_halo_append: call gsc_phase_bounds first");
    if (count == 0) return GSC_OK;
    if (!d_records || (reinterpret_cast<uintptr_t>(d_records) & 15)) return fail(c, GSC_ERR_INVALID, "gsc_halo_append: bad buffer");
    if ((uint64_t)c->n_local + c->n_ghost + count > c->cap)
        return fail(c, GSC_ERR_CAPACITY, "gsc_halo_append: %u local + %u ghost + %u new > max_colliders %u", c->n_local, c->n_ghost,
                    count, c->cap);
    k_halo_append<<<(count + 255) / 256, 256, 0, c->stream>>>(count, (const float4 *)d_records, c->n_local + c->n_ghost, c->d_rec,
                                                              c->d_obb);
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(c->stream));  // the caller may reuse d_records right away
    c->n_ghost += count;
    return GSC_OK;
}

// ---- results -----------------------------------------------------------------------------------------

}  // extern "C"

namespace {
__global__ void k_extract_field(const uint2 *pairs, const uint32_t *perm, uint32_t n, int field, uint32_t *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 p = pairs[perm ? perm[i] : i];
    out[i] = field ? p.y : p.x;
}
__global__ void k_gather_pairs(const uint2 *pairs, const uint32_t *perm, uint32_t n, uint2 *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = pairs[perm[i]];
}
__global__ void k_gather_contacts(const float4 *contacts, const uint32_t *perm, uint32_t n, float4 *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = contacts[perm[i]];
}

// standalone sort of n (key,value) pairs with full control over its scratch; result lands in buffer index *which
int sort_standalone(gsc_ctx *c, cudaStream_t s, uint32_t n, uint32_t *keys[2], uint32_t *vals[2], bool identity, int key_bits,
                    int *which)
{
    const int passes = radix_passes(key_bits);
    *which = passes & 1;
    if (n == 0 || passes == 0) { *which = 0; return GSC_OK; }
    Scratch hold;
    uint32_t *scratch = nullptr;  // [ctl 4 words][tile histogram + digit totals]
    const size_t words = 4 + sort_scratch_words(n);
    CU(c, cudaMalloc((void **)&scratch, words * 4));
    hold.p[0] = scratch;  // freed on every return path
    SortControl *ctl = reinterpret_cast<SortControl *>(scratch);
    SortControl h{ passes, key_bits, radix_digit_bits(key_bits) };
    CU(c, cudaMemcpyAsync(ctl, &h, sizeof h, cudaMemcpyHostToDevice, s));
    launch_sort(s, n, keys, vals, identity, passes, ctl, scratch + 4);
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(s));
    return GSC_OK;
}
}  // namespace

extern "C" {

// permutation that orders the pairs of the last step ascending by (first << 32 | second); cached per step
static int ensure_sorted_perm(gsc_ctx *c)
{
    if (c->perm_valid) return GSC_OK;
    const uint64_t n = c->last_pairs;
    if (n >= (1ull << 30)) return fail(c, GSC_ERR_CAPACITY, "sorted download supports < 2^30 pairs");
    cudaStream_t s = c->stream;
    // LSD over the 64-bit tuple with the 32-bit sort: second entity first, then (stable) the first
    const uint32_t m = (uint32_t)n;
    uint32_t *k[2] = { nullptr, nullptr }, *v[2] = { nullptr, nullptr };
    Scratch hold;  // the four temporaries are freed on every return path
    for (int i = 0; i < 2; ++i) {
        CU(c, dmalloc(&k[i], m));
        hold.p[i] = k[i];
        CU(c, dmalloc(&v[i], m));
        hold.p[2 + i] = v[i];
    }
    if (!c->d_perm) CU(c, dmalloc(&c->d_perm, c->max_pairs));
    const uint32_t blocks = (m + 255) / 256;
    int w1 = 0, w2 = 0;
    k_extract_field<<<blocks, 256, 0, s>>>(c->d_pairs, nullptr, m, 1, k[0]);
    int rc = sort_standalone(c, s, m, k, v, true, 32, &w1);
    if (rc == GSC_OK) {
        if (w1 != 0) { std::swap(k[0], k[1]); std::swap(v[0], v[1]); }
        k_extract_field<<<blocks, 256, 0, s>>>(c->d_pairs, v[0], m, 0, k[0]);
        rc = sort_standalone(c, s, m, k, v, false, 32, &w2);
    }
    if (rc == GSC_OK) {
        cudaError_t e = cudaMemcpyAsync(c->d_perm, v[w2], (size_t)m * 4, cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = fail(c, GSC_ERR_CUDA, "sorted download: %s", cudaGetErrorString(e));
    }
    if (rc == GSC_OK) c->perm_valid = true;
    return rc;
}

int gsc_download_pairs(gsc_ctx *c, uint32_t *dst, uint64_t cap_pairs, int sorted, uint64_t *n_out)
{
    if (int rc = check_ctx(c)) return rc;
    NOT_IN_FLIGHT(c, "gsc_download_pairs");
    const uint64_t n = c->last_pairs;
    if (n_out) *n_out = n;
    if (n == 0 || (!dst && cap_pairs == 0)) return GSC_OK;  // (NULL, 0) only queries the count
    if (!dst || cap_pairs < n) return fail(c, GSC_ERR_CAPACITY, "gsc_download_pairs: need room for %llu pairs", (unsigned long long)n);
    cudaStream_t s = c->stream;
    const uint2 *src = c->d_pairs;
    if (sorted) {
        if (int rc = ensure_sorted_perm(c)) return rc;
        if (!c->d_pairs_sorted) CU(c, dmalloc(&c->d_pairs_sorted, c->max_pairs));
        k_gather_pairs<<<((uint32_t)n + 255) / 256, 256, 0, s>>>(c->d_pairs, c->d_perm, (uint32_t)n, c->d_pairs_sorted);
        src = c->d_pairs_sorted;
    }
    CU(c, cudaMemcpyAsync(dst, src, n * sizeof(uint2), cudaMemcpyDeviceToHost, s));
    CU(c, cudaStreamSynchronize(s));
    return GSC_OK;
}

int gsc_download_contacts(gsc_ctx *c, float *dst, uint64_t cap_pairs, int sorted, uint64_t *n_out)
{
    if (int rc = check_ctx(c)) return rc;
    NOT_IN_FLIGHT(c, "gsc_download_contacts");
    if (!c->d_contacts) return fail(c, GSC_ERR_INVALID, "gsc_download_contacts: the context was created without GSC_FLAG_CONTACTS");
    const uint64_t n = c->last_pairs;
    if (n_out) *n_out = n;
    if (n == 0 || (!dst && cap_pairs == 0)) return GSC_OK;
    if (!dst || cap_pairs < n) return fail(c, GSC_ERR_CAPACITY, "gsc_download_contacts: need room for %llu contacts", (unsigned long long)n);
    cudaStream_t s = c->stream;
    const float4 *src = c->d_contacts;
    if (sorted) {
        if (int rc = ensure_sorted_perm(c)) return rc;
        if (!c->d_contacts_sorted) CU(c, dmalloc(&c->d_contacts_sorted, c->max_pairs));
        k_gather_contacts<<<((uint32_t)n + 255) / 256, 256, 0, s>>>(c->d_contacts, c->d_perm, (uint32_t)n, c->d_contacts_sorted);
        src = c->d_contacts_sorted;
    }
    CU(c, cudaMemcpyAsync(dst, src, n * sizeof(float4), cudaMemcpyDeviceToHost, s));
    CU(c, cudaStreamSynchronize(s));
    return GSC_OK;
}

// ---- adjacency: what CollisionCallbackManager::process_collisions builds (mod.rs:678-708) ---------------------
}  // extern "C"

namespace {
// every pair (a, b) contributes a -> b and b -> a (mod.rs:687-690)
__global__ void k_adj_expand(const uint2 *pairs, uint32_t h, uint32_t *first, uint32_t *second)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h) return;
    const uint2 p = pairs[i];
    first[2 * i] = p.x; second[2 * i] = p.y;
    first[2 * i + 1] = p.y; second[2 * i + 1] = p.x;
}
__global__ void k_gather_u32(const uint32_t *src, const uint32_t *perm, uint32_t n, uint32_t *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[perm[i]];
}
constexpr int kHeadTile = 1024;
// pass 1: number of list heads (first[i] != first[i-1]) per tile; pass 3: their ranks -> entity / offset arrays
__global__ void __launch_bounds__(256) k_adj_heads(const uint32_t *sorted_first, uint32_t m, uint32_t *tile_count,
                                                 const uint32_t *tile_base, uint32_t *entity, uint32_t *offset)
{
    __shared__ uint32_t s_warp[8];
    const uint32_t base = blockIdx.x * kHeadTile + threadIdx.x * 4;
    uint32_t f[4], cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t i = base + k;
        f[k] = (i < m && (i == 0 || sorted_first[i] != sorted_first[i - 1])) ? 1u : 0u;
        cnt += f[k];
    }
    uint32_t total = 0;
    uint32_t run = block_excl_scan256(cnt, s_warp, &total);
    if (!tile_base) {
        if (threadIdx.x == 0) tile_count[blockIdx.x] = total;
        return;
    }
    run += tile_base[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (f[k]) {
            entity[run] = sorted_first[base + k];
            offset[run] = base + k;
            ++run;
        }
}
// pass 2: exclusive scan of the tile counts on one CTA; total -> *n_heads, offset[total] = m
__global__ void __launch_bounds__(256) k_adj_scan_tiles(uint32_t *tile_count, uint32_t tiles, uint32_t m, uint32_t *n_heads,
                                                      uint32_t *offset)
{
    __shared__ uint32_t s_warp[8];
    uint32_t carry = 0;
    for (uint32_t b = 0; b < tiles; b += 256) {
        const uint32_t i = b + threadIdx.x;
        const uint32_t v = i < tiles ? tile_count[i] : 0u;
        uint32_t total = 0;
        const uint32_t e = block_excl_scan256(v, s_warp, &total);
        if (i < tiles) tile_count[i] = carry + e;
        carry += total;
    }
    if (threadIdx.x == 0) {
        *n_heads = carry;
        offset[carry] = m;
    }
}
}  // namespace

extern "C" {

int gsc_build_adjacency(gsc_ctx *c, gsc_adjacency *out)
{
    if (int rc = check_ctx(c)) return rc;
    NOT_IN_FLIGHT(c, "gsc_build_adjacency");
    if (!out) return fail(c, GSC_ERR_INVALID, "gsc_build_adjacency: null output");
    memset(out, 0, sizeof *out);
    const uint64_t h = c->last_pairs;
    if (2 * h >= (1ull << 30)) return fail(c, GSC_ERR_CAPACITY, "gsc_build_adjacency supports < 2^29 pairs");
    const uint32_t m = (uint32_t)(2 * h);
    if (!c->adj_valid) {
        cudaStream_t s = c->stream;
        if (c->adj_cap < m + 1ull) {
            for (uint32_t **p : { &c->d_adj_entity, &c->d_adj_offset, &c->d_adj_other }) {
                if (*p) cudaFree(*p);
                *p = nullptr;
            }
            c->adj_cap = std::max<uint64_t>(m + 1ull, 2 * c->max_pairs / 4 + 1024);
            CU(c, dmalloc(&c->d_adj_entity, c->adj_cap));
            CU(c, dmalloc(&c->d_adj_offset, c->adj_cap + 1));
            CU(c, dmalloc(&c->d_adj_other, c->adj_cap));
        }
        c->adj_entities = 0;
        c->adj_entries = m;
        if (m) {
            uint32_t *first = nullptr, *second = nullptr, *k[2] = { nullptr, nullptr }, *v[2] = { nullptr, nullptr }, *tiles_buf = nullptr;
            const uint32_t tiles = (m + kHeadTile - 1) / kHeadTile;
            Scratch sc[2];
            CU(c, dmalloc(&first, m)); sc[0].p[0] = first;
            CU(c, dmalloc(&second, m)); sc[0].p[1] = second;
            CU(c, dmalloc(&tiles_buf, tiles + 2)); sc[0].p[2] = tiles_buf;
            for (int i = 0; i < 2; ++i) {
                CU(c, dmalloc(&k[i], m)); sc[1].p[i] = k[i];
                CU(c, dmalloc(&v[i], m)); sc[1].p[2 + i] = v[i];
            }
            const uint32_t hb = ((uint32_t)h + 255) / 256, mb = (m + 255) / 256;
            k_adj_expand<<<hb, 256, 0, s>>>(c->d_pairs, (uint32_t)h, first, second);
            // LSD over (first, second): by the neighbour first, then (stable) by the owning entity
            CU(c, cudaMemcpyAsync(k[0], second, (size_t)m * 4, cudaMemcpyDeviceToDevice, s));
            int w1 = 0, w2 = 0;
            if (int rc = sort_standalone(c, s, m, k, v, true, 32, &w1)) return rc;
            if (w1 != 0) { std::swap(k[0], k[1]); std::swap(v[0], v[1]); }
            k_gather_u32<<<mb, 256, 0, s>>>(first, v[0], m, k[0]);
            if (int rc = sort_standalone(c, s, m, k, v, false, 32, &w2)) return rc;
            // k[w2] = owning entity of every entry, v[w2] = its position in the expanded list
            k_gather_u32<<<mb, 256, 0, s>>>(second, v[w2], m, c->d_adj_other);
            k_adj_heads<<<tiles, 256, 0, s>>>(k[w2], m, tiles_buf, nullptr, nullptr, nullptr);
            k_adj_scan_tiles<<<1, 256, 0, s>>>(tiles_buf, tiles, m, tiles_buf + tiles, c->d_adj_offset);
            k_adj_heads<<<tiles, 256, 0, s>>>(k[w2], m, nullptr, tiles_buf, c->d_adj_entity, c->d_adj_offset);
            uint32_t n_heads = 0;
            CU(c, cudaMemcpyAsync(&n_heads, tiles_buf + tiles, 4, cudaMemcpyDeviceToHost, s));
            CU(c, cudaStreamSynchronize(s));
            CU(c, cudaGetLastError());
            c->adj_entities = n_heads;
        }
        c->adj_valid = true;
    }
    out->n_entities = c->adj_entities;
    out->n_entries = c->adj_entries;
    out->d_entity = c->d_adj_entity;
    out->d_offset = c->d_adj_offset;
    out->d_other = c->d_adj_other;
    return GSC_OK;
}

int gsc_download_adjacency(gsc_ctx *c, uint32_t *entity, uint32_t *offset, uint32_t *other, uint64_t cap_entities,
                           uint64_t cap_entries)
{
    gsc_adjacency a;
    if (int rc = gsc_build_adjacency(c, &a)) return rc;
    if (cap_entities < a.n_entities || cap_entries < a.n_entries)
        return fail(c, GSC_ERR_CAPACITY, "gsc_download_adjacency: need room for %llu entities and %llu entries",
                    (unsigned long long)a.n_entities, (unsigned long long)a.n_entries);
    if (a.n_entries == 0) {
        if (offset) offset[0] = 0;
        return GSC_OK;
    }
    if (!entity || !offset || !other) return fail(c, GSC_ERR_INVALID, "gsc_download_adjacency: null array");
    CU(c, cudaMemcpy(entity, a.d_entity, a.n_entities * 4, cudaMemcpyDeviceToHost));
    CU(c, cudaMemcpy(offset, a.d_offset, (a.n_entities + 1) * 4, cudaMemcpyDeviceToHost));
    CU(c, cudaMemcpy(other, a.d_other, a.n_entries * 4, cudaMemcpyDeviceToHost));
    return GSC_OK;
}

// ---- temporal coherence: keep the colliders in the cell order of the last step ----------------------
}  // extern "C"
namespace {
__global__ void k_reorder_static(uint32_t n, const uint32_t *__restrict__ order, const uint8_t *__restrict__ kind,
                                 const float *__restrict__ offset3, const float *__restrict__ size3, const uint32_t *__restrict__ entity,
                                 const uint32_t *__restrict__ gidx, uint8_t *__restrict__ kind_o, float *__restrict__ offset_o,
                                 float *__restrict__ size_o, uint32_t *__restrict__ entity_o, uint32_t *__restrict__ gidx_o)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t i = order[s];
    kind_o[s] = kind[i];
    for (int a = 0; a < 3; ++a) { offset_o[3 * (size_t)s + a] = offset3[3 * (size_t)i + a]; size_o[3 * (size_t)s + a] = size3[3 * (size_t)i + a]; }
    entity_o[s] = entity[i];
    gidx_o[s] = gidx ? gidx[i] : i;  // the position in the reference's dense array keeps orienting every tuple
}
}  // namespace
extern "C" {

int gsc_reorder_colliders(gsc_ctx *c, uint32_t *order, uint32_t cap)
{
    if (int rc = check_ctx(c)) return rc;
    NOT_IN_FLIGHT(c, "gsc_reorder_colliders");
    if (!c->sorted_valid) return fail(c, GSC_ERR_INVALID, "gsc_reorder_colliders: needs a completed step on the current collider set");
    const uint32_t n = c->n_local;
    if (!order || cap < n) return fail(c, GSC_ERR_CAPACITY, "gsc_reorder_colliders: need room for %u entries", n);
    if (n == 0) return GSC_OK;
    cudaStream_t s = c->stream;
    // the sorted index stream of the last step; after a slab step it also holds the ghosts (local index >= n_local),
    // which are dropped here: the local colliders keep their relative cell order
    const uint32_t n_sorted = c->slab_step ? c->h_fb->n_total : n + c->n_ghost;
    const uint32_t *vals = c->d_vals[c->h_fb->sort.num_passes & 1];
    Scratch ord;
    {
        std::vector<uint32_t> all(n_sorted);
        CU(c, cudaMemcpyAsync(all.data(), vals, (size_t)n_sorted * 4, cudaMemcpyDeviceToHost, s));
        CU(c, cudaStreamSynchronize(s));
        uint32_t k = 0;
        for (uint32_t i = 0; i < n_sorted && k < n; ++i)
            if (all[i] < n) order[k++] = all[i];
        if (k != n) return fail(c, GSC_ERR_INVALID, "gsc_reorder_colliders: the last step's order covers %u of %u local colliders", k, n);
        CU(c, cudaMalloc(&ord.p[0], (size_t)n * 4));
        CU(c, cudaMemcpyAsync(ord.p[0], order, (size_t)n * 4, cudaMemcpyHostToDevice, s));
        vals = (const uint32_t *)ord.p[0];
    }
    Scratch tmp[2];
    uint8_t *kind_o = nullptr;
    float *off_o = nullptr, *size_o = nullptr;
    uint32_t *ent_o = nullptr, *gidx_o = nullptr;
    CU(c, cudaMalloc(&tmp[0].p[0], n)); kind_o = (uint8_t *)tmp[0].p[0];
    CU(c, cudaMalloc(&tmp[0].p[1], (size_t)n * 12)); off_o = (float *)tmp[0].p[1];
    CU(c, cudaMalloc(&tmp[0].p[2], (size_t)n * 12)); size_o = (float *)tmp[0].p[2];
    CU(c, cudaMalloc(&tmp[0].p[3], (size_t)n * 4)); ent_o = (uint32_t *)tmp[0].p[3];
    CU(c, cudaMalloc(&tmp[1].p[0], (size_t)n * 4)); gidx_o = (uint32_t *)tmp[1].p[0];
    k_reorder_static<<<(n + 255) / 256, 256, 0, s>>>(n, vals, c->d_kind, c->d_offset, c->d_size, c->d_entity, c->have_gidx ? c->d_gidx : nullptr,
                                                   kind_o, off_o, size_o, ent_o, gidx_o);
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(c->d_kind, kind_o, n, cudaMemcpyDeviceToDevice, s));
    CU(c, cudaMemcpyAsync(c->d_offset, off_o, (size_t)n * 12, cudaMemcpyDeviceToDevice, s));
    CU(c, cudaMemcpyAsync(c->d_size, size_o, (size_t)n * 12, cudaMemcpyDeviceToDevice, s));
    CU(c, cudaMemcpyAsync(c->d_entity, ent_o, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    CU(c, cudaMemcpyAsync(c->d_gidx, gidx_o, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    CU(c, cudaStreamSynchronize(s));
    std::vector<uint8_t> hk(n);
    for (uint32_t k = 0; k < n; ++k) hk[k] = c->h_kind[order[k]];
    c->h_kind.swap(hk);
    c->have_gidx = true;
    if (int rc = rebuild_box_slots(c)) return rc;
    c->have_transforms = false;
    c->rot_scale_valid = false;
    c->sorted_valid = false;
    c->packed = false;
    c->chunked = false;
    c->hier_colliders = 0;
    return GSC_OK;
}

// ---- transform hierarchy (the step before the path) --------------------------------------------

int gsc_hierarchy_upload(gsc_ctx *c, uint32_t n_nodes, uint32_t n_rows, const uint32_t *row_begin, const uint32_t *parent,
                         uint32_t n_colliders, const uint32_t *collider_node)
{
    if (int rc = check_ctx(c)) return rc;
    NOT_IN_FLIGHT(c, "gsc_hierarchy_upload");
    if (!c->have_colliders || n_colliders != c->n_local)
        return fail(c, GSC_ERR_INVALID, "gsc_hierarchy_upload: n_colliders=%u does not match the %u uploaded colliders", n_colliders, c->n_local);
    if (!row_begin || (n_nodes && !parent) || (n_colliders && !collider_node))
        return fail(c, GSC_ERR_INVALID, "gsc_hierarchy_upload: null array");
    if (row_begin[0] != 0 || row_begin[n_rows] != n_nodes)
        return fail(c, GSC_ERR_INVALID, "gsc_hierarchy_upload: row_begin must run from 0 to n_nodes");
    for (uint32_t r = 0; r < n_rows; ++r) {
        if (row_begin[r + 1] < row_begin[r]) return fail(c, GSC_ERR_INVALID, "gsc_hierarchy_upload: row_begin is not ascending");
        for (uint32_t i = row_begin[r]; i < row_begin[r + 1]; ++i) {
            const bool ok = r == 0 ? parent[i] == GSC_NO_PARENT : (parent[i] >= row_begin[r - 1] && parent[i] < row_begin[r]);
            if (!ok) return fail(c, GSC_ERR_INVALID, "gsc_hierarchy_upload: node %u of row %u has parent %u outside row %d", i, r, parent[i], (int)r - 1);
        }
    }
    for (uint32_t k = 0; k < n_colliders; ++k)
        if (collider_node[k] >= n_nodes)
            return fail(c, GSC_ERR_INVALID, "gsc_hierarchy_upload: collider %u names node %u of %u", k, collider_node[k], n_nodes);
    if (n_nodes > c->hier_cap) {
        // (a failure below must not leave a previous hierarchy usable over half-replaced buffers)
        c->hier_cap = 0;
        c->hier_row_begin.clear();
        c->hier_colliders = 0;
        c->hier_derived = false;
        float **fbufs[] = { &c->d_lpos, &c->d_lquat, &c->d_lscale, &c->d_dpos, &c->d_dquat, &c->d_dscale, &c->d_dmat };
        const size_t per[] = { 3, 4, 3, 3, 4, 3, 16 };
        for (int i = 0; i < 7; ++i) {
            if (*fbufs[i]) cudaFree(*fbufs[i]);
            *fbufs[i] = nullptr;
            CU(c, dmalloc(fbufs[i], per[i] * (size_t)n_nodes));
        }
        if (c->d_hier_parent) cudaFree(c->d_hier_parent);
        c->d_hier_parent = nullptr;
        CU(c, dmalloc(&c->d_hier_parent, n_nodes));
        c->hier_cap = n_nodes;
        c->hier_local[0] = c->hier_local[1] = c->hier_local[2] = false;
    }
    if (!c->d_hier_node) CU(c, dmalloc(&c->d_hier_node, c->cap));
    if (n_nodes) CU(c, cudaMemcpyAsync(c->d_hier_parent, parent, (size_t)n_nodes * 4, cudaMemcpyHostToDevice, c->stream));
    if (n_colliders) CU(c, cudaMemcpyAsync(c->d_hier_node, collider_node, (size_t)n_colliders * 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->hier_nodes = n_nodes;
    c->hier_row_begin.assign(row_begin, row_begin + n_rows + 1);
    c->hier_colliders = n_colliders;
    c->hier_derived = false;
    // a new hierarchy: the local transforms on the device belong to the previous one ("NULL keeps the previous
    // values" must not reuse them)
    c->hier_local[0] = c->hier_local[1] = c->hier_local[2] = false;
    return GSC_OK;
}

int gsc_hierarchy_update(gsc_ctx *c, const float *local_pos3, const float *local_quat4, const float *local_scale3)
{
    if (int rc = check_ctx(c)) return rc;
    if (c->hier_row_begin.empty() || c->hier_colliders != c->n_local)
        return fail(c, GSC_ERR_INVALID, "gsc_hierarchy_update: no hierarchy uploaded for the current %u colliders", c->n_local);
    cudaStream_t s = c->stream;
    const size_t n = c->hier_nodes;
    const float *src[3] = { local_pos3, local_quat4, local_scale3 };
    float *dst[3] = { c->d_lpos, c->d_lquat, c->d_lscale };
    const size_t per[3] = { 12, 16, 12 };
    for (int i = 0; i < 3; ++i) {
        if (src[i]) {
            if (n) CU(c, cudaMemcpyAsync(dst[i], src[i], n * per[i], cudaMemcpyHostToDevice, s));
            c->hier_local[i] = true;
        } else if (!c->hier_local[i]) {
            return fail(c, GSC_ERR_INVALID, "gsc_hierarchy_update: the first update needs position, rotation and scale");
        }
    }
    const uint32_t rows = (uint32_t)c->hier_row_begin.size() - 1;
    for (uint32_t r = 0; r < rows; ++r) {  // one launch per row: a row only depends on the one above it
        const uint32_t b = c->hier_row_begin[r], e = c->hier_row_begin[r + 1];
        if (e > b)
            k_transform_row<<<(e - b + kHierThreads - 1) / kHierThreads, kHierThreads, 0, s>>>(
                b, e, c->d_hier_parent, c->d_lpos, reinterpret_cast<const float4 *>(c->d_lquat), c->d_lscale, c->d_dpos,
                reinterpret_cast<float4 *>(c->d_dquat), c->d_dscale, reinterpret_cast<float4 *>(c->d_dmat));
    }
    // a host-to-device transform copy still landing in the owned arrays must not be overtaken
    if (c->chunked) CU(c, cudaStreamWaitEvent(s, c->ev_chunk[gsc_ctx::kChunks - 1], 0));
    if (c->n_local)
        k_hier_bind<<<(c->n_local + 255) / 256, 256, 0, s>>>(c->n_local, c->d_hier_node, c->d_dpos,
                                                            reinterpret_cast<const float4 *>(c->d_dquat), c->d_dscale, c->d_pos,
                                                            reinterpret_cast<float4 *>(c->d_quat), c->d_scale);
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(s));  // the caller's host arrays are borrowed for the call only
    c->hier_derived = true;
    c->rot_scale_valid = true;  // k_hier_bind wrote rotation and scale of every collider
    c->packed = false;
    c->chunked = false;
    c->b_pos = c->d_pos;
    c->b_quat = c->d_quat;
    c->b_scale = c->d_scale;
    c->have_transforms = true;
    return GSC_OK;
}

int gsc_hierarchy_download(gsc_ctx *c, float *pos3, float *quat4, float *scale3, float *matrix16)
{
    if (int rc = check_ctx(c)) return rc;
    if (!c->hier_derived) return fail(c, GSC_ERR_INVALID, "gsc_hierarchy_download before gsc_hierarchy_update");
    const size_t n = c->hier_nodes;
    if (n == 0) return GSC_OK;
    if (pos3) CU(c, cudaMemcpy(pos3, c->d_dpos, n * 12, cudaMemcpyDeviceToHost));
    if (quat4) CU(c, cudaMemcpy(quat4, c->d_dquat, n * 16, cudaMemcpyDeviceToHost));
    if (scale3) CU(c, cudaMemcpy(scale3, c->d_dscale, n * 12, cudaMemcpyDeviceToHost));
    if (matrix16) CU(c, cudaMemcpy(matrix16, c->d_dmat, n * 64, cudaMemcpyDeviceToHost));
    return GSC_OK;
}

int gsc_debug_download(gsc_ctx *c, int what, void *dst, size_t dst_bytes)
{
    if (int rc = check_ctx(c)) return rc;
    NOT_IN_FLIGHT(c, "gsc_debug_download");
    if (!dst) return fail(c, GSC_ERR_INVALID, "gsc_debug_download: null destination");
    const uint32_t n = c->n_local + c->n_ghost;
    const FrameBlock &h = *c->h_fb;
    const int which = h.sort.num_passes & 1;
    const void *src = nullptr;
    size_t bytes = 0;
    switch (what) {
    case GSC_DBG_AABB: {
        bytes = (size_t)n * 24;
        if (dst_bytes < bytes) return fail(c, GSC_ERR_CAPACITY, "gsc_debug_download: need %zu bytes", bytes);
        if (n == 0) return GSC_OK;
        float *tmp = nullptr;
        CU(c, cudaMalloc((void **)&tmp, bytes));
        k_decode_aabb<<<(n + 255) / 256, 256, 0, c->stream>>>(n, c->d_rec, tmp);
        cudaError_t e = cudaMemcpyAsync(dst, tmp, bytes, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        cudaFree(tmp);
        if (e != cudaSuccess) return fail(c, GSC_ERR_CUDA, "gsc_debug_download: %s", cudaGetErrorString(e));
        return GSC_OK;
    }
    case GSC_DBG_SORTED_KEYS: src = c->d_keys[which]; bytes = (size_t)n * 4; break;
    case GSC_DBG_SORTED_IDX: src = c->d_vals[which]; bytes = (size_t)n * 4; break;
    case GSC_DBG_CELL_START: src = c->d_cell_start; bytes = ((size_t)h.grid.num_cells + 1) * 4; break;
    case GSC_DBG_OBB: src = c->d_obb; bytes = (size_t)n * 64; break;
    case GSC_DBG_ENTITY: src = c->d_entity; bytes = (size_t)c->n_local * 4; break;
    case GSC_DBG_KIND: src = c->d_kind; bytes = (size_t)c->n_local; break;
    case GSC_DBG_OFFSET: src = c->d_offset; bytes = (size_t)c->n_local * 12; break;
    case GSC_DBG_SIZE: src = c->d_size; bytes = (size_t)c->n_local * 12; break;
    default: return fail(c, GSC_ERR_INVALID, "gsc_debug_download: unknown item %d", what);
    }
    if (dst_bytes < bytes) return fail(c, GSC_ERR_CAPACITY, "gsc_debug_download: need %zu bytes", bytes);
    CU(c, cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return GSC_OK;
}

// ---- micro-bench batches -------------------------------------------------------------------------------

}  // extern "C"

namespace {
int batch_fail(const char *what, cudaError_t e) { return fail(nullptr, GSC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e)); }
#define BT(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return batch_fail(#expr, e_); } while (0)

template <class Launch>
int run_batch(int device, size_t n, const void *a, size_t a_stride, const void *b, size_t b_stride, void *out,
              size_t out_stride, Launch launch)
{
    if (n == 0) return GSC_OK;
    if (!a || !out) return fail(nullptr, GSC_ERR_INVALID, "batch: null array");
    BT(cudaSetDevice(device));
    Scratch s;
    BT(cudaMalloc(&s.p[0], n * a_stride));
    BT(cudaMemcpy(s.p[0], a, n * a_stride, cudaMemcpyHostToDevice));
    if (b) {
        BT(cudaMalloc(&s.p[1], n * b_stride));
        BT(cudaMemcpy(s.p[1], b, n * b_stride, cudaMemcpyHostToDevice));
    }
    BT(cudaMalloc(&s.p[2], n * out_stride));
    launch(s.p[0], s.p[1], s.p[2], (unsigned)((n + 255) / 256));
    BT(cudaGetLastError());
    BT(cudaMemcpy(out, s.p[2], n * out_stride, cudaMemcpyDeviceToHost));
    return GSC_OK;
}
}  // namespace

extern "C" {

int gsc_test_sphere_sphere(int device, size_t n, const float *a4, const float *b4, uint8_t *out)
{
    return run_batch(device, n, a4, 16, b4, 16, out, 1, [&](void *a, void *b, void *o, unsigned g) {
        k_batch_sphere_sphere<<<g, 256>>>(n, (const float4 *)a, (const float4 *)b, (uint8_t *)o);
    });
}
int gsc_test_sphere_obb(int device, size_t n, const float *s4, const float *obb16, uint8_t *out)
{
    return run_batch(device, n, s4, 16, obb16, 64, out, 1, [&](void *a, void *b, void *o, unsigned g) {
        k_batch_sphere_obb<<<g, 256>>>(n, (const float4 *)a, (const float *)b, (uint8_t *)o);
    });
}
int gsc_test_obb_obb(int device, size_t n, const float *a16, const float *b16, uint8_t *out)
{
    return run_batch(device, n, a16, 64, b16, 64, out, 1, [&](void *a, void *b, void *o, unsigned g) {
        k_batch_obb_obb<<<g, 256>>>(n, (const float *)a, (const float *)b, (uint8_t *)o);
    });
}
int gsc_test_aabb_aabb(int device, size_t n, const float *a6, const float *b6, uint8_t *out)
{
    return run_batch(device, n, a6, 24, b6, 24, out, 1, [&](void *a, void *b, void *o, unsigned g) {
        k_batch_aabb_aabb<<<g, 256>>>(n, (const float *)a, (const float *)b, (uint8_t *)o);
    });
}
int gsc_fnv1a_gridcell(int device, size_t n, const int16_t *cells3, uint64_t *out)
{
    return run_batch(device, n, cells3, 6, nullptr, 0, out, 8, [&](void *a, void *, void *o, unsigned g) {
        k_batch_fnv1a<<<g, 256>>>(n, (const int16_t *)a, (uint64_t *)o);
    });
}

int gsc_contact_sphere_sphere(int device, size_t n, const float *hi4, const float *lo4, float *out4)
{
    return run_batch(device, n, hi4, 16, lo4, 16, out4, 16, [&](void *a, void *b, void *o, unsigned g) {
        k_batch_contact_ss<<<g, 256>>>(n, (const float4 *)a, (const float4 *)b, (float4 *)o);
    });
}
int gsc_contact_obb_obb(int device, size_t n, const float *hi16, const float *lo16, float *out4)
{
    return run_batch(device, n, hi16, 64, lo16, 64, out4, 16, [&](void *a, void *b, void *o, unsigned g) {
        k_batch_contact_oo<<<g, 256>>>(n, (const float *)a, (const float *)b, (float4 *)o);
    });
}
int gsc_contact_sphere_obb(int device, size_t n, const float *s4, const float *obb16, const uint8_t *sphere_is_hi, float *out4)
{
    if (n == 0) return GSC_OK;
    if (!sphere_is_hi) return fail(nullptr, GSC_ERR_INVALID, "batch: null array");
    BT(cudaSetDevice(device));
    Scratch flag;
    BT(cudaMalloc(&flag.p[0], n));
    BT(cudaMemcpy(flag.p[0], sphere_is_hi, n, cudaMemcpyHostToDevice));
    const uint8_t *d_flag = (const uint8_t *)flag.p[0];
    return run_batch(device, n, s4, 16, obb16, 64, out4, 16, [&](void *a, void *b, void *o, unsigned g) {
        k_batch_contact_so<<<g, 256>>>(n, (const float4 *)a, (const float *)b, d_flag, (float4 *)o);
    });
}

int gsc_sort_pairs_u32(int device, size_t n, uint32_t *keys, uint32_t *vals, int key_bits)
{
    if (n == 0) return GSC_OK;
    if (!keys || !vals || key_bits < 1 || key_bits > 32 || n >= (1ull << 30))
        return fail(nullptr, GSC_ERR_INVALID, "gsc_sort_pairs_u32: bad argument");
    BT(cudaSetDevice(device));
    gsc_ctx tmp;  // only for error text plumbing of sort_standalone
    tmp.device = device;
    uint32_t *k[2] = { nullptr, nullptr }, *v[2] = { nullptr, nullptr };
    Scratch s;
    for (int i = 0; i < 2; ++i) {
        BT(cudaMalloc((void **)&k[i], n * 4));
        s.p[i] = k[i];
        BT(cudaMalloc((void **)&v[i], n * 4));
        s.p[2 + i] = v[i];
    }
    BT(cudaMemcpy(k[0], keys, n * 4, cudaMemcpyHostToDevice));
    BT(cudaMemcpy(v[0], vals, n * 4, cudaMemcpyHostToDevice));
    int which = 0;
    if (int rc = sort_standalone(&tmp, 0, (uint32_t)n, k, v, false, key_bits, &which)) {
        g_create_error = tmp.err;
        return rc;
    }
    BT(cudaMemcpy(keys, k[which], n * 4, cudaMemcpyDeviceToHost));
    BT(cudaMemcpy(vals, v[which], n * 4, cudaMemcpyDeviceToHost));
    return GSC_OK;
}

}  // extern "C"

#include "gsc_group.cuh"

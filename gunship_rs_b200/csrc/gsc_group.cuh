// gsc_group.cuh -- gsc_group_*: one process, one context per GPU, the x-slab partition behind one handle.
// (included at the end of gsc_api.cu: it drives the slab phases of the member contexts directly)
//
// The reference decomposes space into 8 octant work units that each scan every volume and dedupe at the merge
// (grid_collision.rs:60-86, 161-193, 262-272).  Here the units are x-slabs, one per GPU; every collider has ONE owner,
// a pair is reported by the owner of its taker, and the one-cell halo travels over NVLink peer memory inside the
// step (k_slab_*).  The group is what a Rust / C++ engine links against: it registers colliders and hands over
// transforms in the engine's dense order and gets the union of the members' pair lists back.
#pragma once

#include <numeric>
#include <thread>

struct gsc_group {
    std::vector<gsc_ctx *> m;
    std::vector<int32_t> devices;
    gsc_config cfg{};
    std::string err;
    uint32_t n = 0;
    bool have_colliders = false, have_transforms = false;
    // the engine's dense arrays, kept for re-partitioning
    std::vector<uint32_t> entity;
    std::vector<uint8_t> kind;
    std::vector<float> offset, size;
    std::vector<std::vector<uint32_t>> owned;            // global indices per member in the member's dense (slot) order
    std::vector<float> cuts;                             // x positions of the world-1 slab cuts
    std::vector<gsc_step_out> last;                      // per member, last step
    std::vector<uint8_t> rot_sent;                       // per member: rotation + scale are on the device for the current set
    uint32_t bank = 0;
    std::vector<uint64_t> pair_scratch;
};

namespace {

int gfail(gsc_group *g, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (g) g->err = buf;
    else g_create_error = buf;
    return code;
}

int gmember_fail(gsc_group *g, uint32_t i, int rc)
{
    g->err = "member " + std::to_string(i) + " (device " + std::to_string(g->devices[i]) + "): " + gsc_last_error(g->m[i]);
    return rc;
}

// owner of every collider from the slab cuts: member r owns cuts[r-1] <= x < cuts[r]
inline uint32_t owner_of(const std::vector<float> &cuts, float x)
{
    return (uint32_t)(std::upper_bound(cuts.begin(), cuts.end(), x) - cuts.begin());
}

// (re)distribute the static collider data according to g->cuts and the positions `pos3`
int group_distribute(gsc_group *g, const float *pos3)
{
    const uint32_t world = (uint32_t)g->m.size(), n = g->n;
    for (auto &o : g->owned) o.clear();
    for (uint32_t i = 0; i < n; ++i) g->owned[owner_of(g->cuts, pos3[3 * (size_t)i])].push_back(i);
    for (uint32_t r = 0; r < world; ++r) {
        const std::vector<uint32_t> &own = g->owned[r];
        const size_t k = own.size();
        if (k > g->cfg.max_colliders)
            return gfail(g, GSC_ERR_CAPACITY, "slab %u owns %zu colliders > max_colliders %u per member", r, k, g->cfg.max_colliders);
        std::vector<uint32_t> e(k);
        std::vector<uint8_t> kd(k);
        std::vector<float> of(3 * k), sz(3 * k);
        for (size_t j = 0; j < k; ++j) {
            const size_t i = own[j];
            e[j] = g->entity[i];
            kd[j] = g->kind[i];
            for (int a = 0; a < 3; ++a) { of[3 * j + a] = g->offset[3 * i + a]; sz[3 * j + a] = g->size[3 * i + a]; }
        }
        if (int rc = gsc_upload_colliders(g->m[r], (uint32_t)k, e.data(), kd.data(), of.data(), sz.data(), own.data())) return gmember_fail(g, r, rc);
        g->rot_sent[r] = 0;
    }
    g->have_transforms = false;
    return GSC_OK;
}

// scatter one frame of dense transforms into the members' pinned staging and hand it over
int group_scatter(gsc_group *g, const float *pos3, const float *quat4, const float *scale3)
{
    const uint32_t world = (uint32_t)g->m.size();
    std::vector<int> rcs(world, GSC_OK);
    std::vector<float *> sp(world), sq(world), ss(world);
    for (uint32_t r = 0; r < world; ++r)
        if (int rc = gsc_map_transform_staging(g->m[r], g->bank, &sp[r], &sq[r], &ss[r])) return gmember_fail(g, r, rc);
    // the gather out of the engine's arrays is host memory work: one thread per member
    auto fill = [&](uint32_t r) {
        const std::vector<uint32_t> &own = g->owned[r];
        const bool rot = quat4 && scale3;
        size_t b = 0;
        for (size_t j = 0; j < own.size(); ++j) {
            const size_t i = own[j];
            sp[r][3 * j] = pos3[3 * i]; sp[r][3 * j + 1] = pos3[3 * i + 1]; sp[r][3 * j + 2] = pos3[3 * i + 2];
            if (g->kind[i] == GSC_BOX) {
                if (rot) {
                    for (int a = 0; a < 4; ++a) sq[r][4 * b + a] = quat4[4 * i + a];
                    for (int a = 0; a < 3; ++a) ss[r][3 * b + a] = scale3[3 * i + a];
                }
                ++b;
            }
        }
    };
    std::vector<std::thread> th;
    for (uint32_t r = 1; r < world; ++r) th.emplace_back(fill, r);
    fill(0);
    for (auto &t : th) t.join();
    for (uint32_t r = 0; r < world; ++r) {
        uint32_t n_box = 0;
        if (int rc = gsc_box_slots(g->m[r], nullptr, 0, &n_box)) return gmember_fail(g, r, rc);
        const int rot = (quat4 && scale3) ? 1 : 0;
        if (!rot && !g->rot_sent[r] && n_box)
            return gfail(g, GSC_ERR_INVALID, "gsc_group_update_transforms: rotation and scale are needed after the colliders (or the partition) changed");
        if (int rc = gsc_commit_transform_staging(g->m[r], g->bank, (uint32_t)g->owned[r].size(), n_box, rot, rot)) return gmember_fail(g, r, rc);
        if (rot) g->rot_sent[r] = 1;
    }
    g->bank ^= 1u;
    g->have_transforms = true;
    return GSC_OK;
}

}  // namespace

extern "C" {

int gsc_group_create(uint32_t n_devices, const int32_t *devices, const gsc_config *cfg, gsc_group **out)
{
    if (!out || !cfg || !devices || n_devices == 0 || n_devices > GSC_MAX_PEERS)
        return gfail(nullptr, GSC_ERR_INVALID, "gsc_group_create: 1..%d devices", GSC_MAX_PEERS);
    *out = nullptr;
    gsc_group *g = new gsc_group();
    g->cfg = *cfg;
    g->devices.assign(devices, devices + n_devices);
    g->owned.resize(n_devices);
    g->last.resize(n_devices);
    g->rot_sent.assign(n_devices, 0);
    for (uint32_t r = 0; r < n_devices; ++r) {
        gsc_config c = *cfg;
        c.device = devices[r];
        gsc_ctx *ctx = nullptr;
        if (int rc = gsc_create(&c, &ctx)) {
            gsc_group_destroy(g);
            return rc;  // (gsc_last_error(NULL) holds the text)
        }
        g->m.push_back(ctx);
    }
    for (uint32_t r = 0; r < n_devices; ++r) {
        for (uint32_t s = 0; s < n_devices; ++s)
            if (int rc = gsc_open_peer_local(g->m[r], s, g->m[s])) {
                gfail(nullptr, rc, "gsc_group_create: %s", gsc_last_error(g->m[r]));
                gsc_group_destroy(g);
                return rc;
            }
        if (int rc = gsc_slab_init(g->m[r], r, n_devices)) {
            gfail(nullptr, rc, "gsc_group_create: %s", gsc_last_error(g->m[r]));
            gsc_group_destroy(g);
            return rc;
        }
    }
    *out = g;
    return GSC_OK;
}

void gsc_group_destroy(gsc_group *g)
{
    if (!g) return;
    // all members stop before any member's arrays (mapped into the others) go away
    for (gsc_ctx *c : g->m)
        if (c) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
    for (gsc_ctx *c : g->m) gsc_destroy(c);
    delete g;
}

const char *gsc_group_last_error(const gsc_group *g) { return g ? g->err.c_str() : g_create_error.c_str(); }

int gsc_group_upload_colliders(gsc_group *g, uint32_t n, const uint32_t *entity, const uint8_t *kind, const float *offset3,
                               const float *size3, const float *pos3)
{
    if (!g) return GSC_ERR_INVALID;
    if (n && (!entity || !kind || !offset3 || !size3 || !pos3)) return gfail(g, GSC_ERR_INVALID, "gsc_group_upload_colliders: null array");
    const uint32_t world = (uint32_t)g->m.size();
    g->n = n;
    g->entity.assign(entity, entity + n);
    g->kind.assign(kind, kind + n);
    g->offset.assign(offset3, offset3 + 3 * (size_t)n);
    g->size.assign(size3, size3 + 3 * (size_t)n);
    // count-balanced x-slabs: the cuts are the world-quantiles of x (gsc_group_rebalance refines them by measured work)
    std::vector<float> xs(n);
    for (uint32_t i = 0; i < n; ++i) xs[i] = pos3[3 * (size_t)i];
    g->cuts.assign(world - 1, 0.0f);
    for (uint32_t r = 1; r < world; ++r) {
        const size_t k = (size_t)((uint64_t)n * r / world);
        if (k < n) {
            std::nth_element(xs.begin(), xs.begin() + k, xs.end());
            g->cuts[r - 1] = xs[k];
        } else {
            g->cuts[r - 1] = n ? *std::max_element(xs.begin(), xs.end()) : 0.0f;
        }
    }
    std::sort(g->cuts.begin(), g->cuts.end());
    g->have_colliders = true;
    return group_distribute(g, pos3);
}

int gsc_group_update_transforms(gsc_group *g, uint32_t n, const float *pos3, const float *quat4, const float *scale3)
{
    if (!g) return GSC_ERR_INVALID;
    if (!g->have_colliders || n != g->n) return gfail(g, GSC_ERR_INVALID, "gsc_group_update_transforms: n=%u does not match the %u uploaded colliders", n, g->n);
    if (n && !pos3) return gfail(g, GSC_ERR_INVALID, "gsc_group_update_transforms: null positions");
    if ((quat4 == nullptr) != (scale3 == nullptr)) return gfail(g, GSC_ERR_INVALID, "gsc_group_update_transforms: pass rotation and scale together");
    return group_scatter(g, pos3, quat4, scale3);
}

int gsc_group_step(gsc_group *g, gsc_step_out *total)
{
    if (!g) return GSC_ERR_INVALID;
    if (!g->have_transforms) return gfail(g, GSC_ERR_INVALID, "gsc_group_step before transforms were provided");
    const uint32_t world = (uint32_t)g->m.size();
    // Three rounds of enqueues, ordered by stream events: a member's gather runs after EVERY member's publish, its
    // wait after every member's done flag -- so the flag kernels never actually spin and members may share a GPU.
    for (uint32_t r = 0; r < world; ++r) {
        if (int rc = check_ctx(g->m[r])) return gmember_fail(g, r, rc);
        if (g->m[r]->in_flight) return gfail(g, GSC_ERR_INVALID, "gsc_group_step: member %u has a step in flight", r);
        if (int rc = slab_phase_a(g->m[r])) return gmember_fail(g, r, rc);
    }
    for (uint32_t r = 0; r < world; ++r) {
        if (int rc = check_ctx(g->m[r])) return gmember_fail(g, r, rc);
        for (uint32_t s = 0; s < world; ++s)
            if (s != r && cudaStreamWaitEvent(g->m[r]->stream, g->m[s]->ev_slab[0], 0) != cudaSuccess)
                return gfail(g, GSC_ERR_CUDA, "cudaStreamWaitEvent: %s", cudaGetErrorString(cudaGetLastError()));
        if (int rc = slab_phase_b(g->m[r])) return gmember_fail(g, r, rc);
    }
    for (uint32_t r = 0; r < world; ++r) {
        if (int rc = check_ctx(g->m[r])) return gmember_fail(g, r, rc);
        for (uint32_t s = 0; s < world; ++s)
            if (s != r && cudaStreamWaitEvent(g->m[r]->stream, g->m[s]->ev_slab[1], 0) != cudaSuccess)
                return gfail(g, GSC_ERR_CUDA, "cudaStreamWaitEvent: %s", cudaGetErrorString(cudaGetLastError()));
        if (int rc = slab_phase_c(g->m[r])) return gmember_fail(g, r, rc);
    }
    int first_rc = GSC_OK;
    uint32_t first_bad = 0;
    gsc_step_out sum;
    memset(&sum, 0, sizeof sum);
    for (uint32_t r = 0; r < world; ++r) {
        if (int rc = check_ctx(g->m[r])) return gmember_fail(g, r, rc);
        const int rc = stage_collect(g->m[r], &g->last[r]);  // (every member is collected even if one failed: no step stays in flight)
        if (rc != GSC_OK && first_rc == GSC_OK) { first_rc = rc; first_bad = r; }
        const gsc_step_out &o = g->last[r];
        sum.n_pairs += o.n_pairs;
        sum.n_sphere_sphere += o.n_sphere_sphere;
        sum.n_sphere_obb += o.n_sphere_obb;
        sum.n_obb_obb += o.n_obb_obb;
        sum.n_cells += o.n_cells;
        sum.n_wide += o.n_wide;
        sum.n_ghost += o.n_ghost;
        sum.status_flags |= o.status_flags;
        sum.cell_size = std::max(sum.cell_size, o.cell_size);
        sum.sort_passes = std::max(sum.sort_passes, o.sort_passes);
        sum.step_ms = std::max(sum.step_ms, o.step_ms);
        for (int i = 0; i < GSC_NUM_STAGES; ++i) sum.stage_ms[i] = std::max(sum.stage_ms[i], o.stage_ms[i]);
    }
    if (total) *total = sum;
    if (first_rc != GSC_OK) return gmember_fail(g, first_bad, first_rc);
    return GSC_OK;
}

int gsc_group_download_pairs(gsc_group *g, uint32_t *dst, uint64_t cap_pairs, int sorted, uint64_t *n_out)
{
    if (!g) return GSC_ERR_INVALID;
    const uint32_t world = (uint32_t)g->m.size();
    uint64_t total = 0;
    std::vector<uint64_t> cnt(world, 0);
    for (uint32_t r = 0; r < world; ++r) {
        if (int rc = gsc_download_pairs(g->m[r], nullptr, 0, 0, &cnt[r])) return gmember_fail(g, r, rc);
        total += cnt[r];
    }
    if (n_out) *n_out = total;
    if (total == 0 || (!dst && cap_pairs == 0)) return GSC_OK;
    if (!dst || cap_pairs < total) return gfail(g, GSC_ERR_CAPACITY, "gsc_group_download_pairs: need room for %llu pairs", (unsigned long long)total);
    uint64_t at = 0;
    for (uint32_t r = 0; r < world; ++r) {
        uint64_t got = 0;
        if (cnt[r])
            if (int rc = gsc_download_pairs(g->m[r], dst + 2 * at, cnt[r], sorted, &got)) return gmember_fail(g, r, rc);
        at += cnt[r];
    }
    if (sorted && world > 1) {
        // the members' lists are sorted runs of (first << 32 | second): merge them on the host
        std::vector<uint64_t> &p = g->pair_scratch;
        p.resize(total);
        for (uint64_t i = 0; i < total; ++i) p[i] = ((uint64_t)dst[2 * i] << 32) | dst[2 * i + 1];
        uint64_t lo = 0, mid = cnt[0];
        for (uint32_t r = 1; r < world; ++r) {
            std::inplace_merge(p.begin() + lo, p.begin() + mid, p.begin() + mid + cnt[r]);
            mid += cnt[r];
        }
        for (uint64_t i = 0; i < total; ++i) { dst[2 * i] = (uint32_t)(p[i] >> 32); dst[2 * i + 1] = (uint32_t)p[i]; }
    }
    return GSC_OK;
}

// Re-cut the slabs so that every member gets the same share of the work of the LAST step, then move the colliders
// whose owner changed (they "migrate": the static arrays of every member are rebuilt from the engine's dense arrays).
// `pos3` = the current positions (dense order), which decide the new owners.  Work model of a slab: colliders + 0.7 x
// candidate pairs (the per-collider stages cost about as much as 26 quantised tests, a candidate stands for ~18), spread
// evenly over the slab's x-extent: the cumulative work is piecewise linear in x and the new cuts are its world-quantiles.
// Afterwards the next gsc_group_update_transforms must carry rotation and scale again (the group keeps no copy of the
// engine's transform arrays).
int gsc_group_rebalance(gsc_group *g, uint32_t n_pos, const float *pos3)
{
    if (!g) return GSC_ERR_INVALID;
    if (!g->have_colliders) return gfail(g, GSC_ERR_INVALID, "gsc_group_rebalance before gsc_group_upload_colliders");
    if (n_pos != g->n || (n_pos && !pos3)) return gfail(g, GSC_ERR_INVALID, "gsc_group_rebalance: positions of all %u colliders are needed", g->n);
    const uint32_t world = (uint32_t)g->m.size(), n = g->n;
    if (world < 2 || n == 0) return GSC_OK;
    float xmin = INFINITY, xmax = -INFINITY;
    for (uint32_t i = 0; i < n; ++i) {
        const float x = pos3[3 * (size_t)i];
        xmin = std::min(xmin, x);
        xmax = std::max(xmax, x);
    }
    if (!(xmax > xmin)) return GSC_OK;
    // current ownership follows the CURRENT positions: count per present slab
    std::vector<double> edge(world + 1), work(world, 0.0);
    edge[0] = xmin;
    edge[world] = xmax;
    for (uint32_t r = 1; r < world; ++r) edge[r] = std::min<double>(std::max<double>(g->cuts[r - 1], xmin), xmax);
    std::vector<uint64_t> count(world, 0);
    for (uint32_t i = 0; i < n; ++i) ++count[owner_of(g->cuts, pos3[3 * (size_t)i])];
    for (uint32_t r = 0; r < world; ++r) {
        const gsc_step_out &o = g->last[r];
        const double cands = (double)(o.n_sphere_sphere + o.n_sphere_obb + o.n_obb_obb);
        // candidates per collider of the last step, applied to the slab's present population
        const double per = g->owned[r].empty() ? 0.0 : cands / (double)g->owned[r].size();
        work[r] = (double)count[r] * (1.0 + 0.7 * per);
    }
    const double total = std::accumulate(work.begin(), work.end(), 0.0);
    if (!(total > 0.0)) return GSC_OK;
    std::vector<float> cuts(world - 1);
    uint32_t r = 0;
    double before = 0.0;
    for (uint32_t k = 1; k < world; ++k) {
        const double want = total * k / world;
        while (r + 1 < world && before + work[r] < want) before += work[r++];
        const double f = work[r] > 0.0 ? (want - before) / work[r] : 0.0;
        cuts[k - 1] = (float)(edge[r] + f * (edge[r + 1] - edge[r]));
    }
    std::sort(cuts.begin(), cuts.end());
    g->cuts = cuts;
    return group_distribute(g, pos3);
}

// Temporal coherence behind the group: every member's dense arrays go into the cell order of the step that just
// completed (gsc_reorder_colliders) and the member's `owned` list is permuted the same way -- the group's scatter of the
// engine's dense transform arrays is a gather anyway, so the engine-facing interface does not change at all.
int gsc_group_reorder(gsc_group *g)
{
    if (!g) return GSC_ERR_INVALID;
    if (!g->have_colliders) return gfail(g, GSC_ERR_INVALID, "gsc_group_reorder before gsc_group_upload_colliders");
    const uint32_t world = (uint32_t)g->m.size();
    for (uint32_t r = 0; r < world; ++r) {
        std::vector<uint32_t> &own = g->owned[r];
        std::vector<uint32_t> order(own.size() + 1), moved(own.size());
        if (int rc = gsc_reorder_colliders(g->m[r], order.data(), (uint32_t)order.size())) return gmember_fail(g, r, rc);
        for (size_t k = 0; k < own.size(); ++k) moved[k] = own[order[k]];
        own.swap(moved);
        g->rot_sent[r] = 0;
    }
    g->have_transforms = false;  // the next gsc_group_update_transforms carries rotation and scale again
    return GSC_OK;
}

int gsc_group_member(gsc_group *g, uint32_t i, gsc_ctx **ctx, uint32_t *n_local, uint32_t *owned)
{
    if (!g || i >= g->m.size()) return g ? gfail(g, GSC_ERR_INVALID, "gsc_group_member: no member %u", i) : GSC_ERR_INVALID;
    if (ctx) *ctx = g->m[i];
    if (n_local) *n_local = (uint32_t)g->owned[i].size();
    if (owned && !g->owned[i].empty()) memcpy(owned, g->owned[i].data(), g->owned[i].size() * sizeof(uint32_t));
    return GSC_OK;
}

}  // extern "C"

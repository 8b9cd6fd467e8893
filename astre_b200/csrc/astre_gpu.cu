// libastre_gpu.so -- context, device mirror and the C ABI of include/astre_gpu.h.
// sm_100a only; there is no CPU path in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/astre_gpu.h"
#include "host_math.hpp"
#include "kernels_merge.cuh"
#include "kernels_mirror.cuh"
#include "kernels_sort.cuh"
#include "kernels_transform.cuh"
#include "kernels_fused.cuh"

using namespace astre;

static_assert(sizeof(astre_host::ViewHost) == sizeof(ViewDev), "host/device view records must match");
static_assert(sizeof(astre_mesh_bounds) == sizeof(BoundsDev), "bounds records must match");

namespace {

thread_local std::string g_create_error;

// ---- guarded device allocations (ASTRE_GUARD=1; the GPU tests run with it) ----------------------------------------
// Every device buffer gets a 4 KB canary on both sides; astre_gpu_debug_guard_check() and every dev_free() look at them.
// A store that leaves its buffer by less than a page would otherwise land in a neighbouring allocation unnoticed.
constexpr size_t kGuardBytes = 4096;
constexpr unsigned char kGuardFill = 0xA5;
std::mutex g_guard_mutex;
std::unordered_map<void *, size_t> g_guarded;      // user pointer -> user bytes
uint64_t g_guard_violations = 0;

bool guard_enabled()
{
    static const bool on = [] { const char *e = std::getenv("ASTRE_GUARD"); return e && *e && *e != '0'; }();
    return on;
}

// Bytes of the two canaries of one allocation that no longer hold the fill value.
uint64_t guard_damage(void *user, size_t bytes)
{
    std::vector<unsigned char> h(2 * kGuardBytes);
    unsigned char *base = static_cast<unsigned char *>(user) - kGuardBytes;
    if (cudaMemcpy(h.data(), base, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(h.data() + kGuardBytes, static_cast<unsigned char *>(user) + bytes, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess)
        return 1;
    uint64_t bad = 0;
    for (size_t i = 0; i < h.size(); ++i)
        if (h[i] != kGuardFill) {
            if (!bad)
                std::fprintf(stderr, "astre_gpu: guard of a %zu-byte device buffer overwritten %s it (offset %zd)\n", bytes,
                             i < kGuardBytes ? "before" : "after", i < kGuardBytes ? (ptrdiff_t)i - (ptrdiff_t)kGuardBytes : (ptrdiff_t)(i - kGuardBytes));
            ++bad;
        }
    return bad;
}

cudaError_t dev_alloc_bytes(void **p, size_t bytes)
{
    if (!guard_enabled()) return cudaMalloc(p, bytes);
    const size_t padded = (bytes + 255) / 256 * 256;      // the rear canary starts where the (rounded) buffer ends
    unsigned char *base = nullptr;
    cudaError_t e = cudaMalloc(&base, padded + 2 * kGuardBytes);
    if (e != cudaSuccess) return e;
    if ((e = cudaMemset(base, kGuardFill, kGuardBytes)) != cudaSuccess ||
        (e = cudaMemset(base + kGuardBytes + bytes, kGuardFill, padded - bytes + kGuardBytes)) != cudaSuccess) {
        cudaFree(base);
        return e;
    }
    *p = base + kGuardBytes;
    std::lock_guard<std::mutex> lock(g_guard_mutex);
    g_guarded[*p] = bytes;
    return cudaSuccess;
}

template <class T>
cudaError_t dev_alloc(T **p, size_t bytes) { return dev_alloc_bytes(reinterpret_cast<void **>(p), bytes); }

cudaError_t dev_free(void *p)
{
    if (!p || !guard_enabled()) return cudaFree(p);
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lock(g_guard_mutex);
        auto it = g_guarded.find(p);
        if (it == g_guarded.end()) return cudaFree(p);
        bytes = it->second;
        g_guarded.erase(it);
    }
    cudaDeviceSynchronize();
    const uint64_t bad = guard_damage(p, bytes);
    if (bad) { std::lock_guard<std::mutex> lock(g_guard_mutex); g_guard_violations += bad; }
    return cudaFree(static_cast<unsigned char *>(p) - kGuardBytes);
}

struct Level {
    uint32_t first = 0, count = 0;  // range [first, first+count) when contiguous
    bool contiguous = true;
    size_t list_offset = 0;         // offset into d_level_slots otherwise
};

}  // namespace

namespace {
// What a frame launches, decided on the host before anything is enqueued.
struct FramePlan {
    uint32_t stage_flags, n, n_counts;
    bool cull, sort, sort_now;
    SortPlan plan;      // LSD fallback digits
    BinPlan bp;         // bucket digit
};
}  // namespace

struct astre_gpu_ctx {
    int device = 0;
    int sm_count = 148;
    uint32_t max_entities = 0, cap_entities = 0, max_views = 0;
    uint64_t max_keys = 0;
    uint32_t n_entities = 0;
    cudaStream_t stream = nullptr;       // stream in use (own_stream unless astre_gpu_set_stream)
    cudaStream_t own_stream = nullptr;

    // mirror
    float *d_tiles = nullptr;        // tile-major mirror: cap_entities/128 records of 1536 words
    float *d_prev = nullptr;         // previous-tick copy of the mirror (lazy; interpolation)
    float4 *d_world = nullptr;
    float *d_basis = nullptr;        // fwd3 | up3 | right3, packed (lazy)
    BoundsDev *d_bounds = nullptr;
    uint32_t n_meshes = 0, cap_meshes = 0;
    ViewDev *d_views = nullptr;
    uint32_t cap_view_records = 0;
    uint32_t n_worlds = 1, entities_per_world = 0, views_per_world = 0, world_bits = 0;

    // frame outputs
    unsigned long long *d_keys[2] = {nullptr, nullptr};
    uint32_t *d_counts = nullptr;
    unsigned long long *d_offsets = nullptr;
    uint32_t cap_counts = 0;
    FrameControl *d_ctl = nullptr;
    SortControl *d_sc = nullptr;
    uint32_t *d_bins = nullptr;           // key sort: bucket counts -> bucket starts -> bucket ends (kernels_sort.cuh)
    uint32_t cap_bin_bits = 0;            // log2 of the bins allocated
    uint32_t *d_chunk_first = nullptr;    // key sort: first bucket of every 1024-key chunk
    CountsMirror counts_mirror{{nullptr, nullptr}};   // astre_gpu_set_counts_mirror
    BoardPeers board_peers{};             // astre_gpu_count_board_attach
    CountBoard *d_board = nullptr;        // this rank's count board (astre_gpu_count_board_create); peers map it through CUDA IPC
    uint32_t board_cap = 0;               // counts per rank and slot it holds
    uint32_t *d_board_table = nullptr;    // [kBoardRanks][board_cap]: where astre_gpu_count_board_offsets puts the table by default
    FrameResultHost *h_res = nullptr;     // mapped host block the frame's last launch writes: key total (also the hint that sizes
    FrameResultHost *d_res = nullptr;     // the next frame's buckets), sort outcome, counts and offsets -- read after one sync, no copy
    uint32_t force_lsd = 0;               // ASTRE_SORT_FORCE_LSD=1: always take the LSD fallback (tests)
    int bin_bits_override = -1;           // ASTRE_SORT_BIN_BITS=n: fixed bucket-digit width (tests)
    uint32_t claim_groups = 8;            // work counters of the fused kernel in a flat frame (ASTRE_CLAIM_GROUPS; 0 = per-CTA shared counters)
    uint32_t shader_or = 0;               // OR of every shader id uploaded (live bits of the key's shader field)
    int sort_grid = 148;                  // co-resident CTAs of k_sort_lsd_all
    unsigned long long *d_lookback = nullptr;
    float4 *d_gather = nullptr;
    uint64_t cap_gather = 0;
    RunHead *d_heads = nullptr;          // draw-run heads (f2), in list order
    uint32_t cap_heads = 0;
    uint32_t *d_head_counts = nullptr;   // run heads per CTA of k_run_count, then their exclusive scan (+ total)
    uint64_t list_serial = 0;            // counts every new draw list (frame, deferred sort, debug sort); never wraps
    uint64_t heads_serial = ~0ull;       // the list the heads were built for
    uint32_t n_heads = 0;
    MeshDrawInfo *d_mesh_info = nullptr; // astre_gpu_set_mesh_draw_info
    uint32_t n_mesh_info = 0, cap_mesh_info = 0;
    DrawIndirect *d_indirect = nullptr;  // astre_gpu_build_indirect: one command per run (cap_heads of them)
    uint32_t n_prev = 0;                 // slots that existed at the last astre_gpu_snapshot_trs
    uint32_t epoch = 0;                  // host mirror of SortControl::epoch (one step per frame)
    uint32_t last_want_bits = 0;         // bucket-digit width of the last plan by key count (hysteresis)
    int bits_bias = 0;                   // correction of that width from the buckets' measured crowding
    bool bits_feedback = true;           // ASTRE_SORT_NO_FEEDBACK=1 switches the correction off (measurement)

    // ASTRE_FRAME_GRAPH: the frame's launches as one instantiated CUDA graph
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraph_t graph_obj = nullptr;                    // kept: its node handles address the exec's kernel nodes
    FramePlan graph_plan;
    // views_version: the CONTENT of the view records changed (same counts, same buffers) -- a single-world frame has
    // them as a kernel parameter of its one fused launch, which is patched in place (cudaGraphExecKernelNodeSetParams);
    // params_version: anything else a captured launch has baked in (buffers reallocated, mesh table, view count, mirror)
    uint64_t graph_views_version = 0, graph_params_version = 0, graph_levels_version = 0;
    uint64_t views_version = 1, params_version = 1, levels_version = 1;
    bool capturing = false;
    uint32_t graph_vs_launches = 0;                     // fused launches of the captured frame that read the ViewSet parameter
    cudaGraphNode_t graph_vs_node = nullptr, graph_vs_node_at_instantiate = nullptr;
    const void *graph_vs_fn_at_instantiate = nullptr;
    bool graph_update_ok = true;                        // ASTRE_GRAPH_NO_UPDATE=1: always instantiate anew (measurement)
    struct { FrameParams P; BinPlan bp; LevelTable LT; const void *fn; unsigned grid; } graph_vs_args;
    cudaStream_t graph_stream = nullptr;
    uint32_t graph_launches = 0, graph_captures = 0, graph_view_patches = 0;
    bool last_gathered = false;          // the last frame ran k_gather_list (ASTRE_FRAME_GATHER) on its sorted list
    bool last_timed = false;             // the last frame recorded the cull events (not when replayed from a graph)

    // f3: global draw list across GPUs
    unsigned char *d_publish[2] = {nullptr, nullptr};        // keys | samples | cut table at a fixed address: what peers map
    uint32_t *d_mcuts = nullptr;                             // scratch of astre_gpu_merge_lists: cut tables
    uint32_t cap_msamples = 0;
    unsigned long long *d_merged_keys = nullptr;
    uint8_t *d_merged_src = nullptr;
    uint64_t cap_merged = 0, n_merged = 0;
    uint32_t *d_splits = nullptr, *d_tile_start = nullptr;
    uint32_t cap_merge_tiles = 0;
    std::vector<void *> imports;
    cudaEvent_t ev_merge0 = nullptr, ev_merge1 = nullptr;
    cudaStream_t merge_stream = nullptr;
    uint32_t *d_merge_error = nullptr;                       // set by k_peer_sync when a peer never signalled
    bool merge_p2p_flags = false;
    MergeDims *d_mdims = nullptr;                            // sizes of the merge when they are only known on the device
    bool merge_dev = false;                                  // last merge used device-side sizes
    uint32_t merge_slice_rank = 0, merge_slice_world = 0;
    bool merge_done = false, merge_begun = false;    // merge_begun: ev_merge0 recorded by astre_gpu_merge_prepare, tail still to come

    // staging for uploads
    unsigned char *d_stage = nullptr;    // two staging halves of stage_entities * 64 bytes
    uint32_t stage_entities = 0;
    cudaStream_t copy_stream = nullptr;  // H2D copies of chunk i+1 overlap the ingest kernel of chunk i
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_ingested[2] = {nullptr, nullptr};

    // hierarchy
    std::vector<uint32_t> h_parent;
    bool has_parents = false, hier_dirty = false;
    std::vector<Level> levels;
    uint32_t *d_level_slots = nullptr;
    size_t cap_level_slots = 0;
    // breadth-first forests: per-CTA descendant ranges ("cones", k_hier_cones)
    uint2 *d_cones = nullptr;
    size_t cap_cones = 0;
    uint32_t n_cones = 0;                // 0: the hierarchy is not breadth-first ordered (or ASTRE_NO_CONES): per-level launches
    bool use_cones = true;

    std::vector<uint64_t> h_entity_id;

    // cached read-backs of the last frame
    bool frame_done = false, results_cached = false, results_cached_total_valid = false;
    uint32_t last_flags = 0;
    bool last_sorted = false;             // the last frame's keys went through the sort
    bool sort_pending = false;            // frame ran with ASTRE_FRAME_SORT_LATER: astre_gpu_sort still to come
    SortPlan pending_plan{};
    BinPlan pending_bins{};
    std::vector<uint32_t> h_counts;
    std::vector<unsigned long long> h_offsets;
    unsigned long long h_total = 0;
    uint32_t h_overflow = 0, h_result_buf = 0;

    cudaEvent_t ev_begin = nullptr, ev_cull0 = nullptr, ev_cull1 = nullptr, ev_end = nullptr;
    cudaEvent_t ev_views = nullptr;      // recorded after the last upload of view records
    cudaStream_t last_stream = nullptr;
    uint64_t launches = 0;
    int fused_blocks_per_sm = 2;
    int cone_blocks_per_sm = 2;
    uint32_t cone_prefetch = 1;      // ASTRE_CONE_PREFETCH=0: the cone kernel does not prefetch its inputs into L2 (measurement)
    bool coop_launch = false;        // device supports cooperative launches (multi-level hierarchy launches)
    bool use_pdl = true;             // programmatic dependent launch for the sort chain (ASTRE_NO_PDL disables)
    uint32_t small_flat = 262144;    // flat frames below this many entities take the thread-per-entity kernel too (ASTRE_SMALL_FLAT; measured: 10k 19.4 -> 9.7 us, 100k 20.6 -> 13.1, 250k 24.4 -> 21.0)
    uint32_t small_level = 32768;    // hierarchy levels below this many entities take the thread-per-entity kernel (measured: config 2, 0.159 -> 0.139 ms)
    ViewSet h_viewset;               // single-world view records, passed as a kernel parameter
    std::string err;

    size_t tile_words() const { return (size_t)(cap_entities / kWarpTile) * kTileFloats; }
};

namespace {

int fail(astre_gpu_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, ASTRE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

#define CK_LAUNCH()                                                                                \
    do {                                                                                           \
        ++ctx->launches;                                                                           \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, ASTRE_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

inline uint32_t ceil_div(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

FrameParams params_of(const astre_gpu_ctx *c, bool cull)
{
    FrameParams P;
    P.tiles = c->d_tiles;
    P.world = c->d_world;
    P.bounds = c->d_bounds;
    P.views = c->d_views;
    P.keys = c->d_keys[0];
    P.counts = c->d_counts;
    P.ctl = c->d_ctl;
    P.bins = c->d_bins;
    P.claim_groups = c->claim_groups;
    P.key_cap = c->max_keys;
    P.n_meshes = c->n_meshes;
    P.views_per_world = cull ? c->views_per_world : 0u;
    P.n_worlds = c->n_worlds;
    P.entities_per_world = c->n_worlds > 1 ? c->entities_per_world : 0u;
    P.world_bits = c->world_bits;
    P.world_cache = (c->n_worlds > 1 && c->entities_per_world % kWarpTile == 0 && c->views_per_world <= (uint32_t)kWorldCacheViews) ? 1u : 0u;
    return P;
}

constexpr uint32_t kEpochWrap = 1u << 27;    // look-back tags hold 27 bits of frame number (x 8 passes + 1 in 30 bits)

__global__ void k_frame_begin(FrameControl *ctl, SortControl *sc, uint32_t *counts, uint32_t n_counts,
                              uint32_t *bins, uint32_t n_bins)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (tid == 0) {
        ctl->key_total = 0; ctl->overflow = 0; ctl->barrier = 0;
        sc->n_keys = 0;
        const uint32_t e = sc->epoch + 1u;             // the host keeps a mirror and clears the tagged words at the wrap
        sc->epoch = e >= kEpochWrap ? 1u : e;
        sc->max_bucket = 0; sc->scan_ticket = 0; sc->sum_sq = 0; sc->lsd_barrier = 0; sc->result_buf = 0; sc->path = kSortNone;
    }
    if (tid < kSortPasses) { sc->skip[tid] = 1; sc->tile_counter[tid] = 0; }
    for (uint32_t i = tid; i < (uint32_t)(kSortPasses * 256); i += nth) (&sc->hist[0][0])[i] = 0;
    for (uint32_t i = tid; i < (uint32_t)(kClaimGroups * kClaimStride); i += nth) ctl->claim[i] = 0;
    for (uint32_t i = tid; i < n_counts; i += nth) counts[i] = 0;
    uint4 *bins4 = reinterpret_cast<uint4 *>(bins);
    for (uint32_t i = tid; i < (n_bins >> 2); i += nth) bins4[i] = make_uint4(0u, 0u, 0u, 0u);
}

// Depth of every slot; returns the level count or -1 (bad index / cycle).
int compute_depths(const std::vector<uint32_t> &parent, uint32_t n, std::vector<uint32_t> &depth)
{
    const uint32_t UNSET = 0xFFFFFFFFu;
    depth.assign(n, UNSET);
    int levels = 0;
    std::vector<uint32_t> chain;
    for (uint32_t i = 0; i < n; ++i) {
        if (depth[i] != UNSET) continue;
        chain.clear();
        uint32_t cur = i;
        while (depth[cur] == UNSET) {
            chain.push_back(cur);
            const uint32_t p = parent[cur];
            if (p == ASTRE_NO_PARENT) break;
            if (p >= n || chain.size() > n) return -1;
            cur = p;
        }
        uint32_t d;
        if (depth[cur] != UNSET) d = depth[cur] + 1;       // stopped at a known ancestor
        else { d = 0; }                                      // stopped at a root (last pushed)
        // chain is child ... ancestor-most; assign from the top down
        for (size_t k = chain.size(); k-- > 0;) {
            const uint32_t s = chain[k];
            if (depth[s] != UNSET) continue;
            depth[s] = d++;
        }
    }
    for (uint32_t i = 0; i < n; ++i) levels = std::max(levels, (int)depth[i] + 1);
    return levels;
}

int rebuild_levels(astre_gpu_ctx *ctx)
{
    const uint32_t n = ctx->n_entities;
    ctx->levels.clear();
    ctx->has_parents = false;
    ++ctx->levels_version;
    for (uint32_t i = 0; i < n; ++i)
        if (ctx->h_parent[i] != ASTRE_NO_PARENT) { ctx->has_parents = true; break; }
    ctx->hier_dirty = false;
    if (!ctx->has_parents) return ASTRE_OK;

    std::vector<uint32_t> depth;
    const int n_levels = compute_depths(ctx->h_parent, n, depth);
    if (n_levels < 0) {
        ctx->hier_dirty = true;
        return fail(ctx, ASTRE_ERR_HIERARCHY, "parent index out of range or cyclic parent links");
    }
    // counting sort of slots by depth (stable: slots ascend inside a level)
    std::vector<size_t> start(n_levels + 1, 0);
    for (uint32_t i = 0; i < n; ++i) ++start[depth[i] + 1];
    for (int l = 0; l < n_levels; ++l) start[l + 1] += start[l];
    std::vector<uint32_t> order(n);
    {
        std::vector<size_t> fill(start.begin(), start.end() - 1);
        for (uint32_t i = 0; i < n; ++i) order[fill[depth[i]]++] = i;
    }
    ctx->levels.resize(n_levels);
    bool any_list = false;
    for (int l = 0; l < n_levels; ++l) {
        Level &L = ctx->levels[l];
        L.count = (uint32_t)(start[l + 1] - start[l]);
        L.first = order[start[l]];
        L.contiguous = (order[start[l + 1] - 1] - L.first + 1 == L.count);
        L.list_offset = start[l];
        any_list |= !L.contiguous;
    }
    // Cones: every level a contiguous range in slot order and parents non-decreasing along each level => the
    // descendants of a run of roots are one contiguous range per level.
    ctx->n_cones = 0;
    if (ctx->use_cones && !any_list && n_levels > 1 && n_levels <= 64) {
        bool bfs = true;
        for (int l = 1; l < n_levels && bfs; ++l) {
            const Level &L = ctx->levels[l], &U = ctx->levels[l - 1];
            uint32_t prev = 0;
            for (uint32_t i = 0; i < L.count; ++i) {
                const uint32_t p = ctx->h_parent[L.first + i];
                if (p < U.first || p >= U.first + U.count || p < prev) { bfs = false; break; }
                prev = p;
            }
        }
        if (bfs) {
            // descendants per root, bottom-up
            std::vector<uint32_t> sub(n, 1u);
            for (int l = n_levels - 1; l >= 1; --l) {
                const Level &L = ctx->levels[l];
                for (uint32_t i = 0; i < L.count; ++i) sub[ctx->h_parent[L.first + i]] += sub[L.first + i];
            }
            const Level &R = ctx->levels[0];
            const uint32_t want = std::min<uint32_t>((uint32_t)(ctx->sm_count * ctx->cone_blocks_per_sm), std::max<uint32_t>(R.count, 1u));
            // cut the roots into `want` runs of about n / want entities each
            std::vector<uint32_t> cut(want + 1, R.first + R.count);
            {
                uint64_t acc = 0;
                uint32_t c = 0;
                cut[0] = R.first;
                for (uint32_t i = 0; i < R.count && c + 1 < want; ++i) {
                    acc += sub[R.first + i];
                    if (acc * want >= (uint64_t)n * (c + 1)) cut[++c] = R.first + i + 1;
                }
                for (uint32_t k = c + 1; k <= want; ++k) cut[k] = R.first + R.count;
            }
            std::vector<uint2> ranges((size_t)want * n_levels);
            for (uint32_t c = 0; c < want; ++c) ranges[(size_t)c * n_levels] = make_uint2(cut[c], cut[c + 1]);
            for (int l = 1; l < n_levels; ++l) {
                const Level &L = ctx->levels[l];
                const uint32_t *par = ctx->h_parent.data() + L.first;
                for (uint32_t c = 0; c < want; ++c) {
                    const uint2 up = ranges[(size_t)c * n_levels + l - 1];         // parents of this cone on the level above
                    const uint32_t lo = (uint32_t)(std::lower_bound(par, par + L.count, up.x) - par);
                    const uint32_t hi = (uint32_t)(std::lower_bound(par, par + L.count, up.y) - par);
                    ranges[(size_t)c * n_levels + l] = up.x < up.y ? make_uint2(L.first + lo, L.first + hi) : make_uint2(L.first, L.first);
                }
            }
            if (ctx->cap_cones < ranges.size()) {
                if (ctx->d_cones) dev_free(ctx->d_cones);
                ctx->d_cones = nullptr;
                CK(dev_alloc(&ctx->d_cones, ranges.size() * sizeof(uint2)));
                ctx->cap_cones = ranges.size();
            }
            CK(cudaMemcpyAsync(ctx->d_cones, ranges.data(), ranges.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            ctx->n_cones = want;
        }
    }
    if (any_list) {
        if (ctx->cap_level_slots < n) {
            if (ctx->d_level_slots) dev_free(ctx->d_level_slots);
            ctx->d_level_slots = nullptr;
            CK(dev_alloc(&ctx->d_level_slots, (size_t)ctx->cap_entities * sizeof(uint32_t)));
            ctx->cap_level_slots = ctx->cap_entities;
        }
        CK(cudaMemcpyAsync(ctx->d_level_slots, order.data(), (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice,
                           ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return ASTRE_OK;
}

int cache_results(astre_gpu_ctx *ctx)
{
    if (!ctx->frame_done) return fail(ctx, ASTRE_ERR_STATE, "no frame has been run");
    if (ctx->results_cached) return ASTRE_OK;
    cudaStream_t s = ctx->last_stream;
    const uint32_t n_counts = ctx->n_worlds * ctx->views_per_world;
    CK(cudaStreamSynchronize(s));
    // the frame's last launch left everything in mapped host memory (counts and offsets too, unless there are many)
    const FrameResultHost &R = *ctx->h_res;
    ctx->h_counts.assign(n_counts, 0);
    ctx->h_offsets.assign((size_t)n_counts + 1, 0);
    if (n_counts) {
        if (R.n_counts == n_counts) {
            std::memcpy(ctx->h_counts.data(), R.counts, (size_t)n_counts * sizeof(uint32_t));
            std::memcpy(ctx->h_offsets.data(), R.offsets, ((size_t)n_counts + 1) * sizeof(unsigned long long));
        } else {
            CK(cudaMemcpy(ctx->h_counts.data(), ctx->d_counts, (size_t)n_counts * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(ctx->h_offsets.data(), ctx->d_offsets, ((size_t)n_counts + 1) * sizeof(unsigned long long),
                          cudaMemcpyDeviceToHost));
        }
    }
    ctx->h_total = n_counts ? R.key_total : 0ull;        // a frame without views emits no keys and publishes nothing
    ctx->h_overflow = n_counts ? ((R.key_total > ctx->max_keys) ? 1u : R.overflow) : 0u;
    ctx->h_result_buf = ctx->last_sorted ? R.result_buf : 0u;
    ctx->results_cached = true;
    ctx->results_cached_total_valid = true;
    return ASTRE_OK;
}

int ensure_stage(astre_gpu_ctx *ctx)
{
    if (ctx->d_stage) return ASTRE_OK;
    ctx->stage_entities = std::min<uint32_t>(std::max<uint32_t>(ctx->max_entities, 1u), 1u << 19);
    CK(dev_alloc(&ctx->d_stage, (size_t)ctx->stage_entities * 64 * 2));
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CK(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_ingested[i], cudaEventDisableTiming));
    }
    return ASTRE_OK;
}

}  // namespace

// ---------------------------------------------------------------------------

extern "C" {

const char *astre_gpu_version(void) { return "astre-b200 0.1 (sm_100a)"; }

const char *astre_gpu_last_error(const astre_gpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

astre_gpu_ctx *astre_gpu_create(int device, uint32_t max_entities, uint32_t max_views, uint64_t max_keys)
{
    astre_gpu_ctx *none = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        fail(none, ASTRE_ERR_CUDA, "no CUDA device available (%s); libastre_gpu has no CPU fallback",
             e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return nullptr;
    }
    if (device < 0 || device >= n_dev) { fail(none, ASTRE_ERR_INVALID, "device %d out of range (0..%d)", device, n_dev - 1); return nullptr; }
    if (max_entities == 0 || max_entities > (1u << ASTRE_KEY_SLOT_BITS)) {
        fail(none, ASTRE_ERR_INVALID, "max_entities must be in 1..2^26 (slot field of the draw key)");
        return nullptr;
    }
    if (max_views > ASTRE_MAX_VIEWS) { fail(none, ASTRE_ERR_INVALID, "max_views must be <= %d", ASTRE_MAX_VIEWS); return nullptr; }
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        fail(none, ASTRE_ERR_CUDA, "cannot select device %d", device);
        return nullptr;
    }
    if (prop.major < 10) {
        fail(none, ASTRE_ERR_CUDA, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
        return nullptr;
    }
    astre_gpu_ctx *ctx = new (std::nothrow) astre_gpu_ctx();
    if (!ctx) { fail(none, ASTRE_ERR_INVALID, "out of host memory"); return nullptr; }
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_entities = max_entities;
    ctx->cap_entities = (uint32_t)((((uint64_t)max_entities + kBlockTile - 1) / kBlockTile) * kBlockTile + kBlockTile);
    ctx->max_views = max_views;
    ctx->max_keys = max_keys ? max_keys : (uint64_t)max_entities * std::max<uint32_t>(max_views, 1u);
    if (ctx->max_keys > 0xFFFFFFF0ull) ctx->max_keys = 0xFFFFFFF0ull;

    auto bail = [&](const char *what, cudaError_t err) -> astre_gpu_ctx * {
        fail(none, ASTRE_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(err));
        astre_gpu_destroy(ctx);
        return nullptr;
    };
#define CKC(call) do { cudaError_t e2_ = (call); if (e2_ != cudaSuccess) return bail(#call, e2_); } while (0)
    const size_t cap = ctx->cap_entities;
    CKC(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    CKC(dev_alloc(&ctx->d_tiles, ctx->tile_words() * sizeof(float)));
    CKC(cudaMemset(ctx->d_tiles, 0, ctx->tile_words() * sizeof(float)));
    k_init_tiles<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->d_tiles, ctx->cap_entities);
    CKC(cudaGetLastError());
    CKC(dev_alloc(&ctx->d_world, cap * 4 * sizeof(float4)));
    ctx->cap_meshes = 16;
    CKC(dev_alloc(&ctx->d_bounds, ctx->cap_meshes * sizeof(BoundsDev)));
    CKC(cudaMemset(ctx->d_bounds, 0, ctx->cap_meshes * sizeof(BoundsDev)));
    ctx->cap_view_records = std::max<uint32_t>(max_views, 1u);
    CKC(dev_alloc(&ctx->d_views, ctx->cap_view_records * sizeof(ViewDev)));
    CKC(cudaMemset(ctx->d_views, 0, ctx->cap_view_records * sizeof(ViewDev)));
    CKC(dev_alloc(&ctx->d_keys[0], ctx->max_keys * sizeof(unsigned long long)));
    CKC(dev_alloc(&ctx->d_keys[1], ctx->max_keys * sizeof(unsigned long long)));
    ctx->cap_counts = std::max<uint32_t>(max_views, 1u);
    CKC(dev_alloc(&ctx->d_counts, ctx->cap_counts * sizeof(uint32_t)));
    CKC(dev_alloc(&ctx->d_offsets, ((size_t)ctx->cap_counts + 1) * sizeof(unsigned long long)));
    CKC(dev_alloc(&ctx->d_ctl, sizeof(FrameControl)));
    CKC(cudaMemset(ctx->d_ctl, 0, sizeof(FrameControl)));
    CKC(dev_alloc(&ctx->d_sc, sizeof(SortControl)));
    CKC(cudaMemset(ctx->d_sc, 0, sizeof(SortControl)));
    const size_t n_sort_tiles = (size_t)((ctx->max_keys + 2048 - 1) / 2048) + 1;
    CKC(dev_alloc(&ctx->d_lookback, n_sort_tiles * 256 * sizeof(unsigned long long)));
    CKC(cudaMemset(ctx->d_lookback, 0, n_sort_tiles * 256 * sizeof(unsigned long long)));
    if (const char *m = std::getenv("ASTRE_CLAIM_GROUPS")) ctx->claim_groups = (uint32_t)std::max(std::atoi(m), 0);
    // key sort: buckets of ~4-8 keys -> up to max_keys / 4 bins (between 2^8 and 2^22)
    ctx->cap_bin_bits = kMinBinBits;
    while (ctx->cap_bin_bits < kMaxBinBits && (1ull << ctx->cap_bin_bits) * 4ull < ctx->max_keys) ++ctx->cap_bin_bits;
    CKC(dev_alloc(&ctx->d_bins, (size_t)(1u << ctx->cap_bin_bits) * sizeof(uint32_t)));
    CKC(dev_alloc(&ctx->d_chunk_first, (size_t)(ctx->max_keys / kChunkKeys + 2) * sizeof(uint32_t)));
    CKC(cudaHostAlloc(&ctx->h_res, sizeof(FrameResultHost), cudaHostAllocMapped));
    std::memset(ctx->h_res, 0, sizeof(FrameResultHost));
    ctx->h_res->key_total = ~0ull;                                   // no frame has finished yet
    CKC(cudaHostGetDevicePointer(&ctx->d_res, ctx->h_res, 0));
    if (std::getenv("ASTRE_SORT_FORCE_LSD")) ctx->force_lsd = 1u;
    if (std::getenv("ASTRE_SORT_NO_FEEDBACK")) ctx->bits_feedback = false;
    if (const char *m = std::getenv("ASTRE_SORT_BIN_BITS")) ctx->bin_bits_override = std::atoi(m);
    {
        int occ_sort = 0;
        CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_sort, k_sort_lsd_all, kSortThreads, 0));
        // the fallback has grid barriers, so its CTAs must all be resident at once -- also while a small kernel of
        // another stream (a count-board receive, a collective) holds a few CTA slots until a peer GPU answers
        ctx->sort_grid = std::max(ctx->sm_count * std::max(occ_sort, 1) - 8, 1);
    }
    CKC(cudaEventCreate(&ctx->ev_begin));
    CKC(cudaEventCreate(&ctx->ev_cull0));
    CKC(cudaEventCreate(&ctx->ev_cull1));
    CKC(cudaEventCreate(&ctx->ev_end));
    int occ = 0;
    CKC(cudaFuncSetAttribute(k_frame_fused<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemBytes));
    CKC(cudaFuncSetAttribute(k_frame_fused<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemBytes));
    CKC(cudaFuncSetAttribute(k_frame_fused<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemBytes));
    CKC(cudaFuncSetAttribute(k_frame_fused<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemBytes));
    CKC(cudaFuncSetAttribute(k_frame_fused<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemBytes));
    CKC(cudaFuncSetAttribute(k_frame_fused<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemBytes));
    {
        // the persistent grid must be co-resident for every instantiation (grid barrier of multi-level launches)
        const void *variants[] = { (const void *)k_frame_fused<false, false, true>, (const void *)k_frame_fused<false, true, true>,
                                   (const void *)k_frame_fused<true, false, true>, (const void *)k_frame_fused<true, true, true>,
                                   (const void *)k_frame_fused<false, false, false>, (const void *)k_frame_fused<true, false, false> };
        int min_occ = 2;
        for (const void *fn : variants) {
            CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kBlockThreads, kFusedSmemBytes));
            min_occ = std::min(min_occ, std::max(occ, 1));
        }
        ctx->fused_blocks_per_sm = min_occ;
        int coop = 0;
        CKC(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
        ctx->coop_launch = coop != 0 && !std::getenv("ASTRE_NO_COOP");
        if (const char *m = std::getenv("ASTRE_SMALL_LEVEL")) ctx->small_level = (uint32_t)std::atol(m);
        if (const char *m = std::getenv("ASTRE_SMALL_FLAT")) ctx->small_flat = (uint32_t)std::atol(m);
        if (std::getenv("ASTRE_NO_CONES")) ctx->use_cones = false;
        if (std::getenv("ASTRE_GRAPH_NO_UPDATE")) ctx->graph_update_ok = false;
        int occ_cones = 0;
        CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cones, k_hier_cones<true>, kBlockThreads, 0));
        ctx->cone_blocks_per_sm = std::max(occ_cones, 1);
        if (const char *m = std::getenv("ASTRE_CONES_PER_SM")) ctx->cone_blocks_per_sm = std::max(std::atoi(m), 1);
        if (const char *m = std::getenv("ASTRE_CONE_PREFETCH")) ctx->cone_prefetch = std::atoi(m) != 0;
        if (std::getenv("ASTRE_NO_PDL")) ctx->use_pdl = false;
    }
    std::memset(&ctx->h_viewset, 0, sizeof(ViewSet));
    CKC(cudaDeviceSynchronize());
#undef CKC
    ctx->h_parent.assign(max_entities, ASTRE_NO_PARENT);
    return ctx;
}

void astre_gpu_destroy(astre_gpu_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    void *ptrs[] = { ctx->d_tiles, ctx->d_prev, ctx->d_world, ctx->d_basis, ctx->d_bounds,
                     ctx->d_views, ctx->d_keys[0], ctx->d_keys[1], ctx->d_counts, ctx->d_offsets, ctx->d_ctl, ctx->d_sc,
                     ctx->d_lookback, ctx->d_gather, ctx->d_stage, ctx->d_level_slots, ctx->d_cones, ctx->d_bins, ctx->d_chunk_first, ctx->d_heads,
                     ctx->d_head_counts, ctx->d_mesh_info, ctx->d_indirect,
                     ctx->d_merged_keys, ctx->d_merged_src, ctx->d_splits, ctx->d_tile_start,
                     ctx->d_mcuts, ctx->d_mdims, ctx->d_merge_error };
    for (void *p : ptrs) if (p) dev_free(p);
    for (unsigned char *p : ctx->d_publish) if (p) cudaFree(p);
    if (ctx->d_board) cudaFree(ctx->d_board);
    if (ctx->d_board_table) dev_free(ctx->d_board_table);
    for (void *p : ctx->imports) cudaIpcCloseMemHandle(p);
    cudaEvent_t evs[] = { ctx->ev_begin, ctx->ev_cull0, ctx->ev_cull1, ctx->ev_end, ctx->ev_merge0, ctx->ev_merge1, ctx->ev_views };
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
        if (ctx->ev_ingested[i]) cudaEventDestroy(ctx->ev_ingested[i]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->h_res) cudaFreeHost(ctx->h_res);
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    if (ctx->graph_obj) cudaGraphDestroy(ctx->graph_obj);
    delete ctx;
}

int astre_gpu_set_entity_count(astre_gpu_ctx *ctx, uint32_t n)
{
    if (!ctx) return ASTRE_ERR_INVALID;
    if (n > ctx->max_entities) return fail(ctx, ASTRE_ERR_INVALID, "entity count %u exceeds max_entities %u", n, ctx->max_entities);
    if (n != ctx->n_entities) ctx->hier_dirty = true;
    ctx->n_entities = n;
    return ASTRE_OK;
}

int astre_gpu_upload_entities(astre_gpu_ctx *ctx, uint32_t first, uint32_t count, const uint64_t *entity_id,
                              const float *pos3, const float *rot4, const float *scl3, const uint32_t *parent_slot,
                              const uint32_t *mesh_id, const uint32_t *shader_id, const uint8_t *flags)
{
    if (!ctx) return ASTRE_ERR_INVALID;
    if ((uint64_t)first + count > ctx->max_entities)
        return fail(ctx, ASTRE_ERR_INVALID, "slots [%u, %u) exceed max_entities %u", first, first + count, ctx->max_entities);
    CK(cudaSetDevice(ctx->device));
    if (int rc = ensure_stage(ctx)) return rc;
    cudaStream_t s = ctx->stream;
    for (uint32_t done = 0; done < count;) {
        const uint32_t c = std::min(ctx->stage_entities, count - done);
        unsigned char *st = ctx->d_stage;
        float *d_pos = (float *)st;                       st += (size_t)c * 12;
        float *d_scl = (float *)st;                       st += (size_t)c * 12;
        st = ctx->d_stage + (((size_t)(st - ctx->d_stage) + 15) & ~(size_t)15);
        float *d_rot = (float *)st;                       st += (size_t)c * 16;
        uint32_t *d_mesh = (uint32_t *)st;                st += (size_t)c * 4;
        uint32_t *d_shader = (uint32_t *)st;              st += (size_t)c * 4;
        uint32_t *d_par = (uint32_t *)st;                 st += (size_t)c * 4;
        uint8_t *d_flags = (uint8_t *)st;
        if (pos3) CK(cudaMemcpyAsync(d_pos, pos3 + 3 * (size_t)done, (size_t)c * 12, cudaMemcpyHostToDevice, s));
        if (rot4) CK(cudaMemcpyAsync(d_rot, rot4 + 4 * (size_t)done, (size_t)c * 16, cudaMemcpyHostToDevice, s));
        if (scl3) CK(cudaMemcpyAsync(d_scl, scl3 + 3 * (size_t)done, (size_t)c * 12, cudaMemcpyHostToDevice, s));
        if (mesh_id) CK(cudaMemcpyAsync(d_mesh, mesh_id + done, (size_t)c * 4, cudaMemcpyHostToDevice, s));
        if (shader_id) CK(cudaMemcpyAsync(d_shader, shader_id + done, (size_t)c * 4, cudaMemcpyHostToDevice, s));
        if (parent_slot) CK(cudaMemcpyAsync(d_par, parent_slot + done, (size_t)c * 4, cudaMemcpyHostToDevice, s));
        if (flags) CK(cudaMemcpyAsync(d_flags, flags + done, (size_t)c, cudaMemcpyHostToDevice, s));
        const uint32_t grid = std::min<uint32_t>(ceil_div(c, 256), ctx->sm_count * 8);
        k_ingest_trs<<<grid, 256, 0, s>>>(ctx->d_tiles, nullptr, first + done, c, pos3 ? d_pos : nullptr, rot4 ? d_rot : nullptr,
                                          scl3 ? d_scl : nullptr, true);
        CK_LAUNCH();
        k_ingest_meta<<<grid, 256, 0, s>>>(ctx->d_tiles, first + done, c, mesh_id ? d_mesh : nullptr,
                                           shader_id ? d_shader : nullptr, flags ? d_flags : nullptr,
                                           parent_slot ? d_par : nullptr);
        CK_LAUNCH();
        done += c;
    }
    CK(cudaStreamSynchronize(s));   // host buffers may be reused by the caller
    for (uint32_t i = 0; i < count; ++i) ctx->h_parent[first + i] = parent_slot ? parent_slot[i] : ASTRE_NO_PARENT;
    if (shader_id) {
        uint32_t acc = 0;
        for (uint32_t i = 0; i < count; ++i) acc |= shader_id[i];
        ctx->shader_or |= acc & 255u;
    }
    if (entity_id) {
        if (ctx->h_entity_id.empty()) {
            ctx->h_entity_id.resize(ctx->max_entities);
            for (uint32_t i = 0; i < ctx->max_entities; ++i) ctx->h_entity_id[i] = i;
        }
        std::memcpy(ctx->h_entity_id.data() + first, entity_id, (size_t)count * sizeof(uint64_t));
    }
    ctx->hier_dirty = true;
    ctx->n_entities = std::max(ctx->n_entities, first + count);
    return ASTRE_OK;
}

static int update_trs_impl(astre_gpu_ctx *ctx, const uint32_t *slots, uint32_t first, uint32_t n,
                           const float *pos3, const float *rot4, const float *scl3)
{
    CK(cudaSetDevice(ctx->device));
    if (int rc = ensure_stage(ctx)) return rc;
    cudaStream_t s = ctx->stream, cs = ctx->copy_stream;
    // Double-buffered: the copies of chunk i+1 (copy stream) run while chunk i is scattered into the
    // tile records (main stream); a staging half is refilled only after its ingest kernel finished.
    int chunk = 0;
    for (uint32_t done = 0; done < n; ++chunk) {
        const int h = chunk & 1;
        const uint32_t c = std::min(ctx->stage_entities, n - done);
        unsigned char *base = ctx->d_stage + (size_t)h * ctx->stage_entities * 64;
        unsigned char *st = base;
        float *d_pos = (float *)st;                       st += (size_t)c * 12;
        float *d_scl = (float *)st;                       st += (size_t)c * 12;
        st = base + (((size_t)(st - base) + 15) & ~(size_t)15);
        float *d_rot = (float *)st;                       st += (size_t)c * 16;
        uint32_t *d_slots = (uint32_t *)st;
        if (chunk >= 2) CK(cudaStreamWaitEvent(cs, ctx->ev_ingested[h], 0));
        if (slots) CK(cudaMemcpyAsync(d_slots, slots + done, (size_t)c * 4, cudaMemcpyHostToDevice, cs));
        if (pos3) CK(cudaMemcpyAsync(d_pos, pos3 + 3 * (size_t)done, (size_t)c * 12, cudaMemcpyHostToDevice, cs));
        if (rot4) CK(cudaMemcpyAsync(d_rot, rot4 + 4 * (size_t)done, (size_t)c * 16, cudaMemcpyHostToDevice, cs));
        if (scl3) CK(cudaMemcpyAsync(d_scl, scl3 + 3 * (size_t)done, (size_t)c * 12, cudaMemcpyHostToDevice, cs));
        CK(cudaEventRecord(ctx->ev_copied[h], cs));
        CK(cudaStreamWaitEvent(s, ctx->ev_copied[h], 0));
        const uint32_t grid = std::min<uint32_t>(ceil_div(c, 256), ctx->sm_count * 8);
        k_ingest_trs<<<grid, 256, 0, s>>>(ctx->d_tiles, slots ? d_slots : nullptr, first + done, c, pos3 ? d_pos : nullptr,
                                          rot4 ? d_rot : nullptr, scl3 ? d_scl : nullptr, false);
        CK_LAUNCH();
        CK(cudaEventRecord(ctx->ev_ingested[h], s));
        done += c;
    }
    CK(cudaStreamSynchronize(s));   // host buffers may be reused by the caller; staging halves are idle again
    return ASTRE_OK;
}

int astre_gpu_update_trs(astre_gpu_ctx *ctx, uint32_t n, const uint32_t *slots, const float *pos3, const float *rot4,
                         const float *scl3)
{
    if (!ctx || (n && !slots)) return ctx ? fail(ctx, ASTRE_ERR_INVALID, "slots is null") : ASTRE_ERR_INVALID;
    for (uint32_t i = 0; i < n; ++i)
        if (slots[i] >= ctx->max_entities) return fail(ctx, ASTRE_ERR_INVALID, "slot %u out of range", slots[i]);
    return update_trs_impl(ctx, slots, 0, n, pos3, rot4, scl3);
}

int astre_gpu_update_trs_range(astre_gpu_ctx *ctx, uint32_t first, uint32_t count, const float *pos3, const float *rot4,
                               const float *scl3)
{
    if (!ctx) return ASTRE_ERR_INVALID;
    if ((uint64_t)first + count > ctx->max_entities) return fail(ctx, ASTRE_ERR_INVALID, "slot range out of bounds");
    return update_trs_impl(ctx, nullptr, first, count, pos3, rot4, scl3);
}

int astre_gpu_update_flags(astre_gpu_ctx *ctx, uint32_t n, const uint32_t *slots, const uint8_t *flags)
{
    if (!ctx || (n && (!slots || !flags))) return ctx ? fail(ctx, ASTRE_ERR_INVALID, "null argument") : ASTRE_ERR_INVALID;
    for (uint32_t i = 0; i < n; ++i)
        if (slots[i] >= ctx->max_entities) return fail(ctx, ASTRE_ERR_INVALID, "slot %u out of range", slots[i]);
    CK(cudaSetDevice(ctx->device));
    if (int rc = ensure_stage(ctx)) return rc;
    cudaStream_t s = ctx->stream;
    for (uint32_t done = 0; done < n;) {
        const uint32_t c = std::min(ctx->stage_entities, n - done);
        uint32_t *d_slots = (uint32_t *)ctx->d_stage;
        uint8_t *d_flags = ctx->d_stage + (size_t)c * 4;
        CK(cudaMemcpyAsync(d_slots, slots + done, (size_t)c * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d_flags, flags + done, (size_t)c, cudaMemcpyHostToDevice, s));
        k_update_flags<<<std::min<uint32_t>(ceil_div(c, 256), ctx->sm_count * 8), 256, 0, s>>>(ctx->d_tiles, d_slots, d_flags, c);
        CK_LAUNCH();
        done += c;
    }
    CK(cudaStreamSynchronize(s));
    return ASTRE_OK;
}

int astre_gpu_set_mesh_bounds(astre_gpu_ctx *ctx, uint32_t n_meshes, const astre_mesh_bounds *bounds)
{
    if (!ctx || (n_meshes && !bounds)) return ctx ? fail(ctx, ASTRE_ERR_INVALID, "bounds is null") : ASTRE_ERR_INVALID;
    if (n_meshes > (1u << ASTRE_KEY_MESH_BITS)) return fail(ctx, ASTRE_ERR_INVALID, "at most 2048 meshes (mesh field of the draw key)");
    CK(cudaSetDevice(ctx->device));
    if (n_meshes > ctx->cap_meshes) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(dev_free(ctx->d_bounds));
        ctx->d_bounds = nullptr;
        CK(dev_alloc(&ctx->d_bounds, (size_t)n_meshes * sizeof(BoundsDev)));
        ctx->cap_meshes = n_meshes;
    }
    if (n_meshes) {
        CK(cudaMemcpyAsync(ctx->d_bounds, bounds, (size_t)n_meshes * sizeof(BoundsDev), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->n_meshes = n_meshes;
    ++ctx->params_version;
    return ASTRE_OK;
}

int astre_gpu_set_world_views(astre_gpu_ctx *ctx, uint32_t n_worlds, uint32_t entities_per_world, uint32_t views_per_world,
                              const float *clip16, const uint8_t *phase_mask)
{
    if (!ctx) return ASTRE_ERR_INVALID;
    if (n_worlds == 0) return fail(ctx, ASTRE_ERR_INVALID, "n_worlds must be >= 1");
    if (views_per_world > ctx->max_views) return fail(ctx, ASTRE_ERR_INVALID, "%u views exceed max_views %u", views_per_world, ctx->max_views);
    if (views_per_world && (!clip16 || !phase_mask)) return fail(ctx, ASTRE_ERR_INVALID, "null view arrays");
    uint32_t wbits = 0;
    while ((1u << wbits) < n_worlds) ++wbits;
    if (n_worlds > 1) {
        if (entities_per_world == 0 || entities_per_world % 4)
            return fail(ctx, ASTRE_ERR_INVALID, "entities_per_world must be a positive multiple of 4");
        if ((uint64_t)n_worlds * entities_per_world > ctx->max_entities)
            return fail(ctx, ASTRE_ERR_INVALID, "n_worlds * entities_per_world exceeds max_entities");
        if (wbits > 20 || entities_per_world > (1u << (ASTRE_KEY_SLOT_BITS - wbits)))
            return fail(ctx, ASTRE_ERR_INVALID, "world index and local slot do not fit the 26-bit key field");
    }
    CK(cudaSetDevice(ctx->device));
    const uint64_t n_rec = (uint64_t)n_worlds * views_per_world;
    if (n_rec > 0xFFFFFFFFull) return fail(ctx, ASTRE_ERR_INVALID, "too many view records");
    const bool reallocated = n_rec > ctx->cap_view_records || n_rec > ctx->cap_counts;
    if (ctx->last_stream && ctx->last_stream != ctx->stream) CK(cudaStreamSynchronize(ctx->last_stream));   // frames enqueued on a caller's stream
    if (reallocated) CK(cudaStreamSynchronize(ctx->stream));
    if (n_rec > ctx->cap_view_records) {
        CK(dev_free(ctx->d_views)); ctx->d_views = nullptr;
        CK(dev_alloc(&ctx->d_views, n_rec * sizeof(ViewDev)));
        ctx->cap_view_records = (uint32_t)n_rec;
    }
    if (n_rec > ctx->cap_counts) {
        CK(dev_free(ctx->d_counts)); ctx->d_counts = nullptr;
        CK(dev_free(ctx->d_offsets)); ctx->d_offsets = nullptr;
        CK(dev_alloc(&ctx->d_counts, n_rec * sizeof(uint32_t)));
        CK(dev_alloc(&ctx->d_offsets, (n_rec + 1) * sizeof(unsigned long long)));
        ctx->cap_counts = (uint32_t)n_rec;
    }
    std::vector<astre_host::ViewHost> recs(n_rec);
    for (uint64_t i = 0; i < n_rec; ++i)
        astre_host::extract_view(clip16 + 16 * i, phase_mask[i % views_per_world] & 15u, &recs[i]);
    // stream-ordered after the frames that still read the old records; `recs` is pageable, so the runtime stages it
    // before returning (no synchronisation: a camera that moves every frame must not stall the pipeline)
    if (n_rec) {
        CK(cudaMemcpyAsync(ctx->d_views, recs.data(), n_rec * sizeof(ViewDev), cudaMemcpyHostToDevice, ctx->stream));
        if (!ctx->ev_views) CK(cudaEventCreateWithFlags(&ctx->ev_views, cudaEventDisableTiming));
        CK(cudaEventRecord(ctx->ev_views, ctx->stream));        // a frame on another stream waits for it
    }
    std::memset(&ctx->h_viewset, 0, sizeof(ViewSet));
    if (n_worlds == 1 && views_per_world) std::memcpy(ctx->h_viewset.v, recs.data(), views_per_world * sizeof(ViewDev));
    // a captured frame graph: new record CONTENT is patched in / read from global memory; anything else is captured again
    ++ctx->views_version;
    if (reallocated || ctx->n_worlds != n_worlds || ctx->views_per_world != views_per_world ||
        ctx->entities_per_world != (n_worlds > 1 ? entities_per_world : 0u)) ++ctx->params_version;
    ctx->n_worlds = n_worlds;
    ctx->entities_per_world = n_worlds > 1 ? entities_per_world : 0;
    ctx->views_per_world = views_per_world;
    ctx->world_bits = wbits;
    return ASTRE_OK;
}

int astre_gpu_set_views(astre_gpu_ctx *ctx, uint32_t n_views, const float *clip16, const uint8_t *phase_mask)
{
    return astre_gpu_set_world_views(ctx, 1, 0, n_views, clip16, phase_mask);
}

// The sort of the frame's keys: bin scan -> scatter -> rank, then the LSD fallback, which returns at once unless
// the scan found the buckets too crowded (all decided on the device).  Programmatic dependent launch: each kernel of
// the chain is made resident while its predecessor drains and blocks in cudaGridDependencySynchronize().
static int launch_sort(astre_gpu_ctx *ctx, const BinPlan &bp, const SortPlan &plan, uint32_t n_counts, cudaStream_t s)
{
    if (!bp.bits) return ASTRE_OK;                 // no key bit can differ: at most one key
    cudaLaunchAttribute pdl;
    pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl.val.programmaticStreamSerializationAllowed = ctx->use_pdl ? 1 : 0;
    cudaLaunchConfig_t cfg{};
    cfg.stream = s; cfg.attrs = &pdl; cfg.numAttrs = 1; cfg.dynamicSmemBytes = 0;
    const unsigned long long cap = ctx->max_keys;
    const uint32_t n_bins = 1u << bp.bits;
    const FrameControl *ctl = ctx->d_ctl;
    const uint32_t wide = (uint32_t)ctx->sm_count * 8u;
    cfg.gridDim = dim3(std::max(n_bins / kScanTile, 1u)); cfg.blockDim = dim3(1024);
    CK(cudaLaunchKernelEx(&cfg, k_bin_scan, ctx->d_bins, n_bins, ctx->d_sc, ctx->d_chunk_first));
    CK_LAUNCH();
    cfg.gridDim = dim3(wide); cfg.blockDim = dim3(256);
    CK(cudaLaunchKernelEx(&cfg, k_bin_scatter, (const unsigned long long *)ctx->d_keys[0], ctx->d_keys[1], ctx->d_bins, bp,
                          (const SortControl *)ctx->d_sc, ctl, cap, ctx->force_lsd));
    CK_LAUNCH();
    CK(cudaLaunchKernelEx(&cfg, k_bin_rank, (const unsigned long long *)ctx->d_keys[1], ctx->d_keys[0], (const uint32_t *)ctx->d_bins,
                          (const uint32_t *)ctx->d_chunk_first, bp, ctx->d_sc, ctl, cap, ctx->force_lsd));
    CK_LAUNCH();
    cfg.gridDim = dim3((unsigned)ctx->sort_grid); cfg.blockDim = dim3(kSortThreads);
    CK(cudaLaunchKernelEx(&cfg, k_sort_lsd_all, ctx->d_keys[0], ctx->d_keys[1], ctx->d_sc, ctx->d_lookback, plan, ctl, cap,
                          ctx->force_lsd, (const uint32_t *)ctx->d_counts, ctx->d_offsets, n_counts, ctx->d_res, ctx->counts_mirror));
    CK_LAUNCH();
    return ASTRE_OK;
}

namespace {

// Every launch of one frame on stream s (also what a CUDA graph of the frame captures: no allocation, no
// synchronisation, no event in here unless `timing`).
int enqueue_frame(astre_gpu_ctx *ctx, const FramePlan &fp, cudaStream_t s, bool timing)
{
    const uint32_t stage_flags = fp.stage_flags, n = fp.n, n_counts = fp.n_counts;
    const bool cull = fp.cull, sort_now = fp.sort_now;
    const SortPlan &plan = fp.plan;
    const BinPlan &bp = fp.bp;
    const uint32_t n_bins = bp.bits ? (1u << bp.bits) : 0u;
    k_frame_begin<<<std::max<uint32_t>(64u, std::min<uint32_t>(n_bins / 4096u, (uint32_t)ctx->sm_count * 4u)), 256, 0, s>>>(
        ctx->d_ctl, ctx->d_sc, ctx->d_counts, n_counts, ctx->d_bins, n_bins);
    CK_LAUNCH();

    const FrameParams P = params_of(ctx, cull);
    if (timing) CK(cudaEventRecord(ctx->ev_cull0, s));
    if (n) {
        const bool worlds = ctx->n_worlds > 1;
        const uint32_t fused_grid = (uint32_t)(ctx->sm_count * ctx->fused_blocks_per_sm);
        // one launch over a list of slot ranges; hier = world[parent] * local (ranges after the first wait for
        // the previous one through a grid barrier, hence the cooperative launch)
        uint32_t barrier_rounds = 0, claim_sets = 0;
        cudaError_t launch_err = cudaSuccess;
        auto launch_ranges = [&](const LevelTable &LT, bool hier) {
            uint32_t tiles = 0;
            for (uint32_t l = 0; l < LT.n; ++l) {
                const uint32_t base0 = LT.first[l] & ~(uint32_t)(kWarpTile - 1);
                tiles = std::max(tiles, ceil_div((uint64_t)LT.end[l] - base0, kBlockTile));
            }
            const uint32_t g = LT.n > 1 ? fused_grid : std::min(tiles, fused_grid);
            const void *fn;
            if (!cull) fn = hier ? (const void *)k_frame_fused<true, false, false> : (const void *)k_frame_fused<false, false, false>;
            else if (hier) fn = worlds ? (const void *)k_frame_fused<true, true, true> : (const void *)k_frame_fused<true, false, true>;
            else fn = worlds ? (const void *)k_frame_fused<false, true, true> : (const void *)k_frame_fused<false, false, true>;
            void *args[] = { (void *)&P, (void *)&ctx->h_viewset, (void *)&bp, (void *)&LT };
            if (ctx->capturing && cull && !worlds) {          // this launch reads the view records from its parameters
                ++ctx->graph_vs_launches;
                ctx->graph_vs_args.P = P; ctx->graph_vs_args.bp = bp; ctx->graph_vs_args.LT = LT;
                ctx->graph_vs_args.fn = fn; ctx->graph_vs_args.grid = g;
            }
            if (LT.n > 1) {
                launch_err = cudaLaunchCooperativeKernel(fn, dim3(g), dim3(kBlockThreads), args, kFusedSmemBytes, s);
            } else {
                cudaLaunchAttribute pdl;
                pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
                pdl.val.programmaticStreamSerializationAllowed = ctx->use_pdl ? 1 : 0;
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(g); cfg.blockDim = dim3(kBlockThreads); cfg.dynamicSmemBytes = kFusedSmemBytes;
                cfg.stream = s; cfg.attrs = &pdl; cfg.numAttrs = 1;
                launch_err = cudaLaunchKernelExC(&cfg, fn, args);
            }
            barrier_rounds += (LT.n - 1) * g;
            claim_sets += LT.n * std::min<uint32_t>(ctx->claim_groups, g);
        };
        if (ctx->has_parents && ctx->n_cones) {
            // breadth-first forest: one launch, a CTA per cone, levels separated by __syncthreads()
            const uint32_t n_levels = (uint32_t)ctx->levels.size();
            cudaLaunchAttribute pdl;
            pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
            pdl.val.programmaticStreamSerializationAllowed = ctx->use_pdl ? 1 : 0;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(ctx->n_cones); cfg.blockDim = dim3(kBlockThreads); cfg.stream = s; cfg.attrs = &pdl; cfg.numAttrs = 1;
            if (cull) CK(cudaLaunchKernelEx(&cfg, k_hier_cones<true>, P, bp, (const uint2 *)ctx->d_cones, n_levels, ctx->cone_prefetch));
            else CK(cudaLaunchKernelEx(&cfg, k_hier_cones<false>, P, bp, (const uint2 *)ctx->d_cones, n_levels, ctx->cone_prefetch));
            CK_LAUNCH();
        } else if (!ctx->has_parents && n < ctx->small_flat) {
            // a frame this small is all latency: one entity per thread finishes in a third of the tiled kernel's first tile
            cudaLaunchAttribute pdl;
            pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
            pdl.val.programmaticStreamSerializationAllowed = ctx->use_pdl ? 1 : 0;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(std::min<uint32_t>(ceil_div(n, kBlockThreads), (uint32_t)ctx->sm_count * 8)); cfg.blockDim = dim3(kBlockThreads);
            cfg.stream = s; cfg.attrs = &pdl; cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, k_transform_cull_indexed, P, bp, (const uint32_t *)nullptr, 0u, n));
            CK_LAUNCH();
        } else if (!ctx->has_parents) {
            LevelTable LT{};
            LT.n = 1; LT.first[0] = 0; LT.end[0] = n;
            launch_ranges(LT, false);
            if (launch_err != cudaSuccess) return fail(ctx, ASTRE_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(launch_err));
            CK_LAUNCH();
        } else {
            // runs of contiguous levels go into one cooperative launch each; a level whose slots are not a
            // contiguous range takes the indexed kernel
            size_t l = 0;
            while (l < ctx->levels.size()) {
                // latency-bound levels (small, or not a contiguous range): thread-per-entity kernel
                if (!ctx->levels[l].contiguous || ctx->levels[l].count < ctx->small_level) {
                    const Level &L = ctx->levels[l];
                    const uint32_t grid = std::min<uint32_t>(ceil_div(L.count, kBlockThreads), (uint32_t)ctx->sm_count * 8);
                    k_transform_cull_indexed<<<grid, kBlockThreads, 0, s>>>(
                        P, bp, L.contiguous ? nullptr : ctx->d_level_slots + L.list_offset, L.first, L.count);
                    CK_LAUNCH();
                    ++l;
                    continue;
                }
                LevelTable LT{};
                LT.barrier_base = barrier_rounds;
                LT.claim_base = claim_sets;                   // the kernel falls back to per-CTA counters when the sets run out
                while (l < ctx->levels.size() && ctx->levels[l].contiguous && ctx->levels[l].count >= ctx->small_level &&
                       LT.n < (uint32_t)kMaxLevelsPerLaunch && (ctx->coop_launch || LT.n == 0)) {
                    LT.first[LT.n] = ctx->levels[l].first;
                    LT.end[LT.n] = ctx->levels[l].first + ctx->levels[l].count;
                    ++LT.n; ++l;
                }
                launch_ranges(LT, true);
                if (launch_err != cudaSuccess) return fail(ctx, ASTRE_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(launch_err));
                CK_LAUNCH();
            }
        }
    }
    if (timing) CK(cudaEventRecord(ctx->ev_cull1, s));

    if (stage_flags & ASTRE_FRAME_BASIS) {
        if (n) {
            float *b = ctx->d_basis;
            const size_t stride = (size_t)ctx->cap_entities * 3;
            k_basis<<<std::min<uint32_t>(ceil_div(n, 256), ctx->sm_count * 8), 256, 0, s>>>(
                ctx->d_tiles, b, b + stride, b + 2 * stride, n);
            CK_LAUNCH();
        }
    }

    const bool counts_in_sort = sort_now && bp.bits != 0;      // the sort's last launch also scans the counts
    if (sort_now) { if (int rc = launch_sort(ctx, bp, plan, n_counts, s)) return rc; }
    if (n_counts && !counts_in_sort) {
        cudaLaunchAttribute pdl;
        pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pdl.val.programmaticStreamSerializationAllowed = ctx->use_pdl ? 1 : 0;
        cudaLaunchConfig_t cfg{};
        cfg.stream = s; cfg.attrs = &pdl; cfg.numAttrs = 1; cfg.gridDim = dim3(1); cfg.blockDim = dim3(256);
        CK(cudaLaunchKernelEx(&cfg, k_scan_counts, (const uint32_t *)ctx->d_counts, ctx->d_offsets, n_counts,
                              (const FrameControl *)ctx->d_ctl, (const SortControl *)ctx->d_sc, ctx->d_res, ctx->counts_mirror));
        CK_LAUNCH();
    }
    if ((stage_flags & ASTRE_FRAME_GATHER) && cull && ctx->d_gather) {
        cudaLaunchAttribute pdl;
        pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pdl.val.programmaticStreamSerializationAllowed = ctx->use_pdl ? 1 : 0;
        cudaLaunchConfig_t cfg{};
        cfg.stream = s; cfg.attrs = &pdl; cfg.numAttrs = 1; cfg.gridDim = dim3((unsigned)ctx->sm_count * 4u); cfg.blockDim = dim3(256);
        CK(cudaLaunchKernelEx(&cfg, k_gather_list, (const unsigned long long *)ctx->d_keys[0], (const unsigned long long *)ctx->d_keys[1],
                              (const SortControl *)ctx->d_sc, (const FrameControl *)ctx->d_ctl, (unsigned long long)ctx->max_keys,
                              (sort_now && bp.bits) ? 1u : 0u, (const float4 *)ctx->d_world, ctx->d_gather,
                              (unsigned long long)ctx->cap_gather, ctx->world_bits, ctx->entities_per_world, ctx->d_res));
        CK_LAUNCH();
    }
    return ASTRE_OK;
}

bool same_plan(const FramePlan &a, const FramePlan &b) { return std::memcmp(&a, &b, sizeof(FramePlan)) == 0; }

}  // namespace

int astre_gpu_frame(astre_gpu_ctx *ctx, uint32_t stage_flags, void *stream)
{
    if (!ctx) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (ctx->hier_dirty) if (int rc = rebuild_levels(ctx)) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    const uint32_t n = ctx->n_entities;
    FramePlan fp;
    std::memset(&fp, 0, sizeof fp);                      // also the padding: plans are compared bytewise
    fp.stage_flags = stage_flags & ~(uint32_t)ASTRE_FRAME_GRAPH;
    fp.n = n;
    fp.cull = (stage_flags & ASTRE_FRAME_CULL) && ctx->views_per_world > 0;
    fp.sort_now = fp.cull && (stage_flags & ASTRE_FRAME_SORT);
    fp.sort = fp.cull && (stage_flags & (ASTRE_FRAME_SORT | ASTRE_FRAME_SORT_LATER));   // bucket counts + plans
    fp.n_counts = ctx->n_worlds * ctx->views_per_world;
    if (ctx->n_worlds > 1 && (uint64_t)ctx->n_worlds * ctx->entities_per_world < n)
        return fail(ctx, ASTRE_ERR_STATE, "entity count %u exceeds n_worlds*entities_per_world", n);

    // The frame number that tags look-back words lives on the device (k_frame_begin bumps it, so a captured graph
    // needs no per-frame parameter); this mirror advances in step and clears the tagged words when the tag wraps.
    ++ctx->list_serial;
    ctx->epoch = ctx->epoch + 1;
    if (ctx->epoch >= kEpochWrap) {
        const size_t n_sort_tiles = (size_t)((ctx->max_keys + 2048 - 1) / 2048) + 1;
        CK(cudaMemsetAsync(ctx->d_lookback, 0, n_sort_tiles * 256 * sizeof(unsigned long long), s));
        CK(cudaMemsetAsync(ctx->d_sc->scan_state, 0, sizeof(ctx->d_sc->scan_state), s));
        ctx->epoch = 1;
    }
    // plans of the sort from the key fields in use (all other key bits are zero in every key)
    if (fp.sort) {
        auto bits_for = [](uint64_t max_value) { uint32_t b = 0; while (b < 32 && (max_value >> b)) ++b; return b; };
        const uint32_t wb = ctx->world_bits;
        const uint64_t local_max = (ctx->n_worlds > 1 ? ctx->entities_per_world : n) ? (uint64_t)(ctx->n_worlds > 1 ? ctx->entities_per_world : n) - 1 : 0;
        unsigned long long live = (1ull << bits_for(local_max)) - 1ull;
        live |= ((1ull << 14) - 1ull) << (26 - wb);
        live |= ((1ull << bits_for(ctx->n_meshes ? ctx->n_meshes - 1 : 0)) - 1ull) << (40 - wb);
        live |= (unsigned long long)ctx->shader_or << (51 - wb);
        live |= ((1ull << bits_for(ctx->views_per_world ? ctx->views_per_world - 1 : 0)) - 1ull) << (59 - wb);
        if (wb) live |= ((1ull << wb) - 1ull) << (64 - wb);
        fp.plan = make_sort_plan(live);
        // buckets of ~4-8 keys: one bin per 8 keys of the last frame that finished (a hint -- any width sorts correctly;
        // before the first frame: one key per 16 entity-view pairs).  The width only follows the hint when it is off by
        // more than one bit, so that steady frames keep their plan (and their captured graph).
        const unsigned long long seen = *(volatile unsigned long long *)&ctx->h_res->key_total;
        const unsigned long long expect = seen != ~0ull ? std::min<unsigned long long>(seen, ctx->max_keys)
                                                        : (unsigned long long)n * ctx->views_per_world / 16ull;
        uint32_t want = kMinBinBits;
        while (want < ctx->cap_bin_bits && (1ull << want) * 8ull < expect) ++want;
        if (ctx->last_want_bits && want + 1 >= ctx->last_want_bits && want <= ctx->last_want_bits + 1) want = ctx->last_want_bits;
        if (want != ctx->last_want_bits) ctx->bits_bias = 0;
        ctx->last_want_bits = want;
        // Feedback from the buckets themselves: sum(m^2) / n is the mean length of a key's rank loop.  Keys rarely fill the
        // digit's value range evenly (two of eight view values hold 99 % of the headline frame's keys: mean 21; depths
        // cluster: mean 190 on a dense scene), so the width is widened while that mean is LONG and narrowed again when it is
        // very short -- one bit per finished frame that ran with the current width.  Only long: a wider table makes the
        // rank kernel cheaper but the cull kernel's REDs dearer (a table that does not stay hot in L2 while 1.8 GB stream
        // through it turns them into DRAM read-modify-writes; measured on the headline frame: 2^19 instead of 2^17 bins
        // = rank -2 us, cull kernel +6 us).
        {
            const FrameResultHost &R = *ctx->h_res;
            const unsigned long long kt = *(volatile unsigned long long *)&R.key_total;
            const uint32_t nb = *(volatile uint32_t *)&R.n_bins;
            const unsigned long long sq = *(volatile unsigned long long *)&R.sum_sq;
            const uint32_t cur = std::min<uint32_t>(std::max<int>((int)want + ctx->bits_bias, (int)kMinBinBits), ctx->cap_bin_bits);
            // (only feedback from a frame that ran with the CURRENT width counts: one step, then wait for its effect)
            if (kt != ~0ull && kt > 1024 && nb == (1u << cur) && ctx->bits_feedback) {
                const unsigned long long mean = sq / kt;
                // ... and only while the wider table would still be hot: ~1.2 M entities stream through L2 between two
                // evictions of a line, and a line holds 32 bins, so at least ~8 REDs per line and window means
                // bins <= 4.7 M x keys / entities (dense 4 M-entity scene: 4.4 M bins; headline frame: 0.2 M; 64 M-entity frame: 0.2 M)
                const unsigned long long hot_limit = kt * 4700000ull / std::max<uint32_t>(n, 1u);
                if (mean > 48 && cur < ctx->cap_bin_bits && (2ull << cur) <= hot_limit) ++ctx->bits_bias;
                else if (mean < 3 && cur > kMinBinBits && ctx->bits_bias > -3) --ctx->bits_bias;
            }
            want = std::min<uint32_t>(std::max<int>((int)want + ctx->bits_bias, (int)kMinBinBits), ctx->cap_bin_bits);
        }
        if (ctx->bin_bits_override >= 0) want = std::max<uint32_t>(2u, std::min<uint32_t>((uint32_t)ctx->bin_bits_override, ctx->cap_bin_bits));
        fp.bp = make_bin_plan(live, want);
    }
    if ((stage_flags & ASTRE_FRAME_BASIS) && !ctx->d_basis)
        CK(dev_alloc(&ctx->d_basis, (size_t)ctx->cap_entities * 9 * sizeof(float)));
    if ((stage_flags & ASTRE_FRAME_GATHER) && !ctx->d_gather) {          // room for a typical list; longer ones are gathered on read
        ctx->cap_gather = std::min<uint64_t>(ctx->max_keys, 1ull << 20);
        CK(dev_alloc(&ctx->d_gather, ctx->cap_gather * 64));
        ++ctx->params_version;
    }

    if (s != ctx->stream && ctx->ev_views) CK(cudaStreamWaitEvent(s, ctx->ev_views, 0));
    CK(cudaEventRecord(ctx->ev_begin, s));
    // Until a frame has finished there is no key count to size the sort's buckets with, and the plan of the first frames is
    // a guess that the next ones correct: those frames are launched kernel by kernel and the capture waits for the real plan.
    // (Measured on the headline frame: an executable instantiated once with the steady plan replays in 0.3265 ms, one
    // instantiated with the guess and updated 0.3284, one destroyed and instantiated a second time 0.335.)
    const bool plan_settled = !fp.sort || *(volatile unsigned long long *)&ctx->h_res->key_total != ~0ull || ctx->bin_bits_override >= 0 ||
                              std::getenv("ASTRE_GRAPH_EAGER_CAPTURE") != nullptr;
    if ((stage_flags & ASTRE_FRAME_GRAPH) && plan_settled) {
        // One cudaGraphLaunch instead of 6+ launches: what matters for small frames, where the launches ARE the frame.
        // Captured when the plan, buffers, the hierarchy or the stream changed (an exec update when the topology is the same).
        const bool stale = !ctx->graph_exec || !same_plan(fp, ctx->graph_plan) || ctx->graph_params_version != ctx->params_version ||
                           ctx->graph_levels_version != ctx->levels_version || ctx->graph_stream != s;
        bool views_changed = ctx->graph_views_version != ctx->views_version;
        if (!stale && views_changed && ctx->graph_vs_launches <= 1) {
            // the camera moved: kernels that read the records from global memory see the new ones already; the one
            // launch that has them as a parameter is patched in place
            if (ctx->graph_vs_launches == 1 && ctx->graph_vs_node) {
                void *args[] = { (void *)&ctx->graph_vs_args.P, (void *)&ctx->h_viewset, (void *)&ctx->graph_vs_args.bp, (void *)&ctx->graph_vs_args.LT };
                cudaKernelNodeParams kp{};
                kp.func = const_cast<void *>(ctx->graph_vs_args.fn);
                kp.gridDim = dim3(ctx->graph_vs_args.grid); kp.blockDim = dim3(kBlockThreads);
                kp.sharedMemBytes = kFusedSmemBytes; kp.kernelParams = args; kp.extra = nullptr;
                if (cudaGraphExecKernelNodeSetParams(ctx->graph_exec, ctx->graph_vs_node, &kp) == cudaSuccess) views_changed = false;
                else cudaGetLastError();
            } else if (ctx->graph_vs_launches == 0) {
                views_changed = false;
            }
            if (!views_changed) { ctx->graph_views_version = ctx->views_version; ++ctx->graph_view_patches; }
        }
        if (stale || views_changed) {
            cudaGraph_t g = nullptr;
            ctx->capturing = true;
            ctx->graph_vs_launches = 0;
            ctx->graph_vs_node = nullptr;
            const cudaError_t be = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
            if (be != cudaSuccess) { ctx->capturing = false; return fail(ctx, ASTRE_ERR_CUDA, "cudaStreamBeginCapture failed: %s", cudaGetErrorString(be)); }
            const uint64_t launches_before = ctx->launches;
            const int rc = enqueue_frame(ctx, fp, s, false);
            const cudaError_t ce = cudaStreamEndCapture(s, &g);
            ctx->capturing = false;
            ctx->graph_launches = (uint32_t)(ctx->launches - launches_before);
            ctx->launches = launches_before;
            if (rc != ASTRE_OK) { if (g) cudaGraphDestroy(g); return rc; }
            if (ce != cudaSuccess) return fail(ctx, ASTRE_ERR_CUDA, "graph capture of the frame failed: %s", cudaGetErrorString(ce));
            bool updated = false;
            if (ctx->graph_exec && ctx->graph_update_ok) {
                // same topology (the usual case: a new bucket width, a new entity count): update the executable in place.
                // Its nodes keep answering to the handles of the graph it was instantiated from, which is kept.
                cudaGraphExecUpdateResultInfo info;
                updated = cudaGraphExecUpdate(ctx->graph_exec, g, &info) == cudaSuccess;
                if (!updated) cudaGetLastError();
            }
            if (updated) {
                cudaGraphDestroy(g);
                g = ctx->graph_obj;
                if (ctx->graph_vs_launches != 1 || ctx->graph_vs_fn_at_instantiate != ctx->graph_vs_args.fn) ctx->graph_vs_node = nullptr;
                else ctx->graph_vs_node = ctx->graph_vs_node_at_instantiate;
            } else {
                if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
                if (ctx->graph_obj) { cudaGraphDestroy(ctx->graph_obj); ctx->graph_obj = nullptr; }
                const cudaError_t ie = cudaGraphInstantiate(&ctx->graph_exec, g, 0);
                if (ie != cudaSuccess) { cudaGraphDestroy(g); ctx->graph_exec = nullptr; return fail(ctx, ASTRE_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); }
                ctx->graph_obj = g;
                ctx->graph_vs_node_at_instantiate = nullptr;
                ctx->graph_vs_fn_at_instantiate = nullptr;
            }
            if (!updated && ctx->graph_vs_launches == 1) {          // find the node of that launch
                size_t n_nodes = 0;
                CK(cudaGraphGetNodes(g, nullptr, &n_nodes));
                std::vector<cudaGraphNode_t> nodes(n_nodes);
                CK(cudaGraphGetNodes(g, nodes.data(), &n_nodes));
                uint32_t found = 0;
                for (cudaGraphNode_t nd : nodes) {
                    cudaGraphNodeType ty;
                    if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
                    cudaKernelNodeParams kp{};
                    if (cudaGraphKernelNodeGetParams(nd, &kp) != cudaSuccess) { cudaGetLastError(); continue; }
                    if (kp.func == ctx->graph_vs_args.fn) { ctx->graph_vs_node = nd; ++found; }
                }
                if (found != 1) ctx->graph_vs_node = nullptr;     // ambiguous: a change of views captures again
                ctx->graph_vs_node_at_instantiate = ctx->graph_vs_node;
                ctx->graph_vs_fn_at_instantiate = ctx->graph_vs_args.fn;
            }
            ctx->graph_plan = fp;
            ctx->graph_views_version = ctx->views_version;
            ctx->graph_params_version = ctx->params_version;
            ctx->graph_levels_version = ctx->levels_version;
            ctx->graph_stream = s;
            ++ctx->graph_captures;
        }
        CK(cudaGraphLaunch(ctx->graph_exec, s));
        ctx->launches += ctx->graph_launches;
        ctx->last_timed = false;
    } else {
        if (int rc = enqueue_frame(ctx, fp, s, true)) return rc;
        ctx->last_timed = true;
    }
    CK(cudaEventRecord(ctx->ev_end, s));
    ctx->sort_pending = fp.sort && !fp.sort_now;
    ctx->pending_plan = fp.plan;
    ctx->pending_bins = fp.bp;
    ctx->last_sorted = fp.sort_now && fp.bp.bits != 0;
    ctx->last_gathered = (stage_flags & ASTRE_FRAME_GATHER) && fp.cull && fp.sort_now;
    ctx->frame_done = true;
    ctx->results_cached = false;
    ctx->last_flags = fp.stage_flags;
    ctx->last_stream = s;
    return ASTRE_OK;
}

int astre_gpu_sort(astre_gpu_ctx *ctx, void *stream)
{
    if (!ctx) return ASTRE_ERR_INVALID;
    if (!ctx->frame_done || !ctx->sort_pending)
        return fail(ctx, ASTRE_ERR_STATE, "astre_gpu_sort needs a frame run with ASTRE_FRAME_CULL | ASTRE_FRAME_SORT_LATER");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    if (int rc = launch_sort(ctx, ctx->pending_bins, ctx->pending_plan, 0u, s)) return rc;
    ctx->last_sorted = ctx->pending_bins.bits != 0;
    ctx->last_gathered = false;
    ctx->last_flags = (ctx->last_flags & ~(uint32_t)ASTRE_FRAME_SORT_LATER) | ASTRE_FRAME_SORT;
    ctx->sort_pending = false;
    ctx->results_cached = false;
    ++ctx->list_serial;
    CK(cudaEventRecord(ctx->ev_end, s));
    ctx->last_stream = s;
    return ASTRE_OK;
}

int astre_gpu_sync(astre_gpu_ctx *ctx)
{
    if (!ctx) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->last_stream && ctx->last_stream != ctx->stream) CK(cudaStreamSynchronize(ctx->last_stream));
    return ASTRE_OK;
}

int astre_gpu_set_stream(astre_gpu_ctx *ctx, void *stream)
{
    if (!ctx) return ASTRE_ERR_INVALID;
    if (int rc = astre_gpu_sync(ctx)) return rc;
    ctx->stream = stream ? (cudaStream_t)stream : ctx->own_stream;
    return ASTRE_OK;
}

uint64_t astre_gpu_launch_count(const astre_gpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

int astre_gpu_last_frame_ms(astre_gpu_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return ASTRE_ERR_INVALID;
    if (!ctx->frame_done) return fail(ctx, ASTRE_ERR_STATE, "no frame has been run");
    CK(cudaEventSynchronize(ctx->ev_end));
    CK(cudaEventElapsedTime(ms, ctx->ev_begin, ctx->ev_end));
    return ASTRE_OK;
}

int astre_gpu_last_cull_ms(astre_gpu_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return ASTRE_ERR_INVALID;
    if (!ctx->frame_done) return fail(ctx, ASTRE_ERR_STATE, "no frame has been run");
    if (!ctx->last_timed) return fail(ctx, ASTRE_ERR_STATE, "the last frame was replayed from a CUDA graph (ASTRE_FRAME_GRAPH): no per-kernel events");
    CK(cudaEventSynchronize(ctx->ev_cull1));
    CK(cudaEventElapsedTime(ms, ctx->ev_cull0, ctx->ev_cull1));
    return ASTRE_OK;
}

int astre_gpu_snapshot_trs(astre_gpu_ctx *ctx, void *stream)
{
    if (!ctx) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    if (!ctx->d_prev) CK(dev_alloc(&ctx->d_prev, ctx->tile_words() * sizeof(float)));
    CK(cudaMemcpyAsync(ctx->d_prev, ctx->d_tiles, ctx->tile_words() * sizeof(float), cudaMemcpyDeviceToDevice, s));
    ctx->n_prev = ctx->n_entities;
    return ASTRE_OK;
}

int astre_gpu_interpolate(astre_gpu_ctx *ctx, float alpha, void *stream)
{
    if (!ctx) return ASTRE_ERR_INVALID;
    if (!ctx->d_prev) return fail(ctx, ASTRE_ERR_STATE, "astre_gpu_interpolate needs a prior astre_gpu_snapshot_trs");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    alpha = alpha < 0.0f ? 0.0f : (alpha > 1.0f ? 1.0f : alpha);       // std::clamp, render.cpp:10 (NaN stays NaN there too)
    const uint32_t n = ctx->n_entities;
    if (n) {
        k_interpolate<<<std::min<uint32_t>(ceil_div(n, 256), (uint32_t)ctx->sm_count * 8), 256, 0, s>>>(
            ctx->d_prev, ctx->d_tiles, ctx->d_world, n, std::min(ctx->n_prev, n), alpha);
        CK_LAUNCH();
    }
    ctx->last_stream = s;
    return ASTRE_OK;
}

// The run heads of the last frame's sorted list in d_heads (found and ordered on the device); once per frame.
static int ensure_run_heads(astre_gpu_ctx *ctx, uint32_t *n_heads_out)
{
    if (int rc = cache_results(ctx)) return rc;
    if (!(ctx->last_flags & ASTRE_FRAME_SORT)) return fail(ctx, ASTRE_ERR_STATE, "draw runs need ASTRE_FRAME_SORT");
    if (ctx->h_overflow) return fail(ctx, ASTRE_ERR_CAPACITY, "frame produced %llu keys, capacity is %llu", ctx->h_total, (unsigned long long)ctx->max_keys);
    const uint32_t n = (uint32_t)ctx->h_total;
    *n_heads_out = 0;
    if (!n) { ctx->n_heads = 0; ctx->heads_serial = ctx->list_serial; return ASTRE_OK; }
    if (ctx->heads_serial == ctx->list_serial && ctx->d_heads) { *n_heads_out = ctx->n_heads; return ASTRE_OK; }
    const uint32_t n_ctas = (uint32_t)ctx->sm_count * 8u;
    if (!ctx->d_heads) {
        ctx->cap_heads = (uint32_t)std::min<uint64_t>(ctx->max_keys, 1u << 22);
        CK(dev_alloc(&ctx->d_heads, (size_t)ctx->cap_heads * sizeof(RunHead)));
        CK(dev_alloc(&ctx->d_head_counts, ((size_t)n_ctas + 1) * sizeof(uint32_t)));
    }
    cudaStream_t s = ctx->last_stream;
    const uint32_t class_shift = 40u - ctx->world_bits;   // below it: depth and slot
    const uint32_t per_cta = ceil_div(n, n_ctas);
    const unsigned long long *keys = ctx->d_keys[ctx->h_result_buf];
    k_run_count<<<n_ctas, kRunThreads, 0, s>>>(keys, n, class_shift, per_cta, ctx->d_head_counts);
    CK_LAUNCH();
    k_run_scan<<<1, 1024, 0, s>>>(ctx->d_head_counts, n_ctas);
    CK_LAUNCH();
    k_run_write<<<n_ctas, kRunThreads, 0, s>>>(keys, n, class_shift, per_cta, ctx->d_head_counts, ctx->d_heads, ctx->cap_heads);
    CK_LAUNCH();
    uint32_t n_heads = 0;
    CK(cudaMemcpyAsync(&n_heads, ctx->d_head_counts + n_ctas, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (n_heads > ctx->cap_heads) return fail(ctx, ASTRE_ERR_CAPACITY, "%u draw runs exceed the internal capacity %u", n_heads, ctx->cap_heads);
    ctx->n_heads = n_heads;
    ctx->heads_serial = ctx->list_serial;
    *n_heads_out = n_heads;
    return ASTRE_OK;
}

int astre_gpu_read_runs(astre_gpu_ctx *ctx, astre_draw_run *runs, uint32_t cap, uint32_t *n_out)
{
    if (!ctx || !n_out || (cap && !runs)) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    uint32_t n_heads = 0;
    *n_out = 0;
    if (int rc = ensure_run_heads(ctx, &n_heads)) return rc;
    if (!n_heads) return ASTRE_OK;
    const uint32_t n = (uint32_t)ctx->h_total;
    std::vector<RunHead> heads(n_heads);
    CK(cudaMemcpy(heads.data(), ctx->d_heads, (size_t)n_heads * sizeof(RunHead), cudaMemcpyDeviceToHost));
    *n_out = n_heads;
    for (uint32_t i = 0; i < n_heads && i < cap; ++i) {          // class = world | view(5) | shader(8) | mesh(11)
        runs[i].first = (uint32_t)heads[i].index;
        runs[i].count = (uint32_t)((i + 1 < n_heads ? heads[i + 1].index : n) - heads[i].index);
        runs[i].world_view = (uint32_t)(heads[i].cls >> 19);
        runs[i].shader_mesh = (uint32_t)(heads[i].cls & ((1ull << 19) - 1ull));
    }
    return ASTRE_OK;
}

int astre_gpu_set_mesh_draw_info(astre_gpu_ctx *ctx, uint32_t n_meshes, const astre_mesh_draw_info *info)
{
    static_assert(sizeof(astre_mesh_draw_info) == sizeof(MeshDrawInfo), "mesh draw records must match");
    if (!ctx || (n_meshes && !info)) return ctx ? fail(ctx, ASTRE_ERR_INVALID, "info is null") : ASTRE_ERR_INVALID;
    if (n_meshes > (1u << ASTRE_KEY_MESH_BITS)) return fail(ctx, ASTRE_ERR_INVALID, "at most 2048 meshes (mesh field of the draw key)");
    CK(cudaSetDevice(ctx->device));
    if (int rc = astre_gpu_sync(ctx)) return rc;
    if (n_meshes > ctx->cap_mesh_info) {
        if (ctx->d_mesh_info) CK(dev_free(ctx->d_mesh_info));
        ctx->d_mesh_info = nullptr; ctx->cap_mesh_info = 0;
        CK(dev_alloc(&ctx->d_mesh_info, (size_t)n_meshes * sizeof(MeshDrawInfo)));
        ctx->cap_mesh_info = n_meshes;
    }
    if (n_meshes) CK(cudaMemcpy(ctx->d_mesh_info, info, (size_t)n_meshes * sizeof(MeshDrawInfo), cudaMemcpyHostToDevice));
    ctx->n_mesh_info = n_meshes;
    return ASTRE_OK;
}

int astre_gpu_build_indirect(astre_gpu_ctx *ctx, uint32_t *n_out, const void **commands_dev)
{
    static_assert(sizeof(astre_draw_indirect) == sizeof(DrawIndirect) && sizeof(DrawIndirect) == 20, "DrawElementsIndirectCommand layout");
    if (!ctx || !n_out) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    uint32_t n_heads = 0;
    *n_out = 0;
    if (commands_dev) *commands_dev = nullptr;
    if (int rc = ensure_run_heads(ctx, &n_heads)) return rc;
    if (!n_heads) return ASTRE_OK;
    if (!ctx->d_indirect) CK(dev_alloc(&ctx->d_indirect, (size_t)ctx->cap_heads * sizeof(DrawIndirect)));
    cudaStream_t s = ctx->last_stream;
    k_indirect_from_runs<<<ceil_div(n_heads, 256u), 256, 0, s>>>(ctx->d_heads, n_heads, (uint32_t)ctx->h_total, ctx->d_mesh_info,
                                                                 ctx->n_mesh_info, ctx->d_indirect);
    CK_LAUNCH();
    *n_out = n_heads;
    if (commands_dev) *commands_dev = ctx->d_indirect;
    return ASTRE_OK;
}

int astre_gpu_read_indirect(astre_gpu_ctx *ctx, astre_draw_indirect *commands, uint32_t cap, uint32_t *n_out)
{
    if (!ctx || !n_out || (cap && !commands)) return ASTRE_ERR_INVALID;
    const void *dev = nullptr;
    if (int rc = astre_gpu_build_indirect(ctx, n_out, &dev)) return rc;
    const uint32_t n = std::min(*n_out, cap);
    if (n) CK(cudaMemcpyAsync(commands, dev, (size_t)n * sizeof(DrawIndirect), cudaMemcpyDeviceToHost, ctx->last_stream));
    CK(cudaStreamSynchronize(ctx->last_stream));
    return ASTRE_OK;
}

int astre_gpu_set_counts_mirror(astre_gpu_ctx *ctx, void *dev_slot0, void *dev_slot1)
{
    if (!ctx || (!dev_slot0) != (!dev_slot1)) return ctx ? fail(ctx, ASTRE_ERR_INVALID, "pass two device buffers, or two NULLs to switch the mirror off") : ASTRE_ERR_INVALID;
    if (int rc = astre_gpu_sync(ctx)) return rc;
    ctx->counts_mirror.slot[0] = static_cast<uint32_t *>(dev_slot0);
    ctx->counts_mirror.slot[1] = static_cast<uint32_t *>(dev_slot1);
    ++ctx->params_version;             // kernel parameters of a captured frame graph
    return ASTRE_OK;
}

// ---- count boards: the 8e exchange as one kernel over peer memory (kernels_mirror.cuh) -----------------------------

int astre_gpu_count_board_create(astre_gpu_ctx *ctx, uint32_t max_counts, void *ipc_handle_out, void **board_dev_out)
{
    if (!ctx || max_counts == 0) return ctx ? fail(ctx, ASTRE_ERR_INVALID, "max_counts must be positive") : ASTRE_ERR_INVALID;
    if (ctx->d_board) return fail(ctx, ASTRE_ERR_STATE, "this context already has a count board");
    CK(cudaSetDevice(ctx->device));
    void *p = nullptr;
    CK(cudaMalloc(&p, board_bytes(max_counts)));      // not guarded: peers map this allocation by its base address (CUDA IPC)
    CK(cudaMemset(p, 0, board_bytes(max_counts)));
    CK(cudaDeviceSynchronize());
    ctx->d_board = static_cast<CountBoard *>(p);
    ctx->board_cap = max_counts;
    CK(dev_alloc(&ctx->d_board_table, (size_t)kBoardRanks * max_counts * sizeof(uint32_t)));
    if (ipc_handle_out) {
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, p));
        std::memcpy(ipc_handle_out, &h, sizeof h);
    }
    if (board_dev_out) *board_dev_out = p;
    return ASTRE_OK;
}

int astre_gpu_count_board_attach(astre_gpu_ctx *ctx, uint32_t world_size, uint32_t rank, void *const *boards)
{
    if (!ctx) return ASTRE_ERR_INVALID;
    if (!ctx->d_board) return fail(ctx, ASTRE_ERR_STATE, "astre_gpu_count_board_create has not been called");
    if (world_size == 0 || world_size > kBoardRanks || rank >= world_size)
        return fail(ctx, ASTRE_ERR_INVALID, "count boards connect 1..%u ranks, got rank %u of %u", kBoardRanks, rank, world_size);
    if (world_size > 1 && !boards) return fail(ctx, ASTRE_ERR_INVALID, "boards is null");
    BoardPeers B{};
    for (uint32_t t = 0; t < world_size; ++t) {
        B.b[t] = t == rank ? ctx->d_board : static_cast<CountBoard *>(boards[t]);
        if (!B.b[t]) return fail(ctx, ASTRE_ERR_INVALID, "board of rank %u is null", t);
    }
    B.own = ctx->d_board; B.n = world_size; B.self = rank; B.cap = ctx->board_cap;
    ctx->board_peers = B;
    return ASTRE_OK;
}

int astre_gpu_count_board_exchange(astre_gpu_ctx *ctx, const uint32_t *counts_dev, uint32_t n_counts, uint32_t rank_major,
                                   uint32_t *table_dev, uint64_t *offsets_dev, uint64_t *view_totals_dev, void *stream)
{
    if (!ctx || (n_counts && !counts_dev)) return ctx ? fail(ctx, ASTRE_ERR_INVALID, "counts_dev is null") : ASTRE_ERR_INVALID;
    const BoardPeers &B = ctx->board_peers;
    if (!B.n) return fail(ctx, ASTRE_ERR_STATE, "no count board attached");
    if (n_counts > B.cap) return fail(ctx, ASTRE_ERR_INVALID, "%u counts exceed the board's capacity %u", n_counts, B.cap);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    k_board_exchange<<<1, 256, 0, s>>>(B, counts_dev, n_counts, rank_major, table_dev ? table_dev : ctx->d_board_table,
                                       reinterpret_cast<unsigned long long *>(offsets_dev),
                                       reinterpret_cast<unsigned long long *>(view_totals_dev));
    CK_LAUNCH();
    return ASTRE_OK;
}

int astre_gpu_count_board_status(astre_gpu_ctx *ctx, uint32_t *exchanges, uint32_t *error)
{
    if (!ctx) return ASTRE_ERR_INVALID;
    if (!ctx->d_board) return fail(ctx, ASTRE_ERR_STATE, "astre_gpu_count_board_create has not been called");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    CountBoard h;
    CK(cudaMemcpy(&h, ctx->d_board, sizeof h, cudaMemcpyDeviceToHost));
    if (exchanges) *exchanges = h.round;
    if (error) *error = h.error;
    return ASTRE_OK;
}

int astre_gpu_frame_index(const astre_gpu_ctx *ctx, uint64_t *index)
{
    if (!ctx || !index) return ASTRE_ERR_INVALID;
    *index = ctx->epoch;
    return ASTRE_OK;
}

namespace {
__global__ void k_debug_adopt_keys(FrameControl *ctl, const unsigned long long *__restrict__ keys, unsigned long long n,
                                   uint32_t *__restrict__ bins, const __grid_constant__ BinPlan bp)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->key_total = n;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
        atomicAdd(&bins[bin_digit(bp, keys[i])], 1u);
}
}  // namespace

int astre_gpu_debug_sort_keys(astre_gpu_ctx *ctx, const uint64_t *keys, uint64_t n, uint64_t live_mask, uint32_t bin_bits,
                              uint64_t *sorted_out, uint32_t *path_out)
{
    if (!ctx || (n && (!keys || !sorted_out))) return ASTRE_ERR_INVALID;
    if (n > ctx->max_keys) return fail(ctx, ASTRE_ERR_CAPACITY, "%llu keys exceed the capacity %llu", (unsigned long long)n, (unsigned long long)ctx->max_keys);
    CK(cudaSetDevice(ctx->device));
    if (int rc = astre_gpu_sync(ctx)) return rc;
    cudaStream_t s = ctx->stream;
    FramePlan fp;
    std::memset(&fp, 0, sizeof fp);
    fp.plan = make_sort_plan(live_mask);
    fp.bp = make_bin_plan(live_mask, std::max<uint32_t>(2u, std::min<uint32_t>(bin_bits, ctx->cap_bin_bits)));
    ++ctx->list_serial;
    ctx->epoch = ctx->epoch + 1;                             // in step with the device-side frame number (k_frame_begin)
    if (ctx->epoch >= kEpochWrap) {
        const size_t n_sort_tiles = (size_t)((ctx->max_keys + 2048 - 1) / 2048) + 1;
        CK(cudaMemsetAsync(ctx->d_lookback, 0, n_sort_tiles * 256 * sizeof(unsigned long long), s));
        CK(cudaMemsetAsync(ctx->d_sc->scan_state, 0, sizeof(ctx->d_sc->scan_state), s));
        ctx->epoch = 1;
    }
    const uint32_t n_bins = fp.bp.bits ? (1u << fp.bp.bits) : 0u;
    k_frame_begin<<<64, 256, 0, s>>>(ctx->d_ctl, ctx->d_sc, ctx->d_counts, 0u, ctx->d_bins, n_bins);
    CK_LAUNCH();
    if (n) CK(cudaMemcpyAsync(ctx->d_keys[0], keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    k_debug_adopt_keys<<<(unsigned)ctx->sm_count * 4u, 256, 0, s>>>(ctx->d_ctl, ctx->d_keys[0], n, ctx->d_bins, fp.bp);
    CK_LAUNCH();
    if (int rc = launch_sort(ctx, fp.bp, fp.plan, 0u, s)) return rc;
    CK(cudaStreamSynchronize(s));
    uint32_t buf = 0, path = 0;
    if (fp.bp.bits) {
        CK(cudaMemcpy(&buf, &ctx->d_sc->result_buf, sizeof buf, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&path, &ctx->d_sc->path, sizeof path, cudaMemcpyDeviceToHost));
    }
    if (n) CK(cudaMemcpy(sorted_out, ctx->d_keys[buf & 1u], n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (path_out) *path_out = path;
    ctx->frame_done = false;                                 // the key buffers no longer hold a frame's list
    ctx->results_cached = false;
    ctx->sort_pending = false;
    return ASTRE_OK;
}

int64_t astre_gpu_debug_guard_check(void)
{
    if (!guard_enabled()) return -1;
    cudaDeviceSynchronize();
    std::vector<std::pair<void *, size_t>> live;
    {
        std::lock_guard<std::mutex> lock(g_guard_mutex);
        live.assign(g_guarded.begin(), g_guarded.end());
    }
    uint64_t bad = 0;
    for (const auto &a : live) bad += guard_damage(a.first, a.second);
    std::lock_guard<std::mutex> lock(g_guard_mutex);
    return (int64_t)(g_guard_violations + bad);
}

int64_t astre_gpu_debug_guard_selftest(void)
{
    if (!guard_enabled()) return -1;
    unsigned char *p = nullptr;
    if (dev_alloc(&p, 1000) != cudaSuccess) return -2;
    cudaMemset(p + 998, 0, 5);           // two bytes inside, three past the end
    cudaMemset(p - 1, 0, 1);             // one before the start
    const uint64_t bad = guard_damage(p, 1000);
    {
        std::lock_guard<std::mutex> lock(g_guard_mutex);
        g_guarded.erase(p);
    }
    cudaFree(p - kGuardBytes);
    return (int64_t)bad;
}

int astre_gpu_graph_stats(const astre_gpu_ctx *ctx, uint32_t *captures, uint32_t *view_patches, uint32_t *bucket_bits)
{
    if (!ctx) return ASTRE_ERR_INVALID;
    if (captures) *captures = ctx->graph_captures;
    if (view_patches) *view_patches = ctx->graph_view_patches;
    if (bucket_bits) *bucket_bits = ctx->pending_bins.bits;
    return ASTRE_OK;
}

int astre_gpu_last_sort_path(astre_gpu_ctx *ctx, uint32_t *path)
{
    if (!ctx || !path) return ASTRE_ERR_INVALID;
    if (!ctx->frame_done) return fail(ctx, ASTRE_ERR_STATE, "no frame has been run");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->last_stream));
    *path = ctx->last_sorted ? ctx->h_res->path : 0u;
    return ASTRE_OK;
}

int astre_gpu_global_offsets_dev(astre_gpu_ctx *ctx, const uint32_t *all_counts_dev, uint32_t world_size, uint32_t rank,
                                 uint32_t n_views, uint32_t rank_major, uint64_t *offsets_dev, uint64_t *view_totals_dev,
                                 void *stream)
{
    if (!ctx || !all_counts_dev || world_size == 0 || rank >= world_size) return ctx ? fail(ctx, ASTRE_ERR_INVALID, "bad argument") : ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    k_global_offsets<<<1, 256, 0, s>>>(all_counts_dev, world_size, rank, n_views, rank_major,
                                       reinterpret_cast<unsigned long long *>(offsets_dev),
                                       reinterpret_cast<unsigned long long *>(view_totals_dev));
    CK_LAUNCH();
    return ASTRE_OK;
}

int astre_gpu_read_world(astre_gpu_ctx *ctx, uint32_t first, uint32_t count, float *out)
{
    if (!ctx || (count && !out)) return ASTRE_ERR_INVALID;
    if ((uint64_t)first + count > ctx->max_entities) return fail(ctx, ASTRE_ERR_INVALID, "slot range out of bounds");
    CK(cudaSetDevice(ctx->device));
    // stream-ordered after the last frame, then one synchronisation of that stream (DMA straight into a pinned
    // destination; a pageable one is staged by the runtime)
    cudaStream_t s = ctx->last_stream ? ctx->last_stream : ctx->stream;
    if (s != ctx->stream) CK(cudaStreamSynchronize(ctx->stream));
    if (count) CK(cudaMemcpyAsync(out, ctx->d_world + (size_t)first * 4, (size_t)count * 64, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ASTRE_OK;
}

int astre_gpu_read_basis(astre_gpu_ctx *ctx, uint32_t first, uint32_t count, float *fwd3, float *up3, float *right3)
{
    if (!ctx) return ASTRE_ERR_INVALID;
    if (!ctx->d_basis) return fail(ctx, ASTRE_ERR_STATE, "no frame with ASTRE_FRAME_BASIS has been run");
    if ((uint64_t)first + count > ctx->max_entities) return fail(ctx, ASTRE_ERR_INVALID, "slot range out of bounds");
    CK(cudaSetDevice(ctx->device));
    if (int rc = astre_gpu_sync(ctx)) return rc;
    const size_t stride = (size_t)ctx->cap_entities * 3;
    float *outs[3] = { fwd3, up3, right3 };
    for (int i = 0; i < 3; ++i)
        if (outs[i] && count)
            CK(cudaMemcpy(outs[i], ctx->d_basis + i * stride + (size_t)first * 3, (size_t)count * 12, cudaMemcpyDeviceToHost));
    return ASTRE_OK;
}

int astre_gpu_read_counts(astre_gpu_ctx *ctx, uint32_t *counts)
{
    if (!ctx || !counts) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (int rc = cache_results(ctx)) return rc;
    std::memcpy(counts, ctx->h_counts.data(), ctx->h_counts.size() * sizeof(uint32_t));
    if (ctx->h_overflow) return fail(ctx, ASTRE_ERR_CAPACITY, "frame produced %llu keys, capacity is %llu", ctx->h_total, (unsigned long long)ctx->max_keys);
    return ASTRE_OK;
}

int astre_gpu_read_total(astre_gpu_ctx *ctx, uint64_t *total)
{
    if (!ctx || !total) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (int rc = cache_results(ctx)) return rc;
    *total = ctx->h_total;
    return ASTRE_OK;
}

int astre_gpu_read_all_keys(astre_gpu_ctx *ctx, uint64_t *keys, uint64_t cap, uint64_t *n_out)
{
    if (!ctx || !n_out) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (int rc = cache_results(ctx)) return rc;
    const uint64_t have = std::min<uint64_t>(ctx->h_total, ctx->max_keys);
    *n_out = have;
    const uint64_t c = std::min(have, cap);
    if (c && keys) CK(cudaMemcpy(keys, ctx->d_keys[ctx->h_result_buf], c * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (ctx->h_overflow) return fail(ctx, ASTRE_ERR_CAPACITY, "frame produced %llu keys, capacity is %llu", ctx->h_total, (unsigned long long)ctx->max_keys);
    return ASTRE_OK;
}

int astre_gpu_read_keys(astre_gpu_ctx *ctx, uint32_t world, uint32_t view, uint64_t *keys, uint64_t cap, uint64_t *n_out)
{
    if (!ctx || !n_out) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (int rc = cache_results(ctx)) return rc;
    if (world >= ctx->n_worlds || view >= ctx->views_per_world) return fail(ctx, ASTRE_ERR_INVALID, "world/view out of range");
    if (!(ctx->last_flags & ASTRE_FRAME_SORT)) return fail(ctx, ASTRE_ERR_STATE, "per-view lists need ASTRE_FRAME_SORT");
    if (ctx->h_overflow) return fail(ctx, ASTRE_ERR_CAPACITY, "frame produced %llu keys, capacity is %llu", ctx->h_total, (unsigned long long)ctx->max_keys);
    const size_t u = (size_t)world * ctx->views_per_world + view;
    const uint64_t off = ctx->h_offsets[u], cnt = ctx->h_counts[u];
    *n_out = cnt;
    const uint64_t c = std::min(cnt, cap);
    if (c && keys) CK(cudaMemcpy(keys, ctx->d_keys[ctx->h_result_buf] + off, c * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return ASTRE_OK;
}

int astre_gpu_read_visible_world(astre_gpu_ctx *ctx, uint64_t first_key, uint64_t n, float *out)
{
    if (!ctx || (n && !out)) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (int rc = cache_results(ctx)) return rc;
    const uint64_t have = std::min<uint64_t>(ctx->h_total, ctx->max_keys);
    if (first_key + n > have) return fail(ctx, ASTRE_ERR_INVALID, "key range [%llu, %llu) exceeds the %llu keys of the frame",
                                          (unsigned long long)first_key, (unsigned long long)(first_key + n), (unsigned long long)have);
    if (!n) return ASTRE_OK;
    if (ctx->cap_gather < n) {
        if (ctx->d_gather) CK(dev_free(ctx->d_gather));
        ctx->d_gather = nullptr;
        CK(dev_alloc(&ctx->d_gather, n * 64));
        ctx->cap_gather = n;
        ++ctx->params_version;                  // a captured frame graph has the old buffer baked in
    }
    cudaStream_t s = ctx->last_stream;
    const uint32_t grid = std::min<uint32_t>(ceil_div(n * 4, 256), (uint32_t)ctx->sm_count * 8);
    k_gather_world<<<grid, 256, 0, s>>>(ctx->d_keys[ctx->h_result_buf], first_key, n, ctx->d_world, ctx->d_gather,
                                        ctx->world_bits, ctx->entities_per_world);
    CK_LAUNCH();
    CK(cudaMemcpyAsync(out, ctx->d_gather, n * 64, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ASTRE_OK;
}

int astre_gpu_read_draw_list(astre_gpu_ctx *ctx, uint64_t *keys, float *world16, uint64_t cap, uint64_t *n_out)
{
    if (!ctx || !n_out) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (int rc = cache_results(ctx)) return rc;                     // one synchronisation: the frame has finished
    if (ctx->h_overflow) return fail(ctx, ASTRE_ERR_CAPACITY, "frame produced %llu keys, capacity is %llu", ctx->h_total, (unsigned long long)ctx->max_keys);
    const uint64_t have = std::min<uint64_t>(ctx->h_total, ctx->max_keys);
    *n_out = have;
    const uint64_t n = std::min(have, cap);
    if (!n) return ASTRE_OK;
    cudaStream_t s = ctx->last_stream;
    if (ctx->last_gathered && ctx->h_res->list_n == have) {          // short list: the frame left it in host memory
        if (keys) std::memcpy(keys, ctx->h_res->list_keys, n * sizeof(uint64_t));
        if (world16) std::memcpy(world16, ctx->h_res->list_world, n * 64);
        return ASTRE_OK;
    }
    if (ctx->last_gathered && have <= ctx->cap_gather) {             // gathered inside the frame: two copies, no launch
        if (keys) CK(cudaMemcpyAsync(keys, ctx->d_keys[ctx->h_result_buf], n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        if (world16) CK(cudaMemcpyAsync(world16, ctx->d_gather, n * 64, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        return ASTRE_OK;
    }
    if (world16) {
        if (ctx->cap_gather < n) {
            if (ctx->d_gather) CK(dev_free(ctx->d_gather));
            ctx->d_gather = nullptr;
            CK(dev_alloc(&ctx->d_gather, n * 64));
            ctx->cap_gather = n;
            ++ctx->params_version;              // a captured frame graph has the old buffer baked in
        }
        const uint32_t grid = std::min<uint32_t>(ceil_div(n * 4, 256), (uint32_t)ctx->sm_count * 8);
        k_gather_world<<<grid, 256, 0, s>>>(ctx->d_keys[ctx->h_result_buf], 0ull, n, ctx->d_world, ctx->d_gather,
                                            ctx->world_bits, ctx->entities_per_world);
        CK_LAUNCH();
    }
    if (keys) CK(cudaMemcpyAsync(keys, ctx->d_keys[ctx->h_result_buf], n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    if (world16) CK(cudaMemcpyAsync(world16, ctx->d_gather, n * 64, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));                                   // and one for both copies
    return ASTRE_OK;
}

int astre_gpu_slot_entity(const astre_gpu_ctx *ctx, uint32_t slot, uint64_t *entity_id)
{
    if (!ctx || !entity_id || slot >= ctx->max_entities) return ASTRE_ERR_INVALID;
    *entity_id = ctx->h_entity_id.empty() ? slot : ctx->h_entity_id[slot];
    return ASTRE_OK;
}

void *astre_gpu_device_world(astre_gpu_ctx *ctx) { return ctx ? ctx->d_world : nullptr; }
void *astre_gpu_device_counts(astre_gpu_ctx *ctx) { return ctx ? ctx->d_counts : nullptr; }
void *astre_gpu_device_sorted_keys(astre_gpu_ctx *ctx)
{
    if (!ctx || cache_results(ctx) != ASTRE_OK) return nullptr;
    return ctx->d_keys[ctx->h_result_buf];
}

// ---- f3: global key-sorted draw list across GPUs ----------------------------

namespace {

// Publish buffer: keys[max_keys] | samples[max_keys/128 + 1] | cuts[kMaxLists * (max_keys/128 + 1)] | PublishHeader.
// Every rank of a job uses the same max_keys, so every rank knows every peer's layout.
struct PublishLayout { size_t samples_off, cuts_off, header_off, bytes; uint32_t sample_cap; };
PublishLayout publish_layout(uint64_t max_keys)
{
    PublishLayout P;
    P.sample_cap = (uint32_t)(max_keys / kSampleStep + 1);
    P.samples_off = (std::max<uint64_t>(max_keys, 1) * 8 + 255) / 256 * 256;
    P.cuts_off = P.samples_off + ((size_t)P.sample_cap * 8 + 255) / 256 * 256;
    P.header_off = (P.cuts_off + (size_t)kMaxLists * P.sample_cap * 4 + 255) / 256 * 256;
    P.bytes = P.header_off + sizeof(PublishHeader);
    return P;
}

int ensure_publish(astre_gpu_ctx *ctx, uint32_t parity)
{
    if (ctx->d_publish[parity]) return ASTRE_OK;
    void *p = nullptr;
    const PublishLayout P = publish_layout(ctx->max_keys);
    CK(cudaMalloc(&p, P.bytes));      // not guarded: peers map this allocation by its base address (CUDA IPC)
    CK(cudaMemset(static_cast<unsigned char *>(p) + P.header_off, 0, sizeof(PublishHeader)));      // flags start at epoch 0
    ctx->d_publish[parity] = static_cast<unsigned char *>(p);
    if (!ctx->d_merge_error) {      // allocated here, not inside the per-frame calls: those must never block on the device
        CK(dev_alloc(&ctx->d_merge_error, sizeof(uint32_t)));
        CK(cudaMemset(ctx->d_merge_error, 0, sizeof(uint32_t)));
    }
    if (!ctx->d_mdims) CK(dev_alloc(&ctx->d_mdims, sizeof(MergeDims)));
    return ASTRE_OK;
}

int fill_lists(astre_gpu_ctx *ctx, MergeLists &L, uint32_t n_lists, const uint64_t *counts)
{
    if (n_lists == 0 || n_lists > (uint32_t)kMaxLists) return fail(ctx, ASTRE_ERR_INVALID, "n_lists must be in 1..%d", kMaxLists);
    if (ctx->world_bits) return fail(ctx, ASTRE_ERR_STATE, "batched worlds shard by whole worlds: their global list is the rank-major concatenation, no merge");
    L = MergeLists{};
    L.n_lists = n_lists;
    L.slot_bits = ASTRE_KEY_SLOT_BITS;
    uint64_t total = 0;
    uint32_t n_samples = 0;
    for (uint32_t t = 0; t < n_lists; ++t) {
        if (counts[t] > (1ull << ASTRE_KEY_SLOT_BITS) * ASTRE_MAX_VIEWS) return fail(ctx, ASTRE_ERR_INVALID, "list %u: count out of range", t);
        L.d.n[t] = (uint32_t)counts[t];
        L.d.sample_base[t] = n_samples;
        n_samples += (uint32_t)((counts[t] + kSampleStep - 1) >> kSampleShift);
        total += counts[t];
    }
    if (total >= (1ull << 24) * kSampleStep) return fail(ctx, ASTRE_ERR_CAPACITY, "global list of %llu keys exceeds 2^31", (unsigned long long)total);
    L.d.sample_base[n_lists] = n_samples;
    L.d.n_lists = n_lists;
    L.d.total = total;
    return ASTRE_OK;
}

// Output and tile-table buffers of a merge.  May free and reallocate (which waits for the device): the flows that
// enqueue kernels waiting on other GPUs call this BEFORE they enqueue anything.
int ensure_merge_buffers(astre_gpu_ctx *ctx, uint64_t out_count, uint32_t n_samples, bool exact)
{
    if (ctx->cap_merged < out_count || !ctx->d_merged_keys) {
        if (ctx->d_merged_keys) CK(dev_free(ctx->d_merged_keys));
        if (ctx->d_merged_src) CK(dev_free(ctx->d_merged_src));
        ctx->d_merged_keys = nullptr; ctx->d_merged_src = nullptr; ctx->cap_merged = 0;
        const uint64_t cap = std::max<uint64_t>(exact ? out_count : out_count + out_count / 4, 1024);
        CK(dev_alloc(&ctx->d_merged_keys, cap * sizeof(unsigned long long)));
        CK(dev_alloc(&ctx->d_merged_src, cap));
        ctx->cap_merged = cap;
    }
    if (ctx->cap_merge_tiles < n_samples + 1 || !ctx->d_splits) {
        if (ctx->d_splits) CK(dev_free(ctx->d_splits));
        if (ctx->d_tile_start) CK(dev_free(ctx->d_tile_start));
        ctx->d_splits = nullptr; ctx->d_tile_start = nullptr; ctx->cap_merge_tiles = 0;
        const uint32_t cap = exact ? n_samples + 1 : n_samples + n_samples / 4 + 64;
        CK(dev_alloc(&ctx->d_splits, (size_t)cap * kMaxLists * sizeof(uint32_t)));
        CK(dev_alloc(&ctx->d_tile_start, (size_t)cap * sizeof(uint32_t)));
        ctx->cap_merge_tiles = cap;
    }
    return ASTRE_OK;
}

// table + tiles, given lists, samples and cut tables.  Host-side sizes (L.dev_dims == nullptr): out_first /
// out_count select the window.  Device-side sizes: n_samples_cap bounds the sample count, the window is slice
// slice_rank of slice_world (0 = whole list) and at most out_count keys.
int merge_tail(astre_gpu_ctx *ctx, const MergeLists &L, uint64_t out_first, uint64_t out_count, uint32_t slice_rank,
               uint32_t slice_world, uint32_t n_samples_cap, cudaStream_t s)
{
    const bool dev = L.dev_dims != nullptr;
    const uint32_t n_samples = dev ? n_samples_cap : L.d.sample_base[L.n_lists];
    if (!dev) {
        out_first = std::min<uint64_t>(out_first, L.d.total);
        out_count = std::min<uint64_t>(out_count, L.d.total - out_first);
    }
    if (int rc = ensure_merge_buffers(ctx, out_count, n_samples, dev)) return rc;
    if (n_samples && out_count) {
        cudaLaunchAttribute pdl;
        pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pdl.val.programmaticStreamSerializationAllowed = ctx->use_pdl ? 1 : 0;
        cudaLaunchConfig_t cfg{};
        cfg.stream = s; cfg.attrs = &pdl; cfg.numAttrs = 1; cfg.dynamicSmemBytes = 0;
        // sizes only known on the device: fixed grids, the kernels stride over the real counts
        const uint32_t table_grid = dev ? (uint32_t)ctx->sm_count * 4u : ceil_div(((uint64_t)n_samples + 1) * 8, 256);
        cfg.gridDim = dim3(table_grid); cfg.blockDim = dim3(256);
        CK(cudaLaunchKernelEx(&cfg, k_merge_table, L, ctx->d_splits, ctx->d_tile_start));
        CK_LAUNCH();
        const uint32_t span = (uint32_t)kMergeBatchKeys - (L.n_lists << kSampleShift);
        const uint32_t tiles_grid = dev ? (uint32_t)ctx->sm_count * 6u : ceil_div(L.d.total, span);
        cfg.gridDim = dim3(tiles_grid); cfg.blockDim = dim3(kMergeThreads);
        const unsigned long long of = out_first, oc = out_count;
        CK(cudaLaunchKernelEx(&cfg, k_merge_tiles, L, (const uint32_t *)ctx->d_splits, (const uint32_t *)ctx->d_tile_start,
                              ctx->d_merged_keys, ctx->d_merged_src, of, oc, slice_rank, slice_world));
        CK_LAUNCH();
    }
    CK(cudaEventRecord(ctx->ev_merge1, s));
    ctx->n_merged = out_count;
    ctx->merge_dev = dev;
    ctx->merge_slice_rank = slice_rank; ctx->merge_slice_world = slice_world;
    ctx->merge_stream = s;
    ctx->merge_done = true;
    ctx->merge_begun = false;
    return ASTRE_OK;
}

int merge_begin(astre_gpu_ctx *ctx, cudaStream_t s)
{
    if (!ctx->ev_merge0) {
        CK(cudaEventCreate(&ctx->ev_merge0));
        CK(cudaEventCreate(&ctx->ev_merge1));
    }
    CK(cudaEventRecord(ctx->ev_merge0, s));
    ctx->merge_begun = true;
    return ASTRE_OK;
}

}  // namespace

int astre_gpu_publish_keys(astre_gpu_ctx *ctx, uint32_t parity, void *stream)
{
    if (!ctx || parity > 1) return ASTRE_ERR_INVALID;
    if (!ctx->frame_done) return fail(ctx, ASTRE_ERR_STATE, "no frame has been run");
    if (ctx->sort_pending || !(ctx->last_flags & ASTRE_FRAME_SORT))
        return fail(ctx, ASTRE_ERR_STATE, "astre_gpu_publish_keys needs a sorted frame (ASTRE_FRAME_SORT, or astre_gpu_sort after ASTRE_FRAME_SORT_LATER)");
    CK(cudaSetDevice(ctx->device));
    if (int rc = ensure_publish(ctx, parity)) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    unsigned char *base = ctx->d_publish[parity];
    k_publish_keys<<<ctx->sm_count * 4, 256, 0, s>>>(ctx->d_keys[0], ctx->d_keys[1], &ctx->d_sc->result_buf, ctx->last_sorted ? 1u : 0u,
                                                     &ctx->d_ctl->key_total, ctx->max_keys,
                                                     reinterpret_cast<unsigned long long *>(base),
                                                     reinterpret_cast<unsigned long long *>(base + publish_layout(ctx->max_keys).samples_off),
                                                     nullptr, 0u, nullptr);
    CK_LAUNCH();
    return ASTRE_OK;
}

void *astre_gpu_device_published_keys(astre_gpu_ctx *ctx, uint32_t parity)
{
    if (!ctx || parity > 1) return nullptr;
    cudaSetDevice(ctx->device);
    return ensure_publish(ctx, parity) == ASTRE_OK ? ctx->d_publish[parity] : nullptr;
}

int astre_gpu_export_handle(astre_gpu_ctx *ctx, uint32_t parity, void *handle_out)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == ASTRE_IPC_HANDLE_BYTES, "IPC handle size");
    if (!ctx || parity > 1 || !handle_out) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (int rc = ensure_publish(ctx, parity)) return rc;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->d_publish[parity]));
    std::memcpy(handle_out, &h, sizeof h);
    return ASTRE_OK;
}

int astre_gpu_import_handle(astre_gpu_ctx *ctx, const void *handle, void **dev_ptr)
{
    if (!ctx || !handle || !dev_ptr) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->imports.push_back(p);
    *dev_ptr = p;
    return ASTRE_OK;
}

int astre_gpu_release_import(astre_gpu_ctx *ctx, void *dev_ptr)
{
    if (!ctx || !dev_ptr) return ASTRE_ERR_INVALID;
    auto it = std::find(ctx->imports.begin(), ctx->imports.end(), dev_ptr);
    if (it == ctx->imports.end()) return fail(ctx, ASTRE_ERR_INVALID, "pointer was not returned by astre_gpu_import_handle");
    CK(cudaSetDevice(ctx->device));
    ctx->imports.erase(it);
    CK(cudaIpcCloseMemHandle(dev_ptr));
    return ASTRE_OK;
}

int astre_gpu_merge_lists(astre_gpu_ctx *ctx, uint32_t n_lists, const void *const *lists, const uint64_t *counts,
                          uint64_t out_first, uint64_t out_count, void *stream)
{
    if (!ctx || !lists || !counts) return ASTRE_ERR_INVALID;
    MergeLists L;
    if (int rc = fill_lists(ctx, L, n_lists, counts)) return rc;
    for (uint32_t t = 0; t < n_lists; ++t) {
        if (counts[t] && !lists[t]) return fail(ctx, ASTRE_ERR_INVALID, "list %u is null", t);
        L.keys[t] = static_cast<const unsigned long long *>(lists[t]);
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    const uint32_t n_samples = L.d.sample_base[n_lists];
    if (ctx->cap_msamples < n_samples || !ctx->d_mcuts) {
        if (ctx->d_mcuts) CK(dev_free(ctx->d_mcuts));
        ctx->d_mcuts = nullptr; ctx->cap_msamples = 0;
        const uint32_t cap = n_samples + n_samples / 4 + 64;
        CK(dev_alloc(&ctx->d_mcuts, (size_t)cap * kMaxLists * 4));
        ctx->cap_msamples = cap;
    }
    if (int rc = merge_begin(ctx, s)) return rc;
    L.sample_stride_shift = kSampleShift;                  // samples are read straight out of the lists
    for (uint32_t t = 0; t < n_lists; ++t) {
        L.samples[t] = L.keys[t];
        L.cuts[t] = ctx->d_mcuts + (size_t)t * ctx->cap_msamples;
    }
    if (n_samples) {
        dim3 grid(std::min<uint32_t>(ceil_div(n_samples, 256), (uint32_t)ctx->sm_count * 8), n_lists);
        k_merge_cuts<<<grid, 256, 0, s>>>(L, 0u, ctx->d_mcuts, (size_t)ctx->cap_msamples);
        CK_LAUNCH();
    }
    return merge_tail(ctx, L, out_first, out_count, 0, 0, 0, s);
}

// Peer-to-peer flow, step 1 (after the counts exchange): rank `rank` ranks every sample of every
// published list inside its own list and leaves the cut table in its publish buffer.
int astre_gpu_merge_prepare(astre_gpu_ctx *ctx, uint32_t n_lists, uint32_t rank, const void *const *publish, const uint64_t *counts,
                            void *stream)
{
    if (!ctx || !publish || !counts || rank >= n_lists) return ASTRE_ERR_INVALID;
    MergeLists L;
    if (int rc = fill_lists(ctx, L, n_lists, counts)) return rc;
    const PublishLayout P = publish_layout(ctx->max_keys);
    for (uint32_t t = 0; t < n_lists; ++t) {
        if (!publish[t]) return fail(ctx, ASTRE_ERR_INVALID, "publish buffer %u is null", t);
        if (counts[t] > ctx->max_keys) return fail(ctx, ASTRE_ERR_CAPACITY, "list %u has %llu keys; publish buffers hold %llu (every rank must use the same max_keys)",
                                                    t, (unsigned long long)counts[t], (unsigned long long)ctx->max_keys);
        const unsigned char *base = static_cast<const unsigned char *>(publish[t]);
        L.keys[t] = reinterpret_cast<const unsigned long long *>(base);
        L.samples[t] = reinterpret_cast<const unsigned long long *>(base + P.samples_off);
        L.cuts[t] = reinterpret_cast<const uint32_t *>(base + P.cuts_off);
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    if (int rc = merge_begin(ctx, s)) return rc;
    const uint32_t n_samples = L.d.sample_base[n_lists];
    if (n_samples) {
        uint32_t *own_cuts = const_cast<uint32_t *>(L.cuts[rank]);
        k_merge_cuts<<<std::min<uint32_t>(ceil_div(n_samples, 256), (uint32_t)ctx->sm_count * 8), 256, 0, s>>>(L, rank, own_cuts, 0);
        CK_LAUNCH();
    }
    return ASTRE_OK;
}

// Step 2 (after a barrier that orders every rank's astre_gpu_merge_prepare before it): tile table
// from all cut tables, then the merge itself, reading keys straight from the peers' buffers.
int astre_gpu_merge_published(astre_gpu_ctx *ctx, uint32_t n_lists, const void *const *publish, const uint64_t *counts,
                              uint64_t out_first, uint64_t out_count, void *stream)
{
    if (!ctx || !publish || !counts) return ASTRE_ERR_INVALID;
    MergeLists L;
    if (int rc = fill_lists(ctx, L, n_lists, counts)) return rc;
    const PublishLayout P = publish_layout(ctx->max_keys);
    for (uint32_t t = 0; t < n_lists; ++t) {
        if (!publish[t]) return fail(ctx, ASTRE_ERR_INVALID, "publish buffer %u is null", t);
        if (counts[t] > ctx->max_keys) return fail(ctx, ASTRE_ERR_CAPACITY, "list %u has %llu keys; publish buffers hold %llu", t,
                                                    (unsigned long long)counts[t], (unsigned long long)ctx->max_keys);
        const unsigned char *base = static_cast<const unsigned char *>(publish[t]);
        L.keys[t] = reinterpret_cast<const unsigned long long *>(base);
        L.samples[t] = reinterpret_cast<const unsigned long long *>(base + P.samples_off);
        L.cuts[t] = reinterpret_cast<const uint32_t *>(base + P.cuts_off);
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    if (!ctx->merge_begun) { if (int rc = merge_begin(ctx, s)) return rc; }     // no prepare on this context this frame
    return merge_tail(ctx, L, out_first, out_count, 0, 0, 0, s);
}

// Same two steps with the key counts left on the device: all_counts_dev is the all-gathered [n_lists][n_views]
// table of per-view visible counts (uint32, device memory), exactly what the 8e exchange produces.  Nothing
// here waits for the host, so the whole frame -> publish -> exchange -> merge chain can be enqueued ahead.
static int lists_from_publish(astre_gpu_ctx *ctx, MergeLists &L, uint32_t n_lists, const void *const *publish)
{
    if (n_lists == 0 || n_lists > (uint32_t)kMaxLists) return fail(ctx, ASTRE_ERR_INVALID, "n_lists must be in 1..%d", kMaxLists);
    if (ctx->world_bits) return fail(ctx, ASTRE_ERR_STATE, "batched worlds shard by whole worlds: their global list is the rank-major concatenation, no merge");
    if ((uint64_t)n_lists * ctx->max_keys >= (1ull << 31)) return fail(ctx, ASTRE_ERR_CAPACITY, "n_lists * max_keys must stay below 2^31 for a merge with device-side counts");
    const PublishLayout P = publish_layout(ctx->max_keys);
    L = MergeLists{};
    L.n_lists = n_lists;
    L.slot_bits = ASTRE_KEY_SLOT_BITS;
    for (uint32_t t = 0; t < n_lists; ++t) {
        if (!publish[t]) return fail(ctx, ASTRE_ERR_INVALID, "publish buffer %u is null", t);
        const unsigned char *base = static_cast<const unsigned char *>(publish[t]);
        L.keys[t] = reinterpret_cast<const unsigned long long *>(base);
        L.samples[t] = reinterpret_cast<const unsigned long long *>(base + P.samples_off);
        L.cuts[t] = reinterpret_cast<const uint32_t *>(base + P.cuts_off);
    }
    if (!ctx->d_mdims) CK(dev_alloc(&ctx->d_mdims, sizeof(MergeDims)));
    L.dev_dims = ctx->d_mdims;
    return ASTRE_OK;
}

int astre_gpu_merge_prepare_dev(astre_gpu_ctx *ctx, uint32_t n_lists, uint32_t rank, const void *const *publish,
                                const uint32_t *all_counts_dev, uint32_t n_views, void *stream)
{
    if (!ctx || !publish || !all_counts_dev || rank >= n_lists || n_views == 0) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    MergeLists L;
    if (int rc = lists_from_publish(ctx, L, n_lists, publish)) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    if (int rc = merge_begin(ctx, s)) return rc;
    k_merge_dims<<<1, 32, 0, s>>>(all_counts_dev, n_lists, n_views, ctx->max_keys, ctx->d_mdims);
    CK_LAUNCH();
    uint32_t *own_cuts = const_cast<uint32_t *>(L.cuts[rank]);
    k_merge_cuts<<<(uint32_t)ctx->sm_count * 8u, 256, 0, s>>>(L, rank, own_cuts, 0);
    CK_LAUNCH();
    return ASTRE_OK;
}

int astre_gpu_merge_published_dev(astre_gpu_ctx *ctx, uint32_t n_lists, const void *const *publish, uint32_t slice_rank,
                                  uint32_t slice_world, uint64_t out_cap, void *stream)
{
    if (!ctx || !publish || (slice_world && slice_rank >= slice_world)) return ASTRE_ERR_INVALID;
    if (!ctx->merge_begun || !ctx->d_mdims) return fail(ctx, ASTRE_ERR_STATE, "astre_gpu_merge_published_dev needs astre_gpu_merge_prepare_dev on this context first");
    CK(cudaSetDevice(ctx->device));
    MergeLists L;
    if (int rc = lists_from_publish(ctx, L, n_lists, publish)) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    const uint64_t whole = (uint64_t)n_lists * ctx->max_keys;
    if (out_cap == 0) out_cap = slice_world ? (whole + slice_world - 1) / slice_world : whole;     // worst case
    const uint32_t n_samples_cap = n_lists * publish_layout(ctx->max_keys).sample_cap;
    return merge_tail(ctx, L, 0, out_cap, slice_rank, slice_world, n_samples_cap, s);
}

// The whole exchange over NVLink peer memory, no NCCL and no host round trip: counts and "ready" flags travel in
// the publish buffers' headers, tiny one-warp kernels signal and poll them.  One call per frame and rank:
//   wait(peers are done reading my buffer of epoch-2) -> publish keys, samples, counts
//   -> signal "published", wait for the peers', list sizes from their counts -> cut table of my own list
//   -> signal "cuts", wait for the peers' -> tile table -> merge (keys pulled from the peers) -> signal "read".
// `epoch` is the frame number: the same on every rank, starting at 1, +1 per call; the buffer used is epoch & 1.
int astre_gpu_merge_p2p(astre_gpu_ctx *ctx, uint32_t n_lists, uint32_t rank, const void *const *publish0,
                        const void *const *publish1, uint32_t epoch, uint32_t slice_rank, uint32_t slice_world,
                        uint64_t out_cap, void *stream)
{
    if (!ctx || !publish0 || !publish1 || rank >= n_lists || epoch == 0 || (slice_world && slice_rank >= slice_world))
        return ASTRE_ERR_INVALID;
    if (!ctx->frame_done) return fail(ctx, ASTRE_ERR_STATE, "no frame has been run");
    if (!(ctx->last_flags & ASTRE_FRAME_SORT)) return fail(ctx, ASTRE_ERR_STATE, "the global list needs a sorted frame (ASTRE_FRAME_SORT)");
    CK(cudaSetDevice(ctx->device));
    const uint32_t parity = epoch & 1u;
    const void *const *publish = parity ? publish1 : publish0;
    MergeLists L;
    if (int rc = lists_from_publish(ctx, L, n_lists, publish)) return rc;
    if (publish[rank] != ctx->d_publish[parity] || !ctx->d_publish[parity])
        return fail(ctx, ASTRE_ERR_INVALID, "publish%u[rank] must be this context's own publish buffer", parity);
    const PublishLayout P = publish_layout(ctx->max_keys);
    PeerHeaders H{};
    H.n = n_lists; H.self = rank;
    for (uint32_t t = 0; t < n_lists; ++t)
        H.h[t] = reinterpret_cast<PublishHeader *>(const_cast<unsigned char *>(static_cast<const unsigned char *>(publish[t])) + P.header_off);
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    const uint32_t n_views = ctx->views_per_world;
    if (n_views == 0 || n_views > ASTRE_MAX_VIEWS) return fail(ctx, ASTRE_ERR_STATE, "no views set");
    const uint64_t whole = (uint64_t)n_lists * ctx->max_keys;
    if (out_cap == 0) out_cap = slice_world ? (whole + slice_world - 1) / slice_world : whole;
    const uint32_t n_samples_cap = n_lists * P.sample_cap;
    if (int rc = ensure_merge_buffers(ctx, out_cap, n_samples_cap, true)) return rc;      // nothing below may block on the device
    if (int rc = merge_begin(ctx, s)) return rc;

    cudaLaunchAttribute pdl;
    pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl.val.programmaticStreamSerializationAllowed = ctx->use_pdl ? 1 : 0;
    cudaLaunchConfig_t cfg{};
    cfg.stream = s; cfg.attrs = &pdl; cfg.numAttrs = 1; cfg.dynamicSmemBytes = 0;
    MergeDims *no_dims = nullptr;
    const unsigned long long cap = ctx->max_keys;
    // my buffer of this parity was last read by the peers for epoch-2
    if (epoch > 2) {
        k_peer_sync<<<1, 32, 0, s>>>(H, kNoFlag, 0u, 2u, epoch - 2u, no_dims, no_dims, 0u, cap, ctx->d_merge_error);
        CK_LAUNCH();
    }
    unsigned char *base = ctx->d_publish[parity];
    cfg.gridDim = dim3((unsigned)ctx->sm_count * 4u); cfg.blockDim = dim3(256);
    CK(cudaLaunchKernelEx(&cfg, k_publish_keys, (const unsigned long long *)ctx->d_keys[0], (const unsigned long long *)ctx->d_keys[1],
                          (const uint32_t *)&ctx->d_sc->result_buf, ctx->last_sorted ? 1u : 0u, (const unsigned long long *)&ctx->d_ctl->key_total, cap,
                          reinterpret_cast<unsigned long long *>(base), reinterpret_cast<unsigned long long *>(base + P.samples_off),
                          (const uint32_t *)ctx->d_counts, n_views, H.h[rank]));
    CK_LAUNCH();
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(32);
    CK(cudaLaunchKernelEx(&cfg, k_peer_sync, H, 0u, epoch, 0u, epoch, ctx->d_mdims, ctx->d_mdims, n_views, cap, ctx->d_merge_error));
    CK_LAUNCH();
    uint32_t *own_cuts = const_cast<uint32_t *>(L.cuts[rank]);
    k_merge_cuts<<<(uint32_t)ctx->sm_count * 8u, 256, 0, s>>>(L, rank, own_cuts, 0);
    CK_LAUNCH();
    CK(cudaLaunchKernelEx(&cfg, k_peer_sync, H, 1u, epoch, 1u, epoch, no_dims, ctx->d_mdims, 0u, cap, ctx->d_merge_error));
    CK_LAUNCH();
    if (int rc = merge_tail(ctx, L, 0, out_cap, slice_rank, slice_world, n_samples_cap, s)) return rc;
    k_peer_sync<<<1, 32, 0, s>>>(H, 2u, epoch, kNoFlag, 0u, no_dims, no_dims, 0u, cap, ctx->d_merge_error);
    CK_LAUNCH();
    CK(cudaEventRecord(ctx->ev_merge1, s));
    ctx->merge_p2p_flags = true;
    return ASTRE_OK;
}

void *astre_gpu_device_merged_keys(astre_gpu_ctx *ctx) { return ctx ? ctx->d_merged_keys : nullptr; }
void *astre_gpu_device_merged_sources(astre_gpu_ctx *ctx) { return ctx ? ctx->d_merged_src : nullptr; }

// Number of keys the last merge kept (synchronises when the sizes were computed on the device).
static int merged_count(astre_gpu_ctx *ctx, uint64_t *n_out)
{
    if (!ctx->merge_done) return fail(ctx, ASTRE_ERR_STATE, "no merge has been run");
    if (ctx->merge_dev) {
        CK(cudaStreamSynchronize(ctx->merge_stream));
        if (ctx->d_merge_error) {
            uint32_t err = 0;
            CK(cudaMemcpy(&err, ctx->d_merge_error, sizeof err, cudaMemcpyDeviceToHost));
            if (err) {
                CK(cudaMemset(ctx->d_merge_error, 0, sizeof err));
                return fail(ctx, ASTRE_ERR_STATE, "peer %u did not signal within the time limit: the merged list is invalid", err - 1u);
            }
        }
        MergeDims d;
        CK(cudaMemcpy(&d, ctx->d_mdims, sizeof d, cudaMemcpyDeviceToHost));
        uint64_t first = 0, count = d.total;
        if (ctx->merge_slice_world) {
            const uint64_t per = (d.total + ctx->merge_slice_world - 1) / ctx->merge_slice_world;
            first = std::min<uint64_t>(per * ctx->merge_slice_rank, d.total);
            count = std::min<uint64_t>(per, d.total - first);
        }
        *n_out = std::min<uint64_t>(count, ctx->n_merged);     // n_merged holds the capacity given to the merge
    } else {
        *n_out = ctx->n_merged;
    }
    return ASTRE_OK;
}

int astre_gpu_merged_count(astre_gpu_ctx *ctx, uint64_t *n_out)
{
    if (!ctx || !n_out) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    return merged_count(ctx, n_out);
}

int astre_gpu_read_merged(astre_gpu_ctx *ctx, uint64_t first, uint64_t n, uint64_t *keys, uint8_t *sources)
{
    if (!ctx) return ASTRE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    uint64_t have = 0;
    if (int rc = merged_count(ctx, &have)) return rc;
    if (first + n > have) return fail(ctx, ASTRE_ERR_INVALID, "range [%llu, %llu) exceeds the %llu merged keys",
                                      (unsigned long long)first, (unsigned long long)(first + n), (unsigned long long)have);
    CK(cudaStreamSynchronize(ctx->merge_stream));
    if (n && keys) CK(cudaMemcpy(keys, ctx->d_merged_keys + first, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (n && sources) CK(cudaMemcpy(sources, ctx->d_merged_src + first, n, cudaMemcpyDeviceToHost));
    return ASTRE_OK;
}

int astre_gpu_last_merge_ms(astre_gpu_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return ASTRE_ERR_INVALID;
    if (!ctx->merge_done) return fail(ctx, ASTRE_ERR_STATE, "no merge has been run");
    CK(cudaEventSynchronize(ctx->ev_merge1));
    CK(cudaEventElapsedTime(ms, ctx->ev_merge0, ctx->ev_merge1));
    return ASTRE_OK;
}

// ---- host-side helpers -------------------------------------------------------

void astre_camera_matrices(const float *pos3, const float *fwd3, const float *up3, float fov_deg, float aspect,
                           float z_near, float z_far, float *view16, float *proj16, float *clip16)
{
    using namespace astre_host;
    float view[16], proj[16];
    const V3 eye = { pos3[0], pos3[1], pos3[2] };
    const V3 center = { pos3[0] + fwd3[0], pos3[1] + fwd3[1], pos3[2] + fwd3[2] };
    look_at(eye, center, V3{ up3[0], up3[1], up3[2] }, view);
    perspective(radians(fov_deg), aspect, z_near, z_far, proj);
    if (view16) std::memcpy(view16, view, sizeof view);
    if (proj16) std::memcpy(proj16, proj, sizeof proj);
    if (clip16) mat_mul(proj, view, clip16);
}

int astre_light_space_matrix(const float *pos3, const float *dir3, int type, float outer_cutoff, float *clip16)
{
    using namespace astre_host;
    float view[16], proj[16];
    const V3 eye = { pos3[0], pos3[1], pos3[2] };
    const V3 d = normalize(V3{ dir3[0], dir3[1], dir3[2] });
    look_at(eye, V3{ eye.x + d.x, eye.y + d.y, eye.z + d.z }, V3{ 0.0f, 1.0f, 0.0f }, view);
    if (type == 1) {
        ortho(-100.0f, 100.0f, -100.0f, 100.0f, 0.1f, 100.0f, proj);
    } else if (type == 3) {
        const float outer = clampf(outer_cutoff, -1.0f, 1.0f);
        float fov = degrees(2.0f * std::acos(clampf(outer, -0.999f, 0.999f)));
        fov = clampf(fov, 5.0f, 179.0f);
        perspective(radians(fov), 1.7f, 0.1f, 1024.0f, proj);
    } else if (type == 2) {
        perspective(radians(90.0f), 1.0f, 0.1f, 100.0f, proj);
    } else {
        return 0;
    }
    mat_mul(proj, view, clip16);
    return 1;
}

void astre_ortho_view_matrix(const float *eye3, const float *center3, const float *up3, float l, float r, float b,
                             float t, float n, float f, float *clip16)
{
    using namespace astre_host;
    float view[16], proj[16];
    look_at(V3{ eye3[0], eye3[1], eye3[2] }, V3{ center3[0], center3[1], center3[2] }, V3{ up3[0], up3[1], up3[2] }, view);
    ortho(l, r, b, t, n, f, proj);
    mat_mul(proj, view, clip16);
}

void astre_global_offsets(const uint32_t *all_counts, uint32_t world_size, uint32_t rank, uint32_t n_views,
                          uint64_t *offsets, uint64_t *view_totals)
{
    uint64_t run = 0;
    for (uint32_t v = 0; v < n_views; ++v) {
        uint64_t total = 0;
        for (uint32_t r = 0; r < world_size; ++r) {
            if (r == rank && offsets) offsets[v] = run + total;
            total += all_counts[(size_t)r * n_views + v];
        }
        if (view_totals) view_totals[v] = total;
        run += total;
    }
}

void astre_global_offsets_rank_major(const uint32_t *all_counts, uint32_t world_size, uint32_t rank, uint32_t n_units,
                                     uint64_t *offsets, uint64_t *segment_len)
{
    uint64_t run = 0;
    for (uint32_t r = 0; r < rank && r < world_size; ++r)
        for (uint32_t u = 0; u < n_units; ++u) run += all_counts[(size_t)r * n_units + u];
    for (uint32_t u = 0; u < n_units; ++u) {
        const uint64_t c = all_counts[(size_t)rank * n_units + u];
        if (offsets) offsets[u] = run;
        if (segment_len) segment_len[u] = c;
        run += c;
    }
}

}  // extern "C"

// K1+K2: fused transform + frustum cull + stream compaction + key emission over a
// slot range (the whole mirror for a flat scene, one hierarchy level otherwise).
//
// Work decomposition: persistent grid (2 CTAs per SM), a block tile is 1024
// consecutive slots, a warp tile 128, and lane L owns slots 4L..4L+3 of its warp
// tile.  Warps are independent: there is no block-wide barrier inside the loop,
// and warp tiles are not bound to warps -- every warp draws its next tile from a
// work counter (global ones for flat frames and batched worlds, a per-CTA one for
// hierarchy levels; see the level loop), which removed the long tail the fixed
// tile = blockIdx + k * grid mapping ended a launch with.  Per warp tile:
//
//  * input: ONE cp.async.bulk (TMA, SASS UBLKCP) of the tile record (5632 or
//    6144 B, kernels_transform.cuh) into the warp's shared buffer, completion on
//    a per-warp mbarrier.  The copy for the warp's next tile is issued as soon as
//    the current one has been read into registers (11 LDS.128 per lane), so HBM
//    latency hides behind a whole tile of arithmetic (1-deep pipeline per warp);
//  * transform: GLM-exact TRS (astre_math.cuh), optional parent product;
//  * output: matrices are staged two entities per lane at a time (lane stride of
//    9 float4: conflict-free STS.128 / LDS.128) and drained with full-line 128-bit
//    streaming stores;
//  * cull, step 1 (all entities x all views): box-vs-box, six compares against
//    constant-bank operands per pair.  What survives is rare and scattered over
//    the lanes, so
//  * cull, step 2 runs on COMPACTED work: survivors are queued as (entity, view)
//    tasks (warp ballot + prefix popc), the entity bounds wait in the now-idle
//    staging buffer, and every 32 tasks one lane-per-task round does the six
//    exact plane tests.  A view that keeps a large part of the tile (>= 40 of
//    128) skips the queue and is tested four entities per lane instead;
//  * compaction: keys are appended to a per-warp shared buffer (ballot + prefix
//    popc) and flushed with one global reservation and coalesced 8-byte stores;
//    the flush also counts every key into its bucket of the key sort (one RED).
#pragma once
#include "kernels_transform.cuh"

namespace astre {

constexpr int kInBytesFlat = 11 * kWarpTile * 4;           // 5632: planes px..meta
constexpr int kInBytesHier = 12 * kWarpTile * 4;           // 6144: + parent
constexpr int kStageBytes = 32 * 9 * 16;                   // 4608: half-tile staging, later 9 words x 128 entities
constexpr int kKeyCap = 144;                               // per-warp key buffer (>= 128 + slack)
constexpr int kQueueCap = 80;                              // per-warp task queue (< 32 + kDenseMin entries live)
constexpr int kDenseMin = 40;                              // survivors per view and warp tile that take the dense path
constexpr int kWorldCacheViews = 1;                        // batched worlds: view records of the warp's world, fetched with the tile
constexpr int kWarpSmemBytes = kInBytesHier + kStageBytes + kKeyCap * 8 + kQueueCap * 2 + 2 * kWorldCacheViews * 256;   // 12576 (cache is double-buffered)
constexpr int kFusedSmemBytes = kWarpsPerBlock * kWarpSmemBytes;                           // 100608
constexpr bool kL2Ahead = false;        // cp.async.bulk.prefetch.L2 of the tile after next: measured r1, no change (0.3376 vs 0.338 ms)
                                        // -- the 1-deep smem pipeline already hides read latency; the ceiling is resident threads per SM
constexpr uint32_t kSmemMeshes = 64;    // bounds tables up to this size are served from shared memory
constexpr uint32_t kSmemViews = 12;     // views up to this index can take the task path

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// Hint the L2 to fetch a tile record ahead of the bulk copy that will need it (no shared memory involved).
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// Slot ranges one launch walks in order.  A flat frame has one range; a hierarchy whose levels are
// contiguous slot ranges passes them all and the kernel (cooperative launch, every CTA resident)
// separates them with a grid-wide barrier instead of one launch per level.
constexpr int kMaxLevelsPerLaunch = 32;
struct LevelTable {
    uint32_t n;
    uint32_t barrier_base;                       // barrier arrivals already counted in this frame
    uint32_t claim_base;                         // first work counter of this launch (level l uses the next n_groups)
    uint32_t first[kMaxLevelsPerLaunch], end[kMaxLevelsPerLaunch];
};

__device__ __forceinline__ void grid_barrier(uint32_t *counter, uint32_t target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        uint32_t v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// View records in shared memory (the warp's world, batched worlds).
struct SmemViews {
    const ViewDev *v;
    __device__ __forceinline__ float4 plane(uint32_t i, int p) const { return v[i].plane[p]; }
    __device__ __forceinline__ float4 aplane(uint32_t i, int p) const { return v[i].aplane[p]; }
    __device__ __forceinline__ float4 depth(uint32_t i) const { return v[i].depth; }
    __device__ __forceinline__ float depth_scale(uint32_t i) const { return v[i].depth_scale; }
    __device__ __forceinline__ uint32_t mask(uint32_t i) const { return v[i].phase_mask; }
    __device__ __forceinline__ uint32_t sym(uint32_t i) const { return v[i].sym; }
    __device__ __forceinline__ float4 lo(uint32_t i) const { return v[i].lo; }
    __device__ __forceinline__ float4 hi(uint32_t i) const { return v[i].hi; }
};

// The registers one lane holds of its warp tile: four entities per component.
struct LaneInputs {
    float4 px, py, pz, qx, qy, qz, qw, sx, sy, sz;
    uint4 meta, parent;
};

// Per-warp output side: the key buffer and where it flushes to.
struct KeySink {
    unsigned long long *kbuf;     // shared, kKeyCap entries
    uint32_t fill;                // warp-uniform
    const FrameParams *P;
    const BinPlan *bp;            // bucket digit of the key sort
    int lane;

    __device__ __forceinline__ void flush()
    {
        __syncwarp();
        if (fill) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&P->ctl->key_total, (unsigned long long)fill);
            base = __shfl_sync(0xffffffffu, base, 0);
            const bool count_bins = bp->bits != 0;
            for (uint32_t i = lane; i < fill; i += 32) {
                const unsigned long long k = kbuf[i], pos = base + i;
                if (pos < P->key_cap) {
                    P->keys[pos] = k;
                    if (count_bins) atomicAdd(&P->bins[bin_digit(*bp, k)], 1u);   // the sort's bucket counts, for free (RED)
                } else {
                    P->ctl->overflow = 1u;
                }
            }
        }
        __syncwarp();
        fill = 0;
    }
    // warp-collective: every lane calls it; lanes with `on` contribute `key`
    __device__ __forceinline__ uint32_t append(bool on, unsigned long long key)
    {
        const uint32_t m = __ballot_sync(0xffffffffu, on);
        const uint32_t c = __popc(m);
        if (c) {
            if (fill + c > (uint32_t)kKeyCap) flush();
            if (on) kbuf[fill + __popc(m & ((1u << lane) - 1u))] = key;
            fill += c;
        }
        return c;
    }
};

// Transform + stage + drain of one warp tile (and the world-space bounds when
// culling).  FULL: the whole block tile lies inside [first, end), no range checks.
template <bool HIER, bool WORLDS, bool CULL, bool FULL>
__device__ __forceinline__ void transform_tile(const FrameParams &P, const LaneInputs &in, float4 *stage,
                                               const BoundsDev *s_bounds, bool small_table,
                                               uint32_t wbase, uint32_t first, uint32_t end, uint32_t world, int lane,
                                               WorldBounds (&wb)[kPerLane], uint32_t (&phases)[kPerLane],
                                               uint32_t (&meta)[kPerLane], uint32_t (&local)[kPerLane])
{
    const uint32_t s0 = wbase + lane * kPerLane;
    float4 *out = P.world + (size_t)wbase * 4;

#pragma unroll
    for (int k = 0; k < kPerLane; ++k) {
        const uint32_t slot = s0 + k;
        const bool active = FULL || (slot >= first && slot < end);
        float m[16];
        float4 *dst = stage + lane * 9 + (k & 1) * 4;
        trs_matrix(f4get(in.px, k), f4get(in.py, k), f4get(in.pz, k),
                   f4get(in.qx, k), f4get(in.qy, k), f4get(in.qz, k), f4get(in.qw, k),
                   f4get(in.sx, k), f4get(in.sy, k), f4get(in.sz, k), m, dst);
        if (HIER) {
            const uint32_t par = u4get(in.parent, k);
            if (active && par != kNoParent) {
                // L2 loads: the parent row was written by another SM (previous level of the same launch, or
                // an earlier launch) and may share a 128-byte line with a matrix this SM read before it existed
                const float4 *pw = P.world + (size_t)par * 4;
                const float4 a0 = __ldcg(pw), a1 = __ldcg(pw + 1), a2 = __ldcg(pw + 2), a3 = __ldcg(pw + 3);
                const float a[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w,
                                     a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w};
                float w[16];
                mat4_mul(a, m, w);
#pragma unroll
                for (int i = 0; i < 16; ++i) m[i] = w[i];
            }
        }
        dst[0] = make_float4(m[0], m[1], m[2], m[3]);
        dst[1] = make_float4(m[4], m[5], m[6], m[7]);
        dst[2] = make_float4(m[8], m[9], m[10], m[11]);
        dst[3] = make_float4(m[12], m[13], m[14], m[15]);

        if (CULL) {
            meta[k] = u4get(in.meta, k);
            local[k] = WORLDS ? slot - world * P.entities_per_world : slot;
            const uint32_t flags = (meta[k] >> 19) & 31u;
            const uint32_t mesh = meta[k] & 2047u;
            const bool cand = active && (flags & 1u) && mesh < P.n_meshes;
            phases[k] = cand ? (flags >> 1) : 0u;
            const uint32_t bi = cand ? mesh : 0u;
            wb[k] = world_bounds(m, small_table ? s_bounds[bi] : P.bounds[bi]);
        }

        if (k & 1) {
            // Drain this half.  Logical float4 j = 8*lane' + t: entity 4*lane' + (k-1) + (t>>2), column t&3.
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int j = i * 32 + lane;
                const int e = 4 * (j >> 3) + (k - 1) + ((j & 7) >> 2);
                const float4 v = stage[(j >> 3) * 9 + (j & 7)];
                if (FULL || (wbase + e >= first && wbase + e < end)) {
                    if (HIER) out[e * 4 + (j & 3)] = v; else __stcs(out + e * 4 + (j & 3), v);
                }
            }
            __syncwarp();
        }
    }
}

// Entity bounds parked in the staging buffer: word w of entity (lane L, k) at sb[w*128 + k*32 + L].
__device__ __forceinline__ WorldBounds load_bounds(const float *sb, int idx)
{
    WorldBounds w;
    w.cx = sb[0 * 128 + idx]; w.cy = sb[1 * 128 + idx]; w.cz = sb[2 * 128 + idx];
    w.ex = sb[3 * 128 + idx]; w.ey = sb[4 * 128 + idx]; w.ez = sb[5 * 128 + idx];
    w.r = sb[6 * 128 + idx];
    return w;
}

template <bool HIER, bool WORLDS, bool CULL>
__global__ void __launch_bounds__(kBlockThreads, 2)
k_frame_fused(const FrameParams P, const __grid_constant__ ViewSet VS, const __grid_constant__ BinPlan bp,
              const __grid_constant__ LevelTable LT)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ __align__(16) BoundsDev s_bounds[CULL ? kSmemMeshes : 1];
    __shared__ __align__(16) ViewDev s_views[(CULL && !WORLDS) ? kSmemViews : 1];
    __shared__ __align__(8) unsigned long long s_bar[kWarpsPerBlock];
    __shared__ uint32_t s_counts[32];
    __shared__ uint32_t s_claim;                                  // next unclaimed warp tile of this CTA's list

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *s_in = s_dyn + warp * kWarpSmemBytes;
    float4 *stage = reinterpret_cast<float4 *>(s_in + kInBytesHier);
    float *sb = reinterpret_cast<float *>(stage);
    unsigned long long *kbuf = reinterpret_cast<unsigned long long *>(s_in + kInBytesHier + kStageBytes);
    unsigned short *queue = reinterpret_cast<unsigned short *>(s_in + kInBytesHier + kStageBytes + kKeyCap * 8);
    const ViewDev *vcache2 = reinterpret_cast<const ViewDev *>(s_in + kInBytesHier + kStageBytes + kKeyCap * 8 + kQueueCap * 2);
    // Batched worlds whose size is a multiple of the warp tile: all 128 entities of a warp tile share one world,
    // so its view records ride along with the tile record (same mbarrier) instead of being read from L2 per plane.
    const bool wcache = WORLDS && P.world_cache != 0;
    const uint32_t bar = smem_u32(&s_bar[warp]);
    const uint32_t s_in_addr = smem_u32(s_in);
    const uint32_t F = P.views_per_world;
    const bool small_table = P.n_meshes <= kSmemMeshes;
    constexpr uint32_t kInBytes = HIER ? kInBytesHier : kInBytesFlat;
    const uint32_t lt_mask = (1u << lane) - 1u;

    if (threadIdx.x < 32) s_counts[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_claim = kWarpsPerBlock;               // warp w starts with warp tile w
    if (CULL && small_table) {
        for (uint32_t i = threadIdx.x; i < P.n_meshes * 2; i += kBlockThreads)   // 32-byte records, two float4 each
            reinterpret_cast<float4 *>(s_bounds)[i] = __ldg(reinterpret_cast<const float4 *>(P.bounds) + i);
    }
    if (CULL && !WORLDS) {
        const uint32_t nv = F < kSmemViews ? F : kSmemViews;
        for (uint32_t i = threadIdx.x; i < nv * 16; i += kBlockThreads)          // 256-byte records, 16 float4 each
            reinterpret_cast<float4 *>(s_views)[i] = reinterpret_cast<const float4 *>(VS.v)[i];
    }
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    KeySink sink{ kbuf, 0u, &P, &bp, lane };
    uint32_t parity = 0;
    // Programmatic dependent launch (single-range launches): everything above overlaps the tail of the frame's
    // reset kernel; nothing that kernel writes (work counters, key total, counts, bins) is touched before here.
    cudaGridDependencySynchronize();

    for (uint32_t level = 0; level < LT.n; ++level) {
    const uint32_t first = LT.first[level], end = LT.end[level];
    const uint32_t base0 = first & ~(uint32_t)(kWarpTile - 1);
    const uint32_t n_tiles = (end - base0 + kBlockTile - 1) / kBlockTile;

    // Work distribution.  Warp tiles are not bound to warps:
    //  * one flat range (the common case): CTA b belongs to group b % n_groups; group g owns block tiles g,
    //    g + n_groups, ... and all warps of the group take the warp tiles of that list from one global counter, two
    //    claims ahead so that the atomic's latency never shows.  Measured on config 3 (ms for transform + cull):
    //    fixed mapping 0.334, per-CTA shared counter 0.319, groups of 2 CTAs 0.315, 4: 0.308, 8: 0.307,
    //    37: 0.304, 74: 0.303, one counter for the grid 0.307 -> 8 groups;
    //    batched worlds the same (config 5: 1.439 -> 1.391 ms);
    //  * hierarchy levels: the CTA owns block tiles blockIdx.x, blockIdx.x + gridDim.x, ... and its 8 warps take the
    //    warp tiles of that list from a shared-memory counter (global counters cost two exposed atomics at the start
    //    of every level: config 2 0.140 -> 0.178 ms).
    // A warp that falls behind therefore does not keep its CTA -- and its SM -- waiting at the end of the launch.
    // (the counters are zeroed once per frame: every level of every launch of a frame gets its own set)
    const uint32_t want_groups = min(min(P.claim_groups, (uint32_t)kClaimGroups), gridDim.x);
    const bool gclaim = want_groups != 0 && !HIER &&
                        LT.claim_base + (level + 1) * want_groups <= (uint32_t)kClaimGroups;
    const uint32_t n_groups = gclaim ? want_groups : gridDim.x;
    const uint32_t group = blockIdx.x % n_groups;
    const uint32_t n_wtiles = group < n_tiles ? ((n_tiles - group + n_groups - 1) / n_groups) * kWarpsPerBlock : 0u;
    uint32_t *gcounter = &P.ctl->claim[(gclaim ? LT.claim_base + level * n_groups + group : 0u) * kClaimStride];
    auto wtile_base = [&](uint32_t j) {
        return base0 + (group + (j / kWarpsPerBlock) * n_groups) * kBlockTile + (j % kWarpsPerBlock) * kWarpTile;
    };
    // lane 0: one bulk copy of the warp tile's record (+ its world's view records into cache half `half`)
    auto issue = [&](uint32_t j, uint32_t half) {
        const uint32_t wb0 = wtile_base(j);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (wcache) {
            const uint32_t w = min(wb0 / P.entities_per_world, P.n_worlds - 1u);
            mbar_expect_tx(bar, kInBytes + F * 256u);
            bulk_g2s(s_in_addr, P.tiles + (size_t)(wb0 >> 7) * kTileFloats, kInBytes, bar);
            bulk_g2s(smem_u32(vcache2 + half * kWorldCacheViews), P.views + (size_t)w * F, F * 256u, bar);
        } else {
            mbar_expect_tx(bar, kInBytes);
            bulk_g2s(s_in_addr, P.tiles + (size_t)(wb0 >> 7) * kTileFloats, kInBytes, bar);
        }
    };
    uint32_t j = (uint32_t)warp, jn = 0, jnn_lane0 = 0;
    if (gclaim) {
        uint32_t a = 0, b = 0;
        if (lane == 0) { a = atomicAdd(gcounter, 1u); b = atomicAdd(gcounter, 1u); }
        j = __shfl_sync(0xffffffffu, a, 0);
        jn = __shfl_sync(0xffffffffu, b, 0);
    }
    if (j < n_wtiles && lane == 0) issue(j, parity);
    auto advance = [&]() {
        j = jn;
        if (gclaim) jn = __shfl_sync(0xffffffffu, jnn_lane0, 0);      // claimed a whole tile ago
    };

    for (; j < n_wtiles; advance()) {
        const uint32_t wbase = wtile_base(j);
        const bool full = wbase >= first && wbase + kWarpTile <= end;   // warp-uniform

        mbar_wait(bar, parity);
        const ViewDev *vcache = vcache2 + parity * kWorldCacheViews;    // this tile's world (batched worlds)
        parity ^= 1u;
        LaneInputs in;
        {
            const float4 *in4 = reinterpret_cast<const float4 *>(s_in) + lane;
            in.px = in4[PX * 32]; in.py = in4[PY * 32]; in.pz = in4[PZ * 32];
            in.qx = in4[QX * 32]; in.qy = in4[QY * 32]; in.qz = in4[QZ * 32]; in.qw = in4[QW * 32];
            in.sx = in4[SX * 32]; in.sy = in4[SY * 32]; in.sz = in4[SZ * 32];
            in.meta = reinterpret_cast<const uint4 *>(s_in)[META * 32 + lane];
            in.parent = make_uint4(kNoParent, kNoParent, kNoParent, kNoParent);
            if (HIER) in.parent = reinterpret_cast<const uint4 *>(s_in)[PARENT * 32 + lane];
        }
        __syncwarp();   // every lane holds its inputs: the buffer may be refilled
        {
            if (gclaim) {
                if (lane == 0) {
                    if (jn < n_wtiles) issue(jn, parity);
                    jnn_lane0 = atomicAdd(gcounter, 1u);           // needed one tile from now
                }
            } else {
                uint32_t claimed = j + kWarpsPerBlock;              // batched worlds without counters: fixed mapping
                if (lane == 0) {
                    if (!WORLDS) claimed = atomicAdd(&s_claim, 1u);
                    if (claimed < n_wtiles) issue(claimed, parity);
                }
                jn = WORLDS ? claimed : __shfl_sync(0xffffffffu, claimed, 0);
            }
        }

        WorldBounds wb[kPerLane];
        uint32_t phases[kPerLane], meta[kPerLane], local[kPerLane];
        const uint32_t world = WORLDS ? min((wbase + lane * kPerLane) / P.entities_per_world, P.n_worlds - 1u) : 0u;
        if (full) transform_tile<HIER, WORLDS, CULL, true>(P, in, stage, s_bounds, small_table, wbase, first, end, world,
                                                           lane, wb, phases, meta, local);
        else transform_tile<HIER, WORLDS, CULL, false>(P, in, stage, s_bounds, small_table, wbase, first, end, world,
                                                       lane, wb, phases, meta, local);
        if (!CULL) continue;

        // ---- step 1 operands; park the bounds in the (drained) staging buffer ----
        WorldBox box[kPerLane];
#pragma unroll
        for (int k = 0; k < kPerLane; ++k) {
            box[k] = world_box(wb[k]);
            const int idx = k * 32 + lane;
            sb[0 * 128 + idx] = wb[k].cx; sb[1 * 128 + idx] = wb[k].cy; sb[2 * 128 + idx] = wb[k].cz;
            sb[3 * 128 + idx] = wb[k].ex; sb[4 * 128 + idx] = wb[k].ey; sb[5 * 128 + idx] = wb[k].ez;
            sb[6 * 128 + idx] = wb[k].r;
            sb[7 * 128 + idx] = __uint_as_float(meta[k]);
            sb[8 * 128 + idx] = __uint_as_float(local[k]);
        }
        __syncwarp();

        uint32_t qcount = 0;   // warp-uniform

        // One lane-per-task round: the top min(32, qcount) tasks of the queue.  View records come from
        // shared memory (one world) or, per task, from the task's world in global memory.
        auto task_round = [&]() {
            const uint32_t n = qcount < 32u ? qcount : 32u;
            const uint32_t qb = qcount - n;
            bool alive = (uint32_t)lane < n;
            uint32_t v = 0, idx = 0, tworld = 0;
            if (alive) {
                const uint32_t t = queue[qb + lane];
                v = t >> 8;
                idx = t & 127u;
                if (WORLDS) tworld = min((wbase + 4u * (idx & 31u) + (idx >> 5)) / P.entities_per_world, P.n_worlds - 1u);
            }
            const WorldBounds w = load_bounds(sb, idx);
            const bool from_global = WORLDS && !wcache;
            const ViewDev *vr = WORLDS ? (wcache ? &vcache[v] : P.views + (size_t)tworld * F + v) : &s_views[v];
#pragma unroll
            for (int p = 0; p < 6; ++p) {
                const float4 pl = from_global ? __ldg(&vr->plane[p]) : vr->plane[p];
                const float4 ap = from_global ? __ldg(&vr->aplane[p]) : vr->aplane[p];
                alive = alive && !(fadd(plane_dist(w, pl), plane_rad(w, ap)) < 0.0f);
            }
            unsigned long long key = 0;
            if (alive) {
                const uint32_t mt = __float_as_uint(sb[7 * 128 + idx]);
                const float4 dp = from_global ? __ldg(&vr->depth) : vr->depth;
                const float ds = from_global ? __ldg(&vr->depth_scale) : vr->depth_scale;
                key = make_key(P.world_bits, tworld, v, (mt >> 11) & 255u, mt & 2047u, depth14(w, dp, ds),
                               __float_as_uint(sb[8 * 128 + idx]));
                if (WORLDS) { if (!wcache) atomicAdd(&P.counts[(size_t)tworld * F + v], 1u); }
                else atomicAdd(&s_counts[v], 1u);
            }
            const uint32_t c = sink.append(alive, key);
            // one world per warp tile and a single cached view: one global atomic per round, not per key
            // (per-key REDs on a few thousand counters saturate the L2 atomic units and stall every load)
            if (WORLDS && wcache && lane == 0 && c) atomicAdd(&P.counts[(size_t)world * F], c);
            qcount = qb;
        };

        for (uint32_t v = 0; v < F; ++v) {
            bool pass[kPerLane];
            uint32_t bal[kPerLane], n_pass = 0;
            if (WORLDS && wcache) {
                const uint32_t mask = vcache[v].phase_mask;
                const float4 lo = vcache[v].lo, hi = vcache[v].hi;
#pragma unroll
                for (int k = 0; k < kPerLane; ++k) pass[k] = (phases[k] & mask) != 0 && box_overlaps(box[k], lo, hi);
            } else if (WORLDS) {
                const GlobalViews GV{ P.views + (size_t)world * F };
                const uint32_t mask = GV.mask(v);
                const float4 lo = GV.lo(v), hi = GV.hi(v);
#pragma unroll
                for (int k = 0; k < kPerLane; ++k) pass[k] = (phases[k] & mask) != 0 && box_overlaps(box[k], lo, hi);
            } else {
                const uint32_t mask = VS.v[v].phase_mask;
                const float4 lo = VS.v[v].lo, hi = VS.v[v].hi;
#pragma unroll
                for (int k = 0; k < kPerLane; ++k) pass[k] = (phases[k] & mask) != 0 && box_overlaps(box[k], lo, hi);
            }
#pragma unroll
            for (int k = 0; k < kPerLane; ++k) { bal[k] = __ballot_sync(0xffffffffu, pass[k]); n_pass += __popc(bal[k]); }
            if (n_pass == 0) continue;

            if (n_pass >= (uint32_t)kDenseMin || (!WORLDS && v >= kSmemViews)) {
                // dense: four entities per lane, warp-uniform plane loop
                WorldBounds w4[kPerLane];
#pragma unroll
                for (int k = 0; k < kPerLane; ++k) w4[k] = load_bounds(sb, k * 32 + lane);
                const GlobalViews GV{ P.views + (size_t)world * F };
                const ConstViews CV{ VS };
                const SmemViews SV{ vcache };
                float4 dp;
                float ds;
                if (WORLDS && wcache) { cull_planes<kPerLane>(SV, v, w4, pass); dp = SV.depth(v); ds = SV.depth_scale(v); }
                else if (WORLDS) { cull_planes<kPerLane>(GV, v, w4, pass); dp = GV.depth(v); ds = GV.depth_scale(v); }
                else { cull_planes<kPerLane>(CV, v, w4, pass); dp = CV.depth(v); ds = CV.depth_scale(v); }
                uint32_t view_total = 0;
#pragma unroll
                for (int k = 0; k < kPerLane; ++k) {
                    unsigned long long key = 0;
                    if (pass[k]) {
                        key = make_key(P.world_bits, world, v, (meta[k] >> 11) & 255u, meta[k] & 2047u,
                                       depth14(w4[k], dp, ds), local[k]);
                        if (WORLDS && !wcache) atomicAdd(&P.counts[(size_t)world * F + v], 1u);
                    }
                    view_total += sink.append(pass[k], key);
                }
                if (lane == 0 && view_total) {
                    if (!WORLDS) atomicAdd(&s_counts[v], view_total);
                    else if (wcache) atomicAdd(&P.counts[(size_t)world * F + v], view_total);
                }
            } else {
                // sparse: queue (entity, view) tasks; a full round runs as soon as 32 are waiting
                uint32_t off = qcount;
#pragma unroll
                for (int k = 0; k < kPerLane; ++k) {
                    if (pass[k]) queue[off + __popc(bal[k] & lt_mask)] = (unsigned short)((v << 8) | (uint32_t)(k * 32 + lane));
                    off += __popc(bal[k]);
                }
                qcount = off;
                __syncwarp();
                while (qcount >= 32u) task_round();
            }
        }
        while (qcount) task_round();
        __syncwarp();   // the staging buffer is reused by the next tile
    }
    // next level reads the matrices this one wrote: every CTA of the (cooperative) grid meets here
    if (level + 1 < LT.n) {
        grid_barrier(&P.ctl->barrier, LT.barrier_base + (level + 1) * gridDim.x);
        if (threadIdx.x == 0) s_claim = kWarpsPerBlock;
        __syncthreads();
    }
    }

    if (CULL) sink.flush();
    __syncthreads();
    if (CULL && !WORLDS && threadIdx.x < F && s_counts[threadIdx.x])
        atomicAdd(&P.counts[threadIdx.x], s_counts[threadIdx.x]);
}

}  // namespace astre

// K1+K2: fused transform + frustum cull + stream compaction + key emission over a
// slot range (the whole mirror for a flat scene, one hierarchy level otherwise).
//
// Work decomposition: persistent grid (2 CTAs per SM), a block tile is 1024
// consecutive slots, a warp tile 128, and lane L owns slots 4L..4L+3 of its warp
// tile.  Per warp tile:
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
//  * cull: views come from the constant bank (kernel parameter) for one world;
//    certain-outcome planes cost 4 FFMA + 2 compares per entity, the exact
//    support radius is only evaluated when a lane is undecided;
//  * compaction: warp ballot/scan -> block scan over 8 warp totals -> one global
//    reservation per block tile -> keys written at base + warp + lane offset.
#pragma once
#include "kernels_transform.cuh"

namespace astre {

constexpr int kInBytesFlat = 11 * kWarpTile * 4;           // 5632: planes px..meta
constexpr int kInBytesHier = 12 * kWarpTile * 4;           // 6144: + parent
constexpr int kHalfStageF4 = 32 * 9;                       // two entities per lane + one pad slot
constexpr int kWarpSmemBytes = kInBytesHier + kHalfStageF4 * 16;   // 10752
constexpr int kFusedSmemBytes = kWarpsPerBlock * kWarpSmemBytes;   // 86016
constexpr uint32_t kSmemMeshes = 256;   // bounds tables up to this size are served from shared memory

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

// The registers one lane holds of its warp tile: four entities per component.
struct LaneInputs {
    float4 px, py, pz, qx, qy, qz, qw, sx, sy, sz;
    uint4 meta, parent;
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
                const float4 *pw = P.world + (size_t)par * 4;
                const float4 a0 = pw[0], a1 = pw[1], a2 = pw[2], a3 = pw[3];
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

template <bool HIER, bool WORLDS, bool CULL>
__global__ void __launch_bounds__(kBlockThreads, 2)
k_frame_fused(const FrameParams P, const __grid_constant__ ViewSet VS, const uint32_t first, const uint32_t end)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ __align__(16) BoundsDev s_bounds[CULL ? kSmemMeshes : 1];
    __shared__ __align__(8) unsigned long long s_bar[kWarpsPerBlock];
    __shared__ uint32_t s_warp_cnt[kWarpsPerBlock];
    __shared__ unsigned long long s_block_base;
    __shared__ uint32_t s_counts[32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *s_in = s_dyn + warp * kWarpSmemBytes;
    float4 *stage = reinterpret_cast<float4 *>(s_in + kInBytesHier);
    const uint32_t bar = smem_u32(&s_bar[warp]);
    const uint32_t s_in_addr = smem_u32(s_in);
    const uint32_t F = P.views_per_world;
    const bool small_table = P.n_meshes <= kSmemMeshes;
    constexpr uint32_t kInBytes = HIER ? kInBytesHier : kInBytesFlat;

    if (threadIdx.x < 32) s_counts[threadIdx.x] = 0;
    if (CULL && small_table) {
        for (uint32_t i = threadIdx.x; i < P.n_meshes * 2; i += kBlockThreads)   // 32-byte records, two float4 each
            reinterpret_cast<float4 *>(s_bounds)[i] = __ldg(reinterpret_cast<const float4 *>(P.bounds) + i);
    }
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t base0 = first & ~(uint32_t)(kWarpTile - 1);
    const uint32_t n_tiles = (end - base0 + kBlockTile - 1) / kBlockTile;

    // lane 0: one bulk copy of the warp tile's record
    auto issue = [&](uint32_t tile) {
        const uint32_t wb0 = base0 + tile * kBlockTile + warp * kWarpTile;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, kInBytes);
        bulk_g2s(s_in_addr, P.tiles + (size_t)(wb0 >> 7) * kTileFloats, kInBytes, bar);
    };

    uint32_t parity = 0;
    if (blockIdx.x < n_tiles && lane == 0) issue(blockIdx.x);

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t tbase = base0 + tile * kBlockTile;
        const uint32_t wbase = tbase + warp * kWarpTile;
        const bool full = tbase >= first && tbase + kBlockTile <= end;   // block-uniform

        mbar_wait(bar, parity);
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
            const uint32_t next = tile + gridDim.x;
            if (next < n_tiles && lane == 0) issue(next);
        }

        WorldBounds wb[kPerLane];
        uint32_t phases[kPerLane], vis[kPerLane], meta[kPerLane], local[kPerLane];
        const uint32_t world = WORLDS ? min((wbase + lane * kPerLane) / P.entities_per_world, P.n_worlds - 1u) : 0u;
        if (full) transform_tile<HIER, WORLDS, CULL, true>(P, in, stage, s_bounds, small_table, wbase, first, end, world,
                                                           lane, wb, phases, meta, local);
        else transform_tile<HIER, WORLDS, CULL, false>(P, in, stage, s_bounds, small_table, wbase, first, end, world,
                                                       lane, wb, phases, meta, local);

        if (CULL) {
#pragma unroll
            for (int k = 0; k < kPerLane; ++k) vis[k] = 0;
            const GlobalViews GV{ P.views + (size_t)world * F };
            const ConstViews CV{ VS };
            if (WORLDS) cull_views<kPerLane>(GV, F, wb, phases, vis);
            else cull_views<kPerLane>(CV, F, wb, phases, vis);

            // block-scan compaction
            uint32_t cnt = 0;
#pragma unroll
            for (int k = 0; k < kPerLane; ++k) cnt += __popc(vis[k]);
            const uint32_t incl = warp_incl_scan(cnt, lane);
            const uint32_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 31) s_warp_cnt[warp] = incl;
            __syncthreads();
            uint32_t block_total = 0, warp_off = 0;
#pragma unroll
            for (int w = 0; w < kWarpsPerBlock; ++w) {
                const uint32_t c = s_warp_cnt[w];
                if (w < warp) warp_off += c;
                block_total += c;
            }
            if (threadIdx.x == 0 && block_total)
                s_block_base = atomicAdd(&P.ctl->key_total, (unsigned long long)block_total);
            __syncthreads();
            if (warp_total) {
                // counts: a few keys -> one shared atomic per key; many -> warp reduction per view
                const bool sparse = !WORLDS && warp_total <= 16u;
                if (!sparse) add_counts<kPerLane>(P, world, vis, s_counts, lane);
                if (cnt) {
                    const unsigned long long pos = s_block_base + warp_off + (incl - cnt);
                    if (WORLDS) emit_keys<kPerLane>(P, GV, world, wb, vis, meta, local, pos, nullptr);
                    else emit_keys<kPerLane>(P, CV, world, wb, vis, meta, local, pos, sparse ? s_counts : nullptr);
                }
            }
        }
    }

    __syncthreads();
    if (CULL && !WORLDS && threadIdx.x < F && s_counts[threadIdx.x])
        atomicAdd(&P.counts[threadIdx.x], s_counts[threadIdx.x]);
}

}  // namespace astre

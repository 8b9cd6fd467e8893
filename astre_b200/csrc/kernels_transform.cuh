// K1+K2: fused transform + frustum cull + stream compaction + key emission.
//
// Replaces the per-entity loops of TransformSystem::run
// (engine/modules/ECS/src/system/transform_system.cpp:21-67), VisualSystem::run
// (engine/modules/ECS/src/system/visual_system.cpp:20-58) and the per-pass
// filters of deferredShadingStage (engine/modules/Pipeline/src/deferred_shading.cpp:115-121,151-156).
//
// Data layout in HBM (the device mirror of the component storages):
//   px,py,pz,qx,qy,qz,qw,sx,sy,sz  one float plane per component, index = slot
//   meta                           u32 per slot: mesh[10:0] | shader[18:11] | flags[23:19]
//   parent                         u32 per slot (ASTRE_NO_PARENT = root)
//   world                          float4[slot*4 + column], GLM column-major mat4
// Every plane is padded to a multiple of kBlockTile entities so that the
// 128-bit loads of a partial tile stay inside the allocation.
//
// Work decomposition: a block tile is 1024 consecutive slots, a warp tile 128,
// and lane L owns slots 4L..4L+3 of its warp tile, so every component is read
// with one coalesced 128-bit load per lane (512 B per warp instruction).  The
// four matrices a lane produces (256 B) are staged in shared memory with a
// 16-byte pad after every 256 B, which makes both the lane-strided STS.128
// and the drain's linear LDS.128 conflict-free; the drain then writes the warp
// tile's 8 KB with fully coalesced 128-bit stores.
//
// Algorithmic HBM bytes per entity: 40 (TRS) + 4 (meta) + 64 (world) = 108,
// + 4 + 64 for a non-root entity (parent index + parent matrix), + 8 per key.
#pragma once
#include "astre_math.cuh"

namespace astre {

constexpr int kBlockThreads = 256;
constexpr int kWarpsPerBlock = kBlockThreads / 32;
constexpr int kPerLane = 4;
constexpr int kWarpTile = 32 * kPerLane;               // 128 entities
constexpr int kBlockTile = kWarpsPerBlock * kWarpTile; // 1024 entities
constexpr int kStageF4PerWarp = kWarpTile * 4 + kWarpTile / 4;  // 512 + 32 pad slots
constexpr uint32_t kNoParent = 0xFFFFFFFFu;

// One view on the device (224 B, 16-byte aligned).
struct ViewDev {
    float4 plane[6];   // (nx,ny,nz,w), not normalised: left,right,bottom,top,near,far
    float4 aplane[6];  // (|nx|,|ny|,|nz|,len)
    float4 depth;      // normalised near plane
    float depth_scale;
    uint32_t phase_mask;
    uint32_t pad[2];
};
static_assert(sizeof(ViewDev) == 224, "ViewDev layout");

struct BoundsDev {   // 32 B
    float cx, cy, cz, ex;
    float ey, ez, r;
    uint32_t reserved;
};

struct FrameControl {            // one per context, zeroed at the start of a frame
    unsigned long long key_total; // keys reserved so far (may exceed capacity)
    uint32_t overflow;
    uint32_t pad;
};

struct FrameParams {
    const float *px, *py, *pz, *qx, *qy, *qz, *qw, *sx, *sy, *sz;
    const uint32_t *meta;
    const uint32_t *parent;
    float4 *world;
    const BoundsDev *bounds;
    const ViewDev *views;          // [n_worlds][views_per_world]
    unsigned long long *keys;
    uint32_t *counts;              // [n_worlds*views_per_world]
    FrameControl *ctl;
    unsigned long long key_cap;
    uint32_t n_meshes;
    uint32_t views_per_world;      // 0 = no culling
    uint32_t n_worlds;
    uint32_t entities_per_world;
    uint32_t world_bits;
};

__device__ __forceinline__ float f4get(const float4 &v, int k)
{
    return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}
__device__ __forceinline__ uint32_t u4get(const uint4 &v, int k)
{
    return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

__device__ __forceinline__ float4 ld_stream(const float *p)
{
    return __ldcs(reinterpret_cast<const float4 *>(p));
}

// World-space bounds (spec; mirrored by the CPU checker).
struct WorldBounds { float cx, cy, cz, ex, ey, ez, r; };

__device__ __forceinline__ WorldBounds world_bounds(const float (&m)[16], const BoundsDev &b)
{
    WorldBounds w;
    w.cx = ffma(m[0], b.cx, ffma(m[4], b.cy, ffma(m[8], b.cz, m[12])));
    w.cy = ffma(m[1], b.cx, ffma(m[5], b.cy, ffma(m[9], b.cz, m[13])));
    w.cz = ffma(m[2], b.cx, ffma(m[6], b.cy, ffma(m[10], b.cz, m[14])));
    w.ex = ffma(fabsf(m[0]), b.ex, ffma(fabsf(m[4]), b.ey, fmul(fabsf(m[8]), b.ez)));
    w.ey = ffma(fabsf(m[1]), b.ex, ffma(fabsf(m[5]), b.ey, fmul(fabsf(m[9]), b.ez)));
    w.ez = ffma(fabsf(m[2]), b.ex, ffma(fabsf(m[6]), b.ey, fmul(fabsf(m[10]), b.ez)));
    const float l0 = ffma(m[0], m[0], ffma(m[1], m[1], fmul(m[2], m[2])));
    const float l1 = ffma(m[4], m[4], ffma(m[5], m[5], fmul(m[6], m[6])));
    const float l2 = ffma(m[8], m[8], ffma(m[9], m[9], fmul(m[10], m[10])));
    w.r = fmul(b.r, __fsqrt_rn(fmaxf(fmaxf(l0, l1), l2)));
    return w;
}

// 1 if the plane rejects the bounds: dist + rad < 0 (false on NaN).
__device__ __forceinline__ bool plane_rejects(const WorldBounds &w, const float4 &pl, const float4 &ap)
{
    const float dist = ffma(pl.x, w.cx, ffma(pl.y, w.cy, ffma(pl.z, w.cz, pl.w)));
    const float rad = ffma(ap.x, w.ex, ffma(ap.y, w.ey, ffma(ap.z, w.ez, fmul(w.r, ap.w))));
    return fadd(dist, rad) < 0.0f;
}

__device__ __forceinline__ uint32_t depth14(const WorldBounds &w, const float4 &dp, float scale)
{
    const float zd = ffma(dp.x, w.cx, ffma(dp.y, w.cy, ffma(dp.z, w.cz, dp.w)));
    const float t = fmul(zd, scale);
    if (!(t >= 0.0f)) return 0u;
    if (t >= 16383.0f) return 16383u;
    return (uint32_t)t;   // cvt.rzi, same as the C cast for 0 <= t < 16383
}

__device__ __forceinline__ unsigned long long make_key(uint32_t wbits, uint32_t world, uint32_t view,
                                                       uint32_t shader, uint32_t mesh, uint32_t depth,
                                                       uint32_t local)
{
    unsigned long long k = world;
    k = (k << 5) | (view & 31u);
    k = (k << 8) | (shader & 255u);
    k = (k << 11) | (mesh & 2047u);
    k = (k << 14) | (depth & 16383u);
    k = (k << (26 - wbits)) | (unsigned long long)local;
    return k;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Tests the (up to) kPerLane entities of one lane against the views of their
// world.  vis[k] gets bit v set when entity k passes view v.  The plane loop
// leaves early when no lane of the warp has a survivor (warp-uniform branch).
template <int K>
__device__ __forceinline__ void cull_views(const FrameParams &P, const ViewDev *vw,
                                           const WorldBounds (&wb)[K], const uint32_t (&phases)[K],
                                           uint32_t (&vis)[K])
{
    const uint32_t F = P.views_per_world;
    for (uint32_t v = 0; v < F; ++v) {
        const uint32_t mask = __ldg(&vw[v].phase_mask);
        uint32_t alive = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) alive |= ((phases[k] & mask) ? 1u : 0u) << k;
        if (!__any_sync(0xffffffffu, alive != 0)) continue;
#pragma unroll 1
        for (int p = 0; p < 6; ++p) {
            const float4 pl = __ldg(&vw[v].plane[p]);
            const float4 ap = __ldg(&vw[v].aplane[p]);
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (plane_rejects(wb[k], pl, ap)) alive &= ~(1u << k);
            if (!__any_sync(0xffffffffu, alive != 0)) break;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) vis[k] |= ((alive >> k) & 1u) << v;
    }
}

// Writes the keys of one lane starting at `pos` (a position in the key buffer).
template <int K>
__device__ __forceinline__ void emit_keys(const FrameParams &P, const ViewDev *vw, uint32_t world,
                                          const WorldBounds (&wb)[K], const uint32_t (&vis)[K],
                                          const uint32_t (&meta)[K], const uint32_t (&local)[K],
                                          unsigned long long pos)
{
#pragma unroll
    for (int k = 0; k < K; ++k) {
        uint32_t m = vis[k];
        while (m) {
            const uint32_t v = __ffs(m) - 1;
            m &= m - 1;
            const float4 dp = __ldg(&vw[v].depth);
            const float sc = __ldg(&vw[v].depth_scale);
            const unsigned long long key = make_key(P.world_bits, world, v, (meta[k] >> 11) & 255u,
                                                    meta[k] & 2047u, depth14(wb[k], dp, sc), local[k]);
            if (pos < P.key_cap) P.keys[pos] = key;
            else P.ctl->overflow = 1u;
            ++pos;
        }
    }
}

// Per-(world,view) counts.  One world: warp-reduce then one shared atomic per
// view per warp; many worlds: one global atomic per lane that has hits.
template <int K>
__device__ __forceinline__ void add_counts(const FrameParams &P, uint32_t world, const uint32_t (&vis)[K],
                                           uint32_t *s_counts, int lane)
{
    const uint32_t F = P.views_per_world;
    if (P.n_worlds == 1) {
        for (uint32_t v = 0; v < F; ++v) {
            uint32_t c = 0;
#pragma unroll
            for (int k = 0; k < K; ++k) c += (vis[k] >> v) & 1u;
            c = __reduce_add_sync(0xffffffffu, c);
            if (lane == 0 && c) atomicAdd(&s_counts[v], c);
        }
    } else {
        for (uint32_t v = 0; v < F; ++v) {
            uint32_t c = 0;
#pragma unroll
            for (int k = 0; k < K; ++k) c += (vis[k] >> v) & 1u;
            if (c) atomicAdd(&P.counts[(size_t)world * F + v], c);
        }
    }
}

// ---------------------------------------------------------------------------
// Range kernel: slots [first, end) in slot order.  HIER = the range is one
// hierarchy level below the roots: world = world[parent] * local.
// Grid: persistent, blocks stride over block tiles starting at first & ~127.
// ---------------------------------------------------------------------------
template <bool HIER>
__global__ void __launch_bounds__(kBlockThreads, 2)
k_transform_cull(const FrameParams P, const uint32_t first, const uint32_t end)
{
    extern __shared__ float4 s_stage[];                 // kWarpsPerBlock * kStageF4PerWarp
    __shared__ uint32_t s_warp_cnt[kWarpsPerBlock];
    __shared__ unsigned long long s_block_base;
    __shared__ uint32_t s_counts[32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 *stage = s_stage + warp * kStageF4PerWarp;
    const bool cull = P.views_per_world != 0;
    if (threadIdx.x < 32) s_counts[threadIdx.x] = 0;
    __syncthreads();

    const uint32_t base0 = first & ~(uint32_t)(kWarpTile - 1);
    const uint32_t n_tiles = (end - base0 + kBlockTile - 1) / kBlockTile;

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t wbase = base0 + tile * kBlockTile + warp * kWarpTile;
        const uint32_t s0 = wbase + lane * kPerLane;

        const float4 PX = ld_stream(P.px + s0), PY = ld_stream(P.py + s0), PZ = ld_stream(P.pz + s0);
        const float4 QX = ld_stream(P.qx + s0), QY = ld_stream(P.qy + s0), QZ = ld_stream(P.qz + s0);
        const float4 QW = ld_stream(P.qw + s0);
        const float4 SX = ld_stream(P.sx + s0), SY = ld_stream(P.sy + s0), SZ = ld_stream(P.sz + s0);
        const uint4 MT = __ldcs(reinterpret_cast<const uint4 *>(P.meta + s0));
        uint4 PAR = make_uint4(kNoParent, kNoParent, kNoParent, kNoParent);
        if (HIER) PAR = __ldcs(reinterpret_cast<const uint4 *>(P.parent + s0));

        WorldBounds wb[kPerLane];
        uint32_t phases[kPerLane], vis[kPerLane], meta[kPerLane], local[kPerLane];
        const uint32_t world = (P.n_worlds > 1) ? min(s0 / P.entities_per_world, P.n_worlds - 1u) : 0u;

#pragma unroll
        for (int k = 0; k < kPerLane; ++k) {
            const uint32_t slot = s0 + k;
            const bool active = slot >= first && slot < end;
            float m[16];
            trs_matrix(f4get(PX, k), f4get(PY, k), f4get(PZ, k),
                       f4get(QX, k), f4get(QY, k), f4get(QZ, k), f4get(QW, k),
                       f4get(SX, k), f4get(SY, k), f4get(SZ, k), m);
            if (HIER) {
                const uint32_t par = u4get(PAR, k);
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
            // stage: entity e = 4*lane+k of the warp tile, one pad slot per 4 entities
            float4 *dst = stage + (lane * kPerLane + k) * 4 + lane;
            dst[0] = make_float4(m[0], m[1], m[2], m[3]);
            dst[1] = make_float4(m[4], m[5], m[6], m[7]);
            dst[2] = make_float4(m[8], m[9], m[10], m[11]);
            dst[3] = make_float4(m[12], m[13], m[14], m[15]);

            meta[k] = u4get(MT, k);
            vis[k] = 0;
            local[k] = slot - world * P.entities_per_world;
            const uint32_t flags = (meta[k] >> 19) & 31u;
            const uint32_t mesh = meta[k] & 2047u;
            const bool cand = cull && active && (flags & 1u) && mesh < P.n_meshes;
            phases[k] = cand ? (flags >> 1) : 0u;
            const BoundsDev b = P.bounds[cand ? mesh : 0u];
            wb[k] = world_bounds(m, b);
        }

        if (cull) {
            const ViewDev *vw = P.views + (size_t)world * P.views_per_world;
            cull_views<kPerLane>(P, vw, wb, phases, vis);

            // block-scan compaction: warp scan -> per-warp totals -> one reservation per block tile
            uint32_t cnt = 0;
#pragma unroll
            for (int k = 0; k < kPerLane; ++k) cnt += __popc(vis[k]);
            const uint32_t incl = warp_incl_scan(cnt, lane);
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
            if (block_total) {
                if (__shfl_sync(0xffffffffu, incl, 31)) add_counts<kPerLane>(P, world, vis, s_counts, lane);
                if (cnt) emit_keys<kPerLane>(P, vw, world, wb, vis, meta, local,
                                             s_block_base + warp_off + (incl - cnt));
            }
        }

        // drain the warp tile: 16 coalesced 128-bit stores per lane
        __syncwarp();
        float4 *out = P.world + (size_t)wbase * 4;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int j = i * 32 + lane;
            const uint32_t slot = wbase + (j >> 2);
            const float4 v = stage[j + (j >> 4)];
            if (slot >= first && slot < end) {
                if (HIER) out[j] = v; else __stcs(out + j, v);
            }
        }
        __syncwarp();
    }

    __syncthreads();
    if (cull && P.n_worlds == 1 && threadIdx.x < P.views_per_world && s_counts[threadIdx.x])
        atomicAdd(&P.counts[threadIdx.x], s_counts[threadIdx.x]);
}

// ---------------------------------------------------------------------------
// Indexed kernel: one hierarchy level given as a slot list (slots of the level
// are not contiguous).  One entity per thread, warp-aggregated compaction.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlockThreads)
k_transform_cull_indexed(const FrameParams P, const uint32_t *__restrict__ slots, const uint32_t n)
{
    const int lane = threadIdx.x & 31;
    const bool cull = P.views_per_world != 0;
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool active = i < n;
        const uint32_t slot = active ? slots[i] : 0u;
        float m[16];
        trs_matrix(P.px[slot], P.py[slot], P.pz[slot], P.qx[slot], P.qy[slot], P.qz[slot], P.qw[slot],
                   P.sx[slot], P.sy[slot], P.sz[slot], m);
        const uint32_t par = P.parent ? P.parent[slot] : kNoParent;
        if (active && par != kNoParent) {
            const float4 *pw = P.world + (size_t)par * 4;
            const float4 a0 = pw[0], a1 = pw[1], a2 = pw[2], a3 = pw[3];
            const float a[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w,
                                 a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w};
            float w[16];
            mat4_mul(a, m, w);
#pragma unroll
            for (int k = 0; k < 16; ++k) m[k] = w[k];
        }
        if (active) {
            float4 *out = P.world + (size_t)slot * 4;
            out[0] = make_float4(m[0], m[1], m[2], m[3]);
            out[1] = make_float4(m[4], m[5], m[6], m[7]);
            out[2] = make_float4(m[8], m[9], m[10], m[11]);
            out[3] = make_float4(m[12], m[13], m[14], m[15]);
        }
        if (!cull) continue;
        const uint32_t world = (P.n_worlds > 1) ? slot / P.entities_per_world : 0u;
        uint32_t meta[1] = { P.meta[slot] }, vis[1] = { 0 }, phases[1], local[1];
        local[0] = slot - world * P.entities_per_world;
        const uint32_t flags = (meta[0] >> 19) & 31u, mesh = meta[0] & 2047u;
        const bool cand = active && (flags & 1u) && mesh < P.n_meshes;
        phases[0] = cand ? (flags >> 1) : 0u;
        WorldBounds wb[1];
        wb[0] = world_bounds(m, P.bounds[cand ? mesh : 0u]);
        const ViewDev *vw = P.views + (size_t)world * P.views_per_world;
        cull_views<1>(P, vw, wb, phases, vis);
        const uint32_t cnt = __popc(vis[0]);
        const uint32_t incl = warp_incl_scan(cnt, lane);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&P.ctl->key_total, (unsigned long long)total);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (uint32_t v = 0; v < P.views_per_world; ++v)
                if ((vis[0] >> v) & 1u) atomicAdd(&P.counts[(size_t)world * P.views_per_world + v], 1u);
            if (cnt) emit_keys<1>(P, vw, world, wb, vis, meta, local, base + (incl - cnt));
        }
    }
}

}  // namespace astre

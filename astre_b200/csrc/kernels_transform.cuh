// Shared definitions of the transform + visibility kernels: the device mirror
// layout, the culling arithmetic (spec: SURVEY.md Appendix B) and the indexed
// fallback kernel.  The main kernel is in kernels_fused.cuh.
//
// Replaces the per-entity loops of TransformSystem::run
// (engine/modules/ECS/src/system/transform_system.cpp:21-67), VisualSystem::run
// (engine/modules/ECS/src/system/visual_system.cpp:20-58) and the per-pass
// filters of deferredShadingStage (engine/modules/Pipeline/src/deferred_shading.cpp:115-121,151-156).
//
// Data layout in HBM -- the mirror of the component storages
// (engine/modules/ECS/include/ecs/component_manager.hpp:34-90), tile-major SoA:
//   tiles   one 6144-byte record per 128 consecutive slots; the record holds twelve
//           128-element planes: px py pz | qx qy qz qw | sx sy sz | meta | parent
//           (meta = mesh[10:0] | shader[18:11] | flags[23:19]; parent = slot or
//           ASTRE_NO_PARENT).  A warp tile's whole input is therefore ONE
//           contiguous 5632-byte (flat) or 6144-byte (hierarchy) block: one TMA
//           bulk copy, one DRAM page run, and every plane inside it is read with
//           conflict-free 128-bit shared loads.
//   world   float4[slot*4 + column], GLM column-major mat4 (the reference's
//           transform_matrix layout, engine/modules/Math/src/math.cpp:70-93)
// The tile array is padded by one block tile beyond max_entities.
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
constexpr uint32_t kNoParent = 0xFFFFFFFFu;

// tile record
constexpr int kTilePlanes = 12;
constexpr int kTileFloats = kTilePlanes * kWarpTile;   // 1536 words = 6144 B
enum Plane { PX = 0, PY, PZ, QX, QY, QZ, QW, SX, SY, SZ, META, PARENT };

__host__ __device__ __forceinline__ size_t tile_index(uint32_t slot, int plane)
{
    return (size_t)(slot >> 7) * kTileFloats + (size_t)plane * kWarpTile + (slot & 127u);
}

// One view on the device (256 B, 16-byte aligned).
struct ViewDev {
    float4 plane[6];   // (nx,ny,nz,w), not normalised: left,right,bottom,top,near,far
    float4 aplane[6];  // (|nx|,|ny|,|nz|,len)
    float4 depth;      // normalised near plane
    float depth_scale;
    uint32_t phase_mask;
    uint32_t sym;      // bit i: planes 2i and 2i+1 have bit-identical (|n|, len) (orthographic pairs)
    uint32_t pad;
    float4 lo, hi;     // world-space box around the view volume (xyz)
};
static_assert(sizeof(ViewDev) == 256, "ViewDev layout");

struct BoundsDev {   // 32 B; extents and radius are >= 0 (checked by astre_gpu_set_mesh_bounds)
    float cx, cy, cz, ex;
    float ey, ez, r;
    uint32_t reserved;
};

constexpr int kClaimGroups = 64, kClaimStride = 32;
struct FrameControl {            // one per context, reset at the start of a frame
    unsigned long long key_total; // keys reserved so far (may exceed capacity)
    uint32_t overflow;
    uint32_t barrier;             // arrivals at the grid barrier of multi-level launches
    uint32_t claim[kClaimGroups * kClaimStride];   // next unclaimed warp tile of every CTA group (one 128-byte line each)
};

struct FrameParams {
    const float *tiles;            // tile-major mirror
    float4 *world;
    const BoundsDev *bounds;
    const ViewDev *views;          // [n_worlds][views_per_world]
    unsigned long long *keys;
    uint32_t *counts;              // [n_worlds*views_per_world]
    FrameControl *ctl;
    uint32_t *bins;                // bucket counts of the key sort (kernels_sort.cuh): one RED per emitted key
    uint32_t claim_groups;         // work counters of a flat frame (0: none); CTA b draws from group b % claim_groups
    unsigned long long key_cap;
    uint32_t n_meshes;
    uint32_t views_per_world;      // 0 = no culling
    uint32_t n_worlds;
    uint32_t entities_per_world;
    uint32_t world_bits;
    uint32_t world_cache;          // batched worlds: 1 = warp tiles never straddle worlds and F fits the per-warp view cache
};

__device__ __forceinline__ float f4get(const float4 &v, int k)
{
    return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}
__device__ __forceinline__ uint32_t u4get(const uint4 &v, int k)
{
    return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

// World-space bounds (spec; the CPU checker restates it).
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

// Axis-aligned box of the bounds: c -+ (e + r) per axis.
struct WorldBox { float lx, ly, lz, hx, hy, hz; };

__device__ __forceinline__ WorldBox world_box(const WorldBounds &w)
{
    const float tx = fadd(w.ex, w.r), ty = fadd(w.ey, w.r), tz = fadd(w.ez, w.r);
    WorldBox b;
    b.lx = fsub(w.cx, tx); b.ly = fsub(w.cy, ty); b.lz = fsub(w.cz, tz);
    b.hx = fadd(w.cx, tx); b.hy = fadd(w.cy, ty); b.hz = fadd(w.cz, tz);
    return b;
}

// Spec step 1: the bounds' box must overlap the view's box (false on NaN = keep).
__device__ __forceinline__ bool box_overlaps(const WorldBox &b, const float4 &lo, const float4 &hi)
{
    return !(b.hx < lo.x) && !(b.lx > hi.x) && !(b.hy < lo.y) && !(b.ly > hi.y) && !(b.hz < lo.z) && !(b.lz > hi.z);
}

__device__ __forceinline__ float plane_dist(const WorldBounds &w, const float4 &pl)
{
    return ffma(pl.x, w.cx, ffma(pl.y, w.cy, ffma(pl.z, w.cz, pl.w)));
}
__device__ __forceinline__ float plane_rad(const WorldBounds &w, const float4 &ap)
{
    return ffma(ap.x, w.ex, ffma(ap.y, w.ey, ffma(ap.z, w.ez, fmul(w.r, ap.w))));
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

// Bucket digit of the key sort (kernels_sort.cuh): the top `bits` key bits that can differ, packed in key
// order.  The cull kernels count every key they emit into bins[bin_digit(key)]; bits == 0: no sort wanted.
constexpr int kBinSegs = 8;
struct BinPlan {
    uint32_t n_segs, bits;
    uint32_t shift[kBinSegs], mask[kBinSegs], lshift[kBinSegs];   // digit = OR of ((key >> shift) & mask) << lshift
};
__host__ __device__ __forceinline__ uint32_t bin_digit(const BinPlan &bp, unsigned long long key)
{
    uint32_t d = 0;
#pragma unroll
    for (int s = 0; s < kBinSegs; ++s)
        if ((uint32_t)s < bp.n_segs) d |= ((uint32_t)(key >> bp.shift[s]) & bp.mask[s]) << bp.lshift[s];
    return d;
}

// Digit plan of the LSD fallback of the key sort (kernels_sort.cuh).
constexpr int kSortPasses = 8;
constexpr int kSortSegs = 4;
constexpr int kHistWords = kSortPasses * 256;
struct SortPlan {
    uint32_t n_passes;
    uint32_t seg[kSortPasses][kSortSegs];   // digit = OR of ((key >> shift) & mask(width)) << lshift; shift | width<<8 | lshift<<16
};
__host__ __device__ __forceinline__ uint32_t sort_digit(const uint32_t (&seg)[kSortSegs], unsigned long long key)
{
    uint32_t d = 0;
#pragma unroll
    for (int s = 0; s < kSortSegs; ++s) {
        const uint32_t w = (seg[s] >> 8) & 255u;
        d |= ((uint32_t)(key >> (seg[s] & 255u)) & ((1u << w) - 1u)) << ((seg[s] >> 16) & 255u);
    }
    return d;
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

// View access: constant bank (one world; the records are a kernel parameter) or
// global memory (batched worlds, one record set per world).
constexpr int kMaxParamViews = 32;            // == ASTRE_MAX_VIEWS
struct ViewSet { ViewDev v[kMaxParamViews]; };

struct ConstViews {
    const ViewSet &vs;
    __device__ __forceinline__ float4 plane(uint32_t v, int p) const { return vs.v[v].plane[p]; }
    __device__ __forceinline__ float4 aplane(uint32_t v, int p) const { return vs.v[v].aplane[p]; }
    __device__ __forceinline__ float4 depth(uint32_t v) const { return vs.v[v].depth; }
    __device__ __forceinline__ float depth_scale(uint32_t v) const { return vs.v[v].depth_scale; }
    __device__ __forceinline__ uint32_t mask(uint32_t v) const { return vs.v[v].phase_mask; }
    __device__ __forceinline__ uint32_t sym(uint32_t v) const { return vs.v[v].sym; }
    __device__ __forceinline__ float4 lo(uint32_t v) const { return vs.v[v].lo; }
    __device__ __forceinline__ float4 hi(uint32_t v) const { return vs.v[v].hi; }
};
struct GlobalViews {
    const ViewDev *vw;
    __device__ __forceinline__ float4 plane(uint32_t v, int p) const { return __ldg(&vw[v].plane[p]); }
    __device__ __forceinline__ float4 aplane(uint32_t v, int p) const { return __ldg(&vw[v].aplane[p]); }
    __device__ __forceinline__ float4 depth(uint32_t v) const { return __ldg(&vw[v].depth); }
    __device__ __forceinline__ float depth_scale(uint32_t v) const { return __ldg(&vw[v].depth_scale); }
    __device__ __forceinline__ uint32_t mask(uint32_t v) const { return __ldg(&vw[v].phase_mask); }
    __device__ __forceinline__ uint32_t sym(uint32_t v) const { return __ldg(&vw[v].sym); }
    __device__ __forceinline__ float4 lo(uint32_t v) const { return __ldg(&vw[v].lo); }
    __device__ __forceinline__ float4 hi(uint32_t v) const { return __ldg(&vw[v].hi); }
};

// Spec step 2 for one view: the six plane tests (reject iff dist + rad < 0) on the
// K entities of one lane; alive[] holds the survivors of step 1 on entry and of
// both steps on exit.  Orthographic pairs share their support radius (sym).
// Leaves as soon as no lane of the warp has a survivor (warp-uniform branch).
template <int K, class Views>
__device__ __forceinline__ void cull_planes(const Views &V, const uint32_t v, const WorldBounds (&wb)[K], bool (&alive)[K])
{
    const uint32_t sym = V.sym(v);
    float rad[K];
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        const float4 pl = V.plane(v, p);
        if (!((p & 1) && ((sym >> (p >> 1)) & 1u))) {
            const float4 ap = V.aplane(v, p);
#pragma unroll
            for (int k = 0; k < K; ++k) rad[k] = plane_rad(wb[k], ap);
        }
        bool any = false;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            alive[k] = alive[k] && !(fadd(plane_dist(wb[k], pl), rad[k]) < 0.0f);
            any |= alive[k];
        }
        if (p < 5 && !__any_sync(0xffffffffu, any)) break;
    }
}

// Both steps over all F views (batched-worlds and indexed paths); vis[k] gets bit v.
template <int K, class Views>
__device__ __forceinline__ void cull_views(const Views &V, const uint32_t F, const WorldBounds (&wb)[K],
                                           const uint32_t (&phases)[K], uint32_t (&vis)[K])
{
    WorldBox box[K];
#pragma unroll
    for (int k = 0; k < K; ++k) box[k] = world_box(wb[k]);
    for (uint32_t v = 0; v < F; ++v) {
        const uint32_t mask = V.mask(v);
        const float4 lo = V.lo(v), hi = V.hi(v);
        bool alive[K], any = false;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            alive[k] = (phases[k] & mask) != 0 && box_overlaps(box[k], lo, hi);
            any |= alive[k];
        }
        if (!__any_sync(0xffffffffu, any)) continue;
        cull_planes<K>(V, v, wb, alive);
#pragma unroll
        for (int k = 0; k < K; ++k) vis[k] |= (alive[k] ? 1u : 0u) << v;
    }
}

// Writes the keys of one lane starting at `pos`; with s_counts != nullptr also
// bumps the per-view shared counters (one world, sparse case).
template <int K, class Views>
__device__ __forceinline__ void emit_keys(const FrameParams &P, const BinPlan &bp, const Views &V, uint32_t world,
                                          const WorldBounds (&wb)[K], const uint32_t (&vis)[K],
                                          const uint32_t (&meta)[K], const uint32_t (&local)[K],
                                          unsigned long long pos, uint32_t *s_counts)
{
#pragma unroll
    for (int k = 0; k < K; ++k) {
        uint32_t m = vis[k];
        while (m) {
            const uint32_t v = __ffs(m) - 1;
            m &= m - 1;
            const unsigned long long key = make_key(P.world_bits, world, v, (meta[k] >> 11) & 255u, meta[k] & 2047u,
                                                    depth14(wb[k], V.depth(v), V.depth_scale(v)), local[k]);
            if (pos < P.key_cap) {
                P.keys[pos] = key;
                if (bp.bits) atomicAdd(&P.bins[bin_digit(bp, key)], 1u);
            } else {
                P.ctl->overflow = 1u;
            }
            ++pos;
            if (s_counts) atomicAdd(&s_counts[v], 1u);
        }
    }
}

// Per-(world,view) counts, dense case.  One world: warp-reduce then one shared
// atomic per view per warp; many worlds: one global atomic per lane with hits.
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
// Thread-per-entity kernel: one hierarchy level given as a slot list (levels whose
// slots are not a contiguous range) or as a small range.  One entity per thread,
// warp-aggregated compaction.  Short dependent chain per warp: a level of a few
// thousand entities finishes in ~5 us, where the throughput kernel (four entities
// per lane, staged stores, task queue) needs ~18 us for its first tile.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlockThreads)
k_transform_cull_indexed(const FrameParams P, const __grid_constant__ BinPlan bp, const uint32_t *__restrict__ slots,
                         const uint32_t first, const uint32_t n)
{
    __shared__ float4 s_scratch[kBlockThreads * 4];
    const int lane = threadIdx.x & 31;
    const bool cull = P.views_per_world != 0;
    const uint32_t n_round = (n + 31u) & ~31u;
    cudaGridDependencySynchronize();       // (no-op unless launched as a programmatic dependent)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool active = i < n;
        const uint32_t slot = active ? (slots ? slots[i] : first + i) : 0u;   // slots == nullptr: the range [first, first+n)
        const float *t = P.tiles + tile_index(slot, 0);
        float m[16];
        trs_matrix(t[PX * kWarpTile], t[PY * kWarpTile], t[PZ * kWarpTile],
                   t[QX * kWarpTile], t[QY * kWarpTile], t[QZ * kWarpTile], t[QW * kWarpTile],
                   t[SX * kWarpTile], t[SY * kWarpTile], t[SZ * kWarpTile], m, s_scratch + threadIdx.x * 4);
        const uint32_t par = __float_as_uint(t[PARENT * kWarpTile]);
        if (active && par != kNoParent) {
            const float4 *pw = P.world + (size_t)par * 4;
            const float4 a0 = __ldcg(pw), a1 = __ldcg(pw + 1), a2 = __ldcg(pw + 2), a3 = __ldcg(pw + 3);
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
        uint32_t meta[1] = { __float_as_uint(t[META * kWarpTile]) }, vis[1] = { 0 }, phases[1], local[1];
        local[0] = slot - world * P.entities_per_world;
        const uint32_t flags = (meta[0] >> 19) & 31u, mesh = meta[0] & 2047u;
        const bool cand = active && (flags & 1u) && mesh < P.n_meshes;
        phases[0] = cand ? (flags >> 1) : 0u;
        WorldBounds wb[1];
        wb[0] = world_bounds(m, P.bounds[cand ? mesh : 0u]);
        const GlobalViews V{ P.views + (size_t)world * P.views_per_world };
        cull_views<1>(V, P.views_per_world, wb, phases, vis);
        const uint32_t cnt = __popc(vis[0]);
        const uint32_t incl = warp_incl_scan(cnt, lane);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&P.ctl->key_total, (unsigned long long)total);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (uint32_t v = 0; v < P.views_per_world; ++v)
                if ((vis[0] >> v) & 1u) atomicAdd(&P.counts[(size_t)world * P.views_per_world + v], 1u);
            if (cnt) emit_keys<1>(P, bp, V, world, wb, vis, meta, local, base + (incl - cnt), nullptr);
        }
    }
}

// ---------------------------------------------------------------------------
// Hierarchies stored in breadth-first order (every level a contiguous slot range, parents non-decreasing along a
// level -- what a level-sorted loader produces, and what configuration 2 is): the descendants of a contiguous run of
// roots are ONE contiguous range per level, a "cone".  The host cuts the forest into cones of equal size; a CTA
// walks its cone level by level, so parent-before-child becomes a __syncthreads() instead of a grid-wide barrier or
// a launch boundary, and a level costs two dependent L2 round trips instead of ~15 us.  Parent rows were written by
// this CTA a level earlier and are read back through L2 (ld.global.cg).  One entity per thread, warp-aggregated
// compaction, per-view counts in shared memory (one global atomic per view and CTA).
// ---------------------------------------------------------------------------
#ifndef ASTRE_CONE_MIN_BLOCKS
#define ASTRE_CONE_MIN_BLOCKS 4      // 62 registers, no spills; measured on config 2: 2 -> 0.083 ms, 3 -> 0.074, 4 -> 0.069
#endif
template <bool CULL>
__global__ void __launch_bounds__(kBlockThreads, ASTRE_CONE_MIN_BLOCKS)
k_hier_cones(const FrameParams P, const __grid_constant__ BinPlan bp, const uint2 *__restrict__ ranges, const uint32_t n_levels,
             const uint32_t prefetch)
{
    constexpr uint32_t kBufKeys = CULL ? 3072u : 1u;          // per-CTA key buffer: ONE global reservation per flush
    __shared__ float4 s_scratch[kBlockThreads * 4];
    __shared__ unsigned long long s_kbuf[kBufKeys];
    __shared__ uint32_t s_counts[32];
    __shared__ uint32_t s_kfill;
    __shared__ unsigned long long s_kbase;
    const int lane = threadIdx.x & 31;
    const uint32_t F = CULL ? P.views_per_world : 0u;
    const bool one_world = P.n_worlds <= 1;
    // every warp reserving its keys with an atomic on the frame's one key counter serialises at the L2 (a third of a
    // million same-address atomics on configuration 2); buffered per CTA it is a few hundred
    const bool buffered = CULL && F * kBlockThreads <= kBufKeys / 2u;
    if (threadIdx.x < 32) s_counts[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_kfill = 0;
    if (prefetch) {
        // the cone's inputs (whole tile records of every level) on their way into L2 before the level chain starts, and
        // before the previous kernel has even finished: neither the mirror nor the ranges are written by a frame
        for (uint32_t level = 0; level < n_levels; ++level) {
            const uint2 r = __ldg(ranges + (size_t)blockIdx.x * n_levels + level);
            if (r.y <= r.x) continue;
            const uint32_t t0 = r.x >> 7, t1 = (r.y - 1u) >> 7;
            const char *rec = reinterpret_cast<const char *>(P.tiles + (size_t)t0 * kTileFloats);
            const uint32_t lines = (t1 - t0 + 1u) * (uint32_t)(kTileFloats * sizeof(float) / 128u);
            for (uint32_t i = threadIdx.x; i < lines; i += kBlockThreads)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + (size_t)i * 128u));
        }
    }
    cudaGridDependencySynchronize();
    __syncthreads();
    auto flush = [&]() {                                       // block-wide; callers synchronise before and after
        const uint32_t fill = s_kfill;
        if (fill) {
            if (threadIdx.x == 0) s_kbase = atomicAdd(&P.ctl->key_total, (unsigned long long)fill);
            __syncthreads();
            const unsigned long long base = s_kbase;
            for (uint32_t i = threadIdx.x; i < fill; i += kBlockThreads) {
                const unsigned long long k = s_kbuf[i], pos = base + i;
                if (pos < P.key_cap) {
                    P.keys[pos] = k;
                    if (bp.bits) atomicAdd(&P.bins[bin_digit(bp, k)], 1u);
                } else {
                    P.ctl->overflow = 1u;
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) s_kfill = 0;
        }
    };
    for (uint32_t level = 0; level < n_levels; ++level) {
        const uint2 r = __ldg(ranges + (size_t)blockIdx.x * n_levels + level);
        const uint32_t first = r.x, n = r.y - r.x;
        const uint32_t n_round = (n + 31u) & ~31u;
        // room in the key buffer: checked once per level when the whole level fits, else before every round of 256
        const bool level_fits = buffered && (unsigned long long)n_round * F <= kBufKeys / 2u;
        if (level_fits) {
            const uint32_t fill_now = s_kfill;                 // (after the barrier that ended the previous level)
            __syncthreads();                                   // read by everyone before anyone appends again: one decision per CTA
            if (fill_now + n_round * F > kBufKeys) { flush(); __syncthreads(); }
        }
        for (uint32_t i0 = 0; i0 < n_round; i0 += kBlockThreads) {
            if (buffered && !level_fits) {
                __syncthreads();
                const uint32_t fill_now = s_kfill;
                __syncthreads();
                if (fill_now + F * kBlockThreads > kBufKeys) { flush(); __syncthreads(); }
            }
            const uint32_t i = i0 + threadIdx.x;
            if (i >= n_round) continue;                        // whole warps only: n_round is a multiple of 32
            const bool active = i < n;
            const uint32_t slot = active ? first + i : first;
            const float *t = P.tiles + tile_index(slot, 0);
            float m[16];
            trs_matrix(t[PX * kWarpTile], t[PY * kWarpTile], t[PZ * kWarpTile],
                       t[QX * kWarpTile], t[QY * kWarpTile], t[QZ * kWarpTile], t[QW * kWarpTile],
                       t[SX * kWarpTile], t[SY * kWarpTile], t[SZ * kWarpTile], m, s_scratch + threadIdx.x * 4);
            const uint32_t par = __float_as_uint(t[PARENT * kWarpTile]);
            if (active && par != kNoParent) {
                const float4 *pw = P.world + (size_t)par * 4;
                const float4 a0 = __ldcg(pw), a1 = __ldcg(pw + 1), a2 = __ldcg(pw + 2), a3 = __ldcg(pw + 3);
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
            if (!CULL) continue;
            const uint32_t world = one_world ? 0u : slot / P.entities_per_world;
            uint32_t meta[1] = { __float_as_uint(t[META * kWarpTile]) }, vis[1] = { 0 }, phases[1], local[1];
            local[0] = slot - world * P.entities_per_world;
            const uint32_t flags = (meta[0] >> 19) & 31u, mesh = meta[0] & 2047u;
            const bool cand = active && (flags & 1u) && mesh < P.n_meshes;
            phases[0] = cand ? (flags >> 1) : 0u;
            WorldBounds wb[1];
            wb[0] = world_bounds(m, P.bounds[cand ? mesh : 0u]);
            const GlobalViews V{ P.views + (size_t)world * F };
            cull_views<1>(V, F, wb, phases, vis);
            const uint32_t cnt = __popc(vis[0]);
            const uint32_t incl = warp_incl_scan(cnt, lane);
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (total) {
                if (!one_world)
                    for (uint32_t v = 0; v < F; ++v)
                        if ((vis[0] >> v) & 1u) atomicAdd(&P.counts[(size_t)world * F + v], 1u);
                if (buffered) {
                    uint32_t sbase = 0;
                    if (lane == 0) sbase = atomicAdd(&s_kfill, total);
                    sbase = __shfl_sync(0xffffffffu, sbase, 0);
                    uint32_t at = sbase + (incl - cnt), mv = vis[0];
                    while (mv) {
                        const uint32_t v = __ffs(mv) - 1;
                        mv &= mv - 1;
                        s_kbuf[at++] = make_key(P.world_bits, world, v, (meta[0] >> 11) & 255u, meta[0] & 2047u,
                                                depth14(wb[0], V.depth(v), V.depth_scale(v)), local[0]);
                        if (one_world) atomicAdd(&s_counts[v], 1u);
                    }
                } else {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(&P.ctl->key_total, (unsigned long long)total);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (cnt) emit_keys<1>(P, bp, V, world, wb, vis, meta, local, base + (incl - cnt), one_world ? s_counts : nullptr);
                }
            }
        }
        __syncthreads();     // the next level reads this level's matrices (same CTA: block-scope ordering is enough)
    }
    if (buffered) { flush(); }
    __syncthreads();
    if (CULL && one_world && threadIdx.x < F && s_counts[threadIdx.x]) atomicAdd(&P.counts[threadIdx.x], s_counts[threadIdx.x]);
}

}  // namespace astre

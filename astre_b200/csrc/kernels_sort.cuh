// K3: device radix sort of the visible draw keys (64-bit, ascending).
//
// The reference draws in absl hash-map iteration order
// (engine/modules/Pipeline/src/deferred_shading.cpp:115,151) -- there is nothing
// to sort there; the order is spec-defined: ascending
// key, keys unique, so any correct sort reproduces the oracle's list.
//
// Least-significant-digit radix sort, 8-bit digits, single-pass ("onesweep")
// scatter with decoupled look-back:
//   k_sort_hist   one read of the keys builds all eight 256-bin histograms
//   k_sort_plan   exclusive scans, skips digits on which all keys agree, fixes
//                 the ping-pong buffer of every pass
//   k_sort_pass   per executed digit: rank inside a 2048-key tile with
//                 warp match-any, chain tile prefixes through 64-bit
//                 (epoch | flag | count) words, scatter
// The key count lives in device memory (FrameControl::key_total); no launch
// depends on a host read-back.
//
// Algorithmic bytes for V keys and P executed passes: V*(8 + 16*P).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels_transform.cuh"

namespace astre {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;   // 2048 keys
constexpr int kSortPasses = 8;

struct SortControl {
    uint32_t hist[kSortPasses][256];     // digit counts, then exclusive prefixes
    uint32_t skip[kSortPasses];          // 1 = all keys share this digit
    uint32_t src[kSortPasses];           // input buffer (0/1) of the pass
    uint32_t tile_counter[kSortPasses];  // dynamic tile ids
    uint32_t result_buf;                 // buffer that holds the sorted keys
    uint32_t n_keys;                     // min(key_total, capacity)
    uint32_t epoch;                      // bumped every frame (< 2^27); tags look-back words
    uint32_t pad;
};

__device__ __forceinline__ uint32_t sort_key_count(const FrameControl *ctl, unsigned long long cap)
{
    const unsigned long long t = ctl->key_total;
    return (uint32_t)(t < cap ? t : cap);
}

__global__ void __launch_bounds__(kSortThreads)
k_sort_hist(const unsigned long long *__restrict__ keys, const FrameControl *ctl, unsigned long long cap,
            SortControl *sc)
{
    __shared__ uint32_t s_hist[kSortPasses][256];
    for (int i = threadIdx.x; i < kSortPasses * 256; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t n = sort_key_count(ctl, cap);
    const uint32_t n_round = (n + 31u) & ~31u;
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool valid = i < n;
        const unsigned long long k = valid ? keys[i] : 0ull;
        // the high digits are nearly constant: aggregate equal digits inside the warp
#pragma unroll
        for (int p = 0; p < kSortPasses; ++p) {
            const uint32_t digit = (uint32_t)(k >> (8 * p)) & 255u;
            const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : (256u + lane));
            if (valid && (peers & lt_mask) == 0) atomicAdd(&s_hist[p][digit], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSortPasses * 256; i += blockDim.x) {
        const uint32_t c = (&s_hist[0][0])[i];
        if (c) atomicAdd(&(&sc->hist[0][0])[i], c);
    }
}

// One block of 8 warps: warp p scans the histogram of pass p (lane l owns bins 8l..8l+7).
__global__ void __launch_bounds__(kSortPasses * 32)
k_sort_plan(const FrameControl *ctl, unsigned long long cap, SortControl *sc)
{
    __shared__ uint32_t s_skip[kSortPasses];
    const uint32_t n = sort_key_count(ctl, cap);
    const int p = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t c[8], sum = 0;
    bool all_same = false;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c[i] = sc->hist[p][lane * 8 + i];
        all_same |= (c[i] == n);
        sum += c[i];
    }
    const uint32_t incl = warp_incl_scan(sum, lane);
    uint32_t run = incl - sum;
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc->hist[p][lane * 8 + i] = run; run += c[i]; }
    const bool skip = __any_sync(0xffffffffu, all_same) || n == 0;
    if (lane == 0) s_skip[p] = skip ? 1u : 0u;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t buf = 0;
        for (int q = 0; q < kSortPasses; ++q) {
            sc->skip[q] = s_skip[q];
            sc->src[q] = buf;
            sc->tile_counter[q] = 0;
            if (!s_skip[q]) buf ^= 1u;
        }
        sc->result_buf = buf;
        sc->n_keys = n;
    }
}

// look-back word: epoch[63:34] | flag[33:32] | value[31:0].  One 64-bit word
// carries flag and value together, so relaxed gpu-scope accesses suffice.
constexpr unsigned long long kFlagAggregate = 1ull << 32;
constexpr unsigned long long kFlagPrefix = 2ull << 32;

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kSortThreads)
k_sort_pass(unsigned long long *__restrict__ buf0, unsigned long long *__restrict__ buf1,
            SortControl *sc, unsigned long long *__restrict__ lookback, const int pass)
{
    if (sc->skip[pass]) return;
    __shared__ uint32_t s_whist[kSortWarps][256];
    __shared__ uint32_t s_tile;

    const unsigned long long *in = sc->src[pass] ? buf1 : buf0;
    unsigned long long *out = sc->src[pass] ? buf0 : buf1;
    const uint32_t n = sc->n_keys;
    const uint32_t n_tiles = (n + kSortTile - 1) / kSortTile;
    const unsigned long long epoch = ((unsigned long long)sc->epoch * kSortPasses + pass + 1ull) << 34;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int shift = pass * 8;
    const uint32_t lt_mask = (1u << lane) - 1u;

    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&sc->tile_counter[pass], 1u);
        for (int i = threadIdx.x; i < kSortWarps * 256; i += blockDim.x) (&s_whist[0][0])[i] = 0;
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_tiles) return;

        // warp-striped load: item i of lane l is key tile*2048 + warp*256 + i*32 + l
        const uint32_t wbase = tile * kSortTile + warp * (kSortItems * 32);
        unsigned long long key[kSortItems];
        uint32_t rank[kSortItems];
#pragma unroll
        for (int i = 0; i < kSortItems; ++i) {
            const uint32_t idx = wbase + i * 32 + lane;
            key[i] = idx < n ? in[idx] : ~0ull;
        }
        // stable rank of every key among the keys of its warp with the same digit
#pragma unroll
        for (int i = 0; i < kSortItems; ++i) {
            const uint32_t idx = wbase + i * 32 + lane;
            const bool valid = idx < n;
            const uint32_t digit = (uint32_t)(key[i] >> shift) & 255u;
            const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : (256u + lane));
            const uint32_t before = s_whist[warp][digit];
            rank[i] = before + __popc(peers & lt_mask);
            __syncwarp();
            if (valid && (peers & lt_mask) == 0) s_whist[warp][digit] = before + __popc(peers);
            __syncwarp();
        }
        __syncthreads();

        // thread d: digit d.  Exclusive prefix over the warps, tile total, look-back.
        {
            const int d = threadIdx.x;
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) {
                const uint32_t c = s_whist[w][d];
                s_whist[w][d] = run;
                run += c;
            }
            unsigned long long *mine = lookback + (size_t)tile * 256 + d;
            uint32_t excl = 0;
            if (tile == 0) {
                st_relaxed(mine, epoch | kFlagPrefix | run);
            } else {
                st_relaxed(mine, epoch | kFlagAggregate | run);
                int t = (int)tile - 1;
                while (true) {
                    const unsigned long long w = ld_relaxed(lookback + (size_t)t * 256 + d);
                    if ((w >> 34) != (epoch >> 34) || !(w & (kFlagAggregate | kFlagPrefix))) continue;  // not published yet
                    excl += (uint32_t)w;
                    if (w & kFlagPrefix) break;
                    --t;
                }
                st_relaxed(mine, epoch | kFlagPrefix | (unsigned long long)(excl + run));
            }
            const uint32_t base = sc->hist[pass][d] + excl;
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) s_whist[w][d] += base;
        }
        __syncthreads();

#pragma unroll
        for (int i = 0; i < kSortItems; ++i) {
            const uint32_t idx = wbase + i * 32 + lane;
            if (idx < n) {
                const uint32_t digit = (uint32_t)(key[i] >> shift) & 255u;
                out[s_whist[warp][digit] + rank[i]] = key[i];
            }
        }
    }
}

// Exclusive scan of counts[0..n) into offsets[0..n] (offsets[n] = total).  One block.
__global__ void __launch_bounds__(1024)
k_scan_counts(const uint32_t *__restrict__ counts, unsigned long long *__restrict__ offsets, const uint32_t n)
{
    __shared__ unsigned long long s_part[1024];
    const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = threadIdx.x * per, hi = min(lo + per, n);
    unsigned long long sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += counts[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (uint32_t t = 0; t < blockDim.x; ++t) { const unsigned long long c = s_part[t]; s_part[t] = run; run += c; }
        offsets[n] = run;
    }
    __syncthreads();
    unsigned long long run = s_part[threadIdx.x];
    for (uint32_t i = lo; i < hi; ++i) { offsets[i] = run; run += counts[i]; }
}

}  // namespace astre

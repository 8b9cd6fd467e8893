// K3: device radix sort of the visible draw keys (64-bit, ascending).
//
// The reference draws in absl hash-map iteration order
// (engine/modules/Pipeline/src/deferred_shading.cpp:115,151) -- there is nothing
// to sort there; the order is spec-defined: ascending
// key, keys unique, so any correct sort reproduces the oracle's list.
//
// Least-significant-digit radix sort restricted to the key bits that actually
// differ between keys.  The cull kernel accumulates the OR and AND of all keys;
// bits where they agree are equal in every key and cannot affect the order.
//   k_sort_plan   packs the varying bits, LSB first, into digits of up to 8 bits
//                 (each digit = up to 4 runs of consecutive key bits); a 16M-entity
//                 5-view frame has 46 varying bits -> 6 passes instead of 8
//   k_sort_hist   one read of the keys builds the histogram of every digit;
//                 per-block partials, no global atomics
//   k_sort_scan   sums the partials and scans them (one block per digit)
//   k_sort_pass   per digit: single-pass ("onesweep") scatter -- rank inside a
//                 2048-key tile with warp match-any, chain tile prefixes through
//                 64-bit (epoch | flag | count) look-back words, scatter
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
constexpr int kSortPasses = 8;                         // upper bound on digits
constexpr int kSortSegs = 4;                           // runs of key bits per digit
constexpr int kHistBlocks = 148;                       // grid of k_sort_hist (one CTA per SM)

struct SortControl {
    uint32_t hist[kSortPasses][256];     // exclusive prefixes per digit (k_sort_scan)
    uint32_t seg[kSortPasses][kSortSegs];// digit = OR of ((key >> shift) & mask(width)) << lshift; shift | width<<8 | lshift<<16
    uint32_t tile_counter[kSortPasses];  // dynamic tile ids
    uint32_t n_passes;
    uint32_t result_buf;                 // buffer that holds the sorted keys
    uint32_t n_keys;                     // min(key_total, capacity)
    uint32_t epoch;                      // bumped every frame (< 2^27); tags look-back words
};

__device__ __forceinline__ uint32_t sort_key_count(const FrameControl *ctl, unsigned long long cap)
{
    const unsigned long long t = ctl->key_total;
    return (uint32_t)(t < cap ? t : cap);
}

struct DigitDesc {
    uint32_t seg[kSortSegs];
    __device__ __forceinline__ uint32_t of(unsigned long long key) const
    {
        uint32_t d = 0;
#pragma unroll
        for (int s = 0; s < kSortSegs; ++s) {
            const uint32_t w = (seg[s] >> 8) & 255u;
            d |= ((uint32_t)(key >> (seg[s] & 255u)) & ((1u << w) - 1u)) << ((seg[s] >> 16) & 255u);
        }
        return d;
    }
};

// Digits from the varying-bit mask (one thread).  Returns the number of passes.
__device__ inline int plan_digits(unsigned long long varying, uint32_t (&seg)[kSortPasses][kSortSegs])
{
    int n_passes = 0;
    bool ok = true;
    int bits = 0, segs = 0;          // of the digit being filled
    for (int p = 0; p < kSortPasses; ++p)
        for (int s = 0; s < kSortSegs; ++s) seg[p][s] = 0;
    int b = 0;
    while (b < 64 && ok) {
        if (!((varying >> b) & 1ull)) { ++b; continue; }
        int run = 0;
        while (b + run < 64 && ((varying >> (b + run)) & 1ull)) ++run;
        while (run > 0) {
            if (bits == 8 || segs == kSortSegs) {            // close the digit
                ++n_passes;
                if (n_passes == kSortPasses) { ok = false; break; }
                bits = 0; segs = 0;
            }
            const int take = run < 8 - bits ? run : 8 - bits;
            seg[n_passes][segs++] = (uint32_t)b | ((uint32_t)take << 8) | ((uint32_t)bits << 16);
            bits += take; b += take; run -= take;
        }
    }
    if (ok) return bits > 0 ? n_passes + 1 : n_passes;
    // more than 8 adaptive digits (very fragmented mask): plain bytes, constant bytes dropped
    n_passes = 0;
    for (int p = 0; p < kSortPasses; ++p)
        for (int s = 0; s < kSortSegs; ++s) seg[p][s] = 0;
    for (int p = 0; p < 8; ++p) {
        if (!((varying >> (8 * p)) & 255ull)) continue;
        seg[n_passes][0] = (uint32_t)(8 * p) | (8u << 8);
        ++n_passes;
    }
    return n_passes;
}

__device__ __forceinline__ unsigned long long varying_bits(const FrameControl *ctl, uint32_t n)
{
    const unsigned long long key_or = ((unsigned long long)ctl->or_hi << 32) | ctl->or_lo;
    const unsigned long long key_and = ((unsigned long long)ctl->and_hi << 32) | ctl->and_lo;
    return n > 1 ? (key_or & ~key_and) : 0ull;
}

__global__ void k_sort_plan(const FrameControl *ctl, unsigned long long cap, SortControl *sc)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t n = sort_key_count(ctl, cap);
    uint32_t seg[kSortPasses][kSortSegs];
    const int n_passes = plan_digits(varying_bits(ctl, n), seg);
    for (int p = 0; p < kSortPasses; ++p) {
        for (int s = 0; s < kSortSegs; ++s) sc->seg[p][s] = seg[p][s];
        sc->tile_counter[p] = 0;
    }
    sc->n_passes = (uint32_t)n_passes;
    sc->result_buf = (uint32_t)(n_passes & 1);
    sc->n_keys = n;
}

// Per-block histograms of every digit: block_hist[block][pass][256], plain stores.
__global__ void __launch_bounds__(kSortThreads)
k_sort_hist(const unsigned long long *__restrict__ keys, const SortControl *sc, uint32_t *__restrict__ block_hist)
{
    __shared__ uint32_t s_hist[kSortPasses][256];
    const uint32_t n_passes = sc->n_passes;
    if (n_passes == 0) return;
    for (int i = threadIdx.x; i < kSortPasses * 256; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    DigitDesc desc[kSortPasses];
#pragma unroll
    for (int p = 0; p < kSortPasses; ++p)
#pragma unroll
        for (int s = 0; s < kSortSegs; ++s) desc[p].seg[s] = sc->seg[p][s];
    __syncthreads();
    const uint32_t n = sc->n_keys;
    const uint32_t n_round = (n + 31u) & ~31u;
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool valid = i < n;
        const unsigned long long k = valid ? keys[i] : 0ull;
        // high digits are nearly constant: aggregate equal digits inside the warp
#pragma unroll
        for (int p = 0; p < kSortPasses; ++p) {
            if (p < (int)n_passes) {
                const uint32_t digit = desc[p].of(k);
                const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : (256u + lane));
                if (valid && (peers & lt_mask) == 0) atomicAdd(&s_hist[p][digit], (uint32_t)__popc(peers));
            }
        }
    }
    __syncthreads();
    uint32_t *out = block_hist + (size_t)blockIdx.x * kSortPasses * 256;
    for (uint32_t i = threadIdx.x; i < n_passes * 256; i += blockDim.x) out[i] = (&s_hist[0][0])[i];
}

// Block p: histogram of digit p = sum of the partials, then exclusive scan over its 256 bins.
__global__ void __launch_bounds__(256)
k_sort_scan(SortControl *sc, const uint32_t *__restrict__ block_hist, const uint32_t n_blocks)
{
    __shared__ uint32_t s_warp[8];
    const uint32_t p = blockIdx.x;
    if (p >= sc->n_passes) return;
    const int d = threadIdx.x, lane = d & 31, warp = d >> 5;
    uint32_t c = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) c += block_hist[((size_t)b * kSortPasses + p) * 256 + d];
    const uint32_t incl = warp_incl_scan(c, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t off = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) if (w < warp) off += s_warp[w];
    sc->hist[p][d] = off + incl - c;
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
    if ((uint32_t)pass >= sc->n_passes) return;
    __shared__ uint32_t s_whist[kSortWarps][256];
    __shared__ uint32_t s_tile;

    const unsigned long long *in = (pass & 1) ? buf1 : buf0;
    unsigned long long *out = (pass & 1) ? buf0 : buf1;
    const uint32_t n = sc->n_keys;
    const uint32_t n_tiles = (n + kSortTile - 1) / kSortTile;
    const unsigned long long epoch = ((unsigned long long)sc->epoch * kSortPasses + pass + 1ull) << 34;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    DigitDesc desc;
#pragma unroll
    for (int s = 0; s < kSortSegs; ++s) desc.seg[s] = sc->seg[pass][s];

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
        uint32_t rank[kSortItems], dig[kSortItems];
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
            dig[i] = desc.of(key[i]);
            const uint32_t peers = __match_any_sync(0xffffffffu, valid ? dig[i] : (256u + lane));
            const uint32_t before = s_whist[warp][dig[i]];
            rank[i] = before + __popc(peers & lt_mask);
            __syncwarp();
            if (valid && (peers & lt_mask) == 0) s_whist[warp][dig[i]] = before + __popc(peers);
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
            if (idx < n) out[s_whist[warp][dig[i]] + rank[i]] = key[i];
        }
    }
}

// ---------------------------------------------------------------------------
// Cooperative sort: ONE persistent launch (all CTAs co-resident) runs the plan
// and every digit pass with grid-wide barriers in between.  For the draw lists
// of a frame (10^5..10^7 keys, L2-resident) a pass is latency-bound, and the
// chained look-back of the onesweep pass costs more than three barriers:
//   A  per-tile digit counts             -> table[tile][256]
//   B  one warp per digit: exclusive prefix over the tiles, digit totals
//   C  scan of the totals, stable in-tile ranks (warp match-any), scatter
// Works for any key count; the onesweep kernels remain the path for lists
// that do not fit L2.
// ---------------------------------------------------------------------------
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

__global__ void __launch_bounds__(kSortThreads)
k_sort_coop(unsigned long long *__restrict__ buf0, unsigned long long *__restrict__ buf1, const FrameControl *ctl,
            unsigned long long cap, SortControl *sc, uint32_t *__restrict__ table, uint32_t *__restrict__ totals,
            uint32_t *barrier)
{
    __shared__ uint32_t s_seg[kSortPasses][kSortSegs];
    __shared__ int s_npass;
    __shared__ uint32_t s_whist[kSortWarps][256];
    __shared__ uint32_t s_base[256];
    __shared__ uint32_t s_warp[8];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t n = sort_key_count(ctl, cap);
    const uint32_t n_tiles = (n + kSortTile - 1) / kSortTile;
    if (threadIdx.x == 0) {
        uint32_t seg[kSortPasses][kSortSegs];
        s_npass = plan_digits(varying_bits(ctl, n), seg);
        for (int p = 0; p < kSortPasses; ++p)
            for (int q = 0; q < kSortSegs; ++q) s_seg[p][q] = seg[p][q];
        if (blockIdx.x == 0) {
            sc->n_passes = (uint32_t)s_npass;
            sc->result_buf = (uint32_t)(s_npass & 1);
            sc->n_keys = n;
        }
    }
    __syncthreads();
    const int n_passes = s_npass;
    uint32_t sync_no = 0;

    for (int pass = 0; pass < n_passes; ++pass) {
        const unsigned long long *in = (pass & 1) ? buf1 : buf0;
        unsigned long long *out = (pass & 1) ? buf0 : buf1;
        DigitDesc desc;
#pragma unroll
        for (int q = 0; q < kSortSegs; ++q) desc.seg[q] = s_seg[pass][q];

        // ---- A: digit counts of every tile ----
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            s_base[threadIdx.x] = 0;
            __syncthreads();
            const uint32_t wbase = tile * kSortTile + warp * (kSortItems * 32);
#pragma unroll
            for (int i = 0; i < kSortItems; ++i) {
                const uint32_t idx = wbase + i * 32 + lane;
                const bool valid = idx < n;
                const uint32_t digit = desc.of(valid ? in[idx] : 0ull);
                const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : (256u + lane));
                if (valid && (peers & lt_mask) == 0) atomicAdd(&s_base[digit], (uint32_t)__popc(peers));
            }
            __syncthreads();
            table[(size_t)tile * 256 + threadIdx.x] = s_base[threadIdx.x];
            __syncthreads();
        }
        grid_barrier(barrier, ++sync_no * gridDim.x);

        // ---- B: per digit, exclusive prefix over the tiles ----
        for (uint32_t d = blockIdx.x * kSortWarps + warp; d < 256u; d += gridDim.x * kSortWarps) {
            uint32_t running = 0;
            for (uint32_t base = 0; base < n_tiles; base += 32) {
                const uint32_t t = base + lane;
                const uint32_t c = t < n_tiles ? table[(size_t)t * 256 + d] : 0u;
                const uint32_t incl = warp_incl_scan(c, lane);
                if (t < n_tiles) table[(size_t)t * 256 + d] = running + incl - c;
                running += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) totals[d] = running;
        }
        grid_barrier(barrier, ++sync_no * gridDim.x);

        // ---- C: digit bases, ranks, scatter ----
        {
            const uint32_t c = totals[threadIdx.x];
            const uint32_t incl = warp_incl_scan(c, lane);
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            uint32_t off = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) if (w < warp) off += s_warp[w];
            s_base[threadIdx.x] = off + incl - c;
            __syncthreads();
        }
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int i = threadIdx.x; i < kSortWarps * 256; i += blockDim.x) (&s_whist[0][0])[i] = 0;
            __syncthreads();
            const uint32_t wbase = tile * kSortTile + warp * (kSortItems * 32);
            unsigned long long key[kSortItems];
            uint32_t rank[kSortItems], dig[kSortItems];
#pragma unroll
            for (int i = 0; i < kSortItems; ++i) {
                const uint32_t idx = wbase + i * 32 + lane;
                key[i] = idx < n ? in[idx] : ~0ull;
            }
#pragma unroll
            for (int i = 0; i < kSortItems; ++i) {
                const uint32_t idx = wbase + i * 32 + lane;
                const bool valid = idx < n;
                dig[i] = desc.of(key[i]);
                const uint32_t peers = __match_any_sync(0xffffffffu, valid ? dig[i] : (256u + lane));
                const uint32_t before = s_whist[warp][dig[i]];
                rank[i] = before + __popc(peers & lt_mask);
                __syncwarp();
                if (valid && (peers & lt_mask) == 0) s_whist[warp][dig[i]] = before + __popc(peers);
                __syncwarp();
            }
            __syncthreads();
            {
                const int d = threadIdx.x;
                uint32_t run = s_base[d] + table[(size_t)tile * 256 + d];
#pragma unroll
                for (int w = 0; w < kSortWarps; ++w) {
                    const uint32_t c = s_whist[w][d];
                    s_whist[w][d] = run;
                    run += c;
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < kSortItems; ++i) {
                const uint32_t idx = wbase + i * 32 + lane;
                if (idx < n) out[s_whist[warp][dig[i]] + rank[i]] = key[i];
            }
            __syncthreads();
        }
        if (pass + 1 < n_passes) grid_barrier(barrier, ++sync_no * gridDim.x);
    }
}

// Exclusive scan of counts[0..n) into offsets[0..n] (offsets[n] = total).  One block of 256.
__global__ void __launch_bounds__(256)
k_scan_counts(const uint32_t *__restrict__ counts, unsigned long long *__restrict__ offsets, const uint32_t n)
{
    __shared__ unsigned long long s_warp[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t per = (n + 255u) / 256u;
    const uint32_t lo = min(threadIdx.x * per, n), hi = min(lo + per, n);
    unsigned long long sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += counts[i];
    unsigned long long incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { if (w < warp) off += s_warp[w]; total += s_warp[w]; }
    unsigned long long run = off + incl - sum;
    for (uint32_t i = lo; i < hi; ++i) { offsets[i] = run; run += counts[i]; }
    if (threadIdx.x == 0) offsets[n] = total;
}

}  // namespace astre

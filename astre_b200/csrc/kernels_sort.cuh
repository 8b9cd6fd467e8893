// K3: device radix sort of the visible draw keys (64-bit, ascending).
//
// The reference draws in absl hash-map iteration order
// (engine/modules/Pipeline/src/deferred_shading.cpp:115,151) -- there is nothing
// to sort there; the order is spec-defined: ascending
// key, keys unique, so any correct sort reproduces the oracle's list.
//
// Least-significant-digit radix sort restricted to the key bits that can differ.
// The host knows the width of every key field in use (slots, meshes, shader ids
// seen, views, worlds); all other bits are zero in every key and cannot affect
// the order.  make_sort_plan packs the live bits, LSB first, into digits of up
// to 8 bits (each digit = up to 4 runs of consecutive key bits): a 16M-entity
// 5-view frame has 46 live bits -> 6 passes instead of 8.
//   (cull kernel) histograms every digit while it flushes its key buffers:
//                 per-CTA partials in shared memory, added to the CTA's own
//                 region of block_hist at kernel end -- no histogram pass
//   k_sort_scan   sums the partials, scans them, marks digits on which all keys
//                 agree as skipped (one block per digit)
//   k_sort_pass   per digit: single-pass ("onesweep") scatter -- rank inside a
//                 tile with warp match-any, chain tile prefixes through 64-bit
//                 (epoch | flag | count) look-back words, scatter
// The key count lives in device memory (FrameControl::key_total); no launch
// depends on a host read-back.
//
// Algorithmic bytes for V keys and P executed passes: V*(8 + 16*P).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstring>
#include "kernels_transform.cuh"

namespace astre {

constexpr int kSortThreads = 256;
constexpr uint32_t kReorderMinKeys = 1u << 20;         // from here on a pass is bandwidth-bound: reorder tiles in shared memory
constexpr int kSortWarps = kSortThreads / 32;
struct SortControl {
    uint32_t hist[kSortPasses][256];     // exclusive prefixes per digit (k_sort_scan)
    uint32_t skip[kSortPasses];          // 1 = every key has the same value of this digit
    uint32_t tile_counter[kSortPasses];  // dynamic tile ids
    uint32_t n_keys;                     // min(key_total, capacity)
    uint32_t epoch;                      // bumped every frame (< 2^27); tags look-back words
    uint32_t n_long;                     // tie runs too long for one thread (k_sort_ties_short -> k_sort_ties_long)
    uint32_t pad;
};

// Packs the bits of `varying`, LSB first, into digits of up to 8 bits, each made of up to
// kSortSegs runs of consecutive key bits.  Falls back to plain bytes for very fragmented masks.
inline SortPlan make_sort_plan(unsigned long long varying)
{
    SortPlan plan;
    std::memset(&plan, 0, sizeof plan);
    int n_passes = 0, bits = 0, segs = 0, b = 0;
    bool ok = true;
    while (b < 64 && ok) {
        if (!((varying >> b) & 1ull)) { ++b; continue; }
        int run = 0;
        while (b + run < 64 && ((varying >> (b + run)) & 1ull)) ++run;
        while (run > 0) {
            if (bits == 8 || segs == kSortSegs) {
                ++n_passes;
                if (n_passes == kSortPasses) { ok = false; break; }
                bits = 0; segs = 0;
            }
            const int take = run < 8 - bits ? run : 8 - bits;
            plan.seg[n_passes][segs++] = (uint32_t)b | ((uint32_t)take << 8) | ((uint32_t)bits << 16);
            bits += take; b += take; run -= take;
        }
    }
    if (ok) { plan.n_passes = (uint32_t)(bits > 0 ? n_passes + 1 : n_passes); return plan; }
    std::memset(&plan, 0, sizeof plan);
    for (int p = 0; p < 8; ++p) {
        if (!((varying >> (8 * p)) & 255ull)) continue;
        plan.seg[plan.n_passes][0] = (uint32_t)(8 * p) | (8u << 8);
        ++plan.n_passes;
    }
    return plan;
}

__device__ __forceinline__ uint32_t sort_key_count(const FrameControl *ctl, unsigned long long cap)
{
    const unsigned long long t = ctl->key_total;
    return (uint32_t)(t < cap ? t : cap);
}

// Block p: histogram of digit p = sum of the per-CTA partials the cull kernel left, then the
// exclusive scan over its 256 bins; a digit on which all keys agree is marked skipped.
// 1024 threads: four groups of 256 bins, each summing a quarter of the regions, 8 loads in flight.
__global__ void __launch_bounds__(1024)
k_sort_scan(SortControl *sc, const uint32_t *__restrict__ block_hist, const uint32_t n_regions,
            const FrameControl *ctl, const unsigned long long cap)
{
    __shared__ uint32_t s_part[4][256];
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_full;
    cudaGridDependencySynchronize();
    const uint32_t p = blockIdx.x;
    const uint32_t n = sort_key_count(ctl, cap);
    const int d = threadIdx.x & 255, g = threadIdx.x >> 8;
    if (threadIdx.x == 0) s_full = 0;
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint32_t *src = block_hist + (size_t)p * 256 + d;
    for (uint32_t b = g; b < n_regions; b += 32) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t r = b + 4u * j;
            if (r < n_regions) acc[j] += src[(size_t)r * kHistWords];
        }
    }
    s_part[g][d] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    __syncthreads();
    if (g == 0) {
        const int lane = d & 31, warp = d >> 5;
        const uint32_t c = (s_part[0][d] + s_part[1][d]) + (s_part[2][d] + s_part[3][d]);
        const uint32_t incl = warp_incl_scan(c, lane);
        if (lane == 31) s_warp[warp] = incl;
        if (c == n) s_full = 1;
        asm volatile("bar.sync 1, 256;");     // only the first 256 threads
        uint32_t off = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) if (w < warp) off += s_warp[w];
        sc->hist[p][d] = off + incl - c;
        if (d == 0) {
            sc->skip[p] = (s_full || n < 2) ? 1u : 0u;
            sc->tile_counter[p] = 0;
            if (p == 0) { sc->n_keys = n; sc->n_long = 0; }
        }
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

// The segments of one digit unpacked into registers (warp-uniform).
struct DigitRegs {
    uint32_t shift[kSortSegs], mask[kSortSegs], lshift[kSortSegs], n;
    __device__ __forceinline__ void load(const uint32_t (&seg)[kSortSegs])
    {
        n = 0;
#pragma unroll
        for (int s = 0; s < kSortSegs; ++s) {
            const uint32_t w = (seg[s] >> 8) & 255u;
            shift[s] = seg[s] & 255u; mask[s] = (1u << w) - 1u; lshift[s] = (seg[s] >> 16) & 255u;
            if (w) n = s + 1;
        }
    }
    __device__ __forceinline__ uint32_t of(unsigned long long key) const
    {
        uint32_t d = (uint32_t)(key >> shift[0]) & mask[0];
        if (n > 1) {
#pragma unroll
            for (int s = 1; s < kSortSegs; ++s) d |= ((uint32_t)(key >> shift[s]) & mask[s]) << lshift[s];
        }
        return d;
    }
};

// Lanes of the warp whose 8-bit digit equals this lane's: eight ballots (MATCH.ANY is slower here).
template <bool BALLOT>
__device__ __forceinline__ uint32_t match_digit(uint32_t d)
{
    if (!BALLOT) return __match_any_sync(0xffffffffu, d);
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        m &= bit ? bal : ~bal;
    }
    return m;
}

// One digit: single-pass scatter with decoupled look-back.  THREADS x ITEMS keys per tile.
// Measured (tools/sortbench.cu, 46 live bits): 256 x 8 is fastest up to a few million keys
// (17 us per pass at 0.7M keys, 10 us at 0.1M); 1024 x 8 wins from ~10M keys (fewer look-backs).
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS)
k_sort_pass(unsigned long long *__restrict__ buf0, unsigned long long *__restrict__ buf1,
            SortControl *sc, unsigned long long *__restrict__ lookback, const __grid_constant__ SortPlan plan, const int pass)
{
    constexpr int kTile = THREADS * ITEMS;
    constexpr int kWarps = THREADS / 32;
    __shared__ uint32_t s_whist[kWarps][256];
    __shared__ uint32_t s_gbase[256], s_dstart[256], s_wtot[8];
    __shared__ unsigned long long s_keys[THREADS == 256 ? kTile : 1];
    // Programmatic dependent launch: this grid may become resident while the previous kernel of the stream is
    // still draining; nothing it wrote is touched before this point.
    cudaGridDependencySynchronize();
    if (sc->skip[pass]) return;

    uint32_t executed = 0;                     // passes before this one that moved the keys
    for (int q = 0; q < pass; ++q) executed += sc->skip[q] ? 0u : 1u;
    const unsigned long long *in = (executed & 1u) ? buf1 : buf0;
    unsigned long long *out = (executed & 1u) ? buf0 : buf1;
    const uint32_t n = sc->n_keys;
    const uint32_t n_tiles = (n + kTile - 1) / kTile;
    const unsigned long long epoch = ((unsigned long long)sc->epoch * kSortPasses + pass + 1ull) << 34;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    DigitRegs digit;
    digit.load(plan.seg[pass]);

    // Static tile striding: the grid never exceeds the number of co-resident CTAs (the host sizes it from
    // the occupancy query), so every tile a CTA looks back to is owned by a CTA that is running or done.
    // (A shared tile counter would serialise ~27 ns per claim on one L2 atomic.)
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < kWarps * 256; i += THREADS) (&s_whist[0][0])[i] = 0;
        __syncthreads();

        // warp-striped load: item i of lane l is key tile*kTile + warp*ITEMS*32 + i*32 + l
        const uint32_t wbase = tile * kTile + warp * (ITEMS * 32);
        unsigned long long key[ITEMS];
        uint32_t rank[ITEMS], dig_of[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t idx = wbase + i * 32 + lane;
            key[i] = idx < n ? in[idx] : ~0ull;
        }
        // stable rank of every key among the keys of its warp with the same digit: the lowest peer lane
        // bumps the warp's counter (shared atomics of one warp retire in order) and hands the old value out
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t idx = wbase + i * 32 + lane;
            const bool valid = idx < n;
            const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
            const uint32_t dig = digit.of(key[i]);
            dig_of[i] = dig;
            const uint32_t peers = match_digit<true>(dig) & vmask;
            const int leader = __ffs(peers) - 1;
            uint32_t before = 0;
            if (valid && lane == leader) before = atomicAdd(&s_whist[warp][dig], (uint32_t)__popc(peers));
            before = __shfl_sync(0xffffffffu, before, leader < 0 ? 0 : leader);
            rank[i] = before + __popc(peers & lt_mask);
        }
        __syncthreads();

        // thread d < 256: digit d.  Exclusive prefix over the warps, tile total, look-back.
        uint32_t my_run = 0;
        if (threadIdx.x < 256) {
            const int d = threadIdx.x;
            uint32_t run = 0;
#pragma unroll 8
            for (int w = 0; w < kWarps; ++w) {
                const uint32_t c = s_whist[w][d];
                s_whist[w][d] = run;
                run += c;
            }
            my_run = run;
            unsigned long long *mine = lookback + (size_t)tile * 256 + d;
            uint32_t excl = 0;
            if (tile == 0) {
                st_relaxed(mine, epoch | kFlagPrefix | run);
            } else {
                st_relaxed(mine, epoch | kFlagAggregate | run);
                // Windowed look-back: kWindow predecessors are fetched at once (one L2 round trip),
                // then consumed nearest-first until a prefix closes the sum or an unpublished word
                // is met (that one is re-fetched as the head of the next window).
                constexpr int kWindow = 8;
                int t = (int)tile - 1;
                bool done = false;
                while (!done) {
                    unsigned long long w[kWindow];
#pragma unroll
                    for (int j = 0; j < kWindow; ++j)
                        w[j] = (t - j >= 0) ? ld_relaxed(lookback + (size_t)(t - j) * 256 + d) : (epoch | kFlagPrefix);
                    int used = 0;
#pragma unroll
                    for (int j = 0; j < kWindow; ++j) {
                        if (done || used != j) continue;
                        const bool ready = (w[j] >> 34) == (epoch >> 34) && (w[j] & (kFlagAggregate | kFlagPrefix));
                        if (!ready) continue;              // used stays j: the window ends here
                        excl += (uint32_t)w[j];
                        used = j + 1;
                        if (w[j] & kFlagPrefix) done = true;
                    }
                    t -= used;
                }
                st_relaxed(mine, epoch | kFlagPrefix | (unsigned long long)(excl + run));
            }
            s_gbase[d] = sc->hist[pass][d] + excl;
        }

        if (THREADS == 256 && n >= kReorderMinKeys) {
            // Bandwidth regime: order the tile by digit in shared memory first, so that consecutive threads
            // write consecutive keys of a digit (runs of full sectors instead of scattered 8-byte stores).
            {
                const uint32_t incl = warp_incl_scan(my_run, lane);       // tile-local start of every digit
                if (lane == 31) s_wtot[warp] = incl;
                __syncthreads();
                uint32_t off = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) if (w < warp) off += s_wtot[w];
                s_dstart[threadIdx.x] = off + incl - my_run;
            }
            __syncthreads();
            const uint32_t tile_keys = min((uint32_t)kTile, n - tile * kTile);
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t idx = wbase + i * 32 + lane;
                if (idx < n) s_keys[s_dstart[dig_of[i]] + s_whist[warp][dig_of[i]] + rank[i]] = key[i];
            }
            __syncthreads();
            for (uint32_t j = threadIdx.x; j < tile_keys; j += THREADS) {
                const unsigned long long k = s_keys[j];
                const uint32_t dg = digit.of(k);
                out[s_gbase[dg] + (j - s_dstart[dg])] = k;
            }
        } else {
            __syncthreads();
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t idx = wbase + i * 32 + lane;
                if (idx < n) out[s_gbase[dig_of[i]] + s_whist[warp][dig_of[i]] + rank[i]] = key[i];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Ties (optional mode, ASTRE_SORT_TIES=1; off by default -- it wins 2 % on the sparse headline frame and loses
// 40 % on a dense one).  The radix passes then run over the bits ABOVE the slot field only (depth, mesh, shader,
// view, world: 22 live bits in a 16M-entity 5-view frame -> 3 passes instead of 6).  They are stable, so keys that
// agree on all those bits end up adjacent, in emission order; the full key order additionally wants them
// ascending in slot.  Such runs are short (a run = same view, shader, mesh AND quantised depth):
//   k_sort_ties_short  one thread per run head: runs of <= 32 keys are insertion-sorted in place, longer
//                      ones are queued;
//   k_sort_ties_long   one CTA per queued run: finds its end by binary search and radix-sorts its slot
//                      field (8 bits per pass) through the spare key buffer.
// Correct for any run length; a scene that puts millions of entities into one depth bucket of one mesh
// degrades to a single CTA sorting that run.
// ---------------------------------------------------------------------------
constexpr int kShortRun = 32;

__device__ __forceinline__ uint32_t sort_result_buffer(const SortControl *sc, uint32_t n_passes)
{
    uint32_t executed = 0;
    for (uint32_t q = 0; q < n_passes; ++q) executed += sc->skip[q] ? 0u : 1u;
    return executed & 1u;
}

__global__ void __launch_bounds__(256)
k_sort_ties_short(unsigned long long *__restrict__ buf0, unsigned long long *__restrict__ buf1, SortControl *sc,
                  const uint32_t n_passes, const uint32_t low_bits, uint32_t *__restrict__ long_list, const uint32_t cap_long)
{
    cudaGridDependencySynchronize();
    unsigned long long *keys = sort_result_buffer(sc, n_passes) ? buf1 : buf0;
    const uint32_t n = sc->n_keys;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long k0 = keys[i];
        const unsigned long long u = k0 >> low_bits;
        if (i > 0 && (keys[i - 1] >> low_bits) == u) continue;          // not a run head
        if (i + 1 >= n || (keys[i + 1] >> low_bits) != u) continue;     // run of one
        unsigned long long a[kShortRun];
        a[0] = k0;
        uint32_t len = 1;
        while (len < (uint32_t)kShortRun && i + len < n) {
            const unsigned long long k = keys[i + len];
            if ((k >> low_bits) != u) break;
            a[len++] = k;
        }
        if (len == (uint32_t)kShortRun && i + len < n && (keys[i + len] >> low_bits) == u) {
            const uint32_t at = atomicAdd(&sc->n_long, 1u);              // longer than one thread should sort
            if (at < cap_long) long_list[at] = i;
            continue;
        }
        for (uint32_t x = 1; x < len; ++x) {                             // insertion sort, keys are unique
            const unsigned long long v = a[x];
            uint32_t y = x;
            while (y > 0 && a[y - 1] > v) { a[y] = a[y - 1]; --y; }
            a[y] = v;
        }
        for (uint32_t x = 0; x < len; ++x) keys[i + x] = a[x];
    }
}

__global__ void __launch_bounds__(256)
k_sort_ties_long(unsigned long long *__restrict__ buf0, unsigned long long *__restrict__ buf1, SortControl *sc,
                 const uint32_t n_passes, const uint32_t low_bits, const uint32_t low_live_bits,
                 const uint32_t *__restrict__ long_list, const uint32_t cap_long)
{
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_whist[8][256];
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_end;
    cudaGridDependencySynchronize();
    const uint32_t n_long = min(sc->n_long, cap_long);
    if (n_long == 0) return;
    const uint32_t res = sort_result_buffer(sc, n_passes);
    unsigned long long *keys = res ? buf1 : buf0, *spare = res ? buf0 : buf1;
    const uint32_t n = sc->n_keys;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;

    for (uint32_t r = blockIdx.x; r < n_long; r += gridDim.x) {
        const uint32_t start = long_list[r];
        if (threadIdx.x == 0) {                       // end of the run: first index whose upper bits differ
            const unsigned long long u = keys[start] >> low_bits;
            uint32_t lo = start + 1, hi = n;          // keys are sorted by upper bits: binary search
            while (lo < hi) {
                const uint32_t mid = lo + (hi - lo) / 2;
                if ((keys[mid] >> low_bits) == u) lo = mid + 1; else hi = mid;
            }
            s_end = lo;
        }
        __syncthreads();
        const uint32_t len = s_end - start;
        unsigned long long *src = keys + start, *dst = spare + start;
        uint32_t passes = 0;
        for (uint32_t shift = 0; shift < low_live_bits; shift += 8, ++passes) {
            s_hist[threadIdx.x] = 0;
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < len; i += 256) atomicAdd(&s_hist[(uint32_t)(src[i] >> shift) & 255u], 1u);
            __syncthreads();
            {   // exclusive scan of the 256 bins -> running output offsets
                const uint32_t c = s_hist[threadIdx.x];
                const uint32_t incl = warp_incl_scan(c, lane);
                if (lane == 31) s_warp[warp] = incl;
                __syncthreads();
                uint32_t off = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) if (w < warp) off += s_warp[w];
                s_hist[threadIdx.x] = off + incl - c;
            }
            __syncthreads();
            // stable scatter, 256 keys at a time: warp ranks (ballot match), warp order, running offsets
            for (uint32_t base = 0; base < len; base += 256) {
                for (int i = threadIdx.x; i < 8 * 256; i += 256) (&s_whist[0][0])[i] = 0;
                __syncthreads();
                const uint32_t i = base + threadIdx.x;
                const bool valid = i < len;
                const unsigned long long k = valid ? src[i] : 0ull;
                const uint32_t dig = (uint32_t)(k >> shift) & 255u;
                const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
                const uint32_t peers = match_digit<true>(dig) & vmask;
                const uint32_t rank = __popc(peers & lt_mask);
                if (valid && rank == 0) s_whist[warp][dig] = __popc(peers);
                __syncthreads();
                uint32_t before = 0;                  // keys of earlier warps of this chunk with my digit
                for (int w = 0; w < warp; ++w) before += s_whist[w][dig];
                if (valid) dst[s_hist[dig] + before + rank] = k;
                __syncthreads();
                {   // advance the running offsets by this chunk's counts
                    uint32_t c = 0;
#pragma unroll
                    for (int w = 0; w < 8; ++w) c += s_whist[w][threadIdx.x];
                    s_hist[threadIdx.x] += c;
                }
                __syncthreads();
            }
            unsigned long long *t = src; src = dst; dst = t;
            __syncthreads();
        }
        if (passes & 1u) {                            // the sorted run sits in the spare buffer: copy it home
            for (uint32_t i = threadIdx.x; i < len; i += 256) keys[start + i] = spare[start + i];
        }
        __syncthreads();
    }
}

// Exclusive scan of counts[0..n) into offsets[0..n] (offsets[n] = total).  One block of 256.
__global__ void __launch_bounds__(256)
k_scan_counts(const uint32_t *__restrict__ counts, unsigned long long *__restrict__ offsets, const uint32_t n)
{
    __shared__ unsigned long long s_warp[8];
    cudaGridDependencySynchronize();
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

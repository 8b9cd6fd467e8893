// K3: device sort of the visible draw keys (64-bit, ascending, keys unique).
//
// The reference draws in absl hash-map iteration order
// (engine/modules/Pipeline/src/deferred_shading.cpp:115,151) -- there is nothing
// to sort there; the order is spec-defined: ascending
// key, keys unique, so any correct sort reproduces the oracle's list bit for bit.
//
// A frame's keys (0.7 M on the headline frame, a few million at most) live in L2, so the sort is bound by
// dependent latencies and launches, not by bytes.  Round 1's six-pass LSD onesweep (one look-back chain per
// pass) took 0.094 ms of a 0.385 ms frame at 12 % of the HBM roofline.  This is a BUCKET SORT on the key bits
// that can differ (the host knows the width of every key field in use; all other bits are zero in every key):
//
//   (cull kernels)  count every key into bins[d], d = the top `bits` live key bits (one RED per key; `bits`
//                   is chosen by the host so that a bucket holds ~4-8 keys, from the previous frame's key count)
//   k_bin_scan      exclusive scan of the bins in place (decoupled look-back over 4096-bin tiles), largest
//                   bucket, sum of squared bucket sizes, and the first bucket of every 1024-key chunk
//   k_bin_scatter   key -> its bucket, position from an atomic cursor: order inside a bucket is arbitrary,
//                   which is fine because
//   k_bin_rank      a CTA per chunk of whole buckets (<= 3072 keys in shared memory) ranks every key among the
//                   keys of its own bucket by counting (keys are unique: the rank is exact) and writes it to
//                   its final position.
// No pass waits for another CTA (except the short scan chain), nothing is stable, nothing needs a histogram
// pass of its own: 3 short kernels.  Cost is sum(m^2) over bucket sizes m; a frame whose keys crowd into a few
// buckets (any bucket > 2048 keys, or sum(m^2) > 256 n) is handed, ON THE DEVICE and without a host round trip,
// to the fallback:
//   k_sort_lsd_all  the round-1 LSD onesweep (digits of <= 8 live bits, tile prefixes chained through 64-bit
//                   (epoch | flag | count) look-back words), all passes in one launch of a co-resident grid with
//                   grid barriers between passes; it returns at once when the bucket sort ran.
// Which path ran is recorded in SortControl::path (astre_gpu_last_sort_path).
//
// Algorithmic bytes for V keys: bucket path V*(8 + 16 + 16) (scatter and rank each read and write the keys once);
// fallback with P executed passes V*(8 + 16*P).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstring>
#include "kernels_transform.cuh"

namespace astre {

constexpr int kSortThreads = 256;
constexpr uint32_t kReorderMinKeys = 1u << 20;         // LSD fallback: from here on a pass is bandwidth-bound, reorder tiles in shared memory
constexpr int kSortWarps = kSortThreads / 32;

constexpr uint32_t kMinBinBits = 8, kMaxBinBits = 22;
constexpr uint32_t kScanTile = 4096;                   // bins per CTA of k_bin_scan (1024 threads x 4)
constexpr uint32_t kMaxScanTiles = (1u << kMaxBinBits) / kScanTile;
constexpr uint32_t kChunkKeys = 1024;                  // chunk c of the rank pass = the buckets that start in [1024 c, 1024 c + 1024)
constexpr uint32_t kMaxBucket = 2048;                  // a larger bucket sends the frame to the LSD fallback
constexpr uint32_t kRankSmemKeys = kChunkKeys + kMaxBucket;
constexpr unsigned long long kMaxMeanSquare = 256;     // fallback when sum(m^2) > this * n

enum SortPath { kSortNone = 0, kSortBins = 1, kSortLsd = 2 };

struct SortControl {
    uint32_t hist[kSortPasses][256];     // LSD fallback: digit counts, then their exclusive prefixes
    uint32_t skip[kSortPasses];          // LSD fallback: 1 = every key has the same value of this digit
    uint32_t tile_counter[kSortPasses];  // LSD fallback: next unclaimed tile of a pass
    uint32_t n_keys;                     // min(key_total, capacity)
    uint32_t epoch;                      // bumped every frame (< 2^27); tags look-back words
    uint32_t max_bucket;                 // bucket sort: largest bucket (k_bin_scan)
    uint32_t scan_ticket;                // bucket sort: next tile of the bin scan
    unsigned long long sum_sq;           // bucket sort: sum of squared bucket sizes
    uint32_t result_buf;                 // key buffer (0 / 1) that holds the sorted list
    uint32_t lsd_barrier;                // arrivals at the fallback's grid barrier
    uint32_t path;                       // SortPath of the last sort
    uint32_t n_bins;                     // bucket sort: bins of this frame (k_bin_scan)
    unsigned long long scan_state[kMaxScanTiles];   // look-back words of the bin scan
};

// ---- plans (host) -----------------------------------------------------------------------------------------

// MSD digit of the bucket sort: the top min(want_bits, live bits) bits of `live`, in key order.
inline BinPlan make_bin_plan(unsigned long long live, uint32_t want_bits)
{
    BinPlan bp;
    std::memset(&bp, 0, sizeof bp);
    uint32_t total = 0;
    for (int b = 0; b < 64; ++b) total += (uint32_t)((live >> b) & 1ull);
    uint32_t bits = want_bits < total ? want_bits : total, remaining = bits;
    int b = 63;
    while (remaining && b >= 0) {
        if (!((live >> b) & 1ull)) { --b; continue; }
        const int hi = b;
        while (b >= 0 && ((live >> b) & 1ull)) --b;                 // the run is bits (b, hi]
        const uint32_t len = (uint32_t)(hi - b), take = len < remaining ? len : remaining;
        if (bp.n_segs == (uint32_t)kBinSegs) break;                 // very fragmented mask: settle for fewer bits
        bp.shift[bp.n_segs] = (uint32_t)hi - take + 1u;
        bp.mask[bp.n_segs] = (1u << take) - 1u;
        bp.lshift[bp.n_segs] = remaining - take;
        ++bp.n_segs;
        remaining -= take;
    }
    if (remaining) {                                                // bits that found no segment: drop them from the digit
        for (uint32_t s = 0; s < bp.n_segs; ++s) bp.lshift[s] -= remaining;
        bits -= remaining;
    }
    bp.bits = bits;
    return bp;
}

// LSD fallback: packs the bits of `varying`, LSB first, into digits of up to 8 bits, each made of up to
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

// ---- shared helpers ---------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t sort_key_count(const FrameControl *ctl, unsigned long long cap)
{
    const unsigned long long t = ctl->key_total;
    return (uint32_t)(t < cap ? t : cap);
}

// Does this frame's sort go to the LSD fallback?  Evaluated identically by every kernel of the chain
// (max_bucket and sum_sq are final when k_bin_scan has ended).
__device__ __forceinline__ bool sort_takes_fallback(const SortControl *sc, uint32_t n, uint32_t force)
{
    return force || sc->max_bucket > kMaxBucket || sc->sum_sq > kMaxMeanSquare * (unsigned long long)n;
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
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Exclusive scan of counts[0..n) into offsets[0..n] (offsets[n] = total) by one block of 256 threads.
__device__ __forceinline__ void scan_counts_block(const uint32_t *__restrict__ counts, unsigned long long *__restrict__ offsets,
                                                  const uint32_t n, unsigned long long *s_warp8)
{
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
    if (lane == 31) s_warp8[warp] = incl;
    __syncthreads();
    unsigned long long off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { if (w < warp) off += s_warp8[w]; total += s_warp8[w]; }
    unsigned long long run = off + incl - sum;
    for (uint32_t i = lo; i < hi; ++i) { offsets[i] = run; run += counts[i]; }
    if (threadIdx.x == 0) offsets[n] = total;
    __syncthreads();
}

// What the host wants to know about a finished frame, mirrored into MAPPED HOST memory by the frame's last launch:
// reading a frame's results then costs one stream synchronisation and no copy (small frames are all latency).
// key_total doubles as the hint that sizes the next frame's buckets.
constexpr uint32_t kInlineCounts = 256;
constexpr uint32_t kInlineList = 4096;                  // draw lists up to this length also land in host memory (ASTRE_FRAME_GATHER)
constexpr uint32_t kNoInlineList = 0xFFFFFFFFu;
struct FrameResultHost {
    unsigned long long key_total;
    uint32_t overflow, result_buf, path;
    uint32_t n_counts;                                  // counts / offsets mirrored below (0: too many, copy them from the device)
    uint32_t counts[kInlineCounts];
    unsigned long long offsets[kInlineCounts + 1];
    uint32_t list_n;                                    // k_gather_list: length of the list mirrored below, or kNoInlineList
    uint32_t n_bins;                                    // bucket sort feedback: bins of this frame and the sum of squared bucket
    unsigned long long sum_sq;                          // sizes (sum_sq / key_total = mean length of a key's rank loop)
    unsigned long long list_keys[kInlineList];
    float4 list_world[kInlineList * 4];
};

// Optional device-side copy of the per-view counts, alternating between two caller buffers by frame parity: what a
// multi-GPU exchange reads while the NEXT frame (which zeroes the live counters) is already running.
struct CountsMirror { uint32_t *slot[2]; };

// by one block of 256 threads, after scan_counts_block
__device__ __forceinline__ void publish_results(FrameResultHost *res, const FrameControl *ctl, const SortControl *sc,
                                                const uint32_t *__restrict__ counts, const unsigned long long *__restrict__ offsets,
                                                const uint32_t n_counts, const bool with_counts, const CountsMirror &mirror)
{
    if (with_counts && mirror.slot[0]) {
        uint32_t *dst = (sc->epoch & 1u) ? mirror.slot[1] : mirror.slot[0];     // (a select: indexing the parameter would put it in local memory)
        for (uint32_t i = threadIdx.x; i < n_counts; i += 256u) dst[i] = counts[i];
    }
    if (!res) return;
    if (threadIdx.x == 0) {
        res->key_total = ctl->key_total;
        res->overflow = ctl->overflow;
        res->result_buf = sc->result_buf;
        res->path = sc->path;
        res->n_bins = sc->n_bins;
        res->sum_sq = sc->sum_sq;
        res->list_n = kNoInlineList;                    // k_gather_list, if the frame has one, runs after this
    }
    if (with_counts) {
        const bool inl = n_counts <= kInlineCounts;
        if (threadIdx.x == 0) res->n_counts = inl ? n_counts : 0u;
        if (inl) {
            if (threadIdx.x < n_counts) res->counts[threadIdx.x] = counts[threadIdx.x];
            for (uint32_t i = threadIdx.x; i <= n_counts; i += 256u) res->offsets[i] = offsets[i];
        }
    }
}

// Frames that do not sort in the same call: the scan of the per-view counts as its own launch
// (sorted frames get it from block 0 of k_sort_lsd_all).
__global__ void __launch_bounds__(256)
k_scan_counts(const uint32_t *__restrict__ counts, unsigned long long *__restrict__ offsets, const uint32_t n,
              const FrameControl *ctl, const SortControl *sc, FrameResultHost *res, const CountsMirror mirror)
{
    __shared__ unsigned long long s_warp[8];
    cudaGridDependencySynchronize();
    scan_counts_block(counts, offsets, n, s_warp);
    publish_results(res, ctl, sc, counts, offsets, n, true, mirror);
}

// ---- bucket sort ------------------------------------------------------------------------------------------

// Exclusive scan of the bins in place (bins[d] becomes the start of bucket d = the scatter's cursor), the
// largest bucket, the sum of squared bucket sizes and chunk_first[c] = first bucket that starts at or after
// key position 1024 c.  Tiles of 4096 bins are taken in ticket order and chained by decoupled look-back
// (one warp reads 32 predecessors per round trip).
__global__ void __launch_bounds__(1024)
k_bin_scan(uint32_t *__restrict__ bins, const uint32_t n_bins, SortControl *sc, uint32_t *__restrict__ chunk_first)
{
    __shared__ uint32_t s_warp[32], s_max[32];
    __shared__ unsigned long long s_sq[32];
    __shared__ uint32_t s_tile, s_base;
    cudaGridDependencySynchronize();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(&sc->scan_ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const unsigned long long tag = (unsigned long long)sc->epoch << 34;
    const uint32_t idx4 = tile * 1024u + threadIdx.x, n4 = n_bins >> 2;
    uint4 c = make_uint4(0, 0, 0, 0);
    if (idx4 < n4) c = reinterpret_cast<const uint4 *>(bins)[idx4];
    const uint32_t mine = (c.x + c.y) + (c.z + c.w);
    const uint32_t incl = warp_incl_scan(mine, lane);
    if (lane == 31) s_warp[warp] = incl;
    {   // largest bucket and sum of squares of this thread's four
        uint32_t mx = max(max(c.x, c.y), max(c.z, c.w));
        unsigned long long sq = (unsigned long long)c.x * c.x + (unsigned long long)c.y * c.y +
                                (unsigned long long)c.z * c.z + (unsigned long long)c.w * c.w;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
            sq += __shfl_xor_sync(0xffffffffu, sq, d);
        }
        if (lane == 0) { s_max[warp] = mx; s_sq[warp] = sq; }
    }
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_warp[lane];
        const uint32_t wincl = warp_incl_scan(w, lane);
        s_warp[lane] = wincl - w;                                     // exclusive prefix of every warp
        const uint32_t total = __shfl_sync(0xffffffffu, wincl, 31);
        uint32_t mx = s_max[lane];
        unsigned long long sq = s_sq[lane];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
            sq += __shfl_xor_sync(0xffffffffu, sq, d);
        }
        if (lane == 0) {
            if (mx) atomicMax(&sc->max_bucket, mx);
            if (sq) atomicAdd(&sc->sum_sq, sq);
        }
        // decoupled look-back: publish this tile's total, then sum the predecessors' words nearest-first
        unsigned long long *state = sc->scan_state;
        uint32_t base = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed(state, tag | kFlagPrefix | total);
        } else {
            if (lane == 0) st_relaxed(state + tile, tag | kFlagAggregate | total);
            int t = (int)tile - 1;
            bool done = false;
            while (!done) {
                const int at = t - lane;
                const unsigned long long wd = at >= 0 ? ld_relaxed(state + at) : (tag | kFlagPrefix);
                const bool ready = (wd >> 34) == (tag >> 34) && (wd & (kFlagAggregate | kFlagPrefix)) != 0;
                const uint32_t rmask = __ballot_sync(0xffffffffu, ready);
                const uint32_t n_ready = rmask == 0xffffffffu ? 32u : (uint32_t)__ffs(~rmask) - 1u;   // ready words, contiguous from the nearest
                const uint32_t in_run = n_ready == 32u ? 0xffffffffu : ((1u << n_ready) - 1u);
                const uint32_t pmask = __ballot_sync(0xffffffffu, ready && (wd & kFlagPrefix)) & in_run;
                const uint32_t use = pmask ? (uint32_t)__ffs(pmask) : n_ready;                          // up to and including the first prefix
                uint32_t v = (uint32_t)lane < use ? (uint32_t)wd : 0u;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                base += v;
                if (pmask) done = true; else t -= (int)use;
            }
            if (lane == 0) st_relaxed(state + tile, tag | kFlagPrefix | (unsigned long long)(base + total));
        }
        if (lane == 0) s_base = base;
    }
    __syncthreads();
    const uint32_t o0 = s_base + s_warp[warp] + incl - mine;        // start of this thread's first bucket
    const uint32_t o1 = o0 + c.x, o2 = o1 + c.y, o3 = o2 + c.z, o4 = o3 + c.w;
    if (idx4 < n4) reinterpret_cast<uint4 *>(bins)[idx4] = make_uint4(o0, o1, o2, o3);
    // bucket e = [start, end) is the last one to start before every chunk boundary 1024 c in (start, end]
    const uint32_t e0 = idx4 * 4u;
    const uint32_t st[5] = { o0, o1, o2, o3, o4 };
    // (a bucket beyond kMaxBucket sends the frame to the fallback, which needs no chunk table: skip its long walk)
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (st[i + 1] - st[i] <= kMaxBucket)
            for (uint32_t cc = st[i] / kChunkKeys + 1u; cc * kChunkKeys <= st[i + 1]; ++cc) chunk_first[cc] = e0 + i + 1u;
    if (tile == 0 && threadIdx.x == 0) { chunk_first[0] = 0u; sc->n_bins = n_bins; }
}

// key -> its bucket; the position inside the bucket comes from an atomic cursor (unordered).
__global__ void __launch_bounds__(256)
k_bin_scatter(const unsigned long long *__restrict__ in, unsigned long long *__restrict__ out, uint32_t *__restrict__ cursor,
              const __grid_constant__ BinPlan bp, const SortControl *sc, const FrameControl *ctl,
              const unsigned long long cap, const uint32_t force_lsd)
{
    cudaGridDependencySynchronize();
    const uint32_t n = sort_key_count(ctl, cap);
    if (sort_takes_fallback(sc, n, force_lsd)) return;
    for (uint32_t base = blockIdx.x * 1024u; base < n; base += gridDim.x * 1024u) {
        unsigned long long k[4];
        uint32_t pos[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t idx = base + j * 256u + threadIdx.x;
            k[j] = idx < n ? in[idx] : 0ull;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t idx = base + j * 256u + threadIdx.x;
            pos[j] = idx < n ? atomicAdd(&cursor[bin_digit(bp, k[j])], 1u) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t idx = base + j * 256u + threadIdx.x;
            if (idx < n) out[pos[j]] = k[j];
        }
    }
}

// After the scatter cursor[d] is the END of bucket d (= the start of bucket d+1).  A CTA takes chunk c = the
// buckets that start in [1024 c, 1024 c + 1024) -- whole buckets, at most 1024 + 2047 keys -- into shared
// memory; every key is ranked among the keys of its own bucket and stored at bucket start + rank.
__global__ void __launch_bounds__(256)
k_bin_rank(const unsigned long long *__restrict__ in, unsigned long long *__restrict__ out, const uint32_t *__restrict__ cursor,
           const uint32_t *__restrict__ chunk_first, const __grid_constant__ BinPlan bp, SortControl *sc,
           const FrameControl *ctl, const unsigned long long cap, const uint32_t force_lsd)
{
    __shared__ unsigned long long s_keys[kRankSmemKeys];
    cudaGridDependencySynchronize();
    const uint32_t n = sort_key_count(ctl, cap);
    if (sort_takes_fallback(sc, n, force_lsd)) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) { sc->result_buf = 0u; sc->path = kSortBins; sc->n_keys = n; }
    const uint32_t n_chunks = n / kChunkKeys + 1u;                    // chunk c exists while 1024 c <= n
    for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uint32_t d0 = chunk_first[c];
        const uint32_t lo = d0 ? cursor[d0 - 1u] : 0u;
        uint32_t hi = n;
        if (c + 1u < n_chunks) { const uint32_t d1 = chunk_first[c + 1u]; hi = d1 ? cursor[d1 - 1u] : 0u; }
        const uint32_t m = hi - lo;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < m; i += 256u) s_keys[i] = in[lo + i];
        __syncthreads();
        for (uint32_t i0 = threadIdx.x; i0 < m; i0 += 4u * 256u) {
            // four keys per thread and round: their bucket bounds (L2 round trips) are fetched together
            unsigned long long key[4];
            uint32_t bs[4], be[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t i = i0 + q * 256u;
                key[q] = i < m ? s_keys[i] : 0ull;
                const uint32_t d = bin_digit(bp, key[q]);
                bs[q] = i < m ? (d ? __ldg(cursor + d - 1u) : 0u) - lo : 0u;
                be[q] = i < m ? __ldg(cursor + d) - lo : 0u;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t r = 0;
                for (uint32_t j = bs[q]; j < be[q]; ++j) r += s_keys[j] < key[q] ? 1u : 0u;
                if (i0 + q * 256u < m) out[lo + bs[q] + r] = key[q];
            }
        }
    }
}

// ---- LSD fallback -----------------------------------------------------------------------------------------

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
__device__ __forceinline__ uint32_t match_digit(uint32_t d)
{
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        m &= bit ? bal : ~bal;
    }
    return m;
}

__device__ __forceinline__ void sort_grid_barrier(uint32_t *counter, uint32_t target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire_u32(counter) < target) { }
        __threadfence();
    }
    __syncthreads();
}

// All passes of the LSD onesweep in one launch.  The grid must be co-resident (the host sizes it from the
// occupancy query; one stream runs one frame at a time, so nothing else competes for the SMs for long).
// Keys are read with ld.global.cg: a key buffer is rewritten by other CTAs two passes later within this launch.
__global__ void __launch_bounds__(kSortThreads)
k_sort_lsd_all(unsigned long long *__restrict__ buf0, unsigned long long *__restrict__ buf1, SortControl *sc,
               unsigned long long *__restrict__ lookback, const __grid_constant__ SortPlan plan, const FrameControl *ctl,
               const unsigned long long cap, const uint32_t force_lsd,
               const uint32_t *__restrict__ counts, unsigned long long *__restrict__ offsets, const uint32_t n_counts,
               FrameResultHost *res, const CountsMirror mirror)
{
    constexpr int THREADS = kSortThreads, ITEMS = 8;
    constexpr int kTile = THREADS * ITEMS;
    constexpr int kWarps = THREADS / 32;
    __shared__ uint32_t s_whist[kWarps][256];
    __shared__ uint32_t s_gbase[256], s_dstart[256], s_wtot[8];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_keys[kTile];
    cudaGridDependencySynchronize();
    const uint32_t n = sort_key_count(ctl, cap);
    // the frame's last launch: block 0 also scans the per-(world, view) counts into list offsets
    if (blockIdx.x == 0 && n_counts) scan_counts_block(counts, offsets, n_counts, s_keys);
    if (!sort_takes_fallback(sc, n, force_lsd)) {
        if (blockIdx.x == 0) publish_results(res, ctl, sc, counts, offsets, n_counts, n_counts != 0, mirror);   // the bucket sort ran
        return;
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t P = plan.n_passes;
    uint32_t round = 0;

    // phase 0: histogram of every digit (the cull kernels only counted the bucket sort's bins)
    for (int i = threadIdx.x; i < kWarps * 256; i += THREADS) (&s_whist[0][0])[i] = 0;
    __syncthreads();
    for (uint32_t idx = blockIdx.x * THREADS + threadIdx.x; idx < n; idx += gridDim.x * THREADS) {
        const unsigned long long k = __ldcg(buf0 + idx);
        for (uint32_t q = 0; q < P; ++q) atomicAdd(&s_whist[q][sort_digit(plan.seg[q], k)], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < P * 256u; i += THREADS) {
        const uint32_t v = (&s_whist[0][0])[i];
        if (v) atomicAdd(&sc->hist[0][0] + i, v);
    }
    sort_grid_barrier(&sc->lsd_barrier, ++round * gridDim.x);

    // phase 1: block p scans digit p; a digit on which all keys agree is skipped
    if (blockIdx.x < P) {
        const uint32_t p = blockIdx.x, d = threadIdx.x;
        const uint32_t c = __ldcg(&sc->hist[p][d]);
        const uint32_t incl = warp_incl_scan(c, lane);
        if (lane == 31) s_wtot[warp] = incl;
        if (threadIdx.x == 0) s_tile = 0;
        __syncthreads();
        if (c == n) s_tile = 1;
        uint32_t off = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) if (w < warp) off += s_wtot[w];
        __syncthreads();
        sc->hist[p][d] = off + incl - c;
        if (d == 0) sc->skip[p] = (s_tile || n < 2) ? 1u : 0u;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) sc->n_keys = n;
    sort_grid_barrier(&sc->lsd_barrier, ++round * gridDim.x);

    uint32_t executed = 0;
    for (uint32_t pass = 0; pass < P; ++pass) {
        if (!__ldcg(&sc->skip[pass])) {
            const unsigned long long *in = (executed & 1u) ? buf1 : buf0;
            unsigned long long *out = (executed & 1u) ? buf0 : buf1;
            ++executed;
            const uint32_t n_tiles = (n + kTile - 1) / kTile;
            const unsigned long long epoch = ((unsigned long long)sc->epoch * kSortPasses + pass + 1ull) << 34;
            DigitRegs digit;
            digit.load(plan.seg[pass]);
            // Tiles are claimed in arrival order: every tile a CTA looks back to has been claimed by a running CTA.
            for (;;) {
                __syncthreads();
                if (threadIdx.x == 0) s_tile = atomicAdd(&sc->tile_counter[pass], 1u);
                for (int i = threadIdx.x; i < kWarps * 256; i += THREADS) (&s_whist[0][0])[i] = 0;
                __syncthreads();
                const uint32_t tile = s_tile;
                if (tile >= n_tiles) break;

                // warp-striped load: item i of lane l is key tile*kTile + warp*ITEMS*32 + i*32 + l
                const uint32_t wbase = tile * kTile + warp * (ITEMS * 32);
                unsigned long long key[ITEMS];
                uint32_t rank[ITEMS], dig_of[ITEMS];
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) {
                    const uint32_t idx = wbase + i * 32 + lane;
                    key[i] = idx < n ? __ldcg(in + idx) : ~0ull;
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
                    const uint32_t peers = match_digit(dig) & vmask;
                    const int leader = __ffs(peers) - 1;
                    uint32_t before = 0;
                    if (valid && lane == leader) before = atomicAdd(&s_whist[warp][dig], (uint32_t)__popc(peers));
                    before = __shfl_sync(0xffffffffu, before, leader < 0 ? 0 : leader);
                    rank[i] = before + __popc(peers & lt_mask);
                }
                __syncthreads();

                // thread d: digit d.  Exclusive prefix over the warps, tile total, look-back.
                uint32_t my_run = 0;
                {
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
                    s_gbase[d] = __ldcg(&sc->hist[pass][d]) + excl;
                }

                if (n >= kReorderMinKeys) {
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
        sort_grid_barrier(&sc->lsd_barrier, ++round * gridDim.x);
    }
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) { sc->result_buf = executed & 1u; sc->path = kSortLsd; }
        __syncthreads();
        publish_results(res, ctl, sc, counts, offsets, n_counts, n_counts != 0, mirror);
    }
}

// ASTRE_FRAME_GATHER: the world matrices of the draw list's entities in list order (what VisualSystem::run puts into
// render_proxies[e].inputs.in_mat4["uModel"], visual_system.cpp:47), produced inside the frame so that reading the list
// back needs no further launch; short lists also land, keys and matrices, in mapped host memory (no copy at all).
__global__ void __launch_bounds__(256)
k_gather_list(const unsigned long long *__restrict__ buf0, const unsigned long long *__restrict__ buf1, const SortControl *sc,
              const FrameControl *ctl, const unsigned long long key_cap, const uint32_t sorted,
              const float4 *__restrict__ world, float4 *__restrict__ out, const unsigned long long out_cap,
              const uint32_t world_bits, const uint32_t entities_per_world, FrameResultHost *res)
{
    cudaGridDependencySynchronize();
    const unsigned long long n = min(ctl->key_total, key_cap);
    const unsigned long long *keys = (sorted && sc->result_buf) ? buf1 : buf0;
    const unsigned long long m = min(n, out_cap);
    const bool inl = res != nullptr && n <= kInlineList && ctl->key_total <= key_cap;
    const uint32_t local_bits = 26u - world_bits;
    for (unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; j < m * 4ull;
         j += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[j >> 2];
        const uint32_t local = (uint32_t)(key & ((1ull << local_bits) - 1ull));
        const uint32_t w = world_bits ? (uint32_t)(key >> (64 - world_bits)) : 0u;
        const float4 v = world[(size_t)(w * entities_per_world + local) * 4 + (j & 3)];
        out[j] = v;
        if (inl) {
            res->list_world[j] = v;
            if ((j & 3) == 0) res->list_keys[j >> 2] = key;
        }
    }
    if (res && blockIdx.x == 0 && threadIdx.x == 0) res->list_n = inl ? (uint32_t)n : kNoInlineList;
}

}  // namespace astre

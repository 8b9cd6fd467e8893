// f3: one globally key-sorted draw list out of the sorted lists of G GPUs.
//
// Spec-defined (the reference has no multi-GPU path).  Rank r owns a contiguous range of the
// scene's entities, so "global entity order" is (rank, local slot) and the global list is the
// G lists merged under the order
//
//     (key >> slot_bits,  rank,  key & slot_mask)            -- class and depth first, then entity
//
// Every rank's list is already sorted by its own 64-bit keys, which is the same order restricted
// to one rank.  Output: the merged keys plus one source byte (the rank) per key.
//
// Shape: a multiway merge partitioned by samples, one pass over the keys.
//   samples         every 128th key of every list (k_publish_keys writes them next to the list for the peers).
//   k_merge_cuts    for ONE list t: the number of its keys that precede each sample of every list
//                   (binary search).  Run by the owner of list t on its own memory: in the peer-to-
//                   peer flow no search ever crosses NVLink, peers read the finished cut table.
//   k_merge_table   the G cuts of a sample sum to its position in the output, and
//                   sum(ceil(cut_t / 128)) is its index among all samples in merged order -- so the
//                   tile table is written without sorting the samples.
//   k_merge_tiles   tile q = everything between sample q and sample q+1: at most 128 keys of each
//                   list.  One CTA per batch of consecutive tiles (<= 2048 keys): the G sub-ranges are
//                   staged in shared memory and merged pairwise in ceil(log2 G) merge-path rounds,
//                   then stored contiguously with their sources.
// The list pointers may be peer memory (CUDA IPC mappings of the other ranks' publish buffers):
// then the tile loads ARE the NVLink transfer -- no gather buffer and no second pass, the exchange
// is fused into the merge, and with out_first/out_count a rank pulls only the keys of its slice.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/astre_gpu.h"

namespace astre {

constexpr int kMaxLists = 8;               // GPUs of one NVSwitch node
constexpr int kSampleShift = 7;            // every 128th key is a sample
constexpr int kSampleStep = 1 << kSampleShift;
constexpr int kMergeWarps = 4;             // warps per CTA of k_merge_tiles
constexpr int kMergeThreads = kMergeWarps * 32;

struct MergeDims {                                  // sizes of one merge
    uint32_t n[kMaxLists];                          // keys per list
    uint32_t sample_base[kMaxLists + 1];            // first sample index of every list; [n_lists] = number of samples
    uint32_t n_lists, pad;
    unsigned long long total;
};

struct MergeLists {
    const unsigned long long *keys[kMaxLists];      // the lists
    const unsigned long long *samples[kMaxLists];   // sample j of list t = keys[t][j * 128] (see sample_stride_shift)
    const uint32_t *cuts[kMaxLists];                // cuts[t][m] = keys of list t preceding sample m
    MergeDims d;                                    // sizes known to the host ...
    const MergeDims *dev_dims;                      // ... or, when not null, computed on the device (k_merge_dims)
    uint32_t n_lists, slot_bits;
    uint32_t sample_stride_shift, pad;              // samples[t][j << shift]: 0 = compact sample array, 7 = the list itself
};

__device__ __forceinline__ const MergeDims &dims_of(const MergeLists &L) { return L.dev_dims ? *L.dev_dims : L.d; }

// Sizes from the all-gathered per-view counts ([n_lists][n_views], as exchanged for 8e): no host round trip.
__global__ void k_merge_dims(const uint32_t *__restrict__ counts, uint32_t n_lists, uint32_t n_views,
                             unsigned long long max_keys, MergeDims *__restrict__ out)
{
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long total = 0;
    uint32_t n_samples = 0;
    for (uint32_t t = 0; t < n_lists; ++t) {
        unsigned long long c = 0;
        for (uint32_t v = 0; v < n_views; ++v) c += counts[(size_t)t * n_views + v];
        if (c > max_keys) c = max_keys;                              // a rank that overflowed published max_keys keys
        out->n[t] = (uint32_t)c;
        out->sample_base[t] = n_samples;
        n_samples += (uint32_t)((c + kSampleStep - 1) >> kSampleShift);
        total += c;
    }
    out->sample_base[n_lists] = n_samples;
    out->n_lists = n_lists;
    out->total = total;
}

// Number of keys of `a[0..n)` whose class (key >> sb) is below `thr`.
__device__ __forceinline__ uint32_t class_lower_bound(const unsigned long long *a, uint32_t n, uint32_t sb,
                                                      unsigned long long thr)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((a[mid] >> sb) < thr) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Cut table of list `t` (blockIdx.y selects it when n_t_lists > 1): out[m] for every sample m.
__global__ void __launch_bounds__(256)
k_merge_cuts(const __grid_constant__ MergeLists L, uint32_t t_first, uint32_t *__restrict__ out, size_t out_stride)
{
    const MergeDims &D = dims_of(L);
    const uint32_t t = t_first + blockIdx.y;
    const uint32_t n_samples = D.sample_base[L.n_lists];
    uint32_t *o = out + (size_t)blockIdx.y * out_stride;
    const unsigned long long *own = L.keys[t];
    const uint32_t n_own = D.n[t];
    for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < n_samples; m += gridDim.x * blockDim.x) {
        uint32_t s = 0;
        while (s + 1 < L.n_lists && m >= D.sample_base[s + 1]) ++s;
        const uint32_t j = m - D.sample_base[s];
        uint32_t c;
        if (s == t) c = j << kSampleShift;
        else c = class_lower_bound(own, n_own, L.slot_bits,
                                   (L.samples[s][(size_t)j << L.sample_stride_shift] >> L.slot_bits) + (t < s ? 1ull : 0ull));
        o[m] = c;
    }
}

// splits[q * kMaxLists + t] = keys of list t that precede sample q (q in merged order);
// tile_start[q] = position of sample q in the output.  Entry q = n_samples is the end sentinel.
__global__ void __launch_bounds__(256)
k_merge_table(const __grid_constant__ MergeLists L, uint32_t *__restrict__ splits, uint32_t *__restrict__ tile_start)
{
    cudaGridDependencySynchronize();                           // launched while the cut kernel drains
    const MergeDims &D = dims_of(L);
    const uint32_t n_samples = D.sample_base[L.n_lists];
    const uint32_t n_items = (((n_samples + 1u) << 3) + 31u) & ~31u;          // 8 lanes per sample; whole warps iterate
    for (uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x; gid < n_items; gid += gridDim.x * blockDim.x) {
        const uint32_t m = gid >> 3, t = gid & 7u;
        uint32_t c = 0;
        if (t < L.n_lists) {
            if (m < n_samples) c = L.cuts[t][m];
            else if (m == n_samples) c = D.n[t];
        }
        uint32_t pos = c, q = (c + kSampleStep - 1) >> kSampleShift;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            pos += __shfl_xor_sync(0xffffffffu, pos, d);
            q += __shfl_xor_sync(0xffffffffu, q, d);
        }
        if (m == n_samples) q = n_samples;
        if (m <= n_samples) {
            splits[(size_t)q * kMaxLists + t] = c;
            if (t == 0) tile_start[q] = pos;
        }
    }
}

// Number of entries of the ascending array a[0..n) that are below `target`; the whole warp searches
// (32 probes per round trip instead of one).
__device__ __forceinline__ uint32_t warp_lower_bound(const uint32_t *__restrict__ a, uint32_t n, uint32_t target, uint32_t lane)
{
    uint32_t lo = 0, hi = n;
    while (hi - lo > 32u) {
        const uint32_t step = (hi - lo + 31u) >> 5;
        const uint32_t idx = lo + lane * step;
        const bool below = idx < hi && a[idx] < target;
        const uint32_t cnt = __popc(__ballot_sync(0xffffffffu, below));       // probes are ascending: a prefix is below
        if (cnt == 0) { hi = lo; break; }
        hi = min(lo + cnt * step, hi);
        lo = lo + (cnt - 1u) * step + 1u;
    }
    const bool below = lo + lane < hi && a[lo + lane] < target;
    return lo + __popc(__ballot_sync(0xffffffffu, below));
}

// One CTA per batch of consecutive tiles: batch b = the tiles whose first output position lies in
// [b * span, (b+1) * span), span = kMergeBatchKeys - n_lists * 128, so a batch never exceeds
// kMergeBatchKeys keys.  The batch's G runs (one contiguous sub-range per list) are staged in shared
// memory (keys + source tags, ping-pong) and merged pairwise in ceil(log2 G) rounds; in a round every
// thread produces K consecutive outputs of one pair: a merge-path search for its starting split, then a
// branch-free sequential two-way merge.  Adjacent runs hold adjacent rank groups, so "left run first on
// equal class" keeps the (class, rank, slot) order.  Output positions are relative to out_first; batches
// outside [out_first, out_first + out_count) are not loaded.  The final stores are contiguous.
constexpr int kMergeBatchKeys = 2048;
// slice_world == 0: keep [out_first, out_first + out_count).  slice_world > 0: keep slice `slice_rank` of
// `slice_world` equal slices of the list (bounds computed here from the total, which may only be known on the
// device), at most out_count keys.  Grid-stride over the batches (the grid may be smaller than their number).
__global__ void __launch_bounds__(kMergeThreads)
k_merge_tiles(const __grid_constant__ MergeLists L, const uint32_t *__restrict__ splits,
              const uint32_t *__restrict__ tile_start, unsigned long long *__restrict__ out_keys,
              uint8_t *__restrict__ out_src, unsigned long long out_first, unsigned long long out_count,
              const uint32_t slice_rank, const uint32_t slice_world)
{
    __shared__ unsigned long long s_keys[2][kMergeBatchKeys];
    __shared__ uint8_t s_tag[2][kMergeBatchKeys];
    __shared__ uint32_t s_run[2][kMaxLists + 2], s_lo[kMaxLists], s_q[2];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t G = L.n_lists;
    const uint32_t span = kMergeBatchKeys - (G << kSampleShift);
    cudaGridDependencySynchronize();                           // launched while k_merge_table drains
    const MergeDims &D = dims_of(L);
    const uint32_t n_tiles = D.sample_base[G];
    const unsigned long long total_keys = D.total;
    if (slice_world) {
        const unsigned long long per = (total_keys + slice_world - 1) / slice_world;
        out_first = min(per * slice_rank, total_keys);
        out_count = min(min(per, total_keys - out_first), out_count);
    }
    const uint32_t n_batches = (uint32_t)((total_keys + span - 1) / span);
    for (uint32_t batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
        __syncthreads();                                       // shared tables of the previous batch are free
        // warps 0 and 1: first tile of this batch and of the next one
        if (warp < 2) {
            const unsigned long long target = (unsigned long long)(batch + warp) * span;
            const uint32_t q = target >= total_keys ? n_tiles : warp_lower_bound(tile_start, n_tiles, (uint32_t)target, lane);
            if (lane == 0) s_q[warp] = q;
        }
        __syncthreads();
        const uint32_t q0 = s_q[0], q1 = s_q[1];
        if (q0 >= q1) continue;
        const unsigned long long t0 = tile_start[q0], t1 = tile_start[q1];
        if (t1 <= out_first || t0 >= out_first + out_count) continue;
        if (tid < G) {
            const uint32_t lo = splits[(size_t)q0 * kMaxLists + tid];
            s_lo[tid] = lo;
            s_run[1][tid] = splits[(size_t)q1 * kMaxLists + tid] - lo;      // run lengths, turned into offsets below
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t off = 0;
            for (uint32_t t = 0; t < G; ++t) { s_run[0][t] = off; off += s_run[1][t]; }
            s_run[0][G] = off;
        }
        __syncthreads();
        const uint32_t total = s_run[0][G];
        // stage the runs: four independent loads in flight per thread (over NVLink one load costs microseconds)
        for (uint32_t e0 = tid; e0 < total; e0 += kMergeThreads * 4u) {
            unsigned long long v[4];
            uint32_t tt[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t e = e0 + (uint32_t)u * kMergeThreads;
                uint32_t t = 0;
                while (t + 1 < G && e >= s_run[0][t + 1]) ++t;
                tt[u] = t;
                v[u] = e < total ? L.keys[t][s_lo[t] + (e - s_run[0][t])] : 0ull;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t e = e0 + (uint32_t)u * kMergeThreads;
                if (e < total) { s_keys[0][e] = v[u]; s_tag[0][e] = (uint8_t)tt[u]; }
            }
        }
        __syncthreads();
        uint32_t cur = 0;
        for (uint32_t R = G; R > 1; R = (R + 1) >> 1) {
            const unsigned long long *src = s_keys[cur];
            unsigned long long *dst = s_keys[cur ^ 1u];
            const uint8_t *stag = s_tag[cur];
            uint8_t *dtag = s_tag[cur ^ 1u];
            const uint32_t *run = s_run[cur];
            // Pair p = runs (2p, 2p+1) (the last pair may be a single run).  The output span of every pair is
            // cut into chunks of K keys, one thread per chunk: K = ceil(total / (threads - P)) leaves room for
            // the P partial chunks, so no thread crosses a pair boundary and control flow stays uniform.
            const uint32_t P = (R + 1) >> 1;
            const uint32_t K = max((total + (kMergeThreads - 1u - P)) / (kMergeThreads - P), 1u);
            uint32_t a0 = 0, a1 = 0, b1 = 0, d = 0, next_chunk = 0;
            bool have = false;
            for (uint32_t pp = 0; pp < P; ++pp) {
                const uint32_t r0 = run[2 * pp], r1 = run[min(2 * pp + 1, R)], r2 = run[min(2 * pp + 2, R)];
                const uint32_t nch = (r2 - r0 + K - 1) / K;
                if (!have && tid < next_chunk + nch) { have = true; a0 = r0; a1 = r1; b1 = r2; d = (tid - next_chunk) * K; }
                next_chunk += nch;
            }
            if (have) {
                const uint32_t na = a1 - a0, nb = b1 - a1;
                uint32_t slo = d > nb ? d - nb : 0u, shi = min(d, na);        // merge path: keys of the left run among the first d
                while (slo < shi) {
                    const uint32_t mid = (slo + shi) >> 1;
                    if ((src[a0 + mid] >> ASTRE_KEY_SLOT_BITS) <= (src[a1 + d - 1 - mid] >> ASTRE_KEY_SLOT_BITS)) slo = mid + 1; else shi = mid;
                }
                uint32_t i = slo, j = d - slo, e = a0 + d;
                const uint32_t end = min(e + K, b1);
                unsigned long long ka = i < na ? src[a0 + i] : ~0ull, kb = j < nb ? src[a1 + j] : ~0ull;
                for (uint32_t k = 0; k < K; ++k, ++e) {                         // branch-free two-way merge
                    // left run (lower ranks) first on equal class; an exhausted run never wins (a real key may have
                    // all class bits set, so the ~0 filler alone does not decide)
                    const bool take = i < na && (!(j < nb) || (ka >> ASTRE_KEY_SLOT_BITS) <= (kb >> ASTRE_KEY_SLOT_BITS));
                    const uint32_t idx = take ? a0 + i : a1 + j;
                    if (e < end) { dst[e] = take ? ka : kb; dtag[e] = stag[idx]; }
                    i += take ? 1u : 0u; j += take ? 0u : 1u;
                    const bool more = take ? i < na : j < nb;
                    const unsigned long long nk = more ? src[take ? a0 + i : a1 + j] : ~0ull;
                    if (take) ka = nk; else kb = nk;
                }
            }
            // runs of the next round: run p = old runs 2p and 2p+1
            if (tid <= P) s_run[cur ^ 1u][tid] = run[min(2 * tid, R)];
            __syncthreads();
            cur ^= 1u;
        }
        for (uint32_t e = tid; e < total; e += kMergeThreads) {
            const unsigned long long at = t0 + e;
            if (at >= out_first && at - out_first < out_count) {
                out_keys[at - out_first] = s_keys[cur][e];
                out_src[at - out_first] = s_tag[cur][e];
            }
        }
    }
}

// ---- exchange without NCCL: flags and counts in the publish buffers, polled over NVLink ------------------
//
// The last 256 bytes of a publish buffer.  Peers read `counts` and poll the flags through their CUDA IPC
// mapping; every flag holds the epoch (frame number, from 1) up to which the statement is true.
struct PublishHeader {
    uint32_t counts[ASTRE_MAX_VIEWS];   // per-view visible counts of the published frame
    uint32_t flag[3];                   // [0] keys + samples + counts complete, [1] cut table complete,
                                        // [2] this rank has finished reading its peers' buffers
    uint32_t pad[29];
};
static_assert(sizeof(PublishHeader) == 256, "PublishHeader layout");

struct PeerHeaders { PublishHeader *h[kMaxLists]; uint32_t n, self; };

constexpr uint32_t kNoFlag = 0xFFu;
constexpr long long kPeerTimeoutCycles = 4000000000ll;       // ~2 s: a peer that never signals must not hang the GPU

// One warp.  (1) signal: own flag[sig] = sig_epoch, after everything this stream ran before.  (2) wait: lane r
// polls peer r's flag[wait] until it reaches wait_epoch.  (3) optionally the list sizes from the peers' counts.
// After a time-out the sizes in `void_dims` are zeroed, which turns the kernels that follow into no-ops.
__global__ void k_peer_sync(const PeerHeaders H, const uint32_t sig, const uint32_t sig_epoch, const uint32_t wait,
                            const uint32_t wait_epoch, MergeDims *__restrict__ dims, MergeDims *__restrict__ void_dims,
                            const uint32_t n_views, const unsigned long long max_keys, uint32_t *__restrict__ error)
{
    cudaGridDependencySynchronize();
    const uint32_t lane = threadIdx.x;
    if (sig != kNoFlag && lane == 0) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&H.h[H.self]->flag[sig]), "r"(sig_epoch) : "memory");
    }
    if (wait != kNoFlag && lane < H.n && lane != H.self) {
        const long long t0 = clock64();
        uint32_t v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(&H.h[lane]->flag[wait]) : "memory");
            if (v >= wait_epoch) break;
            if (clock64() - t0 > kPeerTimeoutCycles) { *error = 1u + lane; break; }
            __nanosleep(200);
        }
    }
    __syncwarp();
    if (*(volatile uint32_t *)error) {               // a peer never signalled: make the rest of the chain a no-op
        if (lane == 0 && void_dims) { void_dims->total = 0; void_dims->sample_base[H.n] = 0; }
        return;
    }
    if (dims && lane == 0) {
        unsigned long long total = 0;
        uint32_t n_samples = 0;
        for (uint32_t t = 0; t < H.n; ++t) {
            unsigned long long c = 0;
            for (uint32_t v = 0; v < n_views; ++v) c += H.h[t]->counts[v];
            if (c > max_keys) c = max_keys;
            dims->n[t] = (uint32_t)c;
            dims->sample_base[t] = n_samples;
            n_samples += (uint32_t)((c + kSampleStep - 1) >> kSampleShift);
            total += c;
        }
        dims->sample_base[H.n] = n_samples;
        dims->n_lists = H.n;
        dims->total = total;
    }
}

// Copies the frame's sorted list (whichever sort buffer it ended in) into a buffer at a fixed
// address, with its samples behind it: what a peer GPU maps.  Everything is read on the device;
// no host round trip.
__global__ void __launch_bounds__(256)
k_publish_keys(const unsigned long long *__restrict__ buf0, const unsigned long long *__restrict__ buf1,
               const uint32_t *__restrict__ result_buf, uint32_t sorted, const unsigned long long *__restrict__ key_total,
               unsigned long long cap, unsigned long long *__restrict__ dst, unsigned long long *__restrict__ samples,
               const uint32_t *__restrict__ counts, uint32_t n_views, PublishHeader *__restrict__ header)
{
    cudaGridDependencySynchronize();
    if (header && blockIdx.x == 0 && threadIdx.x < n_views) header->counts[threadIdx.x] = counts[threadIdx.x];
    const uint32_t which = sorted ? *result_buf : 0u;   // the key buffer the sort ended in (decided on the device)
    const unsigned long long *src = which ? buf1 : buf0;
    const unsigned long long n = min(*key_total, cap);
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long k = src[i];
        dst[i] = k;
        if ((i & (kSampleStep - 1)) == 0) samples[i >> kSampleShift] = k;
    }
}

}  // namespace astre

// Stand-alone timing of the key sort (kernels_sort.cuh) on synthetic draw keys.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I astre_b200/csrc tools/sortbench.cu -o tools/sortbench
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "kernels_sort.cuh"

using namespace astre;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_hist_all(const unsigned long long *keys, uint32_t n, SortPlan plan, uint32_t *block_hist)
{
    __shared__ uint32_t s[kHistWords];
    for (int i = threadIdx.x; i < kHistWords; i += blockDim.x) s[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (uint32_t q = 0; q < plan.n_passes; ++q) atomicAdd(&s[q * 256 + sort_digit(plan.seg[q], keys[i])], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < kHistWords; i += blockDim.x) block_hist[(size_t)blockIdx.x * kHistWords + i] = s[i];
}

template <int THREADS, int ITEMS>
float run(int n, int reps, unsigned long long *d0, unsigned long long *d1, const std::vector<unsigned long long> &h,
          SortControl *sc, FrameControl *ctl, uint32_t *bh, unsigned long long *lookback, const SortPlan &plan, bool check)
{
    const int regions = 296;
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sort_pass<THREADS, ITEMS>, THREADS, 0));
    const int grid = 148 * occ;      // co-resident CTAs only: static tile striding must never wait on an unscheduled CTA
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float total = 0;
    for (int r = 0; r < reps + 2; ++r) {
        CK(cudaMemcpy(d0, h.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
        FrameControl fc = {}; fc.key_total = n;
        CK(cudaMemcpy(ctl, &fc, sizeof fc, cudaMemcpyHostToDevice));
        uint32_t epoch = r + 1;
        CK(cudaMemcpy(&sc->epoch, &epoch, 4, cudaMemcpyHostToDevice));
        k_hist_all<<<regions, 256>>>(d0, n, plan, bh);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(a);
        k_sort_scan<<<plan.n_passes, 1024>>>(sc, bh, regions, ctl, (unsigned long long)n);
        for (int p = 0; p < (int)plan.n_passes; ++p) k_sort_pass<THREADS, ITEMS><<<grid, THREADS>>>(d0, d1, sc, lookback, plan, p);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (r >= 2) total += ms;
    }
    if (check) {
        std::vector<unsigned long long> out(n), ref(h);
        std::sort(ref.begin(), ref.end());
        uint32_t skip[kSortPasses]; CK(cudaMemcpy(skip, sc->skip, sizeof skip, cudaMemcpyDeviceToHost));
        int parity = 0; for (uint32_t p = 0; p < plan.n_passes; ++p) parity ^= skip[p] ? 0 : 1;
        CK(cudaMemcpy(out.data(), parity ? d1 : d0, (size_t)n * 8, cudaMemcpyDeviceToHost));
        printf("   sorted correctly: %s\n", out == ref ? "yes" : "NO");
    }
    return total / reps;
}

int main(int argc, char **argv)
{
    std::vector<int> sizes = {100000, 708104, 2835330, 16777216};
    if (argc > 1) { sizes.clear(); for (int i = 1; i < argc; ++i) sizes.push_back(atoi(argv[i])); }
    const int nmax = *std::max_element(sizes.begin(), sizes.end());
    unsigned long long *d0, *d1, *lookback; SortControl *sc; FrameControl *ctl; uint32_t *bh;
    CK(cudaMalloc(&d0, (size_t)nmax * 8)); CK(cudaMalloc(&d1, (size_t)nmax * 8));
    CK(cudaMalloc(&lookback, ((size_t)nmax / 2048 + 2) * 256 * 8)); CK(cudaMemset(lookback, 0, ((size_t)nmax / 2048 + 2) * 256 * 8));
    CK(cudaMalloc(&sc, sizeof(SortControl))); CK(cudaMemset(sc, 0, sizeof(SortControl)));
    CK(cudaMalloc(&ctl, sizeof(FrameControl))); CK(cudaMalloc(&bh, (size_t)296 * kHistWords * 4));
    // live bits of a 16M-entity, 8-mesh, 4-shader, 5-view frame: slot 24, depth 14, mesh 3, shader 2, view 3
    const unsigned long long live = ((1ull << 24) - 1) | (((1ull << 14) - 1) << 26) | (7ull << 40) | (3ull << 51) | (7ull << 59);
    const SortPlan plan = make_sort_plan(live);
    printf("live bits %d -> %u passes\n", __builtin_popcountll(live), plan.n_passes);
    std::mt19937_64 rng(1);
    for (int n : sizes) {
        std::vector<unsigned long long> h(n);
        for (int i = 0; i < n; ++i) {
            unsigned long long view = rng() % 5, k = (rng() & live & ~(7ull << 59)) | (view << 59);
            h[i] = (k & ~((1ull << 24) - 1)) | (unsigned long long)i % (1u << 24);   // unique slots
        }
        const float a = run<256, 8>(n, 10, d0, d1, h, sc, ctl, bh, lookback, plan, true);
        const float b = run<512, 8>(n, 10, d0, d1, h, sc, ctl, bh, lookback, plan, true);
        const float c = run<1024, 8>(n, 10, d0, d1, h, sc, ctl, bh, lookback, plan, true);
        const float d = run<1024, 16>(n, 10, d0, d1, h, sc, ctl, bh, lookback, plan, true);
        printf("n=%9d  256x8: %.1f us (%.1f/pass)  512x8: %.1f us (%.1f/pass)  1024x8: %.1f us (%.1f/pass)  1024x16: %.1f us (%.1f/pass)\n", n,
               a * 1e3, a * 1e3 / plan.n_passes, b * 1e3, b * 1e3 / plan.n_passes, c * 1e3, c * 1e3 / plan.n_passes, d * 1e3, d * 1e3 / plan.n_passes);
    }
    return 0;
}

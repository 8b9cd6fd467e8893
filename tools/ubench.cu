// Micro-benchmarks of the memory pattern of the fused kernel: 11 read streams of
// 4 B/entity and one write stream of 64 B/entity.  Build: nvcc -arch=sm_100a -O3 ubench.cu -o ubench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Ptrs { const float *p[11]; };

template <int MODE>   // 0: default ld/st, 1: ldcs/stcs, 2: default ld + stcs, 3: ldcs + default st
__global__ void __launch_bounds__(256) k_rw(Ptrs P, float4 *out, uint32_t n4)
{
    const int lane = threadIdx.x & 31;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w * 32 < n4; w += (gridDim.x * blockDim.x) >> 5) {
        const uint32_t g = w * 32 + lane;   // group of 4 entities
        float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < 11; ++i) {
            const float4 v = (MODE == 1 || MODE == 3) ? __ldcs(reinterpret_cast<const float4 *>(P.p[i]) + g)
                                                      : reinterpret_cast<const float4 *>(P.p[i])[g];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float4 *o = out + (size_t)w * 512 + lane;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float4 v = acc; v.x += i;
            if (MODE == 1 || MODE == 2) __stcs(o + i * 32, v); else o[i * 32] = v;
        }
    }
}

__global__ void __launch_bounds__(256) k_read(Ptrs P, float *sink, uint32_t n4)
{
    float s = 0;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += gridDim.x * blockDim.x) {
#pragma unroll
        for (int i = 0; i < 11; ++i) { const float4 v = reinterpret_cast<const float4 *>(P.p[i])[g]; s += v.x + v.y + v.z + v.w; }
    }
    if (s == 12345.678f) *sink = s;
}

__global__ void __launch_bounds__(256) k_write(float4 *out, uint32_t n4)
{
    const int lane = threadIdx.x & 31;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w * 32 < n4; w += (gridDim.x * blockDim.x) >> 5) {
        float4 *o = out + (size_t)w * 512 + lane;
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i * 32] = make_float4(w, i, lane, 1);
    }
}

__global__ void __launch_bounds__(256) k_copy(const float4 *in, float4 *out, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

template <class F> float timeit(F f, int reps = 20)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    CK(cudaGetLastError());
    return ms / reps;
}

int main(int argc, char **argv)
{
    const uint32_t n = argc > 1 ? atoi(argv[1]) : 16777216;
    const uint32_t n4 = n / 4;
    Ptrs P; float *planes; float4 *out; float *sink;
    CK(cudaMalloc(&planes, (size_t)n * 11 * 4)); CK(cudaMemset(planes, 0, (size_t)n * 11 * 4));
    for (int i = 0; i < 11; ++i) P.p[i] = planes + (size_t)i * n;
    CK(cudaMalloc(&out, (size_t)n * 64)); CK(cudaMalloc(&sink, 4));
    const double rw = (double)n * 108, rd = (double)n * 44, wr = (double)n * 64;
    for (int grid : {296, 592, 1184, 2368, 16384}) {
        printf("grid %5d: rw default %.0f GB/s | ldcs+stcs %.0f | ld+stcs %.0f | ldcs+st %.0f | read %.0f | write %.0f\n", grid,
               rw / timeit([&] { k_rw<0><<<grid, 256>>>(P, out, n4); }) / 1e6,
               rw / timeit([&] { k_rw<1><<<grid, 256>>>(P, out, n4); }) / 1e6,
               rw / timeit([&] { k_rw<2><<<grid, 256>>>(P, out, n4); }) / 1e6,
               rw / timeit([&] { k_rw<3><<<grid, 256>>>(P, out, n4); }) / 1e6,
               rd / timeit([&] { k_read<<<grid, 256>>>(P, sink, n4); }) / 1e6,
               wr / timeit([&] { k_write<<<grid, 256>>>(out, n4); }) / 1e6);
    }
    float4 *big; CK(cudaMalloc(&big, (size_t)1 << 30));
    printf("copy 1 GiB->1 GiB (2 GiB traffic): %.0f GB/s\n",
           2.0 * (1 << 30) / 2 * 2 / timeit([&] { k_copy<<<2368, 256>>>(big, big + (1 << 25), (size_t)1 << 25); }) / 1e6 / 2);
    float ms = timeit([&] { cudaMemcpyAsync(out, big, (size_t)n * 32, cudaMemcpyDeviceToDevice); });
    printf("cudaMemcpy D2D %.0f MB: %.0f GB/s (read+write)\n", n * 32 / 1e6, 2.0 * n * 32 / ms / 1e6);
    return 0;
}

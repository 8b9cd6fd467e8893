// Mirror maintenance and small per-entity kernels: ingest of host-layout
// component arrays into the device planes, dirty-set scatter, K4 basis vectors,
// f1 interpolation, gather of visible world matrices.
#pragma once
#include "kernels_transform.cuh"

namespace astre {

struct Planes {
    float *px, *py, *pz, *qx, *qy, *qz, *qw, *sx, *sy, *sz;
};

__device__ __forceinline__ uint32_t pack_meta(uint32_t mesh, uint32_t shader, uint32_t flags)
{
    return (mesh & 2047u) | ((shader & 255u) << 11) | ((flags & 31u) << 19);
}

// Packed host layout (pos xyz, rot xyzw, scale xyz per entity) -> planes.
// Null inputs take TransformSystem's defaults
// (engine/modules/ECS/src/system/transform_system.cpp:24-49).
// slots == nullptr: destination slot = first + i.
__global__ void k_ingest_trs(Planes pl, const uint32_t *__restrict__ slots, uint32_t first, uint32_t n,
                             const float *__restrict__ pos3, const float *__restrict__ rot4,
                             const float *__restrict__ scl3, bool fill_defaults)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t s = slots ? slots[i] : first + i;
        if (pos3) { pl.px[s] = pos3[3 * (size_t)i]; pl.py[s] = pos3[3 * (size_t)i + 1]; pl.pz[s] = pos3[3 * (size_t)i + 2]; }
        else if (fill_defaults) { pl.px[s] = 0.0f; pl.py[s] = 0.0f; pl.pz[s] = 0.0f; }
        if (rot4) {
            const float4 q = reinterpret_cast<const float4 *>(rot4)[i];
            pl.qx[s] = q.x; pl.qy[s] = q.y; pl.qz[s] = q.z; pl.qw[s] = q.w;
        } else if (fill_defaults) { pl.qx[s] = 0.0f; pl.qy[s] = 0.0f; pl.qz[s] = 0.0f; pl.qw[s] = 1.0f; }
        if (scl3) { pl.sx[s] = scl3[3 * (size_t)i]; pl.sy[s] = scl3[3 * (size_t)i + 1]; pl.sz[s] = scl3[3 * (size_t)i + 2]; }
        else if (fill_defaults) { pl.sx[s] = 1.0f; pl.sy[s] = 1.0f; pl.sz[s] = 1.0f; }
    }
}

__global__ void k_ingest_meta(uint32_t *__restrict__ meta, uint32_t *__restrict__ parent, uint32_t first, uint32_t n,
                              const uint32_t *__restrict__ mesh, const uint32_t *__restrict__ shader,
                              const uint8_t *__restrict__ flags, const uint32_t *__restrict__ parent_in)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t s = first + i;
        meta[s] = pack_meta(mesh ? mesh[i] : 0u, shader ? shader[i] : 0u, flags ? flags[i] : 0u);
        parent[s] = parent_in ? parent_in[i] : kNoParent;
    }
}

__global__ void k_update_flags(uint32_t *__restrict__ meta, const uint32_t *__restrict__ slots,
                               const uint8_t *__restrict__ flags, uint32_t n)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t s = slots[i];
        meta[s] = (meta[s] & ~(31u << 19)) | (((uint32_t)flags[i] & 31u) << 19);
    }
}

// K4: forward/up/right of every entity (transform_system.cpp:59-65), packed xyz.
__global__ void k_basis(const float *__restrict__ qx, const float *__restrict__ qy, const float *__restrict__ qz,
                        const float *__restrict__ qw, float *__restrict__ fwd3, float *__restrict__ up3,
                        float *__restrict__ right3, uint32_t n)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = qx[i], y = qy[i], z = qz[i], w = qw[i];
        const Vec3 bf = { 0.0f, 0.0f, -1.0f }, bu = { 0.0f, 1.0f, 0.0f };
        const Vec3 f = normalize3(quat_rotate(x, y, z, w, bf));
        const Vec3 u = normalize3(quat_rotate(x, y, z, w, bu));
        const Vec3 r = normalize3(cross3(f, u));
        fwd3[3 * (size_t)i] = f.x; fwd3[3 * (size_t)i + 1] = f.y; fwd3[3 * (size_t)i + 2] = f.z;
        up3[3 * (size_t)i] = u.x; up3[3 * (size_t)i + 1] = u.y; up3[3 * (size_t)i + 2] = u.z;
        right3[3 * (size_t)i] = r.x; right3[3 * (size_t)i + 1] = r.y; right3[3 * (size_t)i + 2] = r.z;
    }
}

// Gather world matrices in draw-list order: out[i] = world[slot(keys[first+i])].
__global__ void k_gather_world(const unsigned long long *__restrict__ keys, unsigned long long first, unsigned long long n,
                               const float4 *__restrict__ world, float4 *__restrict__ out,
                               uint32_t world_bits, uint32_t entities_per_world)
{
    const unsigned long long total = n * 4;
    for (unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; j < total;
         j += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[first + (j >> 2)];
        const uint32_t local_bits = 26u - world_bits;
        const uint32_t local = (uint32_t)(key & ((1ull << local_bits) - 1ull));
        const uint32_t w = world_bits ? (uint32_t)(key >> (64 - world_bits)) : 0u;
        const uint32_t slot = w * entities_per_world + local;
        out[j] = world[(size_t)slot * 4 + (j & 3)];
    }
}

}  // namespace astre

// Mirror maintenance and small per-entity kernels: ingest of host-layout
// component arrays into the device planes, dirty-set scatter, K4 basis vectors,
// f1 interpolation, gather of visible world matrices.
#pragma once
#include "kernels_transform.cuh"

namespace astre {

__device__ __forceinline__ uint32_t pack_meta(uint32_t mesh, uint32_t shader, uint32_t flags)
{
    return (mesh & 2047u) | ((shader & 255u) << 11) | ((flags & 31u) << 19);
}

// Packed host layout (pos xyz, rot xyzw, scale xyz per entity) -> tile records.
// Null inputs take TransformSystem's defaults
// (engine/modules/ECS/src/system/transform_system.cpp:24-49) when fill_defaults.
// slots == nullptr: destination slot = first + i.
__global__ void k_ingest_trs(float *__restrict__ tiles, const uint32_t *__restrict__ slots, uint32_t first, uint32_t n,
                             const float *__restrict__ pos3, const float *__restrict__ rot4,
                             const float *__restrict__ scl3, bool fill_defaults)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t s = slots ? slots[i] : first + i;
        float *t = tiles + tile_index(s, 0);
        if (pos3) {
            t[PX * kWarpTile] = pos3[3 * (size_t)i]; t[PY * kWarpTile] = pos3[3 * (size_t)i + 1]; t[PZ * kWarpTile] = pos3[3 * (size_t)i + 2];
        } else if (fill_defaults) { t[PX * kWarpTile] = 0.0f; t[PY * kWarpTile] = 0.0f; t[PZ * kWarpTile] = 0.0f; }
        if (rot4) {
            const float4 q = reinterpret_cast<const float4 *>(rot4)[i];
            t[QX * kWarpTile] = q.x; t[QY * kWarpTile] = q.y; t[QZ * kWarpTile] = q.z; t[QW * kWarpTile] = q.w;
        } else if (fill_defaults) { t[QX * kWarpTile] = 0.0f; t[QY * kWarpTile] = 0.0f; t[QZ * kWarpTile] = 0.0f; t[QW * kWarpTile] = 1.0f; }
        if (scl3) {
            t[SX * kWarpTile] = scl3[3 * (size_t)i]; t[SY * kWarpTile] = scl3[3 * (size_t)i + 1]; t[SZ * kWarpTile] = scl3[3 * (size_t)i + 2];
        } else if (fill_defaults) { t[SX * kWarpTile] = 1.0f; t[SY * kWarpTile] = 1.0f; t[SZ * kWarpTile] = 1.0f; }
    }
}

__global__ void k_ingest_meta(float *__restrict__ tiles, uint32_t first, uint32_t n,
                              const uint32_t *__restrict__ mesh, const uint32_t *__restrict__ shader,
                              const uint8_t *__restrict__ flags, const uint32_t *__restrict__ parent_in)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float *t = tiles + tile_index(first + i, 0);
        t[META * kWarpTile] = __uint_as_float(pack_meta(mesh ? mesh[i] : 0u, shader ? shader[i] : 0u, flags ? flags[i] : 0u));
        t[PARENT * kWarpTile] = __uint_as_float(parent_in ? parent_in[i] : kNoParent);
    }
}

__global__ void k_update_flags(float *__restrict__ tiles, const uint32_t *__restrict__ slots,
                               const uint8_t *__restrict__ flags, uint32_t n)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float *t = tiles + tile_index(slots[i], META);
        const uint32_t m = __float_as_uint(*t);
        *t = __uint_as_float((m & ~(31u << 19)) | (((uint32_t)flags[i] & 31u) << 19));
    }
}

// Marks every slot of the padded tile array as "no parent, not drawable".
__global__ void k_init_tiles(float *__restrict__ tiles, uint32_t n_slots)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += gridDim.x * blockDim.x) {
        float *t = tiles + tile_index(i, 0);
        t[QW * kWarpTile] = 1.0f; t[SX * kWarpTile] = 1.0f; t[SY * kWarpTile] = 1.0f; t[SZ * kWarpTile] = 1.0f;
        t[PARENT * kWarpTile] = __uint_as_float(kNoParent);
    }
}

// K4: forward/up/right of every entity (transform_system.cpp:59-65), packed xyz.
__global__ void k_basis(const float *__restrict__ tiles, float *__restrict__ fwd3, float *__restrict__ up3,
                        float *__restrict__ right3, uint32_t n)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float *t = tiles + tile_index(i, 0);
        const float x = t[QX * kWarpTile], y = t[QY * kWarpTile], z = t[QZ * kWarpTile], w = t[QW * kWarpTile];
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

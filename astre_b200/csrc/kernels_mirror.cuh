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

// f1: uModel = translate(mix(pa,pb)) * toMat4(slerp(qa,qb)) * scale(mix(sa,sb))  (Render/src/render.cpp:26-39).
// a = snapshot tiles, b = current tiles.  glm::slerp (ext/quaternion_common.inl): shortest arc, lerp when
// cos > 1 - epsilon.  Slots >= n_prev have no counterpart in `a` and get their plain transform.
__global__ void __launch_bounds__(256)
k_interpolate(const float *__restrict__ tiles_a, const float *__restrict__ tiles_b, float4 *__restrict__ world,
              uint32_t n, uint32_t n_prev, float alpha)
{
    __shared__ float4 s_scratch[256 * 4];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float *b = tiles_b + tile_index(i, 0);
        float p[3], q[4], sc[3];
        if (i < n_prev) {
            const float *a = tiles_a + tile_index(i, 0);
            p[0] = mixf(a[PX * kWarpTile], b[PX * kWarpTile], alpha);
            p[1] = mixf(a[PY * kWarpTile], b[PY * kWarpTile], alpha);
            p[2] = mixf(a[PZ * kWarpTile], b[PZ * kWarpTile], alpha);
            sc[0] = mixf(a[SX * kWarpTile], b[SX * kWarpTile], alpha);
            sc[1] = mixf(a[SY * kWarpTile], b[SY * kWarpTile], alpha);
            sc[2] = mixf(a[SZ * kWarpTile], b[SZ * kWarpTile], alpha);
            const float ax = a[QX * kWarpTile], ay = a[QY * kWarpTile], az = a[QZ * kWarpTile], aw = a[QW * kWarpTile];
            float zx = b[QX * kWarpTile], zy = b[QY * kWarpTile], zz = b[QZ * kWarpTile], zw = b[QW * kWarpTile];
            float c = fadd(fadd(fmul(aw, zw), fmul(ax, zx)), fadd(fmul(ay, zy), fmul(az, zz)));   // compute_dot<qua>
            if (c < 0.0f) { zx = -zx; zy = -zy; zz = -zz; zw = -zw; c = -c; }
            if (c > 1.0f - 1.1920928955078125e-7f) {
                q[3] = mixf(aw, zw, alpha); q[0] = mixf(ax, zx, alpha); q[1] = mixf(ay, zy, alpha); q[2] = mixf(az, zz, alpha);
            } else {
                const float ang = acosf(c);
                const float s0 = sinf(fmul(fsub(1.0f, alpha), ang)), s1 = sinf(fmul(alpha, ang)), sd = sinf(ang);
                q[3] = __fdiv_rn(fadd(fmul(s0, aw), fmul(s1, zw)), sd);
                q[0] = __fdiv_rn(fadd(fmul(s0, ax), fmul(s1, zx)), sd);
                q[1] = __fdiv_rn(fadd(fmul(s0, ay), fmul(s1, zy)), sd);
                q[2] = __fdiv_rn(fadd(fmul(s0, az), fmul(s1, zz)), sd);
            }
        } else {
            p[0] = b[PX * kWarpTile]; p[1] = b[PY * kWarpTile]; p[2] = b[PZ * kWarpTile];
            q[0] = b[QX * kWarpTile]; q[1] = b[QY * kWarpTile]; q[2] = b[QZ * kWarpTile]; q[3] = b[QW * kWarpTile];
            sc[0] = b[SX * kWarpTile]; sc[1] = b[SY * kWarpTile]; sc[2] = b[SZ * kWarpTile];
        }
        float m[16];
        trs_matrix(p[0], p[1], p[2], q[0], q[1], q[2], q[3], sc[0], sc[1], sc[2], m, s_scratch + threadIdx.x * 4);
        float4 *out = world + (size_t)i * 4;
        out[0] = make_float4(m[0], m[1], m[2], m[3]);
        out[1] = make_float4(m[4], m[5], m[6], m[7]);
        out[2] = make_float4(m[8], m[9], m[10], m[11]);
        out[3] = make_float4(m[12], m[13], m[14], m[15]);
    }
}

// f2: heads of the draw runs of the sorted list: key i starts a run when its class
// (key >> class_shift = world|view|shader|mesh) differs from key i-1's.  Unordered emission.
__global__ void k_run_heads(const unsigned long long *__restrict__ keys, uint32_t n, uint32_t class_shift,
                            uint2 *__restrict__ heads, uint32_t cap, uint32_t *__restrict__ counter)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t cls = (uint32_t)(keys[i] >> class_shift);
        if (i == 0 || cls != (uint32_t)(keys[i - 1] >> class_shift)) {
            const uint32_t at = atomicAdd(counter, 1u);
            if (at < cap) heads[at] = make_uint2(i, cls);
        }
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

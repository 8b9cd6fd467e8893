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
                // acos / sin evaluated in double and rounded once: the correctly rounded float in all but ~1e-8 of
                // the cases, i.e. at most 1 ULP from any libm's acosf / sinf (the reference's results depend on
                // the platform's libm here, glm::slerp calls std::acos / std::sin)
                const float ang = (float)acos((double)c);
                const float s0 = (float)sin((double)fmul(fsub(1.0f, alpha), ang)), s1 = (float)sin((double)fmul(alpha, ang)),
                            sd = (float)sin((double)ang);
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

// f2: heads of the draw runs of the sorted list, in list order.  Key i starts a run when its class
// (key >> class_shift = world | view | shader | mesh, up to 44 bits: carried as 64) differs from key i-1's.
// Three small launches: heads per CTA (each CTA owns a contiguous piece of the list), scan of those counts,
// ordered write -- no host-side sort.
struct RunHead { unsigned long long index, cls; };
constexpr int kRunThreads = 256;

__device__ __forceinline__ bool run_head_at(const unsigned long long *__restrict__ keys, uint32_t i, uint32_t class_shift)
{
    return i == 0 || (keys[i] >> class_shift) != (keys[i - 1] >> class_shift);
}

__global__ void __launch_bounds__(kRunThreads)
k_run_count(const unsigned long long *__restrict__ keys, uint32_t n, uint32_t class_shift, uint32_t per_cta,
            uint32_t *__restrict__ cta_counts)
{
    __shared__ uint32_t s_total;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    const uint32_t lo = blockIdx.x * per_cta, hi = min(lo + per_cta, n);
    uint32_t c = 0;
    for (uint32_t i = lo + threadIdx.x; i < hi; i += kRunThreads) c += run_head_at(keys, i, class_shift) ? 1u : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_total, c);
    __syncthreads();
    if (threadIdx.x == 0) cta_counts[blockIdx.x] = s_total;
}

// exclusive scan of cta_counts[0..n_ctas) in place; cta_counts[n_ctas] = total.  One block.
__global__ void __launch_bounds__(1024)
k_run_scan(uint32_t *__restrict__ cta_counts, uint32_t n_ctas)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_ctas; base += 1024u) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t c = i < n_ctas ? cta_counts[i] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t off = s_carry;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (i < n_ctas) cta_counts[i] = off + incl - c;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) cta_counts[n_ctas] = s_carry;
}

__global__ void __launch_bounds__(kRunThreads)
k_run_write(const unsigned long long *__restrict__ keys, uint32_t n, uint32_t class_shift, uint32_t per_cta,
            const uint32_t *__restrict__ cta_offsets, RunHead *__restrict__ heads, uint32_t cap)
{
    __shared__ uint32_t s_warp[kRunThreads / 32];
    __shared__ uint32_t s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lo = blockIdx.x * per_cta, hi = min(lo + per_cta, n);
    if (threadIdx.x == 0) s_base = cta_offsets[blockIdx.x];
    __syncthreads();
    for (uint32_t base = lo; base < hi; base += kRunThreads) {
        const uint32_t i = base + threadIdx.x;
        const bool head = i < hi && run_head_at(keys, i, class_shift);
        const uint32_t bal = __ballot_sync(0xffffffffu, head);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        uint32_t off = s_base, total = 0;
#pragma unroll
        for (int w = 0; w < kRunThreads / 32; ++w) { if (w < warp) off += s_warp[w]; total += s_warp[w]; }
        if (head) {
            const uint32_t at = off + __popc(bal & ((1u << lane) - 1u));
            if (at < cap) { heads[at].index = i; heads[at].cls = keys[i] >> class_shift; }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
}

// f2, one step further: the runs as indirect draw commands that never leave the GPU.  One record per run in the
// layout of OpenGL's DrawElementsIndirectCommand (glMultiDrawElementsIndirect): the mesh's index range, one instance
// per key of the run, baseInstance = the run's first position in the sorted list -- which is also the row of the
// entity's world matrix in the gathered list (ASTRE_FRAME_GATHER), so a vertex shader reads uModel[gl_BaseInstance +
// gl_InstanceID].  Runs whose mesh has no entry in the table become empty commands (count 0).
struct MeshDrawInfo { uint32_t index_count, first_index; int32_t base_vertex; };
struct DrawIndirect { uint32_t count, instance_count, first_index; int32_t base_vertex; uint32_t base_instance; };

__global__ void __launch_bounds__(256)
k_indirect_from_runs(const RunHead *__restrict__ heads, const uint32_t n_heads, const uint32_t n_keys,
                     const MeshDrawInfo *__restrict__ info, const uint32_t n_info, DrawIndirect *__restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_heads) return;
    const unsigned long long first = heads[i].index, next = i + 1u < n_heads ? heads[i + 1u].index : n_keys;
    const uint32_t mesh = (uint32_t)(heads[i].cls & 2047ull);
    DrawIndirect c{0u, (uint32_t)(next - first), 0u, 0, (uint32_t)first};
    if (mesh < n_info) { c.count = info[mesh].index_count; c.first_index = info[mesh].first_index; c.base_vertex = info[mesh].base_vertex; }
    out[i] = c;
}

// 8e: this rank's segment offsets in the global per-view draw lists, on the device.  all_counts is the
// all-gathered table [world_size][n_views]; the global list is ordered (view, rank, key), so
// offsets[v] = sum of every view before v over all ranks + sum of view v over the ranks before this one.
// One block (the table is tiny: n_views <= worlds x views of one shard).
// (`stride` = distance in words between the rows of two ranks; the plain table has stride n_views.)
__device__ __forceinline__ void global_offsets_block(const uint32_t *all_counts, const size_t stride, uint32_t world_size, uint32_t rank,
                                                     uint32_t n_views, uint32_t rank_major, unsigned long long *__restrict__ offsets,
                                                     unsigned long long *__restrict__ view_totals)
{
    __shared__ unsigned long long s_warp[8];
    __shared__ unsigned long long s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    if (rank_major) {
        // whole-world shards: everything of the earlier ranks comes first, then this rank's units in order
        unsigned long long mine = 0;
        for (size_t i = threadIdx.x; i < (size_t)rank * n_views; i += 256u) mine += all_counts[(i / n_views) * stride + i % n_views];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
        if (lane == 0) s_warp[warp] = mine;
        __syncthreads();
        if (threadIdx.x == 0) { unsigned long long t = 0; for (int w = 0; w < 8; ++w) t += s_warp[w]; s_carry = t; }
        __syncthreads();
    }
    for (uint32_t base = 0; base < n_views; base += 256u) {
        const uint32_t v = base + threadIdx.x;
        unsigned long long total = 0, before = 0;
        if (v < n_views) {
            if (rank_major) {
                total = all_counts[(size_t)rank * stride + v];
            } else {
                for (uint32_t r = 0; r < world_size; ++r) {
                    const unsigned long long c = all_counts[(size_t)r * stride + v];
                    if (r < rank) before += c;
                    total += c;
                }
            }
        }
        unsigned long long incl = total;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned long long off = s_carry;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (v < n_views) {
            if (offsets) offsets[v] = off + incl - total + before;
            if (view_totals) view_totals[v] = total;
        }
        __syncthreads();
        if (threadIdx.x == 255) s_carry = off + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
k_global_offsets(const uint32_t *__restrict__ all_counts, uint32_t world_size, uint32_t rank, uint32_t n_views,
                 uint32_t rank_major, unsigned long long *__restrict__ offsets, unsigned long long *__restrict__ view_totals)
{
    global_offsets_block(all_counts, n_views, world_size, rank, n_views, rank_major, offsets, view_totals);
}

// ---- count boards: the 8e exchange (all-gather of the counts + segment offsets) as ONE kernel over peer memory ----
//
// Every rank owns one CountBoard in its own HBM, mapped by its peers through CUDA IPC.  k_board_exchange, one block on
// a side stream, (1) STORES this rank's counts of a frame into the board of every rank over NVLink -- nobody ever
// reads remote memory --, (2) waits until the counts of all ranks have landed in its own board, (3) writes the table
// [n_ranks][n_counts], this rank's segment offsets and the per-view totals (k_global_offsets' arithmetic) and
// (4) acknowledges the frame to every rank.  Every count travels as one 8-byte word {count, exchange number}: an 8-byte
// store arrives whole, so a word whose upper half shows the expected number IS that exchange's count -- no flag, no
// fence and no ordering between the stores is needed (the low-latency protocol of collective libraries).  Exchanges
// are numbered per board on the device (`round`), exchange e uses slot e % kBoardSlots, which a rank rewrites only
// after every peer has acknowledged exchange e - kBoardSlots: ranks may be up to kBoardSlots exchanges apart.
constexpr uint32_t kBoardSlots = 4;
constexpr uint32_t kBoardRanks = 8;             // GPUs of one NVSwitch node
constexpr long long kBoardTimeoutCycles = 4000000000ll;      // ~2 s: a peer that never answers must not hang the GPU

struct CountBoard {
    uint32_t ack[kBoardRanks];                  // ack[t] = e: rank t has finished exchange e (stored by rank t)
    uint32_t round;                             // exchanges the owner has finished
    uint32_t error;                             // 1 + the rank that did not answer in time (sticky)
    uint32_t pad[64 - kBoardRanks - 2];
    // words[kBoardSlots][kBoardRanks][cap] follow: {count, exchange number} of rank t's count i in slot s
};
static_assert(sizeof(CountBoard) == 256, "CountBoard layout");

__host__ __device__ __forceinline__ unsigned long long *board_words(CountBoard *b, uint32_t slot, uint32_t rank, uint32_t cap)
{
    return reinterpret_cast<unsigned long long *>(b + 1) + ((size_t)slot * kBoardRanks + rank) * cap;
}
__host__ __device__ constexpr size_t board_bytes(uint32_t cap) { return sizeof(CountBoard) + (size_t)kBoardSlots * kBoardRanks * cap * 8; }

struct BoardPeers {
    CountBoard *b[kBoardRanks];                 // b[self] is the owner's board (a local pointer), the others are IPC mappings
    CountBoard *own;
    uint32_t n, self, cap;                      // n == 0: no board attached
};

// relaxed, system scope: never served from (or parked in) a non-coherent cache
__device__ __forceinline__ uint32_t ld_sys_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u32(uint32_t *p, uint32_t v) { asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_sys_u64(unsigned long long *p, unsigned long long v) { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

__global__ void __launch_bounds__(256)
k_board_exchange(const BoardPeers B, const uint32_t *__restrict__ counts, const uint32_t n_counts, const uint32_t rank_major,
                 uint32_t *__restrict__ table, unsigned long long *__restrict__ offsets, unsigned long long *__restrict__ view_totals)
{
    const uint32_t e = ld_sys_u32(&B.own->round) + 1u, slot = e % kBoardSlots, n = min(n_counts, B.cap);
    const long long t0 = clock64();
    // (1) the slot still holds exchange e - kBoardSlots on every rank that has not finished it (in steady state all have)
    if (e > kBoardSlots) {
        if (threadIdx.x < B.n)
            while (ld_sys_u32(&B.own->ack[threadIdx.x]) < e - kBoardSlots) {
                if (clock64() - t0 > kBoardTimeoutCycles) { B.own->error = 1u + threadIdx.x; break; }
                __nanosleep(100);
            }
        __syncthreads();
    }
#pragma unroll
    for (uint32_t t = 0; t < kBoardRanks; ++t)          // (static indices: the parameter stays in the constant bank)
        if (t < B.n) {
            unsigned long long *dst = board_words(B.b[t], slot, B.self, B.cap);
            for (uint32_t i = threadIdx.x; i < n; i += 256u) st_sys_u64(dst + i, (unsigned long long)e << 32 | counts[i]);
        }
    // (2) everybody's counts of exchange e, out of this rank's own memory
    for (uint32_t t = 0; t < B.n; ++t) {
        const unsigned long long *src = board_words(B.own, slot, t, B.cap);
        for (uint32_t i = threadIdx.x; i < n; i += 256u) {
            unsigned long long w;
            while ((uint32_t)((w = ld_sys_u64(src + i)) >> 32) != e) {
                if (clock64() - t0 > kBoardTimeoutCycles) { B.own->error = 1u + t; break; }
                __nanosleep(100);
            }
            table[(size_t)t * n + i] = (uint32_t)w;
        }
    }
    __threadfence_block();
    __syncthreads();
    // (3)
    global_offsets_block(table, n, B.n, B.self, n, rank_major, offsets, view_totals);
    __syncthreads();
    // (4)
    if (threadIdx.x == 0) {
        B.own->round = e;
        __threadfence_system();                 // the loads above are done before a peer may rewrite the slot
#pragma unroll
        for (uint32_t t = 0; t < kBoardRanks; ++t)
            if (t < B.n) st_sys_u32(&B.b[t]->ack[B.self], e);
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

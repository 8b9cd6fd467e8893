// Device arithmetic of the transform + visibility path (sm_100a).
//
// Everything that feeds a world matrix uses the explicitly rounded intrinsics
// (__fmul_rn / __fadd_rn / __fsub_rn): nvcc contracts a*b+c into FMA by default
// and the reference's builds do not (cmake/compile_options.cmake:56-70,118-139),
// so contraction would break bit parity with GLM's scalar code.  The culling
// arithmetic is spec-defined and uses __fmaf_rn exactly
// where the specification writes fmaf.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace astre {

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// glm::mat3_cast(q) (gtc/quaternion.inl), reached through math::toMat4 at
// engine/modules/ECS/src/system/transform_system.cpp:54.  r[c*3+row]; no normalisation.
__device__ __forceinline__ void quat_to_rot(float qx, float qy, float qz, float qw, float (&r)[9])
{
    const float qxx = fmul(qx, qx), qyy = fmul(qy, qy), qzz = fmul(qz, qz);
    const float qxz = fmul(qx, qz), qxy = fmul(qx, qy), qyz = fmul(qy, qz);
    const float qwx = fmul(qw, qx), qwy = fmul(qw, qy), qwz = fmul(qw, qz);
    r[0] = fsub(1.0f, fmul(2.0f, fadd(qyy, qzz)));
    r[1] = fmul(2.0f, fadd(qxy, qwz));
    r[2] = fmul(2.0f, fsub(qxz, qwy));
    r[3] = fmul(2.0f, fsub(qxy, qwz));
    r[4] = fsub(1.0f, fmul(2.0f, fadd(qxx, qzz)));
    r[5] = fmul(2.0f, fadd(qyz, qwx));
    r[6] = fmul(2.0f, fadd(qxz, qwy));
    r[7] = fmul(2.0f, fsub(qyz, qwx));
    r[8] = fsub(1.0f, fmul(2.0f, fadd(qxx, qyy)));
}

// glm::operator*(mat4, mat4): R[j] = A[0]*B[j].x + A[1]*B[j].y + A[2]*B[j].z + A[3]*B[j].w,
// left to right, every product and sum rounded separately.
__device__ __forceinline__ void mat4_mul(const float (&a)[16], const float (&b)[16], float (&out)[16])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float s = fadd(fmul(a[r], b[j * 4 + 0]), fmul(a[4 + r], b[j * 4 + 1]));
            s = fadd(s, fmul(a[8 + r], b[j * 4 + 2]));
            out[j * 4 + r] = fadd(s, fmul(a[12 + r], b[j * 4 + 3]));
        }
    }
}

// The op-for-op chain translate(I,p) * toMat4(q) * scale(I,s) of
// transform_system.cpp:51-57, written to `dst` (four float4 columns in shared
// memory).  Used only when the short form below cannot guarantee identical
// bits (a -0 involved, or a non-finite operand).  Out of line and fed by value
// so that the caller keeps its matrix in registers: the arrays here live on
// this function's own stack.
__device__ __noinline__ void trs_chain_store(float px, float py, float pz, float qx, float qy, float qz, float qw,
                                             float sx, float sy, float sz, float4 *dst)
{
    float rot[9];
    quat_to_rot(qx, qy, qz, qw, rot);
    float T[16], R[16], S[16], TR[16], m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { T[i] = 0.0f; R[i] = 0.0f; S[i] = 0.0f; }
    T[0] = T[5] = T[10] = 1.0f;
    // glm::translate: Result[3] = m[0]*v.x + m[1]*v.y + m[2]*v.z + m[3] with m = I
    {
        const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float s = fadd(fmul(I[r], px), fmul(I[4 + r], py));
            s = fadd(s, fmul(I[8 + r], pz));
            T[12 + r] = fadd(s, I[12 + r]);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) R[c * 4 + r] = rot[c * 3 + r];
    R[15] = 1.0f;
    // glm::scale(I, s): Result[i] = I[i]*s[i]
    S[0] = fmul(1.0f, sx); S[1] = fmul(0.0f, sx); S[2] = fmul(0.0f, sx); S[3] = fmul(0.0f, sx);
    S[4] = fmul(0.0f, sy); S[5] = fmul(1.0f, sy); S[6] = fmul(0.0f, sy); S[7] = fmul(0.0f, sy);
    S[8] = fmul(0.0f, sz); S[9] = fmul(0.0f, sz); S[10] = fmul(1.0f, sz); S[11] = fmul(0.0f, sz);
    S[15] = 1.0f;
    mat4_mul(T, R, TR);
    mat4_mul(TR, S, m);
    dst[0] = make_float4(m[0], m[1], m[2], m[3]);
    dst[1] = make_float4(m[4], m[5], m[6], m[7]);
    dst[2] = make_float4(m[8], m[9], m[10], m[11]);
    dst[3] = make_float4(m[12], m[13], m[14], m[15]);
}

// World matrix of one root entity.  For finite inputs every element of
// (T*R)*S is one product plus signed zeros (glm::scale(I,s) leaves 0*s_c, a zero
// carrying the sign of s_c, off the diagonal), so:
//   M[c][r] = R[c][r]*s_c          (c,r < 3; exact unless a -0 is involved)
//   M[c][3] = copysign(0, s_c)     M[3][r] = p_r + 0 (-0 -> +0)     M[3][3] = 1
// A -0 rotation element or product, or Inf/NaN anywhere, takes the full chain,
// where the other zero terms decide the sign.  `scratch` is a 64-byte shared
// memory slot private to the calling thread (the chain's result passes through it).
__device__ __forceinline__ void trs_matrix(float px, float py, float pz,
                                           float qx, float qy, float qz, float qw,
                                           float sx, float sy, float sz, float (&m)[16], float4 *scratch)
{
    float rot[9];
    quat_to_rot(qx, qy, qz, qw, rot);
    m[0] = fmul(rot[0], sx); m[1] = fmul(rot[1], sx); m[2] = fmul(rot[2], sx);
    m[4] = fmul(rot[3], sy); m[5] = fmul(rot[4], sy); m[6] = fmul(rot[5], sy);
    m[8] = fmul(rot[6], sz); m[9] = fmul(rot[7], sz); m[10] = fmul(rot[8], sz);
    m[3] = __int_as_float(__float_as_int(sx) & (int)0x80000000);
    m[7] = __int_as_float(__float_as_int(sy) & (int)0x80000000);
    m[11] = __int_as_float(__float_as_int(sz) & (int)0x80000000);
    m[12] = fadd(px, 0.0f); m[13] = fadd(py, 0.0f); m[14] = fadd(pz, 0.0f); m[15] = 1.0f;
    // 0*x is 0 for finite x and NaN otherwise; a product is non-finite whenever
    // its rotation element or scale is (a finite*finite overflow also lands
    // here, which is merely the slow path).
    float chk = fmul(0.0f, m[0]);
    chk = ffma(0.0f, m[1], chk); chk = ffma(0.0f, m[2], chk);
    chk = ffma(0.0f, m[4], chk); chk = ffma(0.0f, m[5], chk); chk = ffma(0.0f, m[6], chk);
    chk = ffma(0.0f, m[8], chk); chk = ffma(0.0f, m[9], chk); chk = ffma(0.0f, m[10], chk);
    chk = ffma(0.0f, px, chk); chk = ffma(0.0f, py, chk); chk = ffma(0.0f, pz, chk);
    // INT_MIN is the bit pattern of -0 and the unique minimum of all int32.
    int lo = min(min(__float_as_int(m[0]), __float_as_int(m[1])), __float_as_int(m[2]));
    lo = min(lo, min(min(__float_as_int(m[4]), __float_as_int(m[5])), __float_as_int(m[6])));
    lo = min(lo, min(min(__float_as_int(m[8]), __float_as_int(m[9])), __float_as_int(m[10])));
    lo = min(lo, min(min(__float_as_int(rot[0]), __float_as_int(rot[1])), __float_as_int(rot[2])));
    lo = min(lo, min(min(__float_as_int(rot[3]), __float_as_int(rot[4])), __float_as_int(rot[5])));
    lo = min(lo, min(min(__float_as_int(rot[6]), __float_as_int(rot[7])), __float_as_int(rot[8])));
    if (!(chk == 0.0f) || lo == (int)0x80000000) {
        trs_chain_store(px, py, pz, qx, qy, qz, qw, sx, sy, sz, scratch);
        const float4 c0 = scratch[0], c1 = scratch[1], c2 = scratch[2], c3 = scratch[3];
        m[0] = c0.x; m[1] = c0.y; m[2] = c0.z; m[3] = c0.w;
        m[4] = c1.x; m[5] = c1.y; m[6] = c1.z; m[7] = c1.w;
        m[8] = c2.x; m[9] = c2.y; m[10] = c2.z; m[11] = c2.w;
        m[12] = c3.x; m[13] = c3.y; m[14] = c3.z; m[15] = c3.w;
    }
}

// forward/up/right of transform_system.cpp:59-61: normalize(q*(0,0,-1)),
// normalize(q*(0,1,0)), normalize(cross(forward, up)).
struct Vec3 { float x, y, z; };

__device__ __forceinline__ Vec3 cross3(Vec3 a, Vec3 b)
{
    Vec3 r;
    r.x = fsub(fmul(a.y, b.z), fmul(b.y, a.z));
    r.y = fsub(fmul(a.z, b.x), fmul(b.z, a.x));
    r.z = fsub(fmul(a.x, b.y), fmul(b.x, a.y));
    return r;
}

__device__ __forceinline__ Vec3 normalize3(Vec3 v)
{
    const float d = fadd(fadd(fmul(v.x, v.x), fmul(v.y, v.y)), fmul(v.z, v.z));
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(d));
    Vec3 r = { fmul(v.x, inv), fmul(v.y, inv), fmul(v.z, inv) };
    return r;
}

// glm::operator*(qua, vec3): v + ((cross(u,v)*w) + cross(u,cross(u,v))) * 2
__device__ __forceinline__ Vec3 quat_rotate(float qx, float qy, float qz, float qw, Vec3 v)
{
    const Vec3 u = { qx, qy, qz };
    const Vec3 uv = cross3(u, v);
    const Vec3 uuv = cross3(u, uv);
    Vec3 r;
    r.x = fadd(v.x, fmul(fadd(fmul(uv.x, qw), uuv.x), 2.0f));
    r.y = fadd(v.y, fmul(fadd(fmul(uv.y, qw), uuv.y), 2.0f));
    r.z = fadd(v.z, fmul(fadd(fmul(uv.z, qw), uuv.z), 2.0f));
    return r;
}

// glm::mix(x, y, a) = x*(1-a) + y*a
__device__ __forceinline__ float mixf(float x, float y, float a)
{
    return fadd(fmul(x, fsub(1.0f, a)), fmul(y, a));
}

}  // namespace astre

// GLM-API stand-in for building the reference's own sources into oracle/_ref. TEST INFRASTRUCTURE ONLY.
//
// The reference's arithmetic lives in g-truc/glm, tag 1.0.1 (/root/reference/cmake/dependencies.cmake:249-255),
// fetched at configure time and absent from /root/reference and from this image. This header restates, from
// GLM's published scalar (non-SIMD, packed_highp) code paths, exactly the part of its API that the
// reference's Math / ECS / Render sources name, with the same operation order and rounding points, so that
// those sources compile unmodified and compute what they compute with the real library:
//   * configuration as the reference builds it (only GLM_ENABLE_EXPERIMENTAL is defined, Math/include/math/math.hpp:10):
//     right-handed, clip depth -1..1, column-major, quaternion ctor order (w,x,y,z), storage order x,y,z,w;
//   * glm::fma-free scalar paths: every product and sum is rounded on its own (build with -ffp-contract=off).
// Written independently of oracle/astre_oracle.c: agreement between the two is checked, not assumed
// (tests/test_ref_build.py).
#pragma once
#include <cmath>
#include <math.h>      // float overloads of acos etc. in the global namespace, as MSVC's <cmath> gives the reference
#include <cstddef>
#include <cstdint>
#include <limits>

namespace glm
{
    typedef int length_t;
    enum qualifier { packed_highp, packed_mediump, packed_lowp, highp = packed_highp, defaultp = packed_highp };

    template<length_t L, typename T, qualifier Q = defaultp> struct vec;
    template<length_t C, length_t R, typename T, qualifier Q = defaultp> struct mat;
    template<typename T, qualifier Q = defaultp> struct qua;

    // scalar functions GLM takes from the C++ library
    using std::sqrt; using std::cos; using std::sin; using std::tan; using std::acos; using std::asin; using std::abs;

    template<typename T> constexpr T epsilon() { return std::numeric_limits<T>::epsilon(); }
    template<typename T> constexpr T pi() { return static_cast<T>(3.14159265358979323846264338327950288); }
    template<typename T> constexpr T two_pi() { return static_cast<T>(6.28318530717958647692528676655900576); }

    template<typename T> constexpr T min(T a, T b) { return (b < a) ? b : a; }
    template<typename T> constexpr T max(T a, T b) { return (a < b) ? b : a; }
    template<typename T> constexpr T clamp(T x, T lo, T hi) { return min(max(x, lo), hi); }
    template<typename T> constexpr T radians(T deg) { return deg * static_cast<T>(0.01745329251994329576923690768489); }
    template<typename T> constexpr T degrees(T rad) { return rad * static_cast<T>(57.295779513082320876798154814105); }
    template<typename T> inline T inversesqrt(T x) { return static_cast<T>(1) / sqrt(x); }

    // ---------------------------------------------------------------- vec
    template<typename T, qualifier Q>
    struct vec<2, T, Q>
    {
        T x, y;
        constexpr vec() = default;
        constexpr explicit vec(T s) : x(s), y(s) {}
        constexpr vec(T a, T b) : x(a), y(b) {}
        template<typename U, qualifier P> constexpr vec(vec<3, U, P> const & v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
        template<typename U, qualifier P> constexpr vec(vec<4, U, P> const & v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
        static constexpr length_t length() { return 2; }
        constexpr T & operator[](length_t i) { return i == 0 ? x : y; }
        constexpr T const & operator[](length_t i) const { return i == 0 ? x : y; }
    };

    template<typename T, qualifier Q>
    struct vec<3, T, Q>
    {
        T x, y, z;
        constexpr vec() = default;
        constexpr explicit vec(T s) : x(s), y(s), z(s) {}
        constexpr vec(T a, T b, T c) : x(a), y(b), z(c) {}
        template<typename U, qualifier P> constexpr vec(vec<2, U, P> const & v, T c) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(c) {}
        // GLM_EXPLICIT is empty unless GLM_FORCE_EXPLICIT_CTOR is defined: vec4 -> vec3 truncates implicitly
        template<typename U, qualifier P> constexpr vec(vec<4, U, P> const & v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}
        static constexpr length_t length() { return 3; }
        constexpr T & operator[](length_t i) { switch (i) { default: case 0: return x; case 1: return y; case 2: return z; } }
        constexpr T const & operator[](length_t i) const { switch (i) { default: case 0: return x; case 1: return y; case 2: return z; } }
    };

    template<typename T, qualifier Q>
    struct vec<4, T, Q>
    {
        T x, y, z, w;
        constexpr vec() = default;
        constexpr explicit vec(T s) : x(s), y(s), z(s), w(s) {}
        constexpr vec(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
        template<typename U, qualifier P> constexpr vec(vec<3, U, P> const & v, T d) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)), w(d) {}
        static constexpr length_t length() { return 4; }
        constexpr T & operator[](length_t i) { switch (i) { default: case 0: return x; case 1: return y; case 2: return z; case 3: return w; } }
        constexpr T const & operator[](length_t i) const { switch (i) { default: case 0: return x; case 1: return y; case 2: return z; case 3: return w; } }
    };

    // component-wise arithmetic, one rounding per operation
#define REFGLM_VEC_OPS(L, ...)                                                                                          \
    template<typename T, qualifier Q> constexpr vec<L, T, Q> operator+(vec<L, T, Q> const & a, vec<L, T, Q> const & b) { return vec<L, T, Q>(__VA_ARGS__(+)); } \
    template<typename T, qualifier Q> constexpr vec<L, T, Q> operator-(vec<L, T, Q> const & a, vec<L, T, Q> const & b) { return vec<L, T, Q>(__VA_ARGS__(-)); } \
    template<typename T, qualifier Q> constexpr vec<L, T, Q> operator*(vec<L, T, Q> const & a, vec<L, T, Q> const & b) { return vec<L, T, Q>(__VA_ARGS__(*)); } \
    template<typename T, qualifier Q> constexpr vec<L, T, Q> operator/(vec<L, T, Q> const & a, vec<L, T, Q> const & b) { return vec<L, T, Q>(__VA_ARGS__(/)); }
#define REFGLM_V2(op) a.x op b.x, a.y op b.y
#define REFGLM_V3(op) a.x op b.x, a.y op b.y, a.z op b.z
#define REFGLM_V4(op) a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w
    REFGLM_VEC_OPS(2, REFGLM_V2)
    REFGLM_VEC_OPS(3, REFGLM_V3)
    REFGLM_VEC_OPS(4, REFGLM_V4)
#undef REFGLM_VEC_OPS
#undef REFGLM_V2
#undef REFGLM_V3
#undef REFGLM_V4

    template<typename T, qualifier Q> constexpr vec<2, T, Q> operator*(vec<2, T, Q> const & v, T s) { return vec<2, T, Q>(v.x * s, v.y * s); }
    template<typename T, qualifier Q> constexpr vec<3, T, Q> operator*(vec<3, T, Q> const & v, T s) { return vec<3, T, Q>(v.x * s, v.y * s, v.z * s); }
    template<typename T, qualifier Q> constexpr vec<4, T, Q> operator*(vec<4, T, Q> const & v, T s) { return vec<4, T, Q>(v.x * s, v.y * s, v.z * s, v.w * s); }
    template<typename T, qualifier Q> constexpr vec<2, T, Q> operator*(T s, vec<2, T, Q> const & v) { return vec<2, T, Q>(s * v.x, s * v.y); }
    template<typename T, qualifier Q> constexpr vec<3, T, Q> operator*(T s, vec<3, T, Q> const & v) { return vec<3, T, Q>(s * v.x, s * v.y, s * v.z); }
    template<typename T, qualifier Q> constexpr vec<4, T, Q> operator*(T s, vec<4, T, Q> const & v) { return vec<4, T, Q>(s * v.x, s * v.y, s * v.z, s * v.w); }
    template<typename T, qualifier Q> constexpr vec<3, T, Q> operator/(vec<3, T, Q> const & v, T s) { return vec<3, T, Q>(v.x / s, v.y / s, v.z / s); }
    template<typename T, qualifier Q> constexpr vec<4, T, Q> operator/(vec<4, T, Q> const & v, T s) { return vec<4, T, Q>(v.x / s, v.y / s, v.z / s, v.w / s); }
    template<typename T, qualifier Q> constexpr vec<2, T, Q> operator-(vec<2, T, Q> const & v) { return vec<2, T, Q>(-v.x, -v.y); }
    template<typename T, qualifier Q> constexpr vec<3, T, Q> operator-(vec<3, T, Q> const & v) { return vec<3, T, Q>(-v.x, -v.y, -v.z); }
    template<typename T, qualifier Q> constexpr vec<4, T, Q> operator-(vec<4, T, Q> const & v) { return vec<4, T, Q>(-v.x, -v.y, -v.z, -v.w); }

    // geometric functions (detail/func_geometric.inl)
    template<typename T, qualifier Q> constexpr T dot(vec<2, T, Q> const & a, vec<2, T, Q> const & b) { vec<2, T, Q> t(a * b); return t.x + t.y; }
    template<typename T, qualifier Q> constexpr T dot(vec<3, T, Q> const & a, vec<3, T, Q> const & b) { vec<3, T, Q> t(a * b); return t.x + t.y + t.z; }
    template<typename T, qualifier Q> constexpr T dot(vec<4, T, Q> const & a, vec<4, T, Q> const & b) { vec<4, T, Q> t(a * b); return (t.x + t.y) + (t.z + t.w); }

    template<typename T, qualifier Q>
    constexpr vec<3, T, Q> cross(vec<3, T, Q> const & x, vec<3, T, Q> const & y)
    {
        return vec<3, T, Q>(x.y * y.z - y.y * x.z,
                            x.z * y.x - y.z * x.x,
                            x.x * y.y - y.x * x.y);
    }
    template<typename T, qualifier Q> constexpr T cross(vec<2, T, Q> const & x, vec<2, T, Q> const & y) { return x.x * y.y - y.x * x.y; }   // gtx/exterior_product

    template<length_t L, typename T, qualifier Q> inline T length(vec<L, T, Q> const & v) { return sqrt(dot(v, v)); }
    template<length_t L, typename T, qualifier Q> constexpr T length2(vec<L, T, Q> const & v) { return dot(v, v); }
    template<typename T> constexpr T length2(T x) { return x * x; }
    template<length_t L, typename T, qualifier Q> inline vec<L, T, Q> normalize(vec<L, T, Q> const & v) { return v * inversesqrt(dot(v, v)); }
    template<length_t L, typename T, qualifier Q> inline vec<L, T, Q> sqrt(vec<L, T, Q> const & v) { vec<L, T, Q> r(v); for (length_t i = 0; i < L; ++i) r[i] = sqrt(v[i]); return r; }

    // mix (detail/func_common.inl, compute_mix_scalar / compute_mix_vector): x*(1-a) + y*a
    template<typename T, typename U> constexpr T mix(T x, T y, U a)
    {
        return static_cast<T>(static_cast<U>(x) * (static_cast<U>(1) - a) + static_cast<U>(y) * a);
    }
    template<length_t L, typename T, typename U, qualifier Q>
    constexpr vec<L, T, Q> mix(vec<L, T, Q> const & x, vec<L, T, Q> const & y, U a)
    {
        return vec<L, T, Q>(vec<L, U, Q>(x) * (static_cast<U>(1) - a) + vec<L, U, Q>(y) * a);
    }
    template<length_t L, typename T, typename U, qualifier Q>
    constexpr vec<L, T, Q> mix(vec<L, T, Q> const & x, vec<L, T, Q> const & y, vec<L, U, Q> const & a)
    {
        return vec<L, T, Q>(vec<L, U, Q>(x) * (vec<L, U, Q>(static_cast<U>(1)) - a) + vec<L, U, Q>(y) * a);
    }

    // ---------------------------------------------------------------- mat (column-major: m[col][row])
    template<typename T, qualifier Q>
    struct mat<2, 2, T, Q>
    {
        typedef vec<2, T, Q> col_type;
        col_type value[2];
        constexpr mat() = default;
        constexpr explicit mat(T s) : value{col_type(s, 0), col_type(0, s)} {}
        constexpr col_type & operator[](length_t i) { return value[i]; }
        constexpr col_type const & operator[](length_t i) const { return value[i]; }
    };

    template<typename T, qualifier Q>
    struct mat<3, 3, T, Q>
    {
        typedef vec<3, T, Q> col_type;
        col_type value[3];
        constexpr mat() = default;
        constexpr explicit mat(T s) : value{col_type(s, 0, 0), col_type(0, s, 0), col_type(0, 0, s)} {}
        constexpr col_type & operator[](length_t i) { return value[i]; }
        constexpr col_type const & operator[](length_t i) const { return value[i]; }
    };

    template<typename T, qualifier Q>
    struct mat<4, 4, T, Q>
    {
        typedef vec<4, T, Q> col_type;
        col_type value[4];
        constexpr mat() = default;
        constexpr explicit mat(T s) : value{col_type(s, 0, 0, 0), col_type(0, s, 0, 0), col_type(0, 0, s, 0), col_type(0, 0, 0, s)} {}
        constexpr mat(col_type const & a, col_type const & b, col_type const & c, col_type const & d) : value{a, b, c, d} {}
        // mat3 -> mat4: the upper-left block, identity elsewhere (detail/type_mat4x4.inl)
        constexpr explicit mat(mat<3, 3, T, Q> const & m)
            : value{col_type(m[0], 0), col_type(m[1], 0), col_type(m[2], 0), col_type(0, 0, 0, 1)} {}
        constexpr col_type & operator[](length_t i) { return value[i]; }
        constexpr col_type const & operator[](length_t i) const { return value[i]; }
    };

    // mat4 * mat4 (detail/type_mat4x4.inl): column j = A0*B[j].x + A1*B[j].y + A2*B[j].z + A3*B[j].w,
    // evaluated left to right; 1.0.x writes the same chain with glm::fma, which is a*b+c on the scalar path.
    template<typename T, qualifier Q>
    constexpr mat<4, 4, T, Q> operator*(mat<4, 4, T, Q> const & m1, mat<4, 4, T, Q> const & m2)
    {
        typename mat<4, 4, T, Q>::col_type const A0 = m1[0], A1 = m1[1], A2 = m1[2], A3 = m1[3];
        typename mat<4, 4, T, Q>::col_type const B0 = m2[0], B1 = m2[1], B2 = m2[2], B3 = m2[3];
        mat<4, 4, T, Q> R;
        R[0] = A0 * B0[0] + A1 * B0[1] + A2 * B0[2] + A3 * B0[3];
        R[1] = A0 * B1[0] + A1 * B1[1] + A2 * B1[2] + A3 * B1[3];
        R[2] = A0 * B2[0] + A1 * B2[1] + A2 * B2[2] + A3 * B2[3];
        R[3] = A0 * B3[0] + A1 * B3[1] + A2 * B3[2] + A3 * B3[3];
        return R;
    }

    template<typename T, qualifier Q>
    constexpr vec<4, T, Q> operator*(mat<4, 4, T, Q> const & m, vec<4, T, Q> const & v)
    {
        vec<4, T, Q> const Mov0(v[0]), Mov1(v[1]);
        vec<4, T, Q> const Mul0 = m[0] * Mov0, Mul1 = m[1] * Mov1;
        vec<4, T, Q> const Add0 = Mul0 + Mul1;
        vec<4, T, Q> const Mov2(v[2]), Mov3(v[3]);
        vec<4, T, Q> const Mul2 = m[2] * Mov2, Mul3 = m[3] * Mov3;
        vec<4, T, Q> const Add1 = Mul2 + Mul3;
        return Add0 + Add1;
    }

    // ext/matrix_transform.inl
    template<typename T, qualifier Q>
    constexpr mat<4, 4, T, Q> translate(mat<4, 4, T, Q> const & m, vec<3, T, Q> const & v)
    {
        mat<4, 4, T, Q> R(m);
        R[3] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3];
        return R;
    }

    template<typename T, qualifier Q>
    constexpr mat<4, 4, T, Q> scale(mat<4, 4, T, Q> const & m, vec<3, T, Q> const & v)
    {
        mat<4, 4, T, Q> R;
        R[0] = m[0] * v[0];
        R[1] = m[1] * v[1];
        R[2] = m[2] * v[2];
        R[3] = m[3];
        return R;
    }

    // lookAt = lookAtRH with the default (right-handed) configuration
    template<typename T, qualifier Q>
    inline mat<4, 4, T, Q> lookAt(vec<3, T, Q> const & eye, vec<3, T, Q> const & center, vec<3, T, Q> const & up)
    {
        vec<3, T, Q> const f(normalize(center - eye));
        vec<3, T, Q> const s(normalize(cross(f, up)));
        vec<3, T, Q> const u(cross(s, f));
        mat<4, 4, T, Q> R(static_cast<T>(1));
        R[0][0] = s.x; R[1][0] = s.y; R[2][0] = s.z;
        R[0][1] = u.x; R[1][1] = u.y; R[2][1] = u.z;
        R[0][2] = -f.x; R[1][2] = -f.y; R[2][2] = -f.z;
        R[3][0] = -dot(s, eye);
        R[3][1] = -dot(u, eye);
        R[3][2] = dot(f, eye);
        return R;
    }

    // ext/matrix_clip_space.inl: perspective = perspectiveRH_NO, ortho = orthoRH_NO (clip depth -1..1)
    template<typename T>
    inline mat<4, 4, T, defaultp> perspective(T fovy, T aspect, T zNear, T zFar)
    {
        T const tanHalfFovy = tan(fovy / static_cast<T>(2));
        mat<4, 4, T, defaultp> R(static_cast<T>(0));
        R[0][0] = static_cast<T>(1) / (aspect * tanHalfFovy);
        R[1][1] = static_cast<T>(1) / (tanHalfFovy);
        R[2][2] = -(zFar + zNear) / (zFar - zNear);
        R[2][3] = -static_cast<T>(1);
        R[3][2] = -(static_cast<T>(2) * zFar * zNear) / (zFar - zNear);
        return R;
    }

    template<typename T>
    constexpr mat<4, 4, T, defaultp> ortho(T left, T right, T bottom, T top, T zNear, T zFar)
    {
        mat<4, 4, T, defaultp> R(static_cast<T>(1));
        R[0][0] = static_cast<T>(2) / (right - left);
        R[1][1] = static_cast<T>(2) / (top - bottom);
        R[2][2] = -static_cast<T>(2) / (zFar - zNear);
        R[3][0] = -(right + left) / (right - left);
        R[3][1] = -(top + bottom) / (top - bottom);
        R[3][2] = -(zFar + zNear) / (zFar - zNear);
        return R;
    }

    // ---------------------------------------------------------------- qua (storage x,y,z,w; ctor w,x,y,z)
    template<typename T, qualifier Q>
    struct qua
    {
        T x, y, z, w;
        constexpr qua() = default;
        constexpr qua(T w_, T x_, T y_, T z_) : x(x_), y(y_), z(z_), w(w_) {}
        static constexpr qua wxyz(T w_, T x_, T y_, T z_) { return qua(w_, x_, y_, z_); }
        static constexpr length_t length() { return 4; }
        constexpr T & operator[](length_t i) { return (&x)[i]; }
        constexpr T const & operator[](length_t i) const { return (&x)[i]; }
    };

    template<typename T, qualifier Q> constexpr qua<T, Q> operator-(qua<T, Q> const & q) { return qua<T, Q>::wxyz(-q.w, -q.x, -q.y, -q.z); }
    template<typename T, qualifier Q> constexpr qua<T, Q> operator+(qua<T, Q> const & p, qua<T, Q> const & q) { return qua<T, Q>::wxyz(p.w + q.w, p.x + q.x, p.y + q.y, p.z + q.z); }
    template<typename T, qualifier Q> constexpr qua<T, Q> operator*(qua<T, Q> const & q, T const & s) { return qua<T, Q>::wxyz(q.w * s, q.x * s, q.y * s, q.z * s); }
    template<typename T, qualifier Q> constexpr qua<T, Q> operator*(T const & s, qua<T, Q> const & q) { return q * s; }
    template<typename T, qualifier Q> constexpr qua<T, Q> operator/(qua<T, Q> const & q, T const & s) { return qua<T, Q>::wxyz(q.w / s, q.x / s, q.y / s, q.z / s); }

    // quaternion dot (detail/type_quat.inl, compute_dot<qua>): (w*w' + x*x') + (y*y' + z*z')
    template<typename T, qualifier Q>
    constexpr T dot(qua<T, Q> const & a, qua<T, Q> const & b)
    {
        vec<4, T, Q> t(a.w * b.w, a.x * b.x, a.y * b.y, a.z * b.z);
        return (t.x + t.y) + (t.z + t.w);
    }
    template<typename T, qualifier Q> inline T length(qua<T, Q> const & q) { return sqrt(dot(q, q)); }
    template<typename T, qualifier Q>
    inline qua<T, Q> normalize(qua<T, Q> const & q)
    {
        T len = length(q);
        if (len <= static_cast<T>(0)) return qua<T, Q>::wxyz(static_cast<T>(1), static_cast<T>(0), static_cast<T>(0), static_cast<T>(0));
        T oneOverLen = static_cast<T>(1) / len;
        return qua<T, Q>::wxyz(q.w * oneOverLen, q.x * oneOverLen, q.y * oneOverLen, q.z * oneOverLen);
    }

    // rotate a vector (detail/type_quat.inl)
    template<typename T, qualifier Q>
    constexpr vec<3, T, Q> operator*(qua<T, Q> const & q, vec<3, T, Q> const & v)
    {
        vec<3, T, Q> const QuatVector(q.x, q.y, q.z);
        vec<3, T, Q> const uv(cross(QuatVector, v));
        vec<3, T, Q> const uuv(cross(QuatVector, uv));
        return v + ((uv * q.w) + uuv) * static_cast<T>(2);
    }

    // gtc/quaternion.inl: no normalisation of q
    template<typename T, qualifier Q>
    constexpr mat<3, 3, T, Q> mat3_cast(qua<T, Q> const & q)
    {
        mat<3, 3, T, Q> R(T(1));
        T qxx(q.x * q.x);
        T qyy(q.y * q.y);
        T qzz(q.z * q.z);
        T qxz(q.x * q.z);
        T qxy(q.x * q.y);
        T qyz(q.y * q.z);
        T qwx(q.w * q.x);
        T qwy(q.w * q.y);
        T qwz(q.w * q.z);

        R[0][0] = T(1) - T(2) * (qyy + qzz);
        R[0][1] = T(2) * (qxy + qwz);
        R[0][2] = T(2) * (qxz - qwy);

        R[1][0] = T(2) * (qxy - qwz);
        R[1][1] = T(1) - T(2) * (qxx + qzz);
        R[1][2] = T(2) * (qyz + qwx);

        R[2][0] = T(2) * (qxz + qwy);
        R[2][1] = T(2) * (qyz - qwx);
        R[2][2] = T(1) - T(2) * (qxx + qyy);
        return R;
    }
    template<typename T, qualifier Q> constexpr mat<4, 4, T, Q> mat4_cast(qua<T, Q> const & q) { return mat<4, 4, T, Q>(mat3_cast(q)); }
    template<typename T, qualifier Q> constexpr mat<3, 3, T, Q> toMat3(qua<T, Q> const & q) { return mat3_cast(q); }   // gtx/quaternion
    template<typename T, qualifier Q> constexpr mat<4, 4, T, Q> toMat4(qua<T, Q> const & q) { return mat4_cast(q); }   // gtx/quaternion

    // ext/quaternion_common.inl
    template<typename T, qualifier Q>
    constexpr qua<T, Q> lerp(qua<T, Q> const & x, qua<T, Q> const & y, T a) { return x * (static_cast<T>(1) - a) + (y * a); }

    template<typename T, qualifier Q>
    inline qua<T, Q> slerp(qua<T, Q> const & x, qua<T, Q> const & y, T a)
    {
        qua<T, Q> z = y;
        T cosTheta = dot(x, y);
        // the short way round: negate one end when the ends are more than 90 degrees apart
        if (cosTheta < static_cast<T>(0))
        {
            z = -y;
            cosTheta = -cosTheta;
        }
        if (cosTheta > static_cast<T>(1) - epsilon<T>())
        {
            // nearly parallel: component-wise linear interpolation
            return qua<T, Q>::wxyz(mix(x.w, z.w, a), mix(x.x, z.x, a), mix(x.y, z.y, a), mix(x.z, z.z, a));
        }
        else
        {
            T angle = acos(cosTheta);
            return (sin((static_cast<T>(1) - a) * angle) * x + sin(a * angle) * z) / sin(angle);
        }
    }

    // Named by math.hpp's forwarders but never instantiated on this path: declarations only.
    template<typename T, qualifier Q> qua<T, Q> mix(qua<T, Q> const & x, qua<T, Q> const & y, T a);
    template<typename T, qualifier Q> vec<3, T, Q> eulerAngles(qua<T, Q> const & x);
    template<typename T, qualifier Q> qua<T, Q> rotation(vec<3, T, Q> const & orig, vec<3, T, Q> const & dest);
    template<typename T, qualifier Q> qua<T, Q> cross(qua<T, Q> const & q1, qua<T, Q> const & q2);
    template<typename T, qualifier Q> vec<3, T, Q> cross(vec<3, T, Q> const & v, qua<T, Q> const & q);
    template<typename T, qualifier Q> vec<3, T, Q> cross(qua<T, Q> const & q, vec<3, T, Q> const & v);
    template<typename T, qualifier Q> qua<T, Q> sqrt(qua<T, Q> const & q);

    // gtc/type_ptr
    template<length_t L, typename T, qualifier Q> constexpr T const * value_ptr(vec<L, T, Q> const & v) { return &v.x; }
    template<length_t L, typename T, qualifier Q> constexpr T * value_ptr(vec<L, T, Q> & v) { return &v.x; }
    template<length_t C, length_t R, typename T, qualifier Q> constexpr T const * value_ptr(mat<C, R, T, Q> const & m) { return &m[0].x; }
    template<length_t C, length_t R, typename T, qualifier Q> constexpr T * value_ptr(mat<C, R, T, Q> & m) { return &m[0].x; }
    template<typename T, qualifier Q> constexpr T const * value_ptr(qua<T, Q> const & q) { return &q.x; }

    typedef vec<2, float, defaultp> vec2;
    typedef vec<3, float, defaultp> vec3;
    typedef vec<4, float, defaultp> vec4;
    typedef mat<2, 2, float, defaultp> mat2;
    typedef mat<3, 3, float, defaultp> mat3;
    typedef mat<4, 4, float, defaultp> mat4;
    typedef qua<float, defaultp> quat;
}

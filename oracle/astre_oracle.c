/*
 * astre_oracle.c -- CPU restatement of the Astre transform + visibility path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under astre_b200/ (the product) may
 * import, link or execute this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, and only as the checker
 * or the reported CPU baseline.
 *
 * PARITY UNPINNED: the reference (MisterSawyer/astre) ships no test, golden
 * vector or fixture for this path, cannot be built in
 * this image (Windows-only Process/Window modules, 11 un-vendored FetchContent
 * dependencies, cmake/dependencies.cmake:11-366), and its arithmetic lives in
 * g-truc/glm tag 1.0.1 (cmake/dependencies.cmake:249-255), which is absent.
 * The GLM bodies below are restated from the published GLM 1.0.x sources
 * (scalar, non-SIMD packed_highp paths); they are checked here against
 * hand-derived known-answer vectors (tests/golden/kat_*.json) and against an
 * independent numpy float32 restatement (tests/np_reference.py), not against
 * the reference's own outputs.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (no -march), matching the
 * reference's FP environment (cmake/compile_options.cmake:56-70,118-139: no
 * FMA contraction, no fast-math).  Every float op below is a separately
 * rounded IEEE-754 binary32 operation unless fmaf() is written explicitly.
 *
 * Conventions: mat4 is float[16], column-major, m[c*4+r] == glm M[c][r]
 * (engine/modules/Math/src/math.cpp:70-93 serialises in exactly this order).
 * Quaternions are passed x,y,z,w (engine/modules/Math/src/math.cpp:124-133).
 *
 * Reference-defined parts (follow the cited lines op for op):
 *   orc_trs_matrix / orc_transform_flat / orc_basis
 *       engine/modules/ECS/src/system/transform_system.cpp:24-65
 *   orc_camera_view_proj    engine/modules/ECS/src/system/camera_system.cpp:24-37
 *   orc_light_space         engine/modules/ECS/src/system/light_system.cpp:105-160
 *   orc_interpolate_trs     engine/modules/Render/src/render.cpp:26-39
 *   participation predicate engine/modules/Pipeline/src/deferred_shading.cpp:115-121,151-156
 * Spec-defined parts (the reference has no hierarchy, bounds, frustum test,
 * sort key or sort -- SURVEY.md section 0, Appendix B): orc_transform_hierarchy,
 * orc_extract_view, orc_cull_keys, orc_sort_keys, orc_merge_lists.  Their arithmetic is fixed
 * here and the CUDA path must reproduce it bit for bit.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ---- shared layouts (mirrored, not shared, by include/astre_gpu.h) ------- */

typedef struct {
    float center[3];
    float extent[3];   /* box half extents (0 for a sphere)            */
    float radius;      /* inflation radius (0 for a plain AABB)        */
    uint32_t reserved;
} orc_bounds;

typedef struct {
    float plane[6][4];   /* left,right,bottom,top,near,far: (nx,ny,nz,w), not normalised */
    float aplane[6][4];  /* (|nx|,|ny|,|nz|, len) with len = sqrt(nx^2+ny^2+nz^2) */
    float depth[4];      /* normalised near plane (n/len, w/len) */
    float depth_scale;   /* 16384 / (distance between near and far planes) */
    uint32_t phase_mask; /* RenderPhase bits this view draws */
    uint32_t pad[2];
    float lo[4];         /* world-space box around the view volume (xyz; w unused) */
    float hi[4];
} orc_view;

/* ---- GLM 1.0.1 restatements ---------------------------------------------- */

typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } q4;  /* quaternion, glm field names */

static inline v3 v3_make(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_scale(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }

/* glm::dot(vec3): tmp = a*b; tmp.x + tmp.y + tmp.z (detail/func_geometric.inl compute_dot<vec<3>>) */
static inline float v3_dot(v3 a, v3 b)
{
    float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
    return (tx + ty) + tz;
}

/* glm::cross(vec3) (detail/func_geometric.inl compute_cross) */
static inline v3 v3_cross(v3 x, v3 y)
{
    return v3_make(x.y * y.z - y.y * x.z,
                   x.z * y.x - y.z * x.x,
                   x.x * y.y - y.x * x.y);
}

/* glm::normalize(vec3) = v * inversesqrt(dot(v,v)); inversesqrt(x) = 1/sqrt(x) */
static inline v3 v3_normalize(v3 v)
{
    float inv = 1.0f / sqrtf(v3_dot(v, v));
    return v3_scale(v, inv);
}

/* glm::operator*(qua, vec3) (detail/type_quat.inl) */
static inline v3 quat_rotate(q4 q, v3 v)
{
    v3 u = v3_make(q.x, q.y, q.z);
    v3 uv = v3_cross(u, v);
    v3 uuv = v3_cross(u, uv);
    v3 t = v3_add(v3_scale(uv, q.w), uuv);
    return v3_add(v, v3_scale(t, 2.0f));
}

static void mat4_identity(float *m)
{
    for (int i = 0; i < 16; ++i) m[i] = 0.0f;
    m[0] = m[5] = m[10] = m[15] = 1.0f;
}

/* column helpers: r = a*s ; r = a + b */
static inline void col_scale(const float *a, float s, float *r)
{
    r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; r[3] = a[3] * s;
}
static inline void col_add(const float *a, const float *b, float *r)
{
    r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; r[3] = a[3] + b[3];
}

/* glm::operator*(mat4, mat4): R[j] = A[0]*B[j].x + A[1]*B[j].y + A[2]*B[j].z + A[3]*B[j].w,
 * summed left to right (the x->w vs w->x order of the 1.0.1
 * tag is unverified; for T*R*S and for affine operands the result is the same
 * either way because all but one term of every sum is a signed zero). */
ORC_API void orc_mat4_mul(const float *a, const float *b, float *out)
{
    float r[16];
    for (int j = 0; j < 4; ++j) {
        float t0[4], t1[4], t2[4], t3[4], s[4];
        col_scale(a + 0,  b[j * 4 + 0], t0);
        col_scale(a + 4,  b[j * 4 + 1], t1);
        col_scale(a + 8,  b[j * 4 + 2], t2);
        col_scale(a + 12, b[j * 4 + 3], t3);
        col_add(t0, t1, s);
        col_add(s, t2, s);
        col_add(s, t3, r + j * 4);
    }
    memcpy(out, r, sizeof r);
}

/* glm::translate(m, v): Result = m; Result[3] = m[0]*v.x + m[1]*v.y + m[2]*v.z + m[3] */
static void glm_translate(const float *m, v3 v, float *out)
{
    float t0[4], t1[4], t2[4], s[4];
    memcpy(out, m, 16 * sizeof(float));
    col_scale(m + 0, v.x, t0);
    col_scale(m + 4, v.y, t1);
    col_scale(m + 8, v.z, t2);
    col_add(t0, t1, s);
    col_add(s, t2, s);
    col_add(s, m + 12, out + 12);
}

/* glm::scale(m, v): Result[i] = m[i]*v[i]; Result[3] = m[3] */
static void glm_scale(const float *m, v3 v, float *out)
{
    col_scale(m + 0, v.x, out + 0);
    col_scale(m + 4, v.y, out + 4);
    col_scale(m + 8, v.z, out + 8);
    memcpy(out + 12, m + 12, 4 * sizeof(float));
}

/* glm::toMat4(q) = mat4(mat3_cast(q)); no normalisation (gtc/quaternion.inl) */
static void glm_to_mat4(q4 q, float *out)
{
    float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z;
    float qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z;
    float qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
    mat4_identity(out);
    out[0]  = 1.0f - 2.0f * (qyy + qzz);
    out[1]  = 2.0f * (qxy + qwz);
    out[2]  = 2.0f * (qxz - qwy);
    out[4]  = 2.0f * (qxy - qwz);
    out[5]  = 1.0f - 2.0f * (qxx + qzz);
    out[6]  = 2.0f * (qyz + qwx);
    out[8]  = 2.0f * (qxz + qwy);
    out[9]  = 2.0f * (qyz - qwx);
    out[10] = 1.0f - 2.0f * (qxx + qyy);
}

/* transform_system.cpp:51-57: translate(I,pos) * toMat4(rot) * scale(I,scale) */
ORC_API void orc_trs_matrix(const float *pos, const float *rot_xyzw, const float *scl, float *out16)
{
    float I[16], T[16], R[16], S[16], TR[16];
    q4 q = { rot_xyzw[0], rot_xyzw[1], rot_xyzw[2], rot_xyzw[3] };
    mat4_identity(I);
    glm_translate(I, v3_make(pos[0], pos[1], pos[2]), T);
    glm_to_mat4(q, R);
    glm_scale(I, v3_make(scl[0], scl[1], scl[2]), S);
    orc_mat4_mul(T, R, TR);
    orc_mat4_mul(TR, S, out16);
}

/* transform_system.cpp:59-61 with transform_system.hpp:18-19 constants */
ORC_API void orc_basis(const float *rot_xyzw, float *fwd, float *up, float *right)
{
    q4 q = { rot_xyzw[0], rot_xyzw[1], rot_xyzw[2], rot_xyzw[3] };
    v3 f = v3_normalize(quat_rotate(q, v3_make(0.0f, 0.0f, -1.0f)));
    v3 u = v3_normalize(quat_rotate(q, v3_make(0.0f, 1.0f, 0.0f)));
    v3 r = v3_normalize(v3_cross(f, u));
    fwd[0] = f.x; fwd[1] = f.y; fwd[2] = f.z;
    up[0] = u.x; up[1] = u.y; up[2] = u.z;
    right[0] = r.x; right[1] = r.y; right[2] = r.z;
}

/* glm::lookAtRH (ext/matrix_transform.inl) */
ORC_API void orc_look_at(const float *eye3, const float *center3, const float *up3, float *out16)
{
    v3 eye = v3_make(eye3[0], eye3[1], eye3[2]);
    v3 f = v3_normalize(v3_sub(v3_make(center3[0], center3[1], center3[2]), eye));
    v3 s = v3_normalize(v3_cross(f, v3_make(up3[0], up3[1], up3[2])));
    v3 u = v3_cross(s, f);
    mat4_identity(out16);
    out16[0] = s.x;  out16[4] = s.y;  out16[8]  = s.z;
    out16[1] = u.x;  out16[5] = u.y;  out16[9]  = u.z;
    out16[2] = -f.x; out16[6] = -f.y; out16[10] = -f.z;
    out16[12] = -v3_dot(s, eye);
    out16[13] = -v3_dot(u, eye);
    out16[14] = v3_dot(f, eye);
}

/* glm::radians */
static inline float glm_radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
static inline float glm_degrees(float rad) { return rad * 57.295779513082320876798154814105f; }
static inline float glm_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

/* glm::perspectiveRH_NO (ext/matrix_clip_space.inl); fovy in radians */
ORC_API void orc_perspective(float fovy, float aspect, float z_near, float z_far, float *out16)
{
    float t = tanf(fovy / 2.0f);
    for (int i = 0; i < 16; ++i) out16[i] = 0.0f;
    out16[0] = 1.0f / (aspect * t);
    out16[5] = 1.0f / t;
    out16[10] = -(z_far + z_near) / (z_far - z_near);
    out16[11] = -1.0f;
    out16[14] = -(2.0f * z_far * z_near) / (z_far - z_near);
}

/* glm::orthoRH_NO */
ORC_API void orc_ortho(float l, float r, float b, float t, float n, float f, float *out16)
{
    mat4_identity(out16);
    out16[0] = 2.0f / (r - l);
    out16[5] = 2.0f / (t - b);
    out16[10] = -2.0f / (f - n);
    out16[12] = -(r + l) / (r - l);
    out16[13] = -(t + b) / (t - b);
    out16[14] = -(f + n) / (f - n);
}

/* camera_system.cpp:24-37.  fov in degrees as authored. */
ORC_API void orc_camera_view_proj(const float *pos, const float *fwd, const float *up,
                                  float fov_deg, float aspect, float z_near, float z_far,
                                  float *view16, float *proj16)
{
    float center[3] = { pos[0] + fwd[0], pos[1] + fwd[1], pos[2] + fwd[2] };
    orc_look_at(pos, center, up, view16);
    orc_perspective(glm_radians(fov_deg), aspect, z_near, z_far, proj16);
}

/* light_system.cpp:105-160 for one light.  type: 1 DIRECTIONAL, 2 POINT, 3 SPOT
 * (engine/modules/Proto/ECS/components/light_component.proto enum order).
 * Returns 0 and leaves out16 untouched for an unknown type (the `default: continue`). */
ORC_API int orc_light_space(const float *pos3, const float *dir3, int type, float outer_cutoff, float *out16)
{
    float view[16], proj[16];
    v3 d = v3_normalize(v3_make(dir3[0], dir3[1], dir3[2]));
    float center[3] = { pos3[0] + d.x, pos3[1] + d.y, pos3[2] + d.z };
    float upv[3] = { 0.0f, 1.0f, 0.0f };
    orc_look_at(pos3, center, upv, view);
    if (type == 1) {
        orc_ortho(-100.0f, 100.0f, -100.0f, 100.0f, 0.1f, 100.0f, proj);
    } else if (type == 3) {
        float outer = glm_clamp(outer_cutoff, -1.0f, 1.0f);
        float fov = glm_degrees(2.0f * acosf(glm_clamp(outer, -0.999f, 0.999f)));
        fov = glm_clamp(fov, 5.0f, 179.0f);
        orc_perspective(glm_radians(fov), 1.7f, 0.1f, 1024.0f, proj);
    } else if (type == 2) {
        orc_perspective(glm_radians(90.0f), 1.0f, 0.1f, 100.0f, proj);
    } else {
        return 0;
    }
    orc_mat4_mul(proj, view, out16);
    return 1;
}

/* ---- a1: flat transform --------------------------------------------------- */

ORC_API void orc_transform_flat(uint32_t n, const float *pos3, const float *rot4, const float *scl3,
                                float *world16, float *fwd3, float *up3, float *right3)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        orc_trs_matrix(pos3 + 3 * i, rot4 + 4 * i, scl3 + 3 * i, world16 + 16 * i);
        if (fwd3) orc_basis(rot4 + 4 * i, fwd3 + 3 * i, up3 + 3 * i, right3 + 3 * i);
    }
}

/* ---- a9: hierarchy (spec-defined): world = world[parent] * local ----------- */

/* Returns the number of levels (max depth + 1), or -1 if a parent index is out
 * of range or the parent links contain a cycle.  depth_out may be NULL. */
ORC_API int orc_hierarchy_depths(uint32_t n, const uint32_t *parent, uint32_t *depth_out)
{
    uint32_t *depth = depth_out ? depth_out : (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
    const uint32_t UNSET = 0xFFFFFFFFu;
    int levels = 0;
    for (uint32_t i = 0; i < n; ++i) depth[i] = UNSET;
    for (uint32_t i = 0; i < n; ++i) {
        /* walk up to a known ancestor, then unwind */
        uint32_t cur = i, steps = 0;
        while (depth[cur] == UNSET) {
            uint32_t p = parent[cur];
            if (p == 0xFFFFFFFFu) { depth[cur] = 0; break; }
            if (p >= n || ++steps > n) { if (!depth_out) free(depth); return -1; }
            cur = p;
        }
        /* second walk assigns depths */
        uint32_t base = depth[cur];
        uint32_t len = 0;
        for (uint32_t c = i; c != cur; c = parent[c]) ++len;
        uint32_t d = base + len;
        for (uint32_t c = i; c != cur; c = parent[c]) depth[c] = d--;
    }
    for (uint32_t i = 0; i < n; ++i) if ((int)depth[i] + 1 > levels) levels = (int)depth[i] + 1;
    if (!depth_out) free(depth);
    return levels;
}

ORC_API int orc_transform_hierarchy(uint32_t n, const float *pos3, const float *rot4, const float *scl3,
                                    const uint32_t *parent, float *world16)
{
    uint32_t *depth = (uint32_t *)malloc((size_t)(n ? n : 1) * sizeof(uint32_t));
    int levels = orc_hierarchy_depths(n, parent, depth);
    if (levels < 0) { free(depth); return -1; }
    for (int lvl = 0; lvl < levels; ++lvl) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            if (depth[i] != (uint32_t)lvl) continue;
            float local[16];
            orc_trs_matrix(pos3 + 3 * i, rot4 + 4 * i, scl3 + 3 * i, local);
            if (lvl == 0) memcpy(world16 + 16 * i, local, sizeof local);
            else orc_mat4_mul(world16 + 16 * (size_t)parent[i], local, world16 + 16 * i);
        }
    }
    free(depth);
    return levels;
}

/* ---- a10: views, bounds, cull, keys, sort (spec-defined, Appendix B) ------- */

/* World-space axis-aligned box around the view volume: the eight NDC corners
 * (+-1,+-1,+-1) taken through the inverse clip matrix in double precision,
 * widened by 1e-4 of its size + 1e-6, then rounded outwards to float.  A
 * singular matrix or a corner with w <= 0 gives the unbounded box.  (Spec: the
 * box test removes the classic plane-test false positives beyond the corners
 * of the volume.) */
static void view_box(const float *clip16, float *lo4, float *hi4)
{
    double m[16], inv[16];
    for (int i = 0; i < 16; ++i) m[i] = (double)clip16[i];
    inv[0]  =  m[5]*m[10]*m[15] - m[5]*m[11]*m[14] - m[9]*m[6]*m[15] + m[9]*m[7]*m[14] + m[13]*m[6]*m[11] - m[13]*m[7]*m[10];
    inv[4]  = -m[4]*m[10]*m[15] + m[4]*m[11]*m[14] + m[8]*m[6]*m[15] - m[8]*m[7]*m[14] - m[12]*m[6]*m[11] + m[12]*m[7]*m[10];
    inv[8]  =  m[4]*m[9]*m[15]  - m[4]*m[11]*m[13] - m[8]*m[5]*m[15] + m[8]*m[7]*m[13] + m[12]*m[5]*m[11] - m[12]*m[7]*m[9];
    inv[12] = -m[4]*m[9]*m[14]  + m[4]*m[10]*m[13] + m[8]*m[5]*m[14] - m[8]*m[6]*m[13] - m[12]*m[5]*m[10] + m[12]*m[6]*m[9];
    inv[1]  = -m[1]*m[10]*m[15] + m[1]*m[11]*m[14] + m[9]*m[2]*m[15] - m[9]*m[3]*m[14] - m[13]*m[2]*m[11] + m[13]*m[3]*m[10];
    inv[5]  =  m[0]*m[10]*m[15] - m[0]*m[11]*m[14] - m[8]*m[2]*m[15] + m[8]*m[3]*m[14] + m[12]*m[2]*m[11] - m[12]*m[3]*m[10];
    inv[9]  = -m[0]*m[9]*m[15]  + m[0]*m[11]*m[13] + m[8]*m[1]*m[15] - m[8]*m[3]*m[13] - m[12]*m[1]*m[11] + m[12]*m[3]*m[9];
    inv[13] =  m[0]*m[9]*m[14]  - m[0]*m[10]*m[13] - m[8]*m[1]*m[14] + m[8]*m[2]*m[13] + m[12]*m[1]*m[10] - m[12]*m[2]*m[9];
    inv[2]  =  m[1]*m[6]*m[15]  - m[1]*m[7]*m[14]  - m[5]*m[2]*m[15] + m[5]*m[3]*m[14] + m[13]*m[2]*m[7]  - m[13]*m[3]*m[6];
    inv[6]  = -m[0]*m[6]*m[15]  + m[0]*m[7]*m[14]  + m[4]*m[2]*m[15] - m[4]*m[3]*m[14] - m[12]*m[2]*m[7]  + m[12]*m[3]*m[6];
    inv[10] =  m[0]*m[5]*m[15]  - m[0]*m[7]*m[13]  - m[4]*m[1]*m[15] + m[4]*m[3]*m[13] + m[12]*m[1]*m[7]  - m[12]*m[3]*m[5];
    inv[14] = -m[0]*m[5]*m[14]  + m[0]*m[6]*m[13]  + m[4]*m[1]*m[14] - m[4]*m[2]*m[13] - m[12]*m[1]*m[6]  + m[12]*m[2]*m[5];
    inv[3]  = -m[1]*m[6]*m[11]  + m[1]*m[7]*m[10]  + m[5]*m[2]*m[11] - m[5]*m[3]*m[10] - m[9]*m[2]*m[7]   + m[9]*m[3]*m[6];
    inv[7]  =  m[0]*m[6]*m[11]  - m[0]*m[7]*m[10]  - m[4]*m[2]*m[11] + m[4]*m[3]*m[10] + m[8]*m[2]*m[7]   - m[8]*m[3]*m[6];
    inv[11] = -m[0]*m[5]*m[11]  + m[0]*m[7]*m[9]   + m[4]*m[1]*m[11] - m[4]*m[3]*m[9]  - m[8]*m[1]*m[7]   + m[8]*m[3]*m[5];
    inv[15] =  m[0]*m[5]*m[10]  - m[0]*m[6]*m[9]   - m[4]*m[1]*m[10] + m[4]*m[2]*m[9]  + m[8]*m[1]*m[6]   - m[8]*m[2]*m[5];
    const double det = m[0]*inv[0] + m[1]*inv[4] + m[2]*inv[8] + m[3]*inv[12];
    int ok = (det != 0.0) && isfinite(det);
    double lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (int c = 0; c < 8 && ok; ++c) {
        const double x = (c & 1) ? 1.0 : -1.0, y = (c & 2) ? 1.0 : -1.0, z = (c & 4) ? 1.0 : -1.0;
        /* adj * (x,y,z,1); the common 1/det cancels in the divide but fixes the sign of w */
        const double w = (inv[3]*x + inv[7]*y + inv[11]*z + inv[15]) / det;
        if (!(w > 0.0) || !isfinite(w)) { ok = 0; break; }
        const double p[3] = { (inv[0]*x + inv[4]*y + inv[8]*z  + inv[12]) / det / w,
                              (inv[1]*x + inv[5]*y + inv[9]*z  + inv[13]) / det / w,
                              (inv[2]*x + inv[6]*y + inv[10]*z + inv[14]) / det / w };
        for (int i = 0; i < 3; ++i) {
            if (!isfinite(p[i])) ok = 0;
            if (p[i] < lo[i]) lo[i] = p[i];
            if (p[i] > hi[i]) hi[i] = p[i];
        }
    }
    for (int i = 0; i < 3; ++i) {
        if (!ok) { lo4[i] = -INFINITY; hi4[i] = INFINITY; continue; }
        const double d = (hi[i] - lo[i]) * 1e-4 + 1e-6;
        lo4[i] = nextafterf((float)(lo[i] - d), -INFINITY);
        hi4[i] = nextafterf((float)(hi[i] + d), INFINITY);
    }
    lo4[3] = 0.0f; hi4[3] = 0.0f;
}

/* Planes from clip = P*V (GL clip volume -w <= x,y,z <= w), column-major input. */
ORC_API void orc_extract_view(const float *clip16, uint32_t phase_mask, orc_view *v)
{
    float row[4][4];
    for (int i = 0; i < 4; ++i)
        for (int c = 0; c < 4; ++c) row[i][c] = clip16[c * 4 + i];
    for (int k = 0; k < 4; ++k) {
        v->plane[0][k] = row[3][k] + row[0][k];   /* left   */
        v->plane[1][k] = row[3][k] - row[0][k];   /* right  */
        v->plane[2][k] = row[3][k] + row[1][k];   /* bottom */
        v->plane[3][k] = row[3][k] - row[1][k];   /* top    */
        v->plane[4][k] = row[3][k] + row[2][k];   /* near   */
        v->plane[5][k] = row[3][k] - row[2][k];   /* far    */
    }
    for (int p = 0; p < 6; ++p) {
        float nx = v->plane[p][0], ny = v->plane[p][1], nz = v->plane[p][2];
        v->aplane[p][0] = fabsf(nx);
        v->aplane[p][1] = fabsf(ny);
        v->aplane[p][2] = fabsf(nz);
        v->aplane[p][3] = sqrtf((nx * nx + ny * ny) + nz * nz);
    }
    float ln = v->aplane[4][3], lf = v->aplane[5][3];
    v->depth[0] = v->plane[4][0] / ln;
    v->depth[1] = v->plane[4][1] / ln;
    v->depth[2] = v->plane[4][2] / ln;
    v->depth[3] = v->plane[4][3] / ln;
    float range = v->depth[3] + v->plane[5][3] / lf;
    float scale = 16384.0f / range;
    v->depth_scale = (range > 0.0f && scale < INFINITY) ? scale : 0.0f;
    v->phase_mask = phase_mask;
    v->pad[0] = v->pad[1] = 0;
    view_box(clip16, v->lo, v->hi);
}

/* key = world(Wb) | view(5) | shader(8) | mesh(11) | depth(14) | local slot(26-Wb),
 * most significant first; Wb = ceil(log2(n_worlds)). */
static inline uint32_t world_bits_of(uint32_t n_worlds)
{
    uint32_t b = 0;
    while ((1u << b) < n_worlds) ++b;
    return b;
}

static inline uint64_t make_key(uint32_t wbits, uint32_t world, uint32_t view, uint32_t shader,
                                uint32_t mesh, uint32_t depth, uint32_t local)
{
    uint64_t k = world;
    k = (k << 5) | (view & 31u);
    k = (k << 8) | (shader & 255u);
    k = (k << 11) | (mesh & 2047u);
    k = (k << 14) | (depth & 16383u);
    k = (k << (26 - wbits)) | (uint64_t)local;
    return k;
}

/* World-space bounds of one entity: centre c' = M*(c,1); box half extents
 * e'_i = sum_j |M[j][i]|*e_j; radius r' = r * sqrt(max_j |M[j].xyz|^2).
 * fmaf where written; everything else separately rounded. */
typedef struct { float c[3]; float e[3]; float r; } world_bounds;

static inline void world_bounds_of(const float *m, const orc_bounds *b, world_bounds *wb)
{
    for (int i = 0; i < 3; ++i)
        wb->c[i] = fmaf(m[0 + i], b->center[0], fmaf(m[4 + i], b->center[1], fmaf(m[8 + i], b->center[2], m[12 + i])));
    for (int i = 0; i < 3; ++i)
        wb->e[i] = fmaf(fabsf(m[0 + i]), b->extent[0],
                        fmaf(fabsf(m[4 + i]), b->extent[1], fabsf(m[8 + i]) * b->extent[2]));
    float l0 = fmaf(m[0], m[0], fmaf(m[1], m[1], m[2] * m[2]));
    float l1 = fmaf(m[4], m[4], fmaf(m[5], m[5], m[6] * m[6]));
    float l2 = fmaf(m[8], m[8], fmaf(m[9], m[9], m[10] * m[10]));
    float mx = fmaxf(fmaxf(l0, l1), l2);
    wb->r = b->radius * sqrtf(mx);
}

/* Returns 1 if the bounds overlap the view's world-space box and are inside or
 * intersect all six planes.  A NaN never culls (the comparisons are false on
 * unordered operands). */
static inline int inside_view(const world_bounds *wb, const orc_view *v)
{
    for (int i = 0; i < 3; ++i) {
        const float t = wb->e[i] + wb->r;
        if (wb->c[i] + t < v->lo[i]) return 0;
        if (wb->c[i] - t > v->hi[i]) return 0;
    }
    for (int p = 0; p < 6; ++p) {
        const float *pl = v->plane[p], *ap = v->aplane[p];
        float dist = fmaf(pl[0], wb->c[0], fmaf(pl[1], wb->c[1], fmaf(pl[2], wb->c[2], pl[3])));
        float rad = fmaf(ap[0], wb->e[0], fmaf(ap[1], wb->e[1], fmaf(ap[2], wb->e[2], wb->r * ap[3])));
        if (dist + rad < 0.0f) return 0;
    }
    return 1;
}

static inline uint32_t depth14(const world_bounds *wb, const orc_view *v)
{
    float zd = fmaf(v->depth[0], wb->c[0], fmaf(v->depth[1], wb->c[1], fmaf(v->depth[2], wb->c[2], v->depth[3])));
    float t = zd * v->depth_scale;
    if (!(t >= 0.0f)) return 0u;
    if (t >= 16383.0f) return 16383u;
    return (uint32_t)t;
}

/* flags: bit0 = VisualComponent.visible, bits 1..4 = RenderPhase mask
 * (render.hpp:29-36; visual_system.cpp:32,55).  An entity takes part in view v
 * iff visible && (phases & view.phase_mask) -- deferred_shading.cpp:117-120,153-156.
 * World w owns slots [w*E, (w+1)*E) and views[w*F .. w*F+F).  Emits keys
 * world-major, view-major, slot order (unsorted by key).  Returns the number of
 * keys that exist (may exceed cap; only the first cap are stored).
 * counts[w*F+v] = keys of (world w, view v). */
static uint64_t cull_range(uint32_t lo, uint32_t hi, uint32_t w, uint32_t v, uint32_t E, uint32_t F, uint32_t wbits,
                           const float *world16, const uint32_t *mesh_id, const uint32_t *shader_id,
                           const uint8_t *flags, const orc_bounds *bounds, uint32_t n_meshes,
                           const orc_view *views, uint64_t **buf, uint64_t *cap)
{
    const orc_view *vw = views + (size_t)w * F + v;
    uint64_t cnt = 0;
    for (uint32_t i = lo; i < hi; ++i) {
        uint32_t fl = flags[i];
        if (!(fl & 1u)) continue;
        if (!(((fl >> 1) & 15u) & vw->phase_mask)) continue;
        uint32_t mesh = mesh_id[i];
        if (mesh >= n_meshes) continue;
        world_bounds wb;
        world_bounds_of(world16 + 16 * (size_t)i, bounds + mesh, &wb);
        if (!inside_view(&wb, vw)) continue;
        if (cnt == *cap) { *cap = *cap ? *cap * 2 : 1024; *buf = (uint64_t *)realloc(*buf, *cap * sizeof(uint64_t)); }
        (*buf)[cnt++] = make_key(wbits, w, v, shader_id[i], mesh, depth14(&wb, vw), i - w * E);
    }
    return cnt;
}

ORC_API uint64_t orc_cull_keys(uint32_t n, const float *world16, const uint32_t *mesh_id, const uint32_t *shader_id,
                               const uint8_t *flags, const orc_bounds *bounds, uint32_t n_meshes,
                               uint32_t n_worlds, uint32_t entities_per_world, uint32_t views_per_world,
                               const orc_view *views, uint64_t *keys, uint64_t cap, uint32_t *counts)
{
    uint32_t F = views_per_world, W = n_worlds ? n_worlds : 1;
    uint32_t E = (W == 1) ? n : entities_per_world;
    uint32_t wbits = world_bits_of(W);
    uint32_t units = W * F;
    uint64_t **ubuf = (uint64_t **)calloc(units ? units : 1, sizeof(uint64_t *));
    uint64_t *ucnt = (uint64_t *)calloc(units ? units : 1, sizeof(uint64_t));
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t u = 0; u < (int64_t)units; ++u) {
        uint32_t w = (uint32_t)(u / F), v = (uint32_t)(u % F);
        uint64_t lo = (uint64_t)w * E, hi = lo + E;
        if (lo > n) lo = n;
        if (hi > n) hi = n;
        uint64_t c = 0;
        ucnt[u] = cull_range((uint32_t)lo, (uint32_t)hi, w, v, E, F, wbits, world16, mesh_id, shader_id, flags,
                             bounds, n_meshes, views, &ubuf[u], &c);
    }
    uint64_t total = 0;
    for (uint32_t u = 0; u < units; ++u) {
        for (uint64_t k = 0; k < ucnt[u]; ++k) {
            if (total < cap) keys[total] = ubuf[u][k];
            ++total;
        }
        counts[u] = (uint32_t)ucnt[u];
        free(ubuf[u]);
    }
    free(ubuf); free(ucnt);
    return total;
}

/* Same result as orc_cull_keys with W == 1, parallel over entity chunks instead
 * of over (world, view) units: the CPU-baseline variant for one big world. */
ORC_API uint64_t orc_cull_keys_mt(uint32_t n, const float *world16, const uint32_t *mesh_id, const uint32_t *shader_id,
                                  const uint8_t *flags, const orc_bounds *bounds, uint32_t n_meshes,
                                  uint32_t views_per_world, const orc_view *views,
                                  uint64_t *keys, uint64_t cap, uint32_t *counts)
{
    uint32_t F = views_per_world;
    int T = 1;
#ifdef _OPENMP
    T = omp_get_max_threads();
#endif
    uint32_t chunk = (n + (uint32_t)T - 1) / (uint32_t)T;
    chunk = (chunk + 3u) & ~3u;
    uint64_t **tb = (uint64_t **)calloc((size_t)T * (F ? F : 1), sizeof(uint64_t *));
    uint64_t *tc = (uint64_t *)calloc((size_t)T * (F ? F : 1), sizeof(uint64_t));
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < T; ++t) {
        uint64_t lo = (uint64_t)t * chunk, hi = lo + chunk;
        if (lo > n) lo = n;
        if (hi > n) hi = n;
        for (uint32_t v = 0; v < F; ++v) {
            uint64_t c = 0;
            tc[(size_t)v * T + t] = cull_range((uint32_t)lo, (uint32_t)hi, 0, v, n, F, 0, world16, mesh_id, shader_id,
                                               flags, bounds, n_meshes, views, &tb[(size_t)v * T + t], &c);
        }
    }
    uint64_t total = 0;
    for (uint32_t v = 0; v < F; ++v) {
        uint64_t cnt = 0;
        for (int t = 0; t < T; ++t) {
            uint64_t c = tc[(size_t)v * T + t];
            for (uint64_t k = 0; k < c; ++k) {
                if (total < cap) keys[total] = tb[(size_t)v * T + t][k];
                ++total;
            }
            cnt += c;
            free(tb[(size_t)v * T + t]);
        }
        counts[v] = (uint32_t)cnt;
    }
    free(tb); free(tc);
    return total;
}

/* Ascending sort of unique 64-bit keys: LSD radix, 8 passes of 8 bits.  Any
 * correct sort gives the same answer because keys are unique. */
ORC_API void orc_sort_keys(uint64_t *keys, uint64_t n)
{
    if (n < 2) return;
    uint64_t *tmp = (uint64_t *)malloc(n * sizeof(uint64_t));
    uint64_t *src = keys, *dst = tmp;
    for (int pass = 0; pass < 8; ++pass) {
        uint64_t hist[256] = {0};
        int shift = pass * 8;
        for (uint64_t i = 0; i < n; ++i) hist[(src[i] >> shift) & 255u]++;
        if (hist[(src[0] >> shift) & 255u] == n) continue;   /* all keys share this digit */
        uint64_t sum = 0;
        for (int d = 0; d < 256; ++d) { uint64_t c = hist[d]; hist[d] = sum; sum += c; }
        for (uint64_t i = 0; i < n; ++i) dst[hist[(src[i] >> shift) & 255u]++] = src[i];
        uint64_t *t = src; src = dst; dst = t;
    }
    if (src != keys) memcpy(keys, src, n * sizeof(uint64_t));
    free(tmp);
}

/* Whole frame, single shard: transform (flat or hierarchical) -> cull -> sort.
 * parent may be NULL (flat).  Returns total keys, or (uint64_t)-1 on a bad hierarchy. */
ORC_API uint64_t orc_frame(uint32_t n, const float *pos3, const float *rot4, const float *scl3,
                           const uint32_t *parent, const uint32_t *mesh_id, const uint32_t *shader_id,
                           const uint8_t *flags, const orc_bounds *bounds, uint32_t n_meshes,
                           uint32_t n_worlds, uint32_t entities_per_world, uint32_t views_per_world,
                           const orc_view *views,
                           float *world16, uint64_t *keys, uint64_t cap, uint32_t *counts)
{
    if (parent) {
        if (orc_transform_hierarchy(n, pos3, rot4, scl3, parent, world16) < 0) return (uint64_t)-1;
    } else {
        orc_transform_flat(n, pos3, rot4, scl3, world16, NULL, NULL, NULL);
    }
    uint64_t total = (n_worlds <= 1)
        ? orc_cull_keys_mt(n, world16, mesh_id, shader_id, flags, bounds, n_meshes, views_per_world, views, keys, cap, counts)
        : orc_cull_keys(n, world16, mesh_id, shader_id, flags, bounds, n_meshes, n_worlds, entities_per_world,
                        views_per_world, views, keys, cap, counts);
    orc_sort_keys(keys, total < cap ? total : cap);
    return total;
}

/* ---- f3: global key-sorted draw list across shards (spec-defined) --
 *
 * Shard r holds a contiguous range of the scene's entities and a list sorted by
 * its own keys.  The global list orders all keys by
 *     (key >> slot_bits, shard, key & slot_mask)
 * i.e. class and depth first, then global entity order.  Plain G-way merge by
 * repeatedly taking the smallest head; out_src[i] = shard of out_keys[i]. */
ORC_API void orc_merge_lists(uint32_t n_lists, const uint64_t *const *lists, const uint64_t *counts,
                             uint32_t slot_bits, uint64_t *out_keys, uint8_t *out_src)
{
    uint64_t head[64] = {0};
    uint64_t at = 0;
    if (n_lists > 64) return;
    for (;;) {
        int best = -1;
        uint64_t best_cls = 0;
        for (uint32_t t = 0; t < n_lists; ++t) {
            if (head[t] >= counts[t]) continue;
            const uint64_t cls = lists[t][head[t]] >> slot_bits;
            if (best < 0 || cls < best_cls) { best = (int)t; best_cls = cls; }   /* ties: lower shard first */
        }
        if (best < 0) break;
        out_keys[at] = lists[best][head[best]++];
        out_src[at++] = (uint8_t)best;
    }
}

/* ---- f1: interpolateFrame per-proxy step (render.cpp:26-39) ---------------- */

static inline float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }

/* glm::slerp(qua,qua,T) (ext/quaternion_common.inl) */
static q4 glm_slerp(q4 x, q4 y, float a)
{
    q4 z = y;
    /* compute_dot<qua>: (w*w + x*x) + (y*y + z*z) */
    float cosTheta = (x.w * y.w + x.x * y.x) + (x.y * y.y + x.z * y.z);
    if (cosTheta < 0.0f) { z.x = -y.x; z.y = -y.y; z.z = -y.z; z.w = -y.w; cosTheta = -cosTheta; }
    q4 r;
    if (cosTheta > 1.0f - FLT_EPSILON) {
        r.w = mixf(x.w, z.w, a); r.x = mixf(x.x, z.x, a); r.y = mixf(x.y, z.y, a); r.z = mixf(x.z, z.z, a);
    } else {
        float angle = acosf(cosTheta);
        float s0 = sinf((1.0f - a) * angle), s1 = sinf(a * angle), sd = sinf(angle);
        r.w = (s0 * x.w + s1 * z.w) / sd;
        r.x = (s0 * x.x + s1 * z.x) / sd;
        r.y = (s0 * x.y + s1 * z.y) / sd;
        r.z = (s0 * x.z + s1 * z.z) / sd;
    }
    return r;
}

/* alpha is clamped to [0,1] first (render.cpp:10). Outputs the interpolated model matrix. */
ORC_API void orc_interpolate_trs(uint32_t n, float alpha,
                                 const float *pos_a, const float *rot_a, const float *scl_a,
                                 const float *pos_b, const float *rot_b, const float *scl_b,
                                 float *world16)
{
    alpha = alpha < 0.0f ? 0.0f : (alpha > 1.0f ? 1.0f : alpha);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        float p[3], s[3], r[4];
        for (int k = 0; k < 3; ++k) {
            p[k] = mixf(pos_a[3 * i + k], pos_b[3 * i + k], alpha);
            s[k] = mixf(scl_a[3 * i + k], scl_b[3 * i + k], alpha);
        }
        q4 qa = { rot_a[4 * i], rot_a[4 * i + 1], rot_a[4 * i + 2], rot_a[4 * i + 3] };
        q4 qb = { rot_b[4 * i], rot_b[4 * i + 1], rot_b[4 * i + 2], rot_b[4 * i + 3] };
        q4 q = glm_slerp(qa, qb, alpha);
        r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w;
        orc_trs_matrix(p, r, s, world16 + 16 * i);
    }
}

ORC_API int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_threads(int t)
{
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

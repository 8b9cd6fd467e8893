// Host-side view construction for the GPU path: the once-per-view-per-frame
// math that stays on the CPU ("Transcendentals").
//
// Restates the GLM 1.0.1 scalar bodies the reference reaches through
// engine/modules/Math/include/math/math.hpp:92-98 (lookAt, perspective) and
// glm::ortho, in GLM's operation order, float only.  Compiled with
// -ffp-contract=off.  mat4 = float[16], column-major, m[c*4+r].
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace astre_host {

struct V3 { float x, y, z; };

inline V3 sub(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 mul(V3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline float dot(V3 a, V3 b) { const float x = a.x * b.x, y = a.y * b.y, z = a.z * b.z; return (x + y) + z; }
inline V3 cross(V3 a, V3 b) { return { a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y }; }
inline V3 normalize(V3 v) { return mul(v, 1.0f / std::sqrt(dot(v, v))); }

inline void identity(float *m)
{
    for (int i = 0; i < 16; ++i) m[i] = 0.0f;
    m[0] = m[5] = m[10] = m[15] = 1.0f;
}

// glm::operator*(mat4,mat4), columns accumulated x -> w
inline void mat_mul(const float *a, const float *b, float *out)
{
    float r[16];
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            float s = a[i] * b[j * 4] + a[4 + i] * b[j * 4 + 1];
            s = s + a[8 + i] * b[j * 4 + 2];
            r[j * 4 + i] = s + a[12 + i] * b[j * 4 + 3];
        }
    std::memcpy(out, r, sizeof r);
}

// glm::lookAtRH
inline void look_at(V3 eye, V3 center, V3 up, float *m)
{
    const V3 f = normalize(sub(center, eye));
    const V3 s = normalize(cross(f, up));
    const V3 u = cross(s, f);
    identity(m);
    m[0] = s.x; m[4] = s.y; m[8] = s.z;
    m[1] = u.x; m[5] = u.y; m[9] = u.z;
    m[2] = -f.x; m[6] = -f.y; m[10] = -f.z;
    m[12] = -dot(s, eye);
    m[13] = -dot(u, eye);
    m[14] = dot(f, eye);
}

inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
inline float degrees(float rad) { return rad * 57.295779513082320876798154814105f; }
inline float clampf(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }

// glm::perspectiveRH_NO
inline void perspective(float fovy, float aspect, float zn, float zf, float *m)
{
    const float t = std::tan(fovy / 2.0f);
    for (int i = 0; i < 16; ++i) m[i] = 0.0f;
    m[0] = 1.0f / (aspect * t);
    m[5] = 1.0f / t;
    m[10] = -(zf + zn) / (zf - zn);
    m[11] = -1.0f;
    m[14] = -(2.0f * zf * zn) / (zf - zn);
}

// glm::orthoRH_NO
inline void ortho(float l, float r, float b, float t, float zn, float zf, float *m)
{
    identity(m);
    m[0] = 2.0f / (r - l);
    m[5] = 2.0f / (t - b);
    m[10] = -2.0f / (zf - zn);
    m[12] = -(r + l) / (r - l);
    m[13] = -(t + b) / (t - b);
    m[14] = -(zf + zn) / (zf - zn);
}

// Host image of the device view record (same 256-byte layout as astre::ViewDev).
struct ViewHost {
    float plane[6][4];
    float aplane[6][4];
    float depth[4];
    float depth_scale;
    uint32_t phase_mask;
    uint32_t sym;     // bit i: planes 2i, 2i+1 share bit-identical |n| and length
    uint32_t pad;
    float lo[4];      // world-space box around the view volume (xyz)
    float hi[4];
};

// World-space box around the view volume (spec, SURVEY.md Appendix B extension):
// the eight NDC corners through the adjugate of the clip matrix in double
// precision, widened by 1e-4 of its size + 1e-6 and rounded outwards to float;
// unbounded when the matrix is singular or a corner has w <= 0.
inline void view_box(const float *clip, float *lo4, float *hi4)
{
    double m[16], a[16];
    for (int i = 0; i < 16; ++i) m[i] = (double)clip[i];
    a[0]  =  m[5]*m[10]*m[15] - m[5]*m[11]*m[14] - m[9]*m[6]*m[15] + m[9]*m[7]*m[14] + m[13]*m[6]*m[11] - m[13]*m[7]*m[10];
    a[4]  = -m[4]*m[10]*m[15] + m[4]*m[11]*m[14] + m[8]*m[6]*m[15] - m[8]*m[7]*m[14] - m[12]*m[6]*m[11] + m[12]*m[7]*m[10];
    a[8]  =  m[4]*m[9]*m[15]  - m[4]*m[11]*m[13] - m[8]*m[5]*m[15] + m[8]*m[7]*m[13] + m[12]*m[5]*m[11] - m[12]*m[7]*m[9];
    a[12] = -m[4]*m[9]*m[14]  + m[4]*m[10]*m[13] + m[8]*m[5]*m[14] - m[8]*m[6]*m[13] - m[12]*m[5]*m[10] + m[12]*m[6]*m[9];
    a[1]  = -m[1]*m[10]*m[15] + m[1]*m[11]*m[14] + m[9]*m[2]*m[15] - m[9]*m[3]*m[14] - m[13]*m[2]*m[11] + m[13]*m[3]*m[10];
    a[5]  =  m[0]*m[10]*m[15] - m[0]*m[11]*m[14] - m[8]*m[2]*m[15] + m[8]*m[3]*m[14] + m[12]*m[2]*m[11] - m[12]*m[3]*m[10];
    a[9]  = -m[0]*m[9]*m[15]  + m[0]*m[11]*m[13] + m[8]*m[1]*m[15] - m[8]*m[3]*m[13] - m[12]*m[1]*m[11] + m[12]*m[3]*m[9];
    a[13] =  m[0]*m[9]*m[14]  - m[0]*m[10]*m[13] - m[8]*m[1]*m[14] + m[8]*m[2]*m[13] + m[12]*m[1]*m[10] - m[12]*m[2]*m[9];
    a[2]  =  m[1]*m[6]*m[15]  - m[1]*m[7]*m[14]  - m[5]*m[2]*m[15] + m[5]*m[3]*m[14] + m[13]*m[2]*m[7]  - m[13]*m[3]*m[6];
    a[6]  = -m[0]*m[6]*m[15]  + m[0]*m[7]*m[14]  + m[4]*m[2]*m[15] - m[4]*m[3]*m[14] - m[12]*m[2]*m[7]  + m[12]*m[3]*m[6];
    a[10] =  m[0]*m[5]*m[15]  - m[0]*m[7]*m[13]  - m[4]*m[1]*m[15] + m[4]*m[3]*m[13] + m[12]*m[1]*m[7]  - m[12]*m[3]*m[5];
    a[14] = -m[0]*m[5]*m[14]  + m[0]*m[6]*m[13]  + m[4]*m[1]*m[14] - m[4]*m[2]*m[13] - m[12]*m[1]*m[6]  + m[12]*m[2]*m[5];
    a[3]  = -m[1]*m[6]*m[11]  + m[1]*m[7]*m[10]  + m[5]*m[2]*m[11] - m[5]*m[3]*m[10] - m[9]*m[2]*m[7]   + m[9]*m[3]*m[6];
    a[7]  =  m[0]*m[6]*m[11]  - m[0]*m[7]*m[10]  - m[4]*m[2]*m[11] + m[4]*m[3]*m[10] + m[8]*m[2]*m[7]   - m[8]*m[3]*m[6];
    a[11] = -m[0]*m[5]*m[11]  + m[0]*m[7]*m[9]   + m[4]*m[1]*m[11] - m[4]*m[3]*m[9]  - m[8]*m[1]*m[7]   + m[8]*m[3]*m[5];
    a[15] =  m[0]*m[5]*m[10]  - m[0]*m[6]*m[9]   - m[4]*m[1]*m[10] + m[4]*m[2]*m[9]  + m[8]*m[1]*m[6]   - m[8]*m[2]*m[5];
    const double det = m[0]*a[0] + m[1]*a[4] + m[2]*a[8] + m[3]*a[12];
    bool ok = (det != 0.0) && std::isfinite(det);
    double lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (int c = 0; c < 8 && ok; ++c) {
        const double x = (c & 1) ? 1.0 : -1.0, y = (c & 2) ? 1.0 : -1.0, z = (c & 4) ? 1.0 : -1.0;
        const double w = (a[3]*x + a[7]*y + a[11]*z + a[15]) / det;
        if (!(w > 0.0) || !std::isfinite(w)) { ok = false; break; }
        const double p[3] = { (a[0]*x + a[4]*y + a[8]*z  + a[12]) / det / w,
                              (a[1]*x + a[5]*y + a[9]*z  + a[13]) / det / w,
                              (a[2]*x + a[6]*y + a[10]*z + a[14]) / det / w };
        for (int i = 0; i < 3; ++i) {
            if (!std::isfinite(p[i])) ok = false;
            if (p[i] < lo[i]) lo[i] = p[i];
            if (p[i] > hi[i]) hi[i] = p[i];
        }
    }
    for (int i = 0; i < 3; ++i) {
        if (!ok) { lo4[i] = -INFINITY; hi4[i] = INFINITY; continue; }
        const double d = (hi[i] - lo[i]) * 1e-4 + 1e-6;
        lo4[i] = std::nextafterf((float)(lo[i] - d), -INFINITY);
        hi4[i] = std::nextafterf((float)(hi[i] + d), INFINITY);
    }
    lo4[3] = 0.0f; hi4[3] = 0.0f;
}

// Frustum planes of clip = P*V for the GL clip volume -w <= x,y,z <= w
// (engine/resources/shaders/glsl/deferred_shader/vertex.glsl:17-22 fixes the
// P*V*M convention).  Spec: SURVEY.md Appendix B.
inline void extract_view(const float *clip, uint32_t phase_mask, ViewHost *v)
{
    for (int k = 0; k < 4; ++k) {
        const float r0 = clip[k * 4 + 0], r1 = clip[k * 4 + 1], r2 = clip[k * 4 + 2], r3 = clip[k * 4 + 3];
        v->plane[0][k] = r3 + r0;
        v->plane[1][k] = r3 - r0;
        v->plane[2][k] = r3 + r1;
        v->plane[3][k] = r3 - r1;
        v->plane[4][k] = r3 + r2;
        v->plane[5][k] = r3 - r2;
    }
    for (int p = 0; p < 6; ++p) {
        const float x = v->plane[p][0], y = v->plane[p][1], z = v->plane[p][2];
        v->aplane[p][0] = std::fabs(x);
        v->aplane[p][1] = std::fabs(y);
        v->aplane[p][2] = std::fabs(z);
        v->aplane[p][3] = std::sqrt((x * x + y * y) + z * z);
    }
    const float ln = v->aplane[4][3], lf = v->aplane[5][3];
    for (int k = 0; k < 4; ++k) v->depth[k] = v->plane[4][k] / ln;
    const float range = v->depth[3] + v->plane[5][3] / lf;
    const float scale = 16384.0f / range;
    v->depth_scale = (range > 0.0f && scale < INFINITY) ? scale : 0.0f;
    v->phase_mask = phase_mask;
    v->sym = 0;
    for (int i = 0; i < 3; ++i)
        if (std::memcmp(v->aplane[2 * i], v->aplane[2 * i + 1], 4 * sizeof(float)) == 0) v->sym |= 1u << i;
    v->pad = 0;
    view_box(clip, v->lo, v->hi);
}

}  // namespace astre_host

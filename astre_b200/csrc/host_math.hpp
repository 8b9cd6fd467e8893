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

// Host image of the device view record (same 224-byte layout as astre::ViewDev).
struct ViewHost {
    float plane[6][4];
    float aplane[6][4];
    float depth[4];
    float depth_scale;
    uint32_t phase_mask;
    uint32_t sym;     // bit i: planes 2i, 2i+1 share bit-identical |n| and length
    uint32_t pad;
};

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
}

}  // namespace astre_host

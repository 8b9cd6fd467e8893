/* Fast generator of the synthetic entity streams (harness only -- not product, not oracle).
 *
 * Same bits as astre_b200/scenes.py:entity_block (the numpy definition, which stays the specification and the
 * fallback): every value comes from a splitmix64-style hash of (seed, entity index, channel) turned into a
 * 24-bit-mantissa float, so any machine produces identical float32 inputs.  Built with
 * -ffp-contract=off: each float operation below rounds once, as numpy's float32 arithmetic does.
 * tests/test_scenes.py checks this library against the numpy path bit for bit.
 */
#include <math.h>
#include <stdint.h>

#define GOLD 0x9E3779B97F4A7C15ull

static inline uint64_t mix(uint64_t x)
{
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

static inline uint64_t hash3(uint64_t base, uint64_t index, uint64_t channel)
{
    return mix(mix(base + index * 16ull + channel) + GOLD);
}

static inline float unit(uint64_t base, uint64_t index, uint64_t channel)
{
    return (float)(hash3(base, index, channel) >> 40) * 0x1p-24f;
}

static inline float uniform(uint64_t base, uint64_t index, uint64_t channel, float lo, float span)
{
    const float t = span * unit(base, index, channel);      /* two roundings, like numpy */
    return lo + t;
}

/* Entities [lo, hi) of stream `seed`.  pos_lo/pos_span etc. are the float32 values numpy derives
 * (np.float32(lo), np.float32(hi - lo)).  flags: bit0 visible, bits 1..4 phases. */
__attribute__((visibility("default")))
void scenegen_block(uint64_t seed, uint64_t lo, uint64_t hi, float pos_lo, float pos_span, float scl_lo, float scl_span,
                    uint32_t n_mesh, uint32_t n_shader, uint8_t phases,
                    float *pos3, float *rot4, float *scl3, uint32_t *mesh, uint32_t *shader, uint8_t *flags)
{
    const uint64_t base = seed * GOLD;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)(hi - lo); ++k) {
        const uint64_t i = lo + (uint64_t)k;
        for (int c = 0; c < 3; ++c) pos3[3 * k + c] = uniform(base, i, (uint64_t)c, pos_lo, pos_span);
        float q[4];
        for (int c = 0; c < 4; ++c) q[c] = uniform(base, i, (uint64_t)(3 + c), -1.0f, 2.0f);
        const float a = q[0] * q[0], b = q[1] * q[1], c2 = q[2] * q[2], d = q[3] * q[3];
        const float ab = a + b, cd = c2 + d;
        const float n2 = ab + cd;
        const float len = sqrtf(n2);
        for (int c = 0; c < 4; ++c) {
            float v = q[c] / len;
            if (!isfinite(v)) v = 0.0f;
            rot4[4 * k + c] = v;
        }
        if ((i & 63ull) == 63ull) for (int c = 0; c < 4; ++c) rot4[4 * k + c] = q[c];     /* un-normalised */
        if ((i & 1023ull) == 1023ull) for (int c = 0; c < 4; ++c) rot4[4 * k + c] = 0.0f; /* zero quaternion */
        for (int c = 0; c < 3; ++c) scl3[3 * k + c] = uniform(base, i, (uint64_t)(7 + c), scl_lo, scl_span);
        mesh[k] = (uint32_t)((hash3(base, i, 10) >> 33) % n_mesh);
        shader[k] = (uint32_t)((hash3(base, i, 11) >> 33) % n_shader);
        flags[k] = (uint8_t)((((i & 15ull) != 15ull) ? 1u : 0u) | ((phases & 15u) << 1));
    }
}

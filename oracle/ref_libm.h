/* oracle/ref_libm.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Portable restatement of the glibc 2.39 x86-64 single-precision libm routines that the
 * reference reaches through Rust std (`f32::sin/cos/atan2`, akaze/src/descriptors.rs:70-71,
 * akaze/src/scale_space_extrema.rs:242).  On linux-gnu Rust lowers these to the system
 * libm, so "what the reference computes" == "what glibc computes" on the host.
 *
 *  - sinf/cosf : glibc sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, s_sincosf.h (Arm optimized
 *                routines; double-precision polynomial).  x86-64 glibc dispatches (ifunc) to a
 *                variant built with -mfma on every FMA-capable CPU; REF_LIBM_FMA selects which
 *                contraction pattern is restated (1 = the FMA variant, the default).
 *  - atanf/atan2f : glibc sysdeps/ieee754/flt-32/s_atanf.c, e_atan2f.c (fdlibm float code,
 *                built without FMA on x86-64).
 * tests/test_libm.py pins these against the host libm (exhaustively over [0, 2pi] for
 * sin/cos; large random sweeps for atan2f).
 */
#ifndef REF_LIBM_H
#define REF_LIBM_H
#include <stdint.h>
#include <string.h>
#include <math.h>

#ifndef REF_LIBM_FMA
#define REF_LIBM_FMA 1
#endif

static inline uint32_t rl_asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float rl_asfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

#if REF_LIBM_FMA
#define RL_MADD(a, b, c) fma((a), (b), (c))
#else
#define RL_MADD(a, b, c) ((a) * (b) + (c))
#endif

/* cosine polynomial c0..c4, sine polynomial s1..s3 (table entry 0); entry 1 = negated cosine. */
static const double RL_C0 = 0x1p0, RL_C1 = -0x1.ffffffd0c621cp-2, RL_C2 = 0x1.55553e1068f19p-5,
                    RL_C3 = -0x1.6c087e89a359dp-10, RL_C4 = 0x1.99343027bf8c3p-16;
static const double RL_S1 = -0x1.555545995a603p-3, RL_S2 = 0x1.1107605230bc4p-7,
                    RL_S3 = -0x1.994eb3774cf24p-13;
static const double RL_HPI_INV = 0x1.45F306DC9C883p+23; /* 2/pi * 2^24 */
static const double RL_HPI = 0x1.921FB54442D18p0;

static inline float rl_sinf_poly(double x, double x2, int neg_cos, int n) {
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double s1 = RL_MADD(x2, RL_S3, RL_S2);
        double x7 = x3 * x2;
        double s = RL_MADD(x3, RL_S1, x);
        return (float)RL_MADD(x7, s1, s);
    } else {
        double sg = neg_cos ? -1.0 : 1.0;
        double x4 = x2 * x2;
        double c2 = RL_MADD(x2, sg * RL_C4, sg * RL_C3);
        double c1 = RL_MADD(x2, sg * RL_C1, sg * RL_C0);
        double x6 = x4 * x2;
        double c = RL_MADD(x4, sg * RL_C2, c1);
        return (float)RL_MADD(x6, c2, c);
    }
}

static inline uint32_t rl_abstop12(float x) { return (rl_asuint(x) >> 20) & 0x7ff; }

static inline double rl_reduce_fast(double x, int *np) {
    double r = x * RL_HPI_INV;
    int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
#if REF_LIBM_FMA
    return fma(-(double)n, RL_HPI, x);
#else
    return x - n * RL_HPI;
#endif
}

/* valid for |y| < 120 (the only range the reference feeds: angles in [0, 2pi)) */
static inline float rl_sinf(float y) {
    double x = y;
    if (rl_abstop12(y) < rl_abstop12(0x1.921FB6p-1f)) {
        if (rl_abstop12(y) < rl_abstop12(0x1p-12f)) return y;
        return rl_sinf_poly(x, x * x, 0, 0);
    }
    int n;
    x = rl_reduce_fast(x, &n);
    double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return rl_sinf_poly(x * s, x * x, (n & 2) != 0, n);
}

static inline float rl_cosf(float y) {
    double x = y;
    if (rl_abstop12(y) < rl_abstop12(0x1.921FB6p-1f)) {
        if (rl_abstop12(y) < rl_abstop12(0x1p-12f)) return 1.0f;
        return rl_sinf_poly(x, x * x, 0, 1);
    }
    int n;
    x = rl_reduce_fast(x, &n);
    double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return rl_sinf_poly(x * s, x * x, (n & 2) != 0, n ^ 1);
}

/* ---- fdlibm float atanf / atan2f (glibc s_atanf.c / e_atan2f.c) ---- */
static const float rl_atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
static const float rl_atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
static const float rl_aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
                                9.0908870101e-02f, -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f,
                                4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};

static inline float rl_atanf(float x) {
    float w, s1, s2, z;
    int32_t hx = (int32_t)rl_asuint(x), ix = hx & 0x7fffffff, id;
    if (ix >= 0x4c000000) { /* |x| >= 2^25 */
        if (ix > 0x7f800000) return x + x;
        if (hx > 0) return rl_atanhi[3] + rl_atanlo[3];
        return -rl_atanhi[3] - rl_atanlo[3];
    }
    if (ix < 0x3ee00000) { /* |x| < 0.4375 */
        if (ix < 0x31000000) return x; /* |x| < 2^-29 */
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {
            if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else { id = 3; x = -1.0f / x; }
        }
    }
    z = x * x;
    w = z * z;
    s1 = z * (rl_aT[0] + w * (rl_aT[2] + w * (rl_aT[4] + w * (rl_aT[6] + w * (rl_aT[8] + w * rl_aT[10])))));
    s2 = w * (rl_aT[1] + w * (rl_aT[3] + w * (rl_aT[5] + w * (rl_aT[7] + w * rl_aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    z = rl_atanhi[id] - ((x * (s1 + s2) - rl_atanlo[id]) - x);
    return (hx < 0) ? -z : z;
}

static inline float rl_atan2f(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi_o_4 = 7.8539818525e-01f,
                pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    float z;
    int32_t hx = (int32_t)rl_asuint(x), hy = (int32_t)rl_asuint(y);
    int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff, k, m;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return rl_atanf(y);
    m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        switch (m) {
        case 0: case 1: return y;
        case 2: return pi + tiny;
        default: return -pi - tiny;
        }
    }
    if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
            case 0: return pi_o_4 + tiny;
            case 1: return -pi_o_4 - tiny;
            case 2: return 3.0f * pi_o_4 + tiny;
            default: return -3.0f * pi_o_4 - tiny;
            }
        } else {
            switch (m) {
            case 0: return 0.0f;
            case 1: return -0.0f;
            case 2: return pi + tiny;
            default: return -pi - tiny;
            }
        }
    }
    if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    k = (iy - ix) >> 23;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = rl_atanf(fabsf(y / x));
    switch (m) {
    case 0: return z;
    case 1: return rl_asfloat(rl_asuint(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}
#endif

// cv_b200/csrc/device_libm.cuh -- bit-exact device versions of the host libm routines the
// reference reaches through Rust std (f32::sin / cos / atan2 -> glibc 2.39 sinf/cosf/atan2f on
// x86-64 linux-gnu): akaze/src/descriptors.rs:70-71, akaze/src/scale_space_extrema.rs:242.
//
// CUDA's own sinf/cosf/atan2f are <=2 ulp but not bit-identical to glibc, and the angle feeds
// round(sample_x) and hence descriptor bits, so the glibc algorithms are restated here:
//   sinf/cosf     : Arm-optimized-routines double-precision polynomial, in the contraction pattern
//                   of glibc's -mfma ifunc variant (explicit fma()).
//   atanf/atan2f  : fdlibm single-precision code, no contraction.
// This translation unit MUST be compiled with -fmad=false so that no other multiply-add is fused.
#pragma once
#include <stdint.h>

namespace dlm {

__device__ __forceinline__ float poly_sincos(double x, double x2, bool neg_cos, int n) {
    const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
                 C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double s1 = fma(x2, S3, S2);
        double x7 = x3 * x2;
        double s = fma(x3, S1, x);
        return (float)fma(x7, s1, s);
    } else {
        double sg = neg_cos ? -1.0 : 1.0;
        double x4 = x2 * x2;
        double c2 = fma(x2, sg * C4, sg * C3);
        double c1 = fma(x2, sg * C1, sg * C0);
        double x6 = x4 * x2;
        double c = fma(x4, sg * C2, c1);
        return (float)fma(x6, c2, c);
    }
}

__device__ __forceinline__ uint32_t abstop12(float x) { return (__float_as_uint(x) >> 20) & 0x7ff; }

__device__ __forceinline__ double reduce_fast(double x, int *np) {
    const double HPI_INV = 0x1.45F306DC9C883p+23, HPI = 0x1.921FB54442D18p0;
    double r = x * HPI_INV;
    int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return fma(-(double)n, HPI, x);
}

// valid for |y| < 120 (callers pass angles in [0, 2pi))
__device__ __forceinline__ float sinf_glibc(float y) {
    double x = (double)y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        if (abstop12(y) < abstop12(0x1p-12f)) return y;
        return poly_sincos(x, x * x, false, 0);
    }
    int n;
    x = reduce_fast(x, &n);
    double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return poly_sincos(x * s, x * x, (n & 2) != 0, n);
}

__device__ __forceinline__ float cosf_glibc(float y) {
    double x = (double)y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
        return poly_sincos(x, x * x, false, 1);
    }
    int n;
    x = reduce_fast(x, &n);
    double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return poly_sincos(x * s, x * x, (n & 2) != 0, n ^ 1);
}

__device__ __forceinline__ float atanf_glibc(float x) {
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
                aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
                aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
    int32_t hx = (int32_t)__float_as_uint(x), ix = hx & 0x7fffffff, id;
    if (ix >= 0x4c000000) {
        if (ix > 0x7f800000) return x + x;
        if (hx > 0) return atanhi[3] + atanlo[3];
        return -atanhi[3] - atanlo[3];
    }
    if (ix < 0x3ee00000) {
        if (ix < 0x31000000) return x;
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
    float z = x * x;
    float w = z * z;
    float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0) return x - x * (s1 + s2);
    float hi = id == 0 ? atanhi[0] : id == 1 ? atanhi[1] : id == 2 ? atanhi[2] : atanhi[3];
    float lo = id == 0 ? atanlo[0] : id == 1 ? atanlo[1] : id == 2 ? atanlo[2] : atanlo[3];
    z = hi - ((x * (s1 + s2) - lo) - x);
    return (hx < 0) ? -z : z;
}

__device__ __forceinline__ float atan2f_glibc(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi_o_4 = 7.8539818525e-01f, pi = 3.1415927410e+00f,
                pi_lo = -8.7422776573e-08f;
    float z;
    int32_t hx = (int32_t)__float_as_uint(x), hy = (int32_t)__float_as_uint(y);
    int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return atanf_glibc(y);
    int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        if (m < 2) return y;
        return m == 2 ? pi + tiny : -pi - tiny;
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
    int k = (iy - ix) >> 23;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = atanf_glibc(fabsf(y / x));
    switch (m) {
    case 0: return z;
    case 1: return __uint_as_float(__float_as_uint(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}

// scale_space_extrema.rs:242  (y.atan2(x) + 2.*PI).rem_euclid(2.*PI); fmodf is exact on both sides.
__device__ __forceinline__ float fast_atan2_equiv(float y, float x) {
    const float two_pi = 2.0f * 3.14159265358979323846f;
    float v = atan2f_glibc(y, x) + two_pi;
    float r = fmodf(v, two_pi);
    if (r < 0.0f) r = r + fabsf(two_pi);
    return r;
}

}  // namespace dlm

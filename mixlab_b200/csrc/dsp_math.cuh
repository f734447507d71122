// dsp_math.cuh -- scalar f64 building blocks shared by the audio kernels.
//
// Everything here is written with IEEE basic operations and explicit fma() only, so the same
// source gives bit-identical results on the device and on a host compiler (tests/test_dsp_math.py
// compiles it for the host and measures it against glibc's sin, which is what Rust's f64::sin
// resolves to on the reference's Linux target).
//
// The library is compiled with -fmad=false: a*b+c written with separate operators stays two
// roundings, as in the reference (rustc never contracts).  fma() below is always a single rounding.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MXL_HD __host__ __device__ __forceinline__
#else
#define MXL_HD static inline
#endif

namespace mxl {

// std::f64::consts::PI
constexpr double kPi = 3.14159265358979323846264338327950288;

// pi split into three doubles: P1 = RN(pi), P2 = RN(pi - P1), P3 = RN(pi - P1 - P2)
constexpr double kPi_1 = 3.141592653589793;           // 0x400921fb54442d18
constexpr double kPi_2 = 1.2246467991473532e-16;      // 0x3ca1a62633145c07
constexpr double kPi_3 = -2.9947698097183397e-33;     // 0xb92f1976b7ed8fbc
constexpr double kOneOverPi = 0.3183098861837907;     // 0x3fd45f306dc9c883

// Taylor coefficients of sin(r) = r + r z (T1 + z (T2 + ... z T10)), z = r^2: -1/3!, 1/5!, ... 1/21!.
// On |r| <= pi/2 the truncation error r^23/23! stays below 1.3e-18.
constexpr double kT1 = -1.66666666666666666667e-01, kT2 = 8.33333333333333333333e-03,
                 kT3 = -1.98412698412698412698e-04, kT4 = 2.75573192239858906526e-06,
                 kT5 = -2.50521083854417187751e-08, kT6 = 1.60590438368216145994e-10,
                 kT7 = -7.64716373181981647590e-13, kT8 = 2.81145725434552076320e-15,
                 kT9 = -8.22063524662432971696e-18, kT10 = 1.95729410633912612308e-20;

// low / high 32 bits of a double's bit pattern
MXL_HD uint32_t low_word(double t)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(t);
#else
    uint64_t u;
    memcpy(&u, &t, sizeof u);
    return (uint32_t)u;
#endif
}
MXL_HD uint32_t high_word(double t)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2hiint(t);
#else
    uint64_t u;
    memcpy(&u, &t, sizeof u);
    return (uint32_t)(u >> 32);
#endif
}
MXL_HD double flip_sign_if(double v, uint32_t odd)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(v) ^ (int)(odd << 31), __double2loint(v));
#else
    uint64_t u;
    memcpy(&u, &v, sizeof u);
    u ^= (uint64_t)(odd & 1u) << 63;
    memcpy(&v, &u, sizeof u);
    return v;
#endif
}

// sin(x) for 0 < |x| < 2^45 with 1-2 ulp error, branch-free.
//
// Why not CUDA's sin(): its fast path ends at |x| = 105615 and the Payne-Hanek slow path behind it
// spills to local memory.  Oscillator phases 2*pi*f*t/SR pass that bound after seconds of audio
// (oscillator.rs:25-27,74-75), so the common case would be the slow path.  With FMA a three-term
// Cody-Waite reduction stays accurate far beyond that: the first step x - q*P1 is one rounding of
// an exact difference no larger than pi/2, and the P2, P3 terms only add relative errors of 2^-53.
//
// Cost matters: these kernels would be HBM-bound but for the f64 work (64 FP64 lanes/clk/SM) and
// plain instruction issue, so
//   * q = round(x / pi) comes from adding 1.5*2^52 (the sum's low word is q mod 2^32) -- no
//     round-to-integer or float->int conversion, which run on the 16-lane XU pipe;
//   * the reduction is modulo pi, not pi/2: sin(x) = (-1)^q sin(r), |r| <= pi/2, one odd polynomial
//     in a single Horner chain -- no second (cosine) kernel and no per-quadrant selects; the sign
//     is one XOR on the high word.
constexpr double kRoundMagic = 6755399441055744.0;    // 1.5 * 2^52

// On the device the constants live in constant memory: FP64 instructions take a constant-bank operand
// directly, while a literal costs two 32-bit moves per use (measured: a third of the oscillator
// kernel's issue slots went into re-materialising literals, and the kernel was issue-bound).
#if defined(__CUDACC__)
static __constant__ double c_sin_tab[15] = {kOneOverPi, kRoundMagic, kPi_1, kPi_2, kPi_3,
                                            kT1, kT2, kT3, kT4, kT5, kT6, kT7, kT8, kT9, kT10};
#endif

MXL_HD double sin_reduced(double x)
{
#if defined(__CUDA_ARCH__)
    const double* c = c_sin_tab;
#else
    const double c[15] = {kOneOverPi, kRoundMagic, kPi_1, kPi_2, kPi_3, kT1, kT2, kT3, kT4, kT5, kT6, kT7, kT8, kT9, kT10};
#endif
    const double t = fma(x, c[0], c[1]);
    const double q = t - c[1];
    const uint32_t n = low_word(t);
    double r = fma(-q, c[2], x);
    r = fma(-q, c[3], r);
    r = fma(-q, c[4], r);
    const double z = r * r;
    double p = fma(c[14], z, c[13]);
    p = fma(p, z, c[12]);
    p = fma(p, z, c[11]);
    p = fma(p, z, c[10]);
    p = fma(p, z, c[9]);
    p = fma(p, z, c[8]);
    p = fma(p, z, c[7]);
    p = fma(p, z, c[6]);
    p = fma(p, z, c[5]);
    const double res = fma(r * z, p, r);
    return flip_sign_if(res, n & 1u);
}

MXL_HD double sin_f64(double x);

// Four arguments at a time, step by step: each constant is fetched once for the four and the four Horner
// chains interleave (the FP64 pipe has an 8-cycle dependent latency).  Same operations per argument as
// sin_reduced, so the results are identical.
MXL_HD void sin_reduced4(const double x[4], double out[4])
{
#if defined(__CUDA_ARCH__)
    const double* c = c_sin_tab;
#else
    const double c[15] = {kOneOverPi, kRoundMagic, kPi_1, kPi_2, kPi_3, kT1, kT2, kT3, kT4, kT5, kT6, kT7, kT8, kT9, kT10};
#endif
    double q[4], r[4], z[4], p[4];
    uint32_t n[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; i++) {
        const double t = fma(x[i], c[0], c[1]);
        q[i] = t - c[1];
        n[i] = low_word(t);
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; i++) r[i] = fma(-q[i], c[2], x[i]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; i++) r[i] = fma(-q[i], c[3], r[i]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; i++) { r[i] = fma(-q[i], c[4], r[i]); z[i] = r[i] * r[i]; p[i] = fma(c[14], z[i], c[13]); }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 12; k >= 5; k--) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 4; i++) p[i] = fma(p[i], z[i], c[k]);
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; i++) out[i] = flip_sign_if(fma(r[i] * z[i], p[i], r[i]), n[i] & 1u);
}

// sin of four arguments; arguments outside the fast range (or +-0) take the scalar path
MXL_HD void sin_f64x4(const double x[4], double out[4])
{
    bool fast = true;
    for (int i = 0; i < 4; i++) {
        const uint32_t hi = high_word(x[i]) & 0x7fffffffu;
        fast = fast && hi < 0x42c00000u && (hi | low_word(x[i])) != 0u;
    }
    if (fast) { sin_reduced4(x, out); return; }
    for (int i = 0; i < 4; i++) out[i] = sin_f64(x[i]);
}

MXL_HD double sin_f64(double x)
{
    const uint32_t hi = high_word(x) & 0x7fffffffu;     // integer tests: keep them off the FP64 pipe
    if (hi >= 0x42c00000u)                              // |x| >= 2^45, inf or NaN
        return sin(x);
    if ((hi | low_word(x)) == 0u) return x;             // sin(-0.0) = -0.0: Square takes the sign BIT (oscillator.rs:15-23)
    return sin_reduced(x);
}

// oscillator.rs:15-23: `is_sign_positive` tests the sign bit, so -0.0 -> -1.0, NaN by its sign bit.
MXL_HD double sign_bit_f64(double v) { return signbit(v) ? -1.0 : 1.0; }
// oscillator.rs:25-27.  n * 2.0 is exact, so (n * 2.0) * PI and n * (2.0 * PI) round the same real number
constexpr double kTwoPi = 2.0 * kPi;
MXL_HD double wave_sine(double n) { return sin_f64(n * kTwoPi); }
// oscillator.rs:30-32
MXL_HD double wave_saw(double n) { return 2.0 * (n - floor(0.5 + n)); }
// oscillator.rs:35-37
MXL_HD double wave_triangle(double n) { return 2.0 * fabs(wave_saw(n)) - 1.0; }

// Correctly rounded a / b for a fixed positive divisor b with a precomputed y = RN(1/b)
// (Markstein): q0 = RN(a*y); r = a - q0*b exactly (fma); q = RN(q0 + r*y).  Valid when neither
// a/b nor the intermediate products overflow/underflow, which holds for sample indices below 2^53
// divided by a sample rate.  tests/test_dsp_math.py checks it against `/` exhaustively over ranges.
MXL_HD double div_by_const(double a, double b, double inv_b)
{
    double q0 = a * inv_b;
    double r = fma(-q0, b, a);
    return fma(r, inv_b, q0);
}

// envelope.rs:20-28
MXL_HD double clamp01(double x) { return x > 1.0 ? 1.0 : (x < 0.0 ? 0.0 : x); }

}  // namespace mxl

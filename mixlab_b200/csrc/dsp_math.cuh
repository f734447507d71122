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

#if defined(__CUDACC__)
#define MXL_HD __host__ __device__ __forceinline__
#else
#define MXL_HD static inline
#endif

namespace mxl {

// std::f64::consts::PI
constexpr double kPi = 3.14159265358979323846264338327950288;

// pi/2 split into three doubles: P1 = RN(pi/2), P2 = RN(pi/2 - P1), P3 = RN(pi/2 - P1 - P2)
constexpr double kPio2_1 = 1.5707963267948966;        // 0x3ff921fb54442d18
constexpr double kPio2_2 = 6.123233995736766e-17;     // 0x3c91a62633145c07
constexpr double kPio2_3 = -1.4973849048591698e-33;   // 0xb91f1976b7ed8fbc
constexpr double kTwoOverPi = 0.6366197723675814;     // 0x3fe45f306dc9c883

// Minimax coefficients of the classic fdlibm k_sin.c / k_cos.c kernels on [-pi/4, pi/4]
// (public constants; relative error below 2^-57).
constexpr double kS1 = -1.66666666666666324348e-01, kS2 = 8.33333333332248946124e-03,
                 kS3 = -1.98412698298579493134e-04, kS4 = 2.75573137070700676789e-06,
                 kS5 = -2.50507602534068634195e-08, kS6 = 1.58969099521155010221e-10;
constexpr double kC1 = 4.16666666666666019037e-02, kC2 = -1.38888888888741095749e-03,
                 kC3 = 2.48015872894767294178e-05, kC4 = -2.75573143513906633035e-07,
                 kC5 = 2.08757232129817482790e-09, kC6 = -1.13596475577881948265e-11;

// sin(x) for |x| < 2^45 with about 1 ulp error, branch-free apart from the range guard.
//
// Why not CUDA's sin(): its fast path ends at |x| = 105615 and the Payne-Hanek slow path behind it
// spills to local memory.  Oscillator phases 2*pi*f*t/SR pass that bound after seconds of audio
// (oscillator.rs:25-27,74-75), so the common case would be the slow path.  With FMA a three-term
// Cody-Waite reduction stays exact far beyond that: x - q*P1 is exactly representable for
// |x| >= pi/4 (both are multiples of 2^-52 and the difference is below 1), and the P2, P3 terms
// only add relative rounding errors of 2^-53.
MXL_HD double sin_reduced(double x, double* cos_out)
{
    double q = rint(x * kTwoOverPi);
    double r = fma(-q, kPio2_1, x);
    r = fma(-q, kPio2_2, r);
    r = fma(-q, kPio2_3, r);
    double z = r * r;
    // sin kernel
    double w = z * z;
    double ps = fma(z, fma(z, kS4, kS3), kS2) + z * w * fma(z, kS6, kS5);
    double v = z * r;
    double s = fma(v, fma(z, ps, kS1), r);
    // cos kernel
    double pc = z * fma(z, fma(z, kC3, kC2), kC1) + (w * w) * fma(z, fma(z, kC6, kC5), kC4);
    double hz = 0.5 * z;
    double wc = 1.0 - hz;
    double c = wc + (((1.0 - wc) - hz) + z * pc);
    // quadrant
    long long n = (long long)q;
    double rs = (n & 1) ? c : s;
    double rc = (n & 1) ? s : c;
    if (n & 2) rs = -rs;
    if ((n + 1) & 2) rc = -rc;
    if (cos_out) *cos_out = rc;
    return rs;
}

MXL_HD double sin_f64(double x)
{
    if (!(fabs(x) < 35184372088832.0))   // 2^45, also catches NaN / inf
        return sin(x);
    if (x == 0.0) return x;               // sin(-0.0) = -0.0: Square takes the sign BIT (oscillator.rs:15-23)
    return sin_reduced(x, nullptr);
}

// oscillator.rs:15-23: `is_sign_positive` tests the sign bit, so -0.0 -> -1.0, NaN by its sign bit.
MXL_HD double sign_bit_f64(double v) { return signbit(v) ? -1.0 : 1.0; }
// oscillator.rs:25-27
MXL_HD double wave_sine(double n) { return sin_f64(n * 2.0 * kPi); }
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

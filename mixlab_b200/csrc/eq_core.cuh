// eq_core.cuh -- EqThree sample arithmetic (src/module/eq_three.rs:66-86, LowPass::pump 121-128),
// shared by the device kernel (eq_stream.cu) and a host build (tests/host_math) that checks the
// re-scheduled form against the sequential one bit for bit.
//
// The reference advances two 4-pole cascades sample by sample; pole k of sample n needs pole k-1 of
// sample n and pole k of sample n-1, a 13-operation dependent chain per sample (104 cycles on the
// B200 FP64 pipe, 8 cycles per dependent op).  eq_run_chunk_skewed walks the same dependency
// grid along its anti-diagonals: iteration n computes l0[n], l1[n-1], l2[n-2], l3[n-3] (and the
// h-cascade likewise), all from values of iteration n-1, so the eight pole updates of an
// iteration are independent and the loop-carried chain is 4 operations.  Every operation, its
// operands and its order are the reference's -- only the interleaving across samples differs --
// so the results are bit-identical (tests/test_eq_core.py).
#pragma once

#include "dsp_math.cuh"

namespace mxl {

constexpr double kEqVsa = 1.0 / 4294967295.0;   // eq_three.rs:11

struct EqPoles { double l0, l1, l2, l3, h0, h1, h2, h3; };
struct EqGains { double cl, ch, g_lo, g_mid, g_hi; };
struct EqF4 { float x, y, z, w; };

// eq_three.rs:66-86, sequential form.  hist = history[0..2].
MXL_HD float eq_step_seq(EqPoles& p, double hist[3], float xin, const EqGains& g)
{
    const double s = (double)xin;
    p.l0 = p.l0 + (g.cl * (s - p.l0) + kEqVsa);
    p.l1 = p.l1 + g.cl * (p.l0 - p.l1);
    p.l2 = p.l2 + g.cl * (p.l1 - p.l2);
    p.l3 = p.l3 + g.cl * (p.l2 - p.l3);
    p.h0 = p.h0 + (g.ch * (s - p.h0) + kEqVsa);
    p.h1 = p.h1 + g.ch * (p.h0 - p.h1);
    p.h2 = p.h2 + g.ch * (p.h1 - p.h2);
    p.h3 = p.h3 + g.ch * (p.h2 - p.h3);
    double lo = p.l3;
    double hi = hist[0] - p.h3;
    double mid = hist[0] - (hi + lo);
    hist[0] = hist[1]; hist[1] = hist[2]; hist[2] = s;
    lo = lo * g.g_lo;
    mid = mid * g.g_mid;
    hi = hi * g.g_hi;
    return (float)(lo + mid + hi);
}

// One anti-diagonal: pole k advances iff Ak.  Right-hand sides use the values before the iteration.
template <bool A0, bool A1, bool A2, bool A3>
MXL_HD void eq_skew_iter(EqPoles& p, double s, double cl, double ch)
{
    const EqPoles q = p;
    if (A3) { p.l3 = q.l3 + cl * (q.l2 - q.l3); p.h3 = q.h3 + ch * (q.h2 - q.h3); }
    if (A2) { p.l2 = q.l2 + cl * (q.l1 - q.l2); p.h2 = q.h2 + ch * (q.h1 - q.h2); }
    if (A1) { p.l1 = q.l1 + cl * (q.l0 - q.l1); p.h1 = q.h1 + ch * (q.h0 - q.h1); }
    if (A0) { p.l0 = q.l0 + (cl * (s - q.l0) + kEqVsa); p.h0 = q.h0 + (ch * (s - q.h0) + kEqVsa); }
}

// Band split + gains of the sample whose l3/h3 were just completed; x0 = history[0] of that sample.
MXL_HD float eq_bands(double l3, double h3, double x0, const EqGains& g)
{
    double lo = l3;
    double hi = x0 - h3;
    double mid = x0 - (hi + lo);
    lo = lo * g.g_lo;
    mid = mid * g.g_mid;
    hi = hi * g.g_hi;
    return (float)(lo + mid + hi);
}

// A full chunk of LC samples (LC % 4 == 0, LC >= 8).  `io` gives the chunk as LC/4 vectors:
// EqF4 io.load(v), void io.store(v, EqF4); outputs overwrite the inputs in place, vector v is stored
// only after vector v+1 has been loaded.  p: poles before the chunk -> after; hist likewise.
// ROLL: the steady-state loop stays a loop of four samples per trip (~190 instructions) instead of LC / 4 unrolled copies.
// Unrolled is faster once several CTAs share an SM and keep its instruction cache warm (the stand-alone EqThree kernel on
// long lines: 0.125 vs 0.154 ms per 2^25 samples); rolled is faster for a CTA that runs the code once, alone (a live
// one-tick call: the unrolled body stalled its warps on instruction fetch for 5.8 of every 6.8 issue slots).
template <int LC, bool ROLL = false, class Io>
MXL_HD void eq_run_chunk_skewed(EqPoles& p, double hist[3], Io& io, const EqGains& g)
{
    static_assert(LC % 4 == 0 && LC >= 8, "chunk length");
    // d[k] = input of iteration n-1-k (as f64): d[0..2] = the three samples before the chunk
    double d0 = hist[2], d1 = hist[1], d2 = hist[0], d3 = 0.0, d4 = 0.0, d5 = 0.0;
    const double cl = g.cl, ch = g.ch;
    float carry = 0.f;                                   // output of sample 4v, waiting for 4v+1..4v+3
#define MXL_EQ_SHIFT(sv) do { d5 = d4; d4 = d3; d3 = d2; d2 = d1; d1 = d0; d0 = (sv); } while (0)
    // iteration n (input s_n) completes sample n-3, whose history[0] is s_(n-6) = d5 before the shift
    {   // vector 0: pipeline fill, iteration 3 completes sample 0
        const EqF4 x = io.load(0);
        double s;
        s = (double)x.x; eq_skew_iter<true, false, false, false>(p, s, cl, ch); MXL_EQ_SHIFT(s);
        s = (double)x.y; eq_skew_iter<true, true, false, false>(p, s, cl, ch); MXL_EQ_SHIFT(s);
        s = (double)x.z; eq_skew_iter<true, true, true, false>(p, s, cl, ch); MXL_EQ_SHIFT(s);
        s = (double)x.w; eq_skew_iter<true, true, true, true>(p, s, cl, ch);
        carry = eq_bands(p.l3, p.h3, d5, g); MXL_EQ_SHIFT(s);
    }
#if defined(__CUDACC__)
#pragma unroll (ROLL ? 1 : LC / 4)
#endif
    for (int v = 1; v < LC / 4; v++) {
        const EqF4 x = io.load(v);
        EqF4 y;
        double s;
        y.x = carry;
        s = (double)x.x; eq_skew_iter<true, true, true, true>(p, s, cl, ch); y.y = eq_bands(p.l3, p.h3, d5, g); MXL_EQ_SHIFT(s);
        s = (double)x.y; eq_skew_iter<true, true, true, true>(p, s, cl, ch); y.z = eq_bands(p.l3, p.h3, d5, g); MXL_EQ_SHIFT(s);
        s = (double)x.z; eq_skew_iter<true, true, true, true>(p, s, cl, ch); y.w = eq_bands(p.l3, p.h3, d5, g); MXL_EQ_SHIFT(s);
        io.store(v - 1, y);
        s = (double)x.w; eq_skew_iter<true, true, true, true>(p, s, cl, ch); carry = eq_bands(p.l3, p.h3, d5, g); MXL_EQ_SHIFT(s);
    }
    {   // drain: samples LC-3 .. LC-1
        EqF4 y;
        y.x = carry;
        eq_skew_iter<false, true, true, true>(p, 0.0, cl, ch); y.y = eq_bands(p.l3, p.h3, d5, g); MXL_EQ_SHIFT(0.0);
        eq_skew_iter<false, false, true, true>(p, 0.0, cl, ch); y.z = eq_bands(p.l3, p.h3, d5, g); MXL_EQ_SHIFT(0.0);
        eq_skew_iter<false, false, false, true>(p, 0.0, cl, ch); y.w = eq_bands(p.l3, p.h3, d5, g); MXL_EQ_SHIFT(0.0);
        io.store(LC / 4 - 1, y);
    }
#undef MXL_EQ_SHIFT
    // three drain shifts pushed the last real inputs to d3 (s_(LC-1)), d4, d5 (s_(LC-3))
    hist[0] = d5; hist[1] = d4; hist[2] = d3;
}

// The chunk's contribution to the pole state at its end for an all-zero start state:
// acc[e] = K[e] + sum_j x_j * V[j][e].  V[j] = response at the chunk end to a unit input at j, K = the
// VSA terms; both come from the host plan.  Independent FMA chains -- accuracy 1e-16 relative is all
// the carry needs (its error is absorbed by the final `as f32`).
template <int LC, bool ROLL = false, class Io, class Tab>
MXL_HD void eq_zero_state_dot(const Io& io, const Tab& tab, double acc[8])
{
#if defined(__CUDACC__)
#pragma unroll (ROLL ? 1 : LC / 4)
#endif
    for (int v = 0; v < LC / 4; v++) {
        const EqF4 x = io.load(v);
        const double s[4] = {(double)x.x, (double)x.y, (double)x.z, (double)x.w};
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int e = 0; e < 8; e++) acc[e] = fma(s[q], tab.v(v * 4 + q, e), acc[e]);
    }
}

}  // namespace mxl

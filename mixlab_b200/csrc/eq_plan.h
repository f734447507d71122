// eq_plan.h -- host-side constants of the time-parallel EqThree kernel (eq_stream.cu), plain C++ so
// the same code is checked on the CPU (tests/test_eq_core.py).
//
// One 4-pole cascade (LowPass::pump, eq_three.rs:121-128) is the affine map
//     p' = M p + b s + k        M lower-triangular 4x4, s the input sample, k the VSA terms.
// For chunks of Lc samples:
//     A   = M^Lc                      homogeneous chunk transition
//     V_j = M^(Lc-1-j) b              pole state at the chunk end caused by a unit input at sample j
//     K   = sum_j M^j k               pole state at the chunk end caused by the VSA terms alone
// so the end state for a zero start state is  z = K + sum_j s_j V_j  (a dot product per pole), and the
// true start state of chunk c is  sum_(i>=1) A^(i-1) z_(c-i).  The cascades forget geometrically:
// halo = number of chunks after which |A^i| < 2^-75, below anything an f64 state can register.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

namespace mxl {

constexpr int kEqPlanMaxChunk = 64;
constexpr int kEqPlanLevels = 8;

struct EqStreamPlan {
    uint32_t lc = 0;
    uint32_t halo = 0;                 // chunks a CTA recomputes ahead of the ones it owns
    uint32_t lev_lo = 0, lev_hi = 0;   // scan levels that still matter per cascade: 2^lev >= forgetting length
    uint32_t back_lo = 0, back_hi = 0; // warps of 32 chunks before its own that a warp's start states still hear (<= 3)
    double c_lo = 0, c_hi = 0;
    double pow_lo[kEqPlanLevels][10];  // A^(2^d), packed lower-triangular row-major
    double pow_hi[kEqPlanLevels][10];
    double V[kEqPlanMaxChunk][8];      // [j][lo poles 0..3, hi poles 0..3]
    double K[8];
    // A^(l+1) for lane l = 0..31, entry-major [cascade][entry][lane] so a warp reads consecutive doubles
    double lane_pow[2][10][32];
    bool ok = false;
};

namespace eqplan {

inline void mat_mul(const long double X[4][4], const long double Y[4][4], long double Z[4][4])
{
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            long double s = 0;
            for (int q = 0; q < 4; q++) s += X[r][q] * Y[q][c];
            Z[r][c] = s;
        }
}

inline long double abs_max(const long double X[4][4])
{
    long double m = 0;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) m = fmaxl(m, fabsl(X[r][c]));
    return m;
}

// one pump() of a cascade in extended precision: p' = M p + b s + k
inline void pump(long double c, long double p[4], long double s, long double vsa)
{
    p[0] += c * (s - p[0]) + vsa;
    p[1] += c * (p[0] - p[1]);
    p[2] += c * (p[1] - p[2]);
    p[3] += c * (p[2] - p[3]);
}

inline void chunk_transition(long double c, uint32_t steps, long double A[4][4])
{
    for (int col = 0; col < 4; col++) {
        long double p[4] = {0, 0, 0, 0};
        p[col] = 1;
        for (uint32_t s = 0; s < steps; s++) pump(c, p, 0, 0);
        for (int r = 0; r < 4; r++) A[r][col] = p[r];
    }
}

inline void pack_tri(const long double X[4][4], double* out)
{
    int q = 0;
    for (int r = 0; r < 4; r++) for (int c = 0; c <= r; c++) out[q++] = (double)X[r][c];
}

// chunks until |A^i| < 2^-75; 0 if more than `limit`
inline uint32_t forgetting_chunks(const long double A[4][4], uint32_t limit)
{
    long double P[4][4], T[4][4];
    memcpy(P, A, sizeof P);
    const long double tiny = ldexpl(1.0L, -75);
    uint32_t i = 1;
    while (abs_max(P) >= tiny) {
        if (++i > limit) return 0;
        mat_mul(P, A, T);
        memcpy(P, T, sizeof T);
    }
    return i;
}

inline uint32_t levels_for(uint32_t chunks)
{
    uint32_t lev = 0;
    while ((1u << lev) < chunks) lev++;
    return lev;
}

}  // namespace eqplan

// LowPass::set_freq (eq_three.rs:117-119) in f64 exactly as the reference; FREQ_LO/HI eq_three.rs:8-9
inline void eq_coefficients(uint32_t sample_rate, double* c_lo, double* c_hi)
{
    const double pi = 3.14159265358979323846264338327950288;
    *c_lo = 2.0 * sin(pi * 420.0 / (double)sample_rate);
    *c_hi = 2.0 * sin(pi * 2700.0 / (double)sample_rate);
}

inline EqStreamPlan eq_stream_plan(uint32_t sample_rate, uint32_t lc, uint32_t max_halo)
{
    using namespace eqplan;
    EqStreamPlan pl;
    memset(pl.pow_lo, 0, sizeof pl.pow_lo); memset(pl.pow_hi, 0, sizeof pl.pow_hi);
    memset(pl.V, 0, sizeof pl.V); memset(pl.K, 0, sizeof pl.K); memset(pl.lane_pow, 0, sizeof pl.lane_pow);
    pl.lc = lc;
    if (lc < 8 || lc > (uint32_t)kEqPlanMaxChunk || (lc & 3)) return pl;
    eq_coefficients(sample_rate, &pl.c_lo, &pl.c_hi);
    const long double vsa = 1.0L / 4294967295.0L;
    const double cs[2] = {pl.c_lo, pl.c_hi};
    uint32_t forget[2];
    for (int f = 0; f < 2; f++) {
        const long double c = cs[f];
        long double A[4][4], P[4][4], T[4][4];
        chunk_transition(c, lc, A);
        forget[f] = forgetting_chunks(A, max_halo);
        if (forget[f] == 0) return pl;
        memcpy(P, A, sizeof P);
        for (int d = 0; d < kEqPlanLevels; d++) {
            pack_tri(P, f == 0 ? pl.pow_lo[d] : pl.pow_hi[d]);
            mat_mul(P, P, T);
            memcpy(P, T, sizeof T);
        }
        memcpy(P, A, sizeof P);
        for (int l = 0; l < 32; l++) {
            double packed[10];
            pack_tri(P, packed);
            for (int q = 0; q < 10; q++) pl.lane_pow[f][q][l] = packed[q];
            mat_mul(P, A, T);
            memcpy(P, T, sizeof T);
        }
        for (uint32_t j = 0; j < lc; j++) {
            long double p[4] = {0, 0, 0, 0};
            pump(c, p, 1.0L, 0);
            for (uint32_t s = j + 1; s < lc; s++) pump(c, p, 0, 0);
            for (int e = 0; e < 4; e++) pl.V[j][4 * f + e] = (double)p[e];
        }
        long double p[4] = {0, 0, 0, 0};
        for (uint32_t s = 0; s < lc; s++) pump(c, p, 0, vsa);
        for (int e = 0; e < 4; e++) pl.K[4 * f + e] = (double)p[e];
    }
    pl.halo = forget[0] > forget[1] ? forget[0] : forget[1];
    pl.lev_lo = levels_for(forget[0]);
    pl.lev_hi = levels_for(forget[1]);
    // a thread's inclusive value must reach forget-1 chunks back: lane + 32*back >= forget - 1
    pl.back_lo = (forget[0] - 1 + 31) / 32;
    pl.back_hi = (forget[1] - 1 + 31) / 32;
    pl.ok = pl.lev_lo <= (uint32_t)kEqPlanLevels && pl.lev_hi <= (uint32_t)kEqPlanLevels && pl.back_lo <= 3 && pl.back_hi <= 3;
    return pl;
}

}  // namespace mxl

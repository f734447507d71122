// eq_stream.cuh -- device pieces of the time-parallel EqThree scheme shared by eq_stream_kernel
// (eq_stream.cu: the module on its own) and fused_voice_mix_kernel (fused_voice.cu: Oscillator -> EqThree ->
// StereoPanner -> Mixer -> Meter in one launch): the XOR-swizzled tile of 256 chunks, and the scan that
// turns the chunks' zero-state end vectors into their start states.  See eq_stream.cu for the scheme.
#pragma once

#include "eq_core.cuh"
#include "eq_plan.h"
#include "kernels.h"

namespace mxl {
namespace k {
namespace eqs {

constexpr int kT = kEqStreamThreads;

// Shared memory next to the tile: the scan's exchange area and the plan's tables (EqDevTables, brought in by
// load_tables with one round of coalesced loads per CTA).
// HASV = false: the zero-pass table V stays out of shared memory (eq_stream_kernel's long-call variant takes it as
// constant-bank operands from its kernel parameters: 4 KB less per CTA, which is the third resident CTA at LC = 64).
template <int LC, bool HASV = true>
struct alignas(16) EqShared {
    double2 agg[(kT / 32) * 4];                  // warp aggregates of the scan
    double2 fin[(kT / 32) * 4];                  // final value of every warp's lane 31
    double lane_pow[2 * 10 * 32];
    double pow_lo[8][10];
    double pow_hi[8][10];
    double K[8];
    double V[HASV ? LC : 1][8];
};

template <int LC, bool HASV>
__device__ __forceinline__ void load_tables(const EqDevTables* __restrict__ g, EqShared<LC, HASV>* s, int tid)
{
    if (HASV) for (int i = tid; i < LC * 8; i += kT) (&s->V[0][0])[i] = (&g->V[0][0])[i];
    for (int i = tid; i < 2 * 10 * 32; i += kT) s->lane_pow[i] = (&g->lane_pow[0][0][0])[i];
    for (int i = tid; i < 160; i += kT) (&s->pow_lo[0][0])[i] = (&g->pow_lo[0][0])[i];    // pow_lo and pow_hi are adjacent in both
    if (tid < 8) s->K[tid] = g->K[tid];
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src);

// The same with 16-byte cp.async copies (every table is a multiple of 16 bytes and 16-byte aligned in both structs):
// the caller overlaps them with work that does not need the tables, then cp.async.wait_all + __syncthreads.
template <int LC>
__device__ __forceinline__ void load_tables_async(const EqDevTables* __restrict__ g, EqShared<LC>* s, int tid)
{
    for (int i = tid; i < LC * 4; i += kT) cp_async16(&s->V[0][0] + 2 * i, &g->V[0][0] + 2 * i);
    for (int i = tid; i < 10 * 32; i += kT) cp_async16(s->lane_pow + 2 * i, &g->lane_pow[0][0][0] + 2 * i);
    for (int i = tid; i < 80; i += kT) cp_async16(&s->pow_lo[0][0] + 2 * i, &g->pow_lo[0][0] + 2 * i);   // pow_lo and pow_hi are adjacent in both
    if (tid < 4) cp_async16(s->K + 2 * tid, g->K + 2 * tid);
}

__device__ __forceinline__ int tri(int r, int c) { return r * (r + 1) / 2 + c; }

__device__ __forceinline__ void tri_apply(const double* A, const double x[4], double y[4])
{
#pragma unroll
    for (int r = 0; r < 4; r++) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c <= r; c++) acc = fma(A[tri(r, c)], x[c], acc);
        y[r] = acc;
    }
}

// 16-byte slot of (row r, vector v) in a tile of VPR vectors per row.  Eight consecutive rows must
// land in eight different 16-byte bank groups for a fixed v (thread-per-row float4 accesses are served
// eight lanes at a time), and a row's VPR vectors stay inside the row (coalesced staging).
template <int VPR>
__device__ __forceinline__ int slot_of(int r, int v)
{
    constexpr int kRowsPerLine = VPR >= 8 ? 1 : 8 / VPR;          // rows sharing one 128-byte bank line
    constexpr int kMask = VPR >= 8 ? 7 : VPR - 1;
    return r * VPR + (v ^ ((r / kRowsPerLine) & kMask));
}

template <int LC>
struct RowIo {
    float4* tile; int r;
    __device__ __forceinline__ EqF4 load(int v) const
    {
        const float4 x = tile[slot_of<LC / 4>(r, v)];
        return EqF4{x.x, x.y, x.z, x.w};
    }
    __device__ __forceinline__ void store(int v, EqF4 y) { tile[slot_of<LC / 4>(r, v)] = make_float4(y.x, y.y, y.z, y.w); }
};

template <int LC>
struct SharedTab {
    const EqShared<LC>* s;
    __device__ __forceinline__ double v(int j, int e) const { return s->V[j][e]; }
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ bool any_non_finite(const double v[8])
{
    uint32_t bad = 0;
#pragma unroll
    for (int e = 0; e < 8; e++) bad |= ((uint32_t)__double2hiint(v[e]) & 0x7ff00000u) == 0x7ff00000u ? 1u : 0u;
    return bad != 0;
}

// Scan.  In: v = this thread's zero-state end vector (chunk 0 of the call: A p_init + z_0; inactive chunks: 0).
// Inside a warp: Hillis-Steele with shuffles, v_i += A^(2^d) v_(i - 2^d), only the levels the cascade still
// hears.  Across warps: the inclusive value of a warp's last lane is its aggregate; the state entering warp w is
// P = agg[w-1] + A^32 agg[w-2] + A^64 agg[w-3] (as many terms as the cascade still hears), and lane l adds
// A^(l+1) P from a per-lane table.  Out: S = start state of this thread's chunk = inclusive value of the previous
// thread (meaningless for thread 0, which owns a chunk only when it is the call's first: that one starts from the
// module's stored state).  Two block barriers.
template <int LC, bool HASV>
__device__ __forceinline__ void scan_start_states(const EqStreamConsts& b, EqShared<LC, HASV>* sh, double v[8], double S[8], int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int d = 0; d < 5; d++) {
        const bool lo_live = d < (int)b.lev_lo, hi_live = d < (int)b.lev_hi;
        if (!lo_live && !hi_live) break;
        double o[8];
#pragma unroll
        for (int e = 0; e < 4; e++) o[e] = __shfl_up_sync(0xffffffffu, v[e], 1 << d);
        if (hi_live) {
#pragma unroll
            for (int e = 4; e < 8; e++) o[e] = __shfl_up_sync(0xffffffffu, v[e], 1 << d);
        }
        if (lane >= (1 << d)) {
            double y[4];
            if (lo_live) {
                tri_apply(sh->pow_lo[d], o, y);
#pragma unroll
                for (int e = 0; e < 4; e++) v[e] += y[e];
            }
            if (hi_live) {
                tri_apply(sh->pow_hi[d], o + 4, y);
#pragma unroll
                for (int e = 0; e < 4; e++) v[4 + e] += y[e];
            }
        }
    }
    double2* agg = sh->agg;                                      // [8 warps][4 double2]: warp aggregates
    double2* fin = sh->fin;                                      // [8 warps][4 double2]: final value of lane 31
    const double* lane_tab = sh->lane_pow;                       // [2][10][32]
    if (lane == 31) {
        agg[warp * 4 + 0] = make_double2(v[0], v[1]); agg[warp * 4 + 1] = make_double2(v[2], v[3]);
        agg[warp * 4 + 2] = make_double2(v[4], v[5]); agg[warp * 4 + 3] = make_double2(v[6], v[7]);
    }
    __syncthreads();
    if (warp > 0) {
        double P[8];
        {
            const double2 a0 = agg[(warp - 1) * 4 + 0], a1 = agg[(warp - 1) * 4 + 1];
            const double2 a2 = agg[(warp - 1) * 4 + 2], a3 = agg[(warp - 1) * 4 + 3];
            P[0] = a0.x; P[1] = a0.y; P[2] = a1.x; P[3] = a1.y; P[4] = a2.x; P[5] = a2.y; P[6] = a3.x; P[7] = a3.y;
        }
#pragma unroll
        for (int k = 1; k <= 2; k++) {                     // A^(32k) = pow[4 + k]
            if (warp - 1 - k < 0) break;
            const bool lo_live = k < (int)b.back_lo, hi_live = k < (int)b.back_hi;
            if (!lo_live && !hi_live) break;
            const double2* a = agg + (warp - 1 - k) * 4;
            double y[4];
            if (lo_live) {
                const double2 a0 = a[0], a1 = a[1];
                const double o[4] = {a0.x, a0.y, a1.x, a1.y};
                tri_apply(sh->pow_lo[4 + k], o, y);
#pragma unroll
                for (int e = 0; e < 4; e++) P[e] += y[e];
            }
            if (hi_live) {
                const double2 a2 = a[2], a3 = a[3];
                const double o[4] = {a2.x, a2.y, a3.x, a3.y};
                tri_apply(sh->pow_hi[4 + k], o, y);
#pragma unroll
                for (int e = 0; e < 4; e++) P[4 + e] += y[e];
            }
        }
        double Al[10], Ah[10];
#pragma unroll
        for (int q = 0; q < 10; q++) { Al[q] = lane_tab[q * 32 + lane]; Ah[q] = lane_tab[(10 + q) * 32 + lane]; }
        double y[4];
        tri_apply(Al, P, y);
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] += y[e];
        tri_apply(Ah, P + 4, y);
#pragma unroll
        for (int e = 0; e < 4; e++) v[4 + e] += y[e];
    }
    // start state of my chunk = inclusive value of the previous thread
#pragma unroll
    for (int e = 0; e < 8; e++) S[e] = __shfl_up_sync(0xffffffffu, v[e], 1);
    if (lane == 31) {
        fin[warp * 4 + 0] = make_double2(v[0], v[1]); fin[warp * 4 + 1] = make_double2(v[2], v[3]);
        fin[warp * 4 + 2] = make_double2(v[4], v[5]); fin[warp * 4 + 3] = make_double2(v[6], v[7]);
    }
    __syncthreads();
    if (lane == 0 && warp > 0) {
        const double2 a0 = fin[(warp - 1) * 4 + 0], a1 = fin[(warp - 1) * 4 + 1];
        const double2 a2 = fin[(warp - 1) * 4 + 2], a3 = fin[(warp - 1) * 4 + 3];
        S[0] = a0.x; S[1] = a0.y; S[2] = a1.x; S[3] = a1.y; S[4] = a2.x; S[5] = a2.y; S[6] = a3.x; S[7] = a3.y;
    }
}

}  // namespace eqs
}  // namespace k
}  // namespace mxl

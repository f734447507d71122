// eq_stream.cu -- EqThree (src/module/eq_three.rs:58-89,106-125) as one launch, parallel along time.
//
// The module is two 4-pole cascades of one-pole low-passes in f64, strictly serial in the
// reference.  Here the time axis is cut into chunks of LC samples, one thread per chunk, one CTA of
// 256 threads per 256 consecutive chunks of one instance:
//
//   1. stage     the CTA's 256*LC input samples go global -> shared with 16-byte cp.async, into a
//                tile whose 16-byte slots are XOR-swizzled so that both the coalesced staging
//                accesses and the thread-per-row accesses of steps 2 and 4 are bank-conflict free;
//   2. zero pass every thread forms the end state its chunk would reach from an all-zero start
//                state, z = K + sum_j s_j V_j: eight independent FMA chains per thread, operands V_j
//                straight from the constant bank (kernel parameters);
//   3. scan      start state of chunk i = sum_(k>=1) A^(k-1) z_(i-k): Hillis-Steele with warp shuffles
//                inside each warp (only the levels the cascade has not forgotten), then one hop
//                across warps through their aggregates -- two block barriers in all;
//   4. exact     every owned chunk is re-run from its start state with the reference's operations
//                in the reference's order (eq_core.cuh: skewed so that one warp per scheduler already
//                saturates the FP64 pipe), outputs overwrite the tile and leave coalesced.
//
// The first `halo` chunks of a CTA belong to the previous CTA's range and only feed the scan: after
// `halo` chunks the cascades have forgotten their start state to below 2^-75, so CTAs never
// communicate and the whole module is a single launch.  The chunk that starts the call runs from the
// module's stored state, exactly, so successive calls continue bit-exactly.
//
// Non-finite samples: in the reference a NaN or an infinity entering the poles stays there for ever (every
// later output is NaN, in this call and the next).  The carry above forgets by construction, so it would let
// chunks further than `halo` behind such a sample recover.  Every thread therefore checks its zero-pass
// vector; the first chunk with a non-finite carry is recorded (atomicMin), and the CTA of an instance that
// finishes last rewrites everything after that chunk as NaN, poles included -- inside the chunk itself the
// exact pass already did the right thing sample by sample.  Cost when nothing is wrong: one atomic per CTA.
//
// Roofline: 8 B/sample of line traffic against ~47 FP64 operations per sample (34 exact + 8 zero
// pass + ~4 scan + halo): at the measured 64 FP64 lanes/clk/SM (tools/pipe_rates.cu) the FP64 pipe
// allows 0.40 of the HBM copy peak, so this kernel is FP64-issue-bound, not HBM-bound.
#include "eq_core.cuh"
#include "eq_plan.h"
#include "kernels.h"

namespace mxl {
namespace k {

namespace {

constexpr int kT = kEqStreamThreads;
constexpr int kEqXchDoubles = 2 * (kT / 32) * 8 + 2 * 10 * 32;    // warp aggregates, final lane-31 values, lane table

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

struct ParamTab {
    const EqStreamBatch& b;
    __device__ __forceinline__ double v(int j, int e) const { return b.V[j][e]; }
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}

template <int LC>
__global__ void __launch_bounds__(kT) eq_stream_kernel(const __grid_constant__ EqStreamBatch b)
{
    constexpr int VPR = LC / 4;
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char eq_smem[];
    float4* tile = reinterpret_cast<float4*>(eq_smem);                              // [256][VPR], swizzled
    double* xch = reinterpret_cast<double*>(eq_smem + (size_t)kT * LC * sizeof(float));   // kEqXchDoubles
    const EqStreamInst& in = b.inst[blockIdx.y];
    const int tid = threadIdx.x;
    const int halo = (int)b.halo;
    const int U = kT - halo;
    const int64_t c0 = (int64_t)blockIdx.x * U - halo;        // chunk of thread 0
    const int64_t c = c0 + tid;
    const bool active = c >= 0 && c < (int64_t)b.n_chunks;

    // ---- 1. stage ----
    {
        const float* src = in.in;
        const bool vec_ok = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
#pragma unroll
        for (int it = 0; it < VPR; it++) {
            const int idx = it * kT + tid;
            const int r = idx / VPR, v = idx % VPR;
            const int64_t g = (c0 + r) * LC + v * 4;               // first sample of this vector
            float4* dst = tile + slot_of<VPR>(r, v);
            if (src && g >= 0 && g + 4 <= (int64_t)b.frames && vec_ok) {
                cp_async16(dst, src + g);
            } else {
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (src && g >= 0 && g < (int64_t)b.frames) {
                    x.x = src[g];
                    if (g + 1 < (int64_t)b.frames) x.y = src[g + 1];
                    if (g + 2 < (int64_t)b.frames) x.z = src[g + 2];
                    if (g + 3 < (int64_t)b.frames) x.w = src[g + 3];
                }
                *dst = x;
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();

    const double* st = in.state;
    RowIo<LC> row{tile, tid};

    // ---- 2. zero pass ----
    double v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = 0.0;
    if (active) {
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = b.K[e];
        ParamTab tab{b};
        eq_zero_state_dot<LC>(row, tab, v);
        if (c == 0) {                                      // v_0 = A p_init + z_0
            const double pl[4] = {st[0], st[1], st[2], st[3]}, ph[4] = {st[4], st[5], st[6], st[7]};
            double yl[4], yh[4];
            tri_apply(b.pow_lo[0], pl, yl);
            tri_apply(b.pow_hi[0], ph, yh);
#pragma unroll
            for (int e = 0; e < 4; e++) { v[e] += yl[e]; v[4 + e] += yh[e]; }
        }
    }

    // a non-finite carry (NaN / inf input in my chunk, or a non-finite stored state): remember the first such chunk
    if (active) {
        uint32_t bad = 0;
#pragma unroll
        for (int e = 0; e < 8; e++) bad |= ((uint32_t)__double2hiint(v[e]) & 0x7ff00000u) == 0x7ff00000u ? 1u : 0u;
        if (bad) atomicMin(in.poison, (uint32_t)c);
    }

    // ---- 3. scan.  Inside a warp: Hillis-Steele with shuffles, v_i += A^(2^d) v_(i - 2^d), only the
    //         levels the cascade still hears.  Across warps: the inclusive value of a warp's last lane
    //         is its aggregate; the state entering warp w is P = agg[w-1] + A^32 agg[w-2] + A^64 agg[w-3]
    //         (as many terms as the cascade still hears), and lane l adds A^(l+1) P from a per-lane
    //         table.  Two block barriers in all. ----
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
                tri_apply(b.pow_lo[d], o, y);
#pragma unroll
                for (int e = 0; e < 4; e++) v[e] += y[e];
            }
            if (hi_live) {
                tri_apply(b.pow_hi[d], o + 4, y);
#pragma unroll
                for (int e = 0; e < 4; e++) v[4 + e] += y[e];
            }
        }
    }
    double2* agg = reinterpret_cast<double2*>(xch);              // [8 warps][4 double2]: warp aggregates
    double2* fin = agg + (kT / 32) * 4;                          // [8 warps][4 double2]: final value of lane 31
    double* lane_tab = xch + 2 * (kT / 32) * 8;                  // [2][10][32]
    if (lane == 31) {
        agg[warp * 4 + 0] = make_double2(v[0], v[1]); agg[warp * 4 + 1] = make_double2(v[2], v[3]);
        agg[warp * 4 + 2] = make_double2(v[4], v[5]); agg[warp * 4 + 3] = make_double2(v[6], v[7]);
    }
    for (int i = tid; i < 2 * 10 * 32; i += kT) lane_tab[i] = b.lane_pow[i];
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
                tri_apply(b.pow_lo[4 + k], o, y);
#pragma unroll
                for (int e = 0; e < 4; e++) P[e] += y[e];
            }
            if (hi_live) {
                const double2 a2 = a[2], a3 = a[3];
                const double o[4] = {a2.x, a2.y, a3.x, a3.y};
                tri_apply(b.pow_hi[4 + k], o, y);
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
    double S[8];
#pragma unroll
    for (int e = 0; e < 8; e++) S[e] = __shfl_up_sync(0xffffffffu, v[e], 1);
    if (lane == 31) {
        fin[warp * 4 + 0] = make_double2(v[0], v[1]); fin[warp * 4 + 1] = make_double2(v[2], v[3]);
        fin[warp * 4 + 2] = make_double2(v[4], v[5]); fin[warp * 4 + 3] = make_double2(v[6], v[7]);
    }
    __syncthreads();

    // ---- 4. exact re-run of the owned chunks ----
    const bool owner = active && tid >= halo;
    EqPoles p;
    double hist[3];
    uint32_t count = 0;
    if (owner) {
        if (c == 0) {
            p = EqPoles{st[0], st[1], st[2], st[3], st[4], st[5], st[6], st[7]};
            hist[0] = st[8]; hist[1] = st[9]; hist[2] = st[10];
        } else {                                           // c > 0 and tid >= halo >= 1
            if (lane == 0) {
                const double2 a0 = fin[(warp - 1) * 4 + 0], a1 = fin[(warp - 1) * 4 + 1];
                const double2 a2 = fin[(warp - 1) * 4 + 2], a3 = fin[(warp - 1) * 4 + 3];
                S[0] = a0.x; S[1] = a0.y; S[2] = a1.x; S[3] = a1.y; S[4] = a2.x; S[5] = a2.y; S[6] = a3.x; S[7] = a3.y;
            }
            p = EqPoles{S[0], S[1], S[2], S[3], S[4], S[5], S[6], S[7]};
            const float4 prev = tile[slot_of<VPR>(tid - 1, VPR - 1)];      // last vector of the chunk before mine
            hist[0] = (double)prev.y; hist[1] = (double)prev.z; hist[2] = (double)prev.w;
        }
        const uint64_t s0 = (uint64_t)c * LC;
        count = (uint32_t)((s0 + LC <= b.frames) ? LC : (b.frames - s0));
    }
    __syncthreads();                                       // every history read precedes any overwrite
    if (owner) {
        const EqGains g{b.c_lo, b.c_hi, in.g_lo, in.g_mid, in.g_hi};
        if (count == LC) {
            eq_run_chunk_skewed<LC>(p, hist, row, g);
        } else {                                           // ragged end of the call: one thread, sequential form
            float* mine = reinterpret_cast<float*>(tile);
            for (uint32_t j = 0; j < count; j++) {
                float* cell = mine + slot_of<VPR>(tid, j >> 2) * 4 + (j & 3);
                *cell = eq_step_seq(p, hist, *cell, g);
            }
        }
        if (c + 1 == (int64_t)b.n_chunks) {                // state after this call
            double* so = in.state_out;
            so[0] = p.l0; so[1] = p.l1; so[2] = p.l2; so[3] = p.l3;
            so[4] = p.h0; so[5] = p.h1; so[6] = p.h2; so[7] = p.h3;
            so[8] = hist[0]; so[9] = hist[1]; so[10] = hist[2];
        }
    }
    __syncthreads();

    // ---- 5. coalesced store of the owned rows ----
    {
        float* dst = in.out;
        const bool vec_ok = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
        const int total = U * VPR;
        for (int idx = tid; idx < total; idx += kT) {
            const int r = halo + idx / VPR, vv = idx % VPR;
            const int64_t ci = c0 + r;
            if (ci < 0 || ci >= (int64_t)b.n_chunks) continue;
            const uint64_t g = (uint64_t)ci * LC + vv * 4;
            const float4 y = tile[slot_of<VPR>(r, vv)];
            if (g + 4 <= b.frames && vec_ok) {
                *reinterpret_cast<float4*>(dst + g) = y;
            } else {
                if (g < b.frames) dst[g] = y.x;
                if (g + 1 < b.frames) dst[g + 1] = y.y;
                if (g + 2 < b.frames) dst[g + 2] = y.z;
                if (g + 3 < b.frames) dst[g + 3] = y.w;
            }
        }
    }

    // ---- 6. poison fix-up by the CTA of this instance that finishes last ----
    __shared__ uint32_t s_first_bad;
    __threadfence();                                       // my outputs are visible before I count myself done
    __syncthreads();
    if (tid == 0) {
        uint32_t first_bad = 0xffffffffu;
        if (atomicAdd(in.poison + 1, 1u) == gridDim.x - 1) {
            first_bad = atomicExch(in.poison, 0xffffffffu) ;  // read and re-arm for the next launch
            in.poison[1] = 0u;
            __threadfence();
        }
        s_first_bad = first_bad;
    }
    __syncthreads();
    if (s_first_bad != 0xffffffffu) {
        const float nan = __int_as_float(0x7fc00000);
        for (uint64_t i = ((uint64_t)s_first_bad + 1) * LC + tid; i < b.frames; i += kT) in.out[i] = nan;
        if (tid == 0 && s_first_bad + 1 < b.n_chunks) {     // (in the last chunk the exact pass already left the true state)
            double* so = in.state_out;
#pragma unroll
            for (int e = 0; e < 8; e++) so[e] = __longlong_as_double(0x7ff8000000000000ll);
            for (int j = 0; j < 3; j++) {                  // history = the last three inputs, whatever they are
                const int64_t idx2 = (int64_t)b.frames - 3 + j;
                so[8 + j] = idx2 >= 0 ? (in.in ? (double)in.in[idx2] : 0.0) : st[8 + 3 + idx2];
            }
        }
    }
}

template <int LC>
int launch_lc(mxl_ctx* ctx, const EqStreamBatch& b)
{
    const size_t smem = (size_t)kT * LC * sizeof(float) + (size_t)kEqXchDoubles * sizeof(double);
    if (smem > 48 * 1024 && !(ctx->eq_stream_smem_set & (1u << (LC / 16)))) {
        MXL_CUDA(cudaFuncSetAttribute(eq_stream_kernel<LC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->eq_stream_smem_set |= 1u << (LC / 16);
    }
    const uint32_t U = kT - b.halo;
    dim3 grid((b.n_chunks + U - 1) / U, b.n);
    MXL_TIMED(ctx, "eq_stream_kernel");
    launch_chained(ctx, eq_stream_kernel<LC>, grid, dim3(kT), smem, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) MXL_FAIL(MXL_ERR_CUDA, "launch of eq_stream_kernel<%d> failed: %s", LC, cudaGetErrorString(e));
    ctx->launches++;
    return MXL_OK;
}

}  // namespace

int launch_eq_stream(mxl_ctx* ctx, const EqStreamBatch& b)
{
    if (!ctx || !ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_TRY(ctx->activate());
    if (b.n <= 0 || b.frames == 0) return MXL_OK;
    if (b.halo == 0 || b.halo > kT / 2 || b.lev_lo > (uint32_t)kEqPlanLevels || b.lev_hi > (uint32_t)kEqPlanLevels ||
        b.back_lo > 3 || b.back_hi > 3 || !b.lane_pow)
        MXL_FAIL(MXL_ERR_INVALID, "eq_stream_kernel: bad plan (chunk %u, halo %u, levels %u/%u)", b.chunk, b.halo, b.lev_lo, b.lev_hi);
    switch (b.chunk) {
    case 16: return launch_lc<16>(ctx, b);
    case 32: return launch_lc<32>(ctx, b);
    case 64: return launch_lc<64>(ctx, b);
    default: MXL_FAIL(MXL_ERR_INVALID, "eq_stream_kernel: unsupported chunk length %u", b.chunk);
    }
}

}  // namespace k
}  // namespace mxl

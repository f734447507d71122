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
// module's stored state, exactly.  Every other chunk starts from a CARRIED state (FMA dot products and a scan),
// equal to the reference's serially rounded poles to ~1e-16 relative, not bit for bit: the f32 output differs
// from the reference's only where a sum lies within that noise of an f32 rounding boundary, ~1e-8 per sample
// (0 of the golden vector's 355 285 samples; tests/test_parity_audio.py::test_eq_three_long_differential_run
// counts them on 2^23-sample calls).  "Bit-exact" for this module means that, not a theorem; the two-launch
// kernel with a large MXL_EQ_CHUNK (few carried starts) is the strict mode.
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
#include <stdlib.h>
#include <string.h>

#include "eq_stream.cuh"

namespace mxl {
namespace k {

namespace {

using namespace eqs;

// zero-pass table of the long-call variant: kernel parameters = constant-bank operands of the unrolled FMA chains
struct EqVParam { double K[8]; double V[64][8]; };
struct EqNoParam {};
struct ParamTab {
    const EqVParam& p;
    __device__ __forceinline__ double v(int j, int e) const { return p.V[j][e]; }
};

// BIG = the long-call variant: chunk loops unrolled (eq_core.cuh), V / K from the kernel parameters, no V in shared memory.
template <int LC, bool BIG, class VP>
__device__ __forceinline__ void eq_stream_body(const EqStreamBatch& b, const VP& vp)
{
    constexpr int VPR = LC / 4;
    constexpr bool ROLL = !BIG;
    using Shared = EqShared<LC, !BIG>;
    extern __shared__ __align__(16) unsigned char eq_smem[];
    float4* tile = reinterpret_cast<float4*>(eq_smem);                              // [256][VPR], swizzled
    Shared* sh = reinterpret_cast<Shared*>(eq_smem + (size_t)kT * LC * sizeof(float));
    const EqStreamInst& in = b.inst[blockIdx.y];
    const int tid = threadIdx.x;
    const EqStreamConsts& q = b.eq;
    const int halo = (int)q.halo;
    const int U = kT - halo;
    const int64_t c0 = (int64_t)blockIdx.x * U - halo;        // chunk of thread 0
    const int64_t c = c0 + tid;
    const bool active = c >= 0 && c < (int64_t)b.n_chunks;
    // every chunk of the tile exists and is full, lines 16-byte aligned: staging and store are straight copies
    const bool interior = c0 >= 0 && (uint64_t)(c0 + kT) * LC <= b.frames && in.in &&
                          ((reinterpret_cast<uintptr_t>(in.in) | reinterpret_cast<uintptr_t>(in.out)) & 15) == 0;
    // vector idx = it * kT + tid of the tile is (row it * (kT / VPR) + tid / VPR, vector tid % VPR): the swizzle term depends
    // on the row's low bits only, which `it` does not touch, so slots and addresses advance by constants
    static_assert((kT / VPR) % 8 == 0, "rows per staging round");

    // ---- 1. stage ----
    if (interior) {
        const float* src = in.in + c0 * LC + 4 * tid;
        float4* dst = tile + slot_of<VPR>(tid / VPR, tid % VPR);
#pragma unroll
        for (int it = 0; it < VPR; it++) cp_async16(dst + it * kT, src + it * kT * 4);
    } else {
        const float* src = in.in;
        const bool vec_ok = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
#pragma unroll 1
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
    }
    load_tables<LC>(q.tab, sh, tid);                            // in flight with the staging copies
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    const double* st = in.state;
    RowIo<LC> row{tile, tid};

    // ---- 2. zero pass ----
    double v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = 0.0;
    if (active) {
        if constexpr (BIG) {
#pragma unroll
            for (int e = 0; e < 8; e++) v[e] = vp.K[e];
            ParamTab tab{vp};
            eq_zero_state_dot<LC, ROLL>(row, tab, v);
        } else {
#pragma unroll
            for (int e = 0; e < 8; e++) v[e] = sh->K[e];
            SharedTab<LC> tab{sh};
            eq_zero_state_dot<LC, ROLL>(row, tab, v);
        }
        if (c == 0) {                                      // v_0 = A p_init + z_0
            const double pl[4] = {st[0], st[1], st[2], st[3]}, ph[4] = {st[4], st[5], st[6], st[7]};
            double yl[4], yh[4];
            tri_apply(sh->pow_lo[0], pl, yl);
            tri_apply(sh->pow_hi[0], ph, yh);
#pragma unroll
            for (int e = 0; e < 4; e++) { v[e] += yl[e]; v[4 + e] += yh[e]; }
        }
    }

    // a non-finite carry (NaN / inf input in my chunk, or a non-finite stored state): remember the first such chunk
    if (active && any_non_finite(v)) atomicMin(in.poison, (uint32_t)c);

    // ---- 3. scan (eq_stream.cuh): two block barriers ----
    double S[8];
    scan_start_states<LC>(q, sh, v, S, tid);

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
            p = EqPoles{S[0], S[1], S[2], S[3], S[4], S[5], S[6], S[7]};
            const float4 prev = tile[slot_of<VPR>(tid - 1, VPR - 1)];      // last vector of the chunk before mine
            hist[0] = (double)prev.y; hist[1] = (double)prev.z; hist[2] = (double)prev.w;
        }
        const uint64_t s0 = (uint64_t)c * LC;
        count = (uint32_t)((s0 + LC <= b.frames) ? LC : (b.frames - s0));
    }
    __syncthreads();                                       // every history read precedes any overwrite
    if (owner) {
        const EqGains g{q.c_lo, q.c_hi, in.g_lo, in.g_mid, in.g_hi};
        if (count == LC) {
            eq_run_chunk_skewed<LC, ROLL>(p, hist, row, g);
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
    if (interior) {
        // vector idx of the owned part is (row halo + idx / VPR, vector idx % VPR): constants per round, as in the staging
        const int total = U * VPR;
        float* dst = in.out + (c0 + halo) * LC + 4 * tid;
        const float4* srcv = tile + slot_of<VPR>(halo + tid / VPR, tid % VPR);
#pragma unroll
        for (int it = 0; it < VPR; it++)
            if (it * kT + tid < total) *reinterpret_cast<float4*>(dst + it * kT * 4) = srcv[it * kT];
    } else {
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
    // The barrier orders every thread's stores before thread 0's fence (fences are cumulative), so one fence per CTA makes
    // the tile's outputs visible before the CTA counts itself done.
    __shared__ uint32_t s_first_bad;
    __syncthreads();
    if (tid == 0) {
        __threadfence();
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

// (A persistent form -- one resident CTA per SM walking the tiles with two tile buffers, the next tile's cp.async copies in
// flight during the current tile's passes, tables loaded once -- gave the same bits and 0.180 ms instead of 0.127: with 8
// warps per SM the FP64 pipe is 50 % busy, with 24 (three one-tile CTAs) 73 %; in the exact pass a warp's stalls are fixed-
// latency dependency waits, so the pipe is fed by the NUMBER of warps, and the 64 KB tile caps that at 24.)
// (Starting the k-th resident CTA of an SM k * 1-6 us late, so that the CTAs sharing an SM do not pass through their phases
// in step, changed nothing: 0.125-0.127 ms at every delay.)
// (128-thread CTAs for the long-call variant -- more, smaller CTAs per SM passing through their phases at different
// times -- measured slower: 0.143 against 0.126 ms per 2^25 samples at LC = 64, twice the halo share.)
// short calls (a CTA or two per SM, code run once): rolled loops, every table in shared memory
template <int LC>
__global__ void __launch_bounds__(kT) eq_stream_kernel(const __grid_constant__ EqStreamBatch b)
{
    pdl_prologue();
    eq_stream_body<LC, false>(b, EqNoParam{});
}

// long calls (several waves of CTAs): unrolled loops, zero-pass table in the constant bank
template <int LC>
__global__ void __launch_bounds__(kT, LC == 64 ? 3 : 4) eq_stream_kernel_long(const __grid_constant__ EqStreamBatch b, const __grid_constant__ EqVParam vp)
{
    pdl_prologue();
    eq_stream_body<LC, true>(b, vp);
}

template <int LC>
int launch_lc(mxl_ctx* ctx, const EqStreamBatch& b)
{
    const size_t smem_short = (size_t)kT * LC * sizeof(float) + sizeof(EqShared<LC, true>);
    const size_t smem_long = (size_t)kT * LC * sizeof(float) + sizeof(EqShared<LC, false>);
    if (!(ctx->eq_stream_smem_set & (1u << (LC / 16)))) {
        MXL_CUDA(cudaFuncSetAttribute(eq_stream_kernel<LC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_short));
        MXL_CUDA(cudaFuncSetAttribute(eq_stream_kernel_long<LC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_long));
        MXL_CUDA(cudaFuncSetAttribute(eq_stream_kernel_long<LC>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        ctx->eq_stream_smem_set |= 1u << (LC / 16);
    }
    const uint32_t U = kT - b.eq.halo;
    dim3 grid((b.n_chunks + U - 1) / U, b.n);
    // the long-call variant once the SMs hold several CTAs each (see eq_core.cuh), the short one otherwise
    const uint64_t ctas = (uint64_t)grid.x * grid.y, sms = (uint64_t)(ctx->sm_count > 0 ? ctx->sm_count : 148);
    static const char* force = getenv("MXL_EQ_ROLL");
    const bool roll = force ? atoi(force) != 0 : ctas < 2 * sms;
    MXL_TIMED(ctx, "eq_stream_kernel");
    if (roll || !b.eq.host_V || !b.eq.host_K) {
        launch_chained(ctx, eq_stream_kernel<LC>, grid, dim3(kT), smem_short, b);
    } else {
        static thread_local EqVParam vp;                   // copied into the launch by cudaLaunchKernelEx
        memcpy(vp.K, b.eq.host_K, sizeof vp.K);
        memcpy(vp.V, b.eq.host_V, sizeof(double) * 8 * LC);
        launch_chained(ctx, eq_stream_kernel_long<LC>, grid, dim3(kT), smem_long, b, vp);
    }
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
    const EqStreamConsts& q = b.eq;
    if (q.halo == 0 || q.halo > (uint32_t)eqs::kT / 2 || q.lev_lo > (uint32_t)kEqPlanLevels || q.lev_hi > (uint32_t)kEqPlanLevels ||
        q.back_lo > 3 || q.back_hi > 3 || !q.tab)
        MXL_FAIL(MXL_ERR_INVALID, "eq_stream_kernel: bad plan (chunk %u, halo %u, levels %u/%u)", q.chunk, q.halo, q.lev_lo, q.lev_hi);
    switch (q.chunk) {
    case 16: return launch_lc<16>(ctx, b);
    case 32: return launch_lc<32>(ctx, b);
    case 64: return launch_lc<64>(ctx, b);
    default: MXL_FAIL(MXL_ERR_INVALID, "eq_stream_kernel: unsupported chunk length %u", q.chunk);
    }
}

}  // namespace k
}  // namespace mxl

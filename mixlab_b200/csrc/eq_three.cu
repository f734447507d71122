// eq_three.cu -- EqThree (src/module/eq_three.rs:58-89,106-125) parallelised along time: the generic
// two-launch scheme.  The product path is the single-launch kernel in eq_stream.cu; this one serves
// sample rates at which that kernel's cascades would not forget inside half a CTA (above ~190 kHz) and
// stays selectable with MXL_EQ_CHUNK for A/B runs.
//
// The module is two 4-pole cascades of one-pole low-passes in f64: a recurrence that is strictly
// serial in the reference.  One call here may cover millions of samples (many ticks per launch),
// so the time axis is cut into chunks of Lc samples, one thread per chunk:
//
//   phase A (eq_zero_state_kernel)  each chunk is run from an all-zero pole state; its end state
//           z_k is the contribution of the chunk's own inputs (and VSA terms) to later states.
//           Only ~1e-16 relative accuracy is needed, so the poles are advanced in FMA form.
//   carry   the true pole state at the start of chunk k is  sum_j A^(j-1) z_(k-j)  (+ A^k p_init
//           for the first chunks), A = M^Lc the homogeneous Lc-sample transition.  The 4-pole
//           cascade forgets geometrically, so the sum is cut after J terms once |A^J| < 2^-75;
//           each thread of phase C forms its own start state -- no scan, no inter-block traffic.
//   phase C (eq_exact_kernel)  each chunk is re-run from its start state with the reference's
//           exact operation order (no FMA) and the f32 outputs are written.
//
// Chunk 0 starts from the module's stored state, so successive calls continue bit-exactly.  For
// k > 0 the start state differs from the sequential one by f64 rounding noise (a few ulp, decaying
// at the filter's own rate), which the final `as f32` rounding absorbs: the reference's golden
// vector is reproduced bit-exactly at every chunk length tested (tests/test_parity_audio.py).
#include "dsp_math.cuh"
#include "kernels.h"

namespace mxl {
namespace k {

namespace {

constexpr int kEqThreads = 128;
constexpr double kVsa = 1.0 / 4294967295.0;   // eq_three.rs:11

__device__ __forceinline__ int tri(int r, int c) { return r * (r + 1) / 2 + c; }

// y += A * x for a packed lower-triangular 4x4 (FMA is fine here: carry accuracy only)
__device__ __forceinline__ void tri_mac(const double* A, const double x[4], double y[4])
{
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c <= r; c++) y[r] = fma(A[tri(r, c)], x[c], y[r]);
}

__global__ void __launch_bounds__(kEqThreads) eq_zero_state_kernel(const __grid_constant__ EqBatch b)
{
    const EqInst& in = b.inst[blockIdx.y];
    const uint32_t k = blockIdx.x * kEqThreads + threadIdx.x;
    if (k + 1 >= b.n_chunks) return;                 // the last chunk feeds nobody
    const float* src = in.in ? in.in + (uint64_t)k * b.chunk : nullptr;
    const double cl = b.c_lo, ch = b.c_hi, al = 1.0 - b.c_lo, ah = 1.0 - b.c_hi;
    double l0 = 0, l1 = 0, l2 = 0, l3 = 0, h0 = 0, h1 = 0, h2 = 0, h3 = 0;
    for (uint32_t i = 0; i < b.chunk; i += 4) {
        float4 x = src ? *reinterpret_cast<const float4*>(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const double s = (double)xs[j];
            l0 = fma(al, l0, fma(cl, s, kVsa));
            l1 = fma(cl, l0 - l1, l1);
            l2 = fma(cl, l1 - l2, l2);
            l3 = fma(cl, l2 - l3, l3);
            h0 = fma(ah, h0, fma(ch, s, kVsa));
            h1 = fma(ch, h0 - h1, h1);
            h2 = fma(ch, h1 - h2, h2);
            h3 = fma(ch, h2 - h3, h3);
        }
    }
    double* z = in.zend + (uint64_t)k * 8;
    z[0] = l0; z[1] = l1; z[2] = l2; z[3] = l3;
    z[4] = h0; z[5] = h1; z[6] = h2; z[7] = h3;
    // a NaN / inf that entered the poles never leaves them in the reference: remember the first such chunk
    // (eq_stream.cu explains; eq_exact_kernel's last CTA rewrites what follows it)
    if (!isfinite(l3 + h3 + l0 + h0)) atomicMin(in.poison, k);
}

struct EqRegs {
    double l0, l1, l2, l3, h0, h1, h2, h3;   // poles
    double x0, x1, x2;                       // history[0..2]
};

// eq_three.rs:66-86 with LowPass::pump (121-128) inlined; operation order is the reference's.
__device__ __forceinline__ float eq_step(EqRegs& r, float xin, double cl, double ch, double g_lo, double g_mid, double g_hi)
{
    const double s = (double)xin;
    r.l0 = r.l0 + (cl * (s - r.l0) + kVsa);
    r.l1 = r.l1 + cl * (r.l0 - r.l1);
    r.l2 = r.l2 + cl * (r.l1 - r.l2);
    r.l3 = r.l3 + cl * (r.l2 - r.l3);
    r.h0 = r.h0 + (ch * (s - r.h0) + kVsa);
    r.h1 = r.h1 + ch * (r.h0 - r.h1);
    r.h2 = r.h2 + ch * (r.h1 - r.h2);
    r.h3 = r.h3 + ch * (r.h2 - r.h3);
    double lo = r.l3;
    double hi = r.x0 - r.h3;
    double mid = r.x0 - (hi + lo);
    r.x0 = r.x1; r.x1 = r.x2; r.x2 = s;
    lo = lo * g_lo;
    mid = mid * g_mid;
    hi = hi * g_hi;
    return (float)(lo + mid + hi);
}

__global__ void __launch_bounds__(kEqThreads) eq_exact_kernel(const __grid_constant__ EqBatch b)
{
    const EqInst& in = b.inst[blockIdx.y];
    const uint32_t k = blockIdx.x * kEqThreads + threadIdx.x;
    if (k >= b.n_chunks) return;
    const uint64_t s0 = (uint64_t)k * b.chunk;
    const uint64_t s1 = (s0 + b.chunk < b.frames) ? s0 + b.chunk : b.frames;
    const double* st = in.state;                       // state before this call
    EqRegs r;
    if (k == 0) {
        r.l0 = st[0]; r.l1 = st[1]; r.l2 = st[2]; r.l3 = st[3];
        r.h0 = st[4]; r.h1 = st[5]; r.h2 = st[6]; r.h3 = st[7];
    } else {
        double pl[4] = {0, 0, 0, 0}, ph[4] = {0, 0, 0, 0};
        const uint32_t terms = k < b.carry_terms ? k : b.carry_terms;
        for (uint32_t j = 1; j <= terms; j++) {
            const double* z = in.zend + (uint64_t)(k - j) * 8;
            const double zl[4] = {z[0], z[1], z[2], z[3]}, zh[4] = {z[4], z[5], z[6], z[7]};
            tri_mac(b.pow_lo[j - 1], zl, pl);
            tri_mac(b.pow_hi[j - 1], zh, ph);
        }
        if (k <= b.carry_terms) {                      // initial state still audible: + A^k p_init
            const double il[4] = {st[0], st[1], st[2], st[3]}, ih[4] = {st[4], st[5], st[6], st[7]};
            tri_mac(b.pow_lo[k], il, pl);
            tri_mac(b.pow_hi[k], ih, ph);
        }
        r.l0 = pl[0]; r.l1 = pl[1]; r.l2 = pl[2]; r.l3 = pl[3];
        r.h0 = ph[0]; r.h1 = ph[1]; r.h2 = ph[2]; r.h3 = ph[3];
    }
    // history = the three inputs before s0 (inputs before the call live in the stored state)
    {
        double hist[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int64_t idx = (int64_t)s0 - 3 + j;
            hist[j] = idx >= 0 ? (in.in ? (double)in.in[idx] : 0.0) : st[8 + 3 + idx];
        }
        r.x0 = hist[0]; r.x1 = hist[1]; r.x2 = hist[2];
    }
    const double cl = b.c_lo, ch = b.c_hi, g_lo = in.g_lo, g_mid = in.g_mid, g_hi = in.g_hi;
    uint64_t i = s0;
    for (; i + 4 <= s1; i += 4) {
        float4 x = in.in ? *reinterpret_cast<const float4*>(in.in + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 y;
        y.x = eq_step(r, x.x, cl, ch, g_lo, g_mid, g_hi);
        y.y = eq_step(r, x.y, cl, ch, g_lo, g_mid, g_hi);
        y.z = eq_step(r, x.z, cl, ch, g_lo, g_mid, g_hi);
        y.w = eq_step(r, x.w, cl, ch, g_lo, g_mid, g_hi);
        *reinterpret_cast<float4*>(in.out + i) = y;
    }
    for (; i < s1; i++) in.out[i] = eq_step(r, in.in ? in.in[i] : 0.f, cl, ch, g_lo, g_mid, g_hi);
    if (k + 1 == b.n_chunks) {                         // state after this call (other half of the double buffer)
        double* so = in.state_out;
        so[0] = r.l0; so[1] = r.l1; so[2] = r.l2; so[3] = r.l3;
        so[4] = r.h0; so[5] = r.h1; so[6] = r.h2; so[7] = r.h3;
        so[8] = r.x0; so[9] = r.x1; so[10] = r.x2;
        // a non-finite sample in the last chunk (which has no zero-state run) or a poisoned stored state
        if (!isfinite(r.l3 + r.h3 + r.l0 + r.h0)) atomicMin(in.poison, k);
    }
    if (k == 0 && !isfinite(st[0] + st[3] + st[4] + st[7])) atomicMin(in.poison, 0u);
}

// Runs after eq_exact_kernel, one small CTA per instance: everything after the first poisoned chunk becomes NaN.
__global__ void __launch_bounds__(kEqThreads) eq_poison_fix_kernel(const __grid_constant__ EqBatch b)
{
    const EqInst& in = b.inst[blockIdx.x];
    __shared__ uint32_t s_first_bad;
    if (threadIdx.x == 0) s_first_bad = atomicExch(in.poison, 0xffffffffu);      // read and re-arm
    __syncthreads();
    const uint32_t first_bad = s_first_bad;
    if (first_bad == 0xffffffffu || first_bad + 1 >= b.n_chunks) return;
    const float nan = __int_as_float(0x7fc00000);
    for (uint64_t i = ((uint64_t)first_bad + 1) * b.chunk + threadIdx.x; i < b.frames; i += kEqThreads) in.out[i] = nan;
    if (threadIdx.x < 8) in.state_out[threadIdx.x] = __longlong_as_double(0x7ff8000000000000ll);
}

}  // namespace

int launch_eq_three(mxl_ctx* ctx, const EqBatch& b)
{
    if (!ctx || !ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_TRY(ctx->activate());
    if (b.n <= 0 || b.frames == 0) return MXL_OK;
    dim3 grid((b.n_chunks + kEqThreads - 1) / kEqThreads, b.n);
    if (b.n_chunks > 1) {
        MXL_TIMED(ctx, "eq_zero_state_kernel");
        eq_zero_state_kernel<<<grid, kEqThreads, 0, ctx->stream>>>(b);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) MXL_FAIL(MXL_ERR_CUDA, "launch of eq_zero_state_kernel failed: %s", cudaGetErrorString(e));
        ctx->launches++;
    }
    {
        MXL_TIMED(ctx, "eq_exact_kernel");
        eq_exact_kernel<<<grid, kEqThreads, 0, ctx->stream>>>(b);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) MXL_FAIL(MXL_ERR_CUDA, "launch of eq_exact_kernel failed: %s", cudaGetErrorString(e));
    ctx->launches++;
    {
        MXL_TIMED(ctx, "eq_poison_fix_kernel");
        eq_poison_fix_kernel<<<b.n, kEqThreads, 0, ctx->stream>>>(b);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) MXL_FAIL(MXL_ERR_CUDA, "launch of eq_poison_fix_kernel failed: %s", cudaGetErrorString(e));
    ctx->launches++;
    return MXL_OK;
}

}  // namespace k
}  // namespace mxl

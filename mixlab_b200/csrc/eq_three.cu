// eq_three.cu -- EqThree (src/module/eq_three.rs:58-89,106-125) parallelised along time.
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
    }
}

// ------------------------------------------------------------------------------------------------
// Single-launch variant.  One CTA of 256 threads owns 256 consecutive chunks of Lc samples of one
// instance; thread i owns chunk c = blockIdx.x*U - Hc + i, U = 256 - Hc.  The first Hc chunks are a
// halo re-computed from the previous CTA's range: after Hc chunks the 4-pole cascades have forgotten
// their start state to below 2^-75, so CTAs never talk to each other.
//   1. the CTA's 256*Lc input samples are staged coalesced into a padded shared-memory tile
//      (row stride Lc+1 words: thread-per-row reads are bank-conflict free);
//   2. every thread runs its chunk from zero state (FMA form) -> z_i, as four sub-chunks advanced
//      in lock step (four independent FP64 dependency chains per thread);  the chunk that starts
//      the call runs from the module's stored state instead;
//   3. inclusive scan of v_i = A v_(i-1) + z_i over the CTA: Hillis-Steele with A^(2^d) through
//      shared memory;  the start state of chunk i is v_(i-1);
//   4. the non-halo chunks are re-run from their start state in the reference's exact operation
//      order, outputs overwrite the tile row and leave with coalesced float4 stores.
// ------------------------------------------------------------------------------------------------
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

template <int NSUB>
__global__ void __launch_bounds__(kEqBlockThreads) eq_block_kernel(const __grid_constant__ EqBlockBatch b)
{
    extern __shared__ __align__(16) unsigned char eq_smem[];
    const EqBlockInst& in = b.inst[blockIdx.y];
    const int tid = threadIdx.x;
    const uint32_t Lc = b.chunk, row = Lc + 1;
    float* tile = reinterpret_cast<float*>(eq_smem);                                   // [256][Lc+1]
    double* xch = reinterpret_cast<double*>(eq_smem + (((size_t)kEqBlockThreads * row * 4 + 15) & ~(size_t)15));   // [256][8]
    const uint32_t U = kEqBlockThreads - b.halo_chunks;
    const int64_t c0 = (int64_t)blockIdx.x * U - (int64_t)b.halo_chunks;               // chunk of thread 0
    const int64_t c = c0 + tid;
    const bool active = c >= 0 && c < (int64_t)b.n_chunks;

    // ---- 1. stage the inputs (loads of four iterations in flight before the first shared store) ----
    {
        const uint32_t vec_per_row = Lc >> 2;
        const uint32_t total = kEqBlockThreads * vec_per_row;          // multiple of 4 * kEqBlockThreads (Lc % 16 == 0)
        for (uint32_t base = tid; base < total; base += 4 * kEqBlockThreads) {
            float4 v[4];
            uint32_t dst[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t idx = base + u * kEqBlockThreads;
                const uint32_t i = idx / vec_per_row, j = (idx - i * vec_per_row) << 2;
                const int64_t ci = c0 + i;
                dst[u] = i * row + j;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (in.in && ci >= 0) {
                    const uint64_t g = (uint64_t)ci * Lc + j;
                    if (g + 4 <= b.frames) v[u] = *reinterpret_cast<const float4*>(in.in + g);
                    else if (g < b.frames) {
                        v[u].x = in.in[g];
                        if (g + 1 < b.frames) v[u].y = in.in[g + 1];
                        if (g + 2 < b.frames) v[u].z = in.in[g + 2];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                float* d = tile + dst[u];
                d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
            }
        }
    }
    __syncthreads();

    const double cl = b.c_lo, ch = b.c_hi, al = 1.0 - b.c_lo, ah = 1.0 - b.c_hi;
    const float* mine = tile + tid * row;
    const double* st = in.state;

    // ---- 2. zero-state run; the chunk is cut into 4 sub-chunks advanced in lock step (4 independent
    //         dependency chains per thread hide the FP64 latency).  w[q] = state at the start of
    //         sub-chunk q for a zero state at the start of the chunk; the chunk that starts the call
    //         runs its first sub-chunk from the module's stored state instead. ----
    const uint32_t Ls = Lc / NSUB;
    double w[NSUB][8];
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < NSUB; q++)
#pragma unroll
        for (int e = 0; e < 8; e++) w[q][e] = 0.0;
    if (active) {
        double z[NSUB][8];
#pragma unroll
        for (int q = 0; q < NSUB; q++)
#pragma unroll
            for (int e = 0; e < 8; e++) z[q][e] = (q == 0 && c == 0) ? st[e] : 0.0;
        for (uint32_t j = 0; j < Ls; j++) {
#pragma unroll
            for (int q = 0; q < NSUB; q++) {
                const double s = (double)mine[q * Ls + j];
                z[q][0] = fma(al, z[q][0], fma(cl, s, kVsa));
                z[q][1] = fma(cl, z[q][0] - z[q][1], z[q][1]);
                z[q][2] = fma(cl, z[q][1] - z[q][2], z[q][2]);
                z[q][3] = fma(cl, z[q][2] - z[q][3], z[q][3]);
                z[q][4] = fma(ah, z[q][4], fma(ch, s, kVsa));
                z[q][5] = fma(ch, z[q][4] - z[q][5], z[q][5]);
                z[q][6] = fma(ch, z[q][5] - z[q][6], z[q][6]);
                z[q][7] = fma(ch, z[q][6] - z[q][7], z[q][7]);
            }
        }
        // w[q+1] = B w[q] + z[q];  v = w[NSUB]
#pragma unroll
        for (int q = 0; q < NSUB; q++) {
            double yl[4], yh[4];
            tri_apply(b.sub_lo[0], w[q], yl);
            tri_apply(b.sub_hi[0], w[q] + 4, yh);
            double* dst = q < NSUB - 1 ? w[q + 1 < NSUB ? q + 1 : 0] : v;
#pragma unroll
            for (int e = 0; e < 4; e++) { dst[e] = yl[e] + z[q][e]; dst[4 + e] = yh[e] + z[q][4 + e]; }
        }
    }

    // ---- 3. inclusive scan over the CTA: Hillis-Steele, v_i += A^(2^d) v_(i - 2^d) ----
    // (the partner's value must cover exactly the 2^d chunks before mine, so every step crosses warp
    // boundaries: all steps go through shared memory)
#pragma unroll
    for (int d = 0; d < kEqBlockLevels; d++) {
#pragma unroll
        for (int q = 0; q < 8; q++) xch[tid * 8 + q] = v[q];
        __syncthreads();
        if (tid >= (1 << d)) {
            double o[8], yl[4], yh[4];
#pragma unroll
            for (int q = 0; q < 8; q++) o[q] = xch[(tid - (1 << d)) * 8 + q];
            tri_apply(b.pow_lo[d], o, yl);
            tri_apply(b.pow_hi[d], o + 4, yh);
#pragma unroll
            for (int q = 0; q < 4; q++) { v[q] += yl[q]; v[4 + q] += yh[q]; }
        }
        __syncthreads();
    }
    // start state of my chunk = inclusive result of the previous thread
#pragma unroll
    for (int q = 0; q < 8; q++) xch[tid * 8 + q] = v[q];
    __syncthreads();

    // ---- 4. exact re-run of the chunks this CTA owns: 4 sub-chunks in lock step, each from its own
    //         start state  B^q S + w[q]  (S = start state of the chunk) ----
    const bool owner = active && (tid >= (int)b.halo_chunks || blockIdx.x == 0);
    EqRegs r[NSUB];
    uint32_t count = 0;
    if (owner) {
        double S[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (c != 0) {
            const double* p = xch + (tid - 1) * 8;       // c > 0 => tid >= 1
#pragma unroll
            for (int e = 0; e < 8; e++) S[e] = p[e];
        }
#pragma unroll
        for (int q = 0; q < NSUB; q++) {
            double pl[4], ph[4];
            if (q == 0) {
#pragma unroll
                for (int e = 0; e < 4; e++) { pl[e] = S[e]; ph[e] = S[4 + e]; }
            } else {
                tri_apply(b.sub_lo[q > 0 ? q - 1 : 0], S, pl);
                tri_apply(b.sub_hi[q > 0 ? q - 1 : 0], S + 4, ph);
#pragma unroll
                for (int e = 0; e < 4; e++) { pl[e] += w[q][e]; ph[e] += w[q][4 + e]; }
            }
            r[q].l0 = pl[0]; r[q].l1 = pl[1]; r[q].l2 = pl[2]; r[q].l3 = pl[3];
            r[q].h0 = ph[0]; r[q].h1 = ph[1]; r[q].h2 = ph[2]; r[q].h3 = ph[3];
            // history = the three inputs before the sub-chunk
            double hist[3];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int64_t idx = c * (int64_t)Lc + (int64_t)(q * Ls) - 3 + j;   // absolute sample index
                if (idx >= 0) {
                    const int64_t rel = idx - c0 * (int64_t)Lc;                    // position inside the tile
                    hist[j] = (double)tile[(rel / Lc) * row + (rel % Lc)];
                } else {
                    hist[j] = st[8 + 3 + idx];                                     // inputs of the previous call
                }
            }
            r[q].x0 = hist[0]; r[q].x1 = hist[1]; r[q].x2 = hist[2];
        }
        if (c == 0) {                                    // the call starts from the stored state, exactly
            r[0].l0 = st[0]; r[0].l1 = st[1]; r[0].l2 = st[2]; r[0].l3 = st[3];
            r[0].h0 = st[4]; r[0].h1 = st[5]; r[0].h2 = st[6]; r[0].h3 = st[7];
        }
        const uint64_t s0 = (uint64_t)c * Lc;
        count = (uint32_t)((s0 + Lc <= b.frames) ? Lc : (b.frames - s0));
    }
    __syncthreads();                                     // every history read precedes any overwrite
    if (owner) {
        float* wr = tile + tid * row;
        const double g_lo = in.g_lo, g_mid = in.g_mid, g_hi = in.g_hi;
        if (count == Lc) {
            for (uint32_t j = 0; j < Ls; j++) {
#pragma unroll
                for (int q = 0; q < NSUB; q++) wr[q * Ls + j] = eq_step(r[q], wr[q * Ls + j], cl, ch, g_lo, g_mid, g_hi);
            }
        } else {
            for (uint32_t j = 0; j < Ls; j++) {
#pragma unroll
                for (int q = 0; q < NSUB; q++)
                    if (q * Ls + j < count) wr[q * Ls + j] = eq_step(r[q], wr[q * Ls + j], cl, ch, g_lo, g_mid, g_hi);
            }
        }
        if (c + 1 == (int64_t)b.n_chunks) {              // state after this call: the sub-chunk holding the last sample
            const uint32_t ql = (count - 1) / Ls;
            EqRegs f = r[0];
#pragma unroll
            for (int q = 1; q < NSUB; q++) if (ql == (uint32_t)q) f = r[q];
            double* so = in.state_out;
            so[0] = f.l0; so[1] = f.l1; so[2] = f.l2; so[3] = f.l3;
            so[4] = f.h0; so[5] = f.h1; so[6] = f.h2; so[7] = f.h3;
            so[8] = f.x0; so[9] = f.x1; so[10] = f.x2;
        }
    }
    __syncthreads();

    // ---- coalesced store of the owned rows ----
    {
        const uint32_t first = b.halo_chunks;            // tile rows below Hc are halo (or precede the call)
        const uint32_t vec_per_row = Lc >> 2;
        const uint32_t total = (kEqBlockThreads - first) * vec_per_row;
        for (uint32_t idx = tid; idx < total; idx += kEqBlockThreads) {
            const uint32_t i = first + idx / vec_per_row, j = (idx % vec_per_row) << 2;
            const int64_t ci = c0 + i;
            if (ci < 0 || ci >= (int64_t)b.n_chunks) continue;
            const uint64_t g = (uint64_t)ci * Lc + j;
            const float* sp = tile + i * row + j;
            if (g + 4 <= b.frames) *reinterpret_cast<float4*>(in.out + g) = make_float4(sp[0], sp[1], sp[2], sp[3]);
            else for (uint32_t q = 0; q < 4 && g + q < b.frames; q++) in.out[g + q] = sp[q];
        }
    }
}

}  // namespace

int launch_eq_three_block(mxl_ctx* ctx, const EqBlockBatch& b)
{
    if (!ctx || !ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_TRY(ctx->activate());
    if (b.n <= 0 || b.frames == 0) return MXL_OK;
    if (b.chunk == 0 || (b.chunk & 15) || b.chunk > kEqBlockMaxChunk || b.halo_chunks == 0 || b.halo_chunks > kEqBlockThreads / 2)
        MXL_FAIL(MXL_ERR_INVALID, "eq_block_kernel: bad plan (chunk %u, halo %u)", b.chunk, b.halo_chunks);
    const size_t tile_bytes = ((size_t)kEqBlockThreads * (b.chunk + 1) * sizeof(float) + 15) & ~(size_t)15;
    const size_t smem = tile_bytes + (size_t)kEqBlockThreads * 8 * sizeof(double);
    const int nsub = b.subs == 4 ? 4 : (b.subs == 2 ? 2 : 1);
    if (smem > 48 * 1024 && smem > ctx->eq_block_smem) {
        MXL_CUDA(cudaFuncSetAttribute(eq_block_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MXL_CUDA(cudaFuncSetAttribute(eq_block_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MXL_CUDA(cudaFuncSetAttribute(eq_block_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->eq_block_smem = smem;
    }
    const uint32_t U = kEqBlockThreads - b.halo_chunks;
    dim3 grid((b.n_chunks + U - 1) / U, b.n);
    if (nsub == 4) eq_block_kernel<4><<<grid, kEqBlockThreads, smem, ctx->stream>>>(b);
    else if (nsub == 2) eq_block_kernel<2><<<grid, kEqBlockThreads, smem, ctx->stream>>>(b);
    else eq_block_kernel<1><<<grid, kEqBlockThreads, smem, ctx->stream>>>(b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) MXL_FAIL(MXL_ERR_CUDA, "launch of eq_block_kernel failed: %s", cudaGetErrorString(e));
    ctx->launches++;
    return MXL_OK;
}

int launch_eq_three(mxl_ctx* ctx, const EqBatch& b)
{
    if (!ctx || !ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_TRY(ctx->activate());
    if (b.n <= 0 || b.frames == 0) return MXL_OK;
    dim3 grid((b.n_chunks + kEqThreads - 1) / kEqThreads, b.n);
    if (b.n_chunks > 1) {
        eq_zero_state_kernel<<<grid, kEqThreads, 0, ctx->stream>>>(b);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) MXL_FAIL(MXL_ERR_CUDA, "launch of eq_zero_state_kernel failed: %s", cudaGetErrorString(e));
        ctx->launches++;
    }
    eq_exact_kernel<<<grid, kEqThreads, 0, ctx->stream>>>(b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) MXL_FAIL(MXL_ERR_CUDA, "launch of eq_exact_kernel failed: %s", cudaGetErrorString(e));
    ctx->launches++;
    return MXL_OK;
}

}  // namespace k
}  // namespace mxl

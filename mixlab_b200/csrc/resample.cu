// resample.cu -- audio sample-rate converter for the ingest side of the path (NEW, self-specified: the reference has
// only the TODO -- Icecast ingest drops every stream that is not 44.1 kHz, src/icecast/mod.rs:94-97, RTMP ingest panics,
// src/rtmp/mod.rs:229-232).  Parity unpinned: oracle/mixlab_oracle.c (orc_resampler_*) is the definition.
//
// Polyphase windowed sinc for a rational ratio L / M (44.1 k -> 48 k: 160 / 147):
//   output m sits at input position m * M / L:  n0 = floor(m * M / L), phase = (m * M) mod L
//   y[m] = f32( sum_{k=0}^{2H-1} c[phase][k] * f64(x[n0 - H + 1 + k]) ),  H = 16, accumulated with fma in ascending k
//   c[phase][k] = h(k - H + 1 - phase / L) / (sum of the row),  h(x) = w sinc(w x) bh(x / H),  w = 0.92 min(1, L / M),
//   bh = 4-term Blackman-Harris; inputs before the start of the stream are zero.
// The stream is a pure function of everything pushed so far: any split into calls gives the same bits.  After N input
// frames, ceil((N - H) L / M) output frames are determined (a latency of H input frames, 0.36 ms at 44.1 kHz).
//
// Kernel: outputs m and m + L share a phase, so a thread OWNS a phase: it loads its 32 coefficients once into registers
// and walks the outputs m0 + p + L r, r = 0 .. R-1, of its block (consecutive lanes = consecutive outputs: coalesced
// stores).  The input window the block touches (R M + 2H frames) is staged once in shared memory AS f64 (one conversion per
// input sample instead of one per tap); i16 sources are unpacked on the way (s / 32768, stream_input.rs:167-173).
// FP64-bound by design: 64 fma per stereo output frame against 15.4 bytes of line traffic.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "kernels.h"

namespace mxl {
namespace k {

namespace {

constexpr int kRsThreads = 256;
constexpr int kRsH = 16;                   // taps on each side
constexpr int kRsTaps = 2 * kRsH;

struct ResampleLaunch {
    const float* in_f32;                   // device: new input frames as f32, or nullptr
    const short* in_i16;                   // device: new input frames as i16, or nullptr
    const float* hist;                     // device: the kRsTaps input frames before the new ones (f32), zeros at stream start
    float* out;
    const double* coef;                    // device: [L][kRsTaps]
    uint64_t in_base;                      // absolute index of the first new input frame
    uint64_t n_new;                        // new input frames
    uint64_t out_base;                     // absolute index of out[0]
    uint64_t n_out;
    uint64_t block0;                       // absolute index (a multiple of L) of the first output of block 0
    uint32_t L, M, channels;
    uint32_t repeats;                      // R: outputs per thread; a block is L * R consecutive outputs
    uint32_t window;                       // input frames a block stages: R * M + kRsTaps (+ slack)
};

__device__ __forceinline__ float fetch(const ResampleLaunch& p, int64_t n, uint32_t ch)
{
    // n = absolute input frame index; before the stream: zero; the kRsTaps frames before in_base: history
    if (n < 0) return 0.f;
    const int64_t rel = n - (int64_t)p.in_base;
    if (rel < 0) return rel >= -(int64_t)kRsTaps ? p.hist[(size_t)(rel + kRsTaps) * p.channels + ch] : 0.f;
    if ((uint64_t)rel >= p.n_new) return 0.f;
    if (p.in_i16) return (float)p.in_i16[(size_t)rel * p.channels + ch] / 32768.0f;     // stream_input.rs:167-173
    return p.in_f32[(size_t)rel * p.channels + ch];
}

// grid (blocks, ceil(L / 256)): thread = phase.
template <int CH>
__global__ void __launch_bounds__(kRsThreads) resample_kernel(const ResampleLaunch p)
{
    extern __shared__ __align__(16) double rs_x[];         // [window][CH]
    const uint64_t m0 = p.block0 + (uint64_t)blockIdx.x * p.L * p.repeats;     // a multiple of L
    const int64_t n_first = (int64_t)(m0 / p.L) * p.M - kRsH + 1;              // first input frame this block touches
    // Staging, one FRAME per thread and trip.  The window [n_first, n_first + window) is, in order: frames before the stream
    // (zeros), the kRsTaps frames of history, the new input, frames not pushed yet (zeros, never used by a determined output);
    // the boundaries are block-uniform, the new input of an interior block is one straight vector copy.
    {
        const int64_t rel0 = n_first - (int64_t)p.in_base;                     // window frame i is new-input frame rel0 + i
        const bool interior = rel0 >= 0 && (uint64_t)rel0 + p.window <= p.n_new;
        if (interior && CH == 2 && p.in_f32) {
            const float2* src = reinterpret_cast<const float2*>(p.in_f32) + rel0;
            for (uint32_t i = threadIdx.x; i < p.window; i += blockDim.x) {
                const float2 v = __ldg(src + i);
                *reinterpret_cast<double2*>(rs_x + 2 * i) = make_double2((double)v.x, (double)v.y);
            }
        } else if (interior && CH == 2 && p.in_i16) {
            const short2* src = reinterpret_cast<const short2*>(p.in_i16) + rel0;
            for (uint32_t i = threadIdx.x; i < p.window; i += blockDim.x) {
                const short2 v = __ldg(src + i);
                *reinterpret_cast<double2*>(rs_x + 2 * i) = make_double2((double)((float)v.x / 32768.0f), (double)((float)v.y / 32768.0f));
            }
        } else if (interior && CH == 1 && p.in_f32) {
            for (uint32_t i = threadIdx.x; i < p.window; i += blockDim.x) rs_x[i] = (double)__ldg(p.in_f32 + rel0 + i);
        } else if (interior && CH == 1 && p.in_i16) {
            for (uint32_t i = threadIdx.x; i < p.window; i += blockDim.x) rs_x[i] = (double)((float)__ldg(p.in_i16 + rel0 + i) / 32768.0f);
        } else {
            for (uint32_t i = threadIdx.x; i < p.window * CH; i += blockDim.x)
                rs_x[i] = (double)fetch(p, n_first + i / CH, i % CH);
        }
    }
    const uint32_t ph_idx = blockIdx.y * blockDim.x + threadIdx.x;             // output offset inside a run of L
    double c[kRsTaps];
    uint32_t x_off = 0;
    if (ph_idx < p.L) {
        const uint64_t pos = (uint64_t)ph_idx * p.M;                           // (m0 + ph_idx) * M = (m0 / L) * M * L + pos
        const double* row = p.coef + (size_t)(pos % p.L) * kRsTaps;
#pragma unroll
        for (int kk = 0; kk < kRsTaps; kk++) c[kk] = __ldg(row + kk);
        x_off = (uint32_t)(pos / p.L);                                         // n0 - (m0 / L) * M for r = 0
    }
    __syncthreads();
    if (ph_idx >= p.L) return;
    const uint64_t out_end = p.out_base + p.n_out;
#pragma unroll 1
    for (uint32_t r = 0; r < p.repeats; r++) {
        const uint64_t m = m0 + ph_idx + (uint64_t)r * p.L;
        if (m < p.out_base || m >= out_end) continue;
        const double* x = rs_x + (size_t)(x_off + r * p.M) * CH;              // x[0] = input frame n0 - H + 1
        if (CH == 2) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int kk = 0; kk < kRsTaps; kk++) {
                const double2 xv = *reinterpret_cast<const double2*>(x + 2 * kk);
                a0 = fma(c[kk], xv.x, a0);
                a1 = fma(c[kk], xv.y, a1);
            }
            *reinterpret_cast<float2*>(p.out + 2 * (m - p.out_base)) = make_float2((float)a0, (float)a1);
        } else {
            double a0 = 0.0;
#pragma unroll
            for (int kk = 0; kk < kRsTaps; kk++) a0 = fma(c[kk], x[kk], a0);
            p.out[m - p.out_base] = (float)a0;
        }
    }
}

// the last kRsTaps input frames of (history + new input) become the next call's history
__global__ void resample_history_kernel(const ResampleLaunch p, float* hist_out)
{
    const uint32_t i = threadIdx.x;
    if (i >= kRsTaps * p.channels) return;
    const uint32_t fr = i / p.channels, ch = i % p.channels;
    const int64_t n = (int64_t)(p.in_base + p.n_new) - kRsTaps + fr;
    hist_out[i] = fetch(p, n, ch);
}

uint64_t gcd_u64(uint64_t a, uint64_t b) { while (b) { const uint64_t t = a % b; a = b; b = t; } return a; }

double sinc(double u) { const double pi = 3.14159265358979323846264338327950288; return u == 0.0 ? 1.0 : sin(pi * u) / (pi * u); }

double blackman_harris(double u)           // u in [-1, 1]
{
    const double pi = 3.14159265358979323846264338327950288;
    return 0.35875 + 0.48829 * cos(pi * u) + 0.14128 * cos(2.0 * pi * u) + 0.01168 * cos(3.0 * pi * u);
}

}  // namespace
}  // namespace k
}  // namespace mxl

using namespace mxl;

struct mxl_resampler {
    mxl_ctx* ctx = nullptr;
    uint32_t in_rate = 0, out_rate = 0, channels = 0, L = 1, M = 1;
    uint64_t total_in = 0, total_out = 0;
    double* coef = nullptr;                // device [L][32]
    float* hist[2] = {nullptr, nullptr};   // device, double-buffered [32][channels]
    int cur = 0;
    short* stage = nullptr;                // device staging of pushed i16
    size_t stage_cap = 0;
};

static uint64_t determined(const mxl_resampler* r, uint64_t n_in)
{
    if (n_in <= (uint64_t)k::kRsH) return 0;
    return ((n_in - k::kRsH) * r->L + r->M - 1) / r->M;   // ceil((N - H) L / M)
}

extern "C" {

mxl_resampler* mxl_resampler_create(mxl_ctx* ctx, uint32_t in_rate, uint32_t out_rate, uint32_t channels)
{
    if (!ctx || !ctx->has_device()) { set_error("mxl_resampler_create: needs a device context (there is no CPU fallback)"); return nullptr; }
    if (in_rate == 0 || out_rate == 0 || (channels != 1 && channels != 2)) { set_error("mxl_resampler_create: rates must be positive, channels 1 or 2"); return nullptr; }
    if (ctx->activate() != MXL_OK) return nullptr;
    mxl_resampler* r = new mxl_resampler();
    r->ctx = ctx; r->in_rate = in_rate; r->out_rate = out_rate; r->channels = channels;
    const uint64_t g = k::gcd_u64(in_rate, out_rate);
    r->L = (uint32_t)(out_rate / g);
    r->M = (uint32_t)(in_rate / g);
    if (r->L > 4096) { set_error("mxl_resampler_create: %u / %u needs %u phases (limit 4096)", out_rate, in_rate, r->L); delete r; return nullptr; }
    // coefficient table: the definition at the head of this file, in f64 on the host
    std::vector<double> tab((size_t)r->L * k::kRsTaps);
    const double ratio = (double)r->L / (double)r->M;
    const double w = 0.92 * (ratio < 1.0 ? ratio : 1.0);
    for (uint32_t ph = 0; ph < r->L; ph++) {
        double row[k::kRsTaps], sum = 0.0;
        for (int kk = 0; kk < k::kRsTaps; kk++) {
            const double x = (double)(kk - k::kRsH + 1) - (double)ph / (double)r->L;
            row[kk] = w * k::sinc(w * x) * k::blackman_harris(x / (double)k::kRsH);
            sum += row[kk];
        }
        for (int kk = 0; kk < k::kRsTaps; kk++) tab[(size_t)ph * k::kRsTaps + kk] = row[kk] / sum;
    }
    bool ok = cudaMalloc(&r->coef, tab.size() * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMemcpy(r->coef, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
    for (int i = 0; i < 2 && ok; i++) {
        ok = cudaMalloc(&r->hist[i], (size_t)k::kRsTaps * channels * sizeof(float)) == cudaSuccess;
        ok = ok && cudaMemset(r->hist[i], 0, (size_t)k::kRsTaps * channels * sizeof(float)) == cudaSuccess;
    }
    if (!ok) { set_error("mxl_resampler_create: CUDA allocation failed"); mxl_resampler_destroy(r); return nullptr; }
    return r;
}

void mxl_resampler_destroy(mxl_resampler* r)
{
    if (!r) return;
    if (r->ctx && r->ctx->has_device()) { r->ctx->activate(); cudaStreamSynchronize(r->ctx->stream); }
    if (r->coef) cudaFree(r->coef);
    for (int i = 0; i < 2; i++) if (r->hist[i]) cudaFree(r->hist[i]);
    if (r->stage) cudaFree(r->stage);
    delete r;
}

int mxl_resampler_reset(mxl_resampler* r)
{
    if (!r) MXL_FAIL(MXL_ERR_INVALID, "NULL resampler");
    MXL_TRY(r->ctx->activate());
    for (int i = 0; i < 2; i++) MXL_CUDA(cudaMemsetAsync(r->hist[i], 0, (size_t)k::kRsTaps * r->channels * sizeof(float), r->ctx->stream));
    r->total_in = r->total_out = 0;
    return MXL_OK;
}

uint64_t mxl_resampler_output_frames(const mxl_resampler* r, uint64_t in_frames)
{
    if (!r) return 0;
    return determined(r, r->total_in + in_frames) - r->total_out;
}

static int64_t resampler_push(mxl_resampler* r, const float* in_f32, const short* in_i16, uint64_t in_frames, mxl_line* out)
{
    mxl_ctx* ctx = r->ctx;
    if (!out || out->ctx != ctx) MXL_FAIL(MXL_ERR_INVALID, "mxl_resampler: the output line is NULL or of another context");
    if (out->type != (r->channels == 2 ? MXL_LINE_STEREO : MXL_LINE_MONO)) MXL_FAIL(MXL_ERR_LINE_TYPE, "mxl_resampler: output line type does not match %u channels", r->channels);
    const uint64_t n_out = determined(r, r->total_in + in_frames) - r->total_out;
    MXL_TRY(line_resize(out, n_out));
    k::ResampleLaunch p{};
    p.in_f32 = in_f32; p.in_i16 = in_i16; p.hist = r->hist[r->cur]; p.out = out->dev; p.coef = r->coef;
    p.in_base = r->total_in; p.n_new = in_frames; p.out_base = r->total_out; p.n_out = n_out;
    p.L = r->L; p.M = r->M; p.channels = r->channels;
    // a block = L * R consecutive outputs starting at a multiple of L; R so that the staged window stays near 32 KB
    p.repeats = std::max<uint32_t>(1u, std::min<uint32_t>(32u, 2048u / r->M));
    p.window = p.repeats * r->M + k::kRsTaps + 2;
    p.block0 = r->total_out / r->L * r->L;
    const size_t smem = (size_t)p.window * r->channels * sizeof(double);
    if (smem > 200 * 1024) MXL_FAIL(MXL_ERR_UNSUPPORTED, "mxl_resampler: ratio %u / %u needs a %zu-byte window", r->L, r->M, smem);
    MXL_TRY(ctx->compute_begin());
    if (n_out) {
        const uint64_t per_block = (uint64_t)r->L * p.repeats;
        const unsigned blocks = (unsigned)((r->total_out + n_out - p.block0 + per_block - 1) / per_block);
        // a thread per phase: L rounded up to whole warps, at most 256 per block (44.1 k -> 48 k: 160 threads, no idle warps)
        const unsigned threads = std::min<unsigned>(k::kRsThreads, (r->L + 31u) / 32u * 32u);
        dim3 grid(blocks, (r->L + threads - 1) / threads);
        MXL_TIMED(ctx, "resample_kernel");
        if (r->channels == 2) {
            if (smem > 48 * 1024) MXL_CUDA(cudaFuncSetAttribute(k::resample_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k::resample_kernel<2><<<grid, threads, smem, ctx->stream>>>(p);
        } else {
            if (smem > 48 * 1024) MXL_CUDA(cudaFuncSetAttribute(k::resample_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k::resample_kernel<1><<<grid, threads, smem, ctx->stream>>>(p);
        }
        if (cudaGetLastError() != cudaSuccess) MXL_FAIL(MXL_ERR_CUDA, "launch of resample_kernel failed");
        ctx->launches++;
    }
    if (in_frames) {
        k::resample_history_kernel<<<1, 64, 0, ctx->stream>>>(p, r->hist[r->cur ^ 1]);
        if (cudaGetLastError() != cudaSuccess) MXL_FAIL(MXL_ERR_CUDA, "launch of resample_history_kernel failed");
        ctx->launches++;
        r->cur ^= 1;
    }
    MXL_TRY(ctx->compute_end());
    r->total_in += in_frames;
    r->total_out += n_out;
    return (int64_t)n_out;
}

int64_t mxl_resampler_push_i16(mxl_resampler* r, const int16_t* host_pcm, uint64_t in_frames, mxl_line* out)
{
    if (!r || (in_frames && !host_pcm)) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    mxl_ctx* ctx = r->ctx;
    MXL_TRY(ctx->activate());
    const size_t n = (size_t)in_frames * r->channels;
    if (n > r->stage_cap) {
        if (r->stage) { MXL_CUDA(cudaStreamSynchronize(ctx->stream)); cudaFree(r->stage); r->stage = nullptr; r->stage_cap = 0; }
        MXL_CUDA(cudaMalloc(&r->stage, (n + n / 4 + 64) * sizeof(short)));
        r->stage_cap = n + n / 4 + 64;
    }
    if (n) {
        MXL_CUDA(cudaMemcpyAsync(r->stage, host_pcm, n * sizeof(short), cudaMemcpyHostToDevice, ctx->stream));
        ctx->h2d_bytes += n * sizeof(short);
    }
    return resampler_push(r, nullptr, r->stage, in_frames, out);
}

int64_t mxl_resampler_push_line(mxl_resampler* r, const mxl_line* in, mxl_line* out)
{
    if (!r || !in) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if (in->ctx != r->ctx || in->type != (r->channels == 2 ? MXL_LINE_STEREO : MXL_LINE_MONO)) MXL_FAIL(MXL_ERR_LINE_TYPE, "mxl_resampler: input line type does not match %u channels", r->channels);
    MXL_TRY(r->ctx->activate());
    return resampler_push(r, in->dev, nullptr, in->frames, out);
}

}  // extern "C"

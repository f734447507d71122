// envelope.cu -- Envelope (src/module/envelope.rs:16-58,91-120) without the serial state machine.
//
// The reference walks a 3-state machine sample by sample.  Its transitions depend only on whether
// the machine is "on" (TriggerOn) or not (Initial / TriggerOff): a gate sample == 1.0 switches
// not-on to on, a gate sample == 0.0 switches on to not-on, every other sample is inert.  Hence
//   * the class after sample i is the class of the last sample <= i whose value is exactly 1.0 or
//     0.0 (an "event"), or the incoming class if there is none;
//   * an event is a TRANSITION iff its class differs from the class just before it;
//   * the state at sample i is fixed by the last transition p <= i (TriggerOn{on: p} or
//     TriggerOff{off: p, ..}); off_amplitude is amplitude(TriggerOn{on: q}, p) with q the transition
//     before p (or the incoming `on`).
// Both "last event <= i" and "last transition <= i" are inclusive max-scans over sample indices:
// block-local warp-shuffle scans plus a scan of per-block maxima.  Amplitudes are then evaluated
// independently per sample with the reference's f64 expressions (bit-exact: only IEEE +,-,*,/).
#include "dsp_math.cuh"
#include "kernels.h"

namespace mxl {
namespace k {

namespace {

constexpr int kEnvThreads = 256;
constexpr int kEnvPerThread = 4;
constexpr int kEnvTile = kEnvThreads * kEnvPerThread;    // samples per block

__device__ __forceinline__ uint32_t warp_incl_max(uint32_t v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v = max(v, n);
    }
    return v;
}

// In-block inclusive max-scan of 4 consecutive items per thread; returns the exclusive prefix of the
// thread's first item (within the block) and the block maximum through *block_max.
__device__ __forceinline__ uint32_t block_excl_max(uint32_t thread_max, uint32_t* block_max)
{
    __shared__ uint32_t warp_max[kEnvThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = warp_incl_max(thread_max);
    if (lane == 31) warp_max[w] = incl;
    __syncthreads();
    uint32_t prefix = 0;
    for (int i = 0; i < w; i++) prefix = max(prefix, warp_max[i]);
    uint32_t total = 0;
    for (int i = 0; i < kEnvThreads / 32; i++) total = max(total, warp_max[i]);
    uint32_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 0;
    __syncthreads();
    *block_max = total;
    return max(prefix, excl);
}

// event class of a gate sample: 1 = "on" event (== 1.0), 2 = "off" event (== 0.0, either sign), 0 = inert
__device__ __forceinline__ int event_of(float v) { return v == 1.0f ? 1 : (v == 0.0f ? 2 : 0); }

// pass 1: local[i] = max index+1 of an event in [tile_start, i]
__global__ void __launch_bounds__(kEnvThreads) env_events_kernel(const __grid_constant__ EnvLaunch p)
{
    const uint32_t base = blockIdx.x * kEnvTile + threadIdx.x * kEnvPerThread;
    uint32_t v[kEnvPerThread], run = 0;
#pragma unroll
    for (int j = 0; j < kEnvPerThread; j++) {
        const uint32_t i = base + j;
        const float x = (i < p.frames && p.in) ? p.in[i] : (i < p.frames ? 0.0f : 2.0f);
        run = max(run, event_of(x) ? i + 1 : 0u);
        v[j] = run;
    }
    uint32_t bmax;
    const uint32_t excl = block_excl_max(run, &bmax);
#pragma unroll
    for (int j = 0; j < kEnvPerThread; j++)
        if (base + j < p.frames) p.scratch_a[base + j] = max(v[j], excl);
    if (threadIdx.x == 0) p.block_a[blockIdx.x] = bmax;
}

// exclusive max-scan of the per-block maxima, in place, one block
__global__ void __launch_bounds__(kEnvThreads) env_block_scan_kernel(uint32_t* blocks, uint32_t n)
{
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += kEnvThreads) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t mine = i < n ? blocks[i] : 0u;
        uint32_t bmax;
        const uint32_t excl = block_excl_max(mine, &bmax);
        const uint32_t carry = carry_s;
        if (i < n) blocks[i] = max(excl, carry);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = max(carry, bmax);
        __syncthreads();
    }
}

__device__ __forceinline__ int class_at(const EnvLaunch& p, uint32_t idx_plus1, int incoming_on)
{
    // class (1 = on, 0 = not on) established by the event at idx_plus1-1, or the incoming class
    if (idx_plus1 == 0) return incoming_on;
    const float x = p.in ? p.in[idx_plus1 - 1] : 0.0f;
    return x == 1.0f ? 1 : 0;
}

__device__ __forceinline__ uint32_t last_event_upto(const EnvLaunch& p, uint32_t i)
{
    return max(p.scratch_a[i], p.block_a[i / kEnvTile]);
}

// pass 2: local[i] = max index+1 of a transition in [tile_start, i]
__global__ void __launch_bounds__(kEnvThreads) env_transitions_kernel(const __grid_constant__ EnvLaunch p)
{
    const int incoming_on = p.state->state == 1;
    const uint32_t base = blockIdx.x * kEnvTile + threadIdx.x * kEnvPerThread;
    uint32_t v[kEnvPerThread], run = 0;
#pragma unroll
    for (int j = 0; j < kEnvPerThread; j++) {
        const uint32_t i = base + j;
        uint32_t mark = 0;
        if (i < p.frames) {
            const int ev = event_of(p.in ? p.in[i] : 0.0f);
            if (ev) {
                const int before = class_at(p, i > 0 ? last_event_upto(p, i - 1) : 0u, incoming_on);
                if ((ev == 1) != (before == 1)) mark = i + 1;
            }
        }
        run = max(run, mark);
        v[j] = run;
    }
    uint32_t bmax;
    const uint32_t excl = block_excl_max(run, &bmax);
#pragma unroll
    for (int j = 0; j < kEnvPerThread; j++)
        if (base + j < p.frames) p.scratch_b[base + j] = max(v[j], excl);
    if (threadIdx.x == 0) p.block_b[blockIdx.x] = bmax;
}

__device__ __forceinline__ uint32_t last_transition_upto(const EnvLaunch& p, uint32_t i)
{
    return max(p.scratch_b[i], p.block_b[i / kEnvTile]);
}

// envelope.rs:16-18
__device__ __forceinline__ double duration_ms(uint64_t first, uint64_t last, double sr)
{
    return (double)(last - first) / sr * 1000.0;
}

// envelope.rs:38-51, TriggerOn arm
__device__ __forceinline__ double amp_on(const EnvLaunch& p, uint64_t on, uint64_t t)
{
    const double ms_since_on = duration_ms(on, t, p.sample_rate);
    if (ms_since_on < p.attack_ms) return 1.0 / p.attack_ms * ms_since_on;
    const double ms_since_decay_started = ms_since_on - p.attack_ms;
    const double decay_amplitude = 1.0 - clamp01(1.0 / p.decay_ms * ms_since_decay_started);
    return p.sustain + ((1.0 - p.sustain) * decay_amplitude);
}

// envelope.rs:52-57, TriggerOff arm
__device__ __forceinline__ double amp_off(const EnvLaunch& p, uint64_t off, double off_amplitude, uint64_t t)
{
    const double ms_since_off = duration_ms(off, t, p.sample_rate);
    const double release_amplitude = 1.0 - clamp01(1.0 / p.release_ms * ms_since_off);
    return off_amplitude * release_amplitude;
}

// pass 3: resolve the state at every sample and evaluate the amplitude (envelope.rs:116)
__global__ void __launch_bounds__(kEnvThreads) env_apply_kernel(const __grid_constant__ EnvLaunch p, EnvState* state_out)
{
    const EnvState s0 = *p.state;
    const uint32_t base = (blockIdx.x * kEnvThreads + threadIdx.x) * kEnvPerThread;
    float y[kEnvPerThread];
#pragma unroll
    for (int j = 0; j < kEnvPerThread; j++) {
        const uint32_t i = base + j;
        if (i >= p.frames) { y[j] = 0.f; continue; }
        const uint64_t t = p.t0 + i;
        const uint32_t tr = last_transition_upto(p, i);
        EnvState s = s0;
        if (tr != 0) {
            const uint32_t pidx = tr - 1;
            const bool to_on = (p.in ? p.in[pidx] : 0.0f) == 1.0f;
            s.seq = p.t0 + pidx;
            if (to_on) {
                s.state = 1;
            } else {
                // envelope.rs:108-112: off_amplitude = amplitude(TriggerOn{on}, off)
                const uint32_t q = pidx > 0 ? last_transition_upto(p, pidx - 1) : 0u;
                const uint64_t on = q != 0 ? p.t0 + (q - 1) : s0.seq;
                s.state = 2;
                s.off_amplitude = amp_on(p, on, s.seq);
            }
        }
        double a = 0.0;                                   // Initial
        if (s.state == 1) a = amp_on(p, s.seq, t);
        else if (s.state == 2) a = amp_off(p, s.seq, s.off_amplitude, t);
        y[j] = (float)a;
        if (i + 1 == p.frames) *state_out = s;
    }
    if (base + kEnvPerThread <= p.frames && (reinterpret_cast<uintptr_t>(p.out + base) & 15) == 0) {
        *reinterpret_cast<float4*>(p.out + base) = make_float4(y[0], y[1], y[2], y[3]);
    } else {
#pragma unroll
        for (int j = 0; j < kEnvPerThread; j++)
            if (base + j < p.frames) p.out[base + j] = y[j];
    }
}

__global__ void env_commit_kernel(EnvState* dst, const EnvState* src) { *dst = *src; }

}  // namespace

uint32_t envelope_blocks(uint32_t frames) { return (frames + kEnvTile - 1) / kEnvTile; }

int launch_envelope(mxl_ctx* ctx, const EnvLaunch& p)
{
    if (!ctx || !ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_TRY(ctx->activate());
    if (p.frames == 0) return MXL_OK;
    const uint32_t nb = envelope_blocks(p.frames);
    EnvState* staged = p.state_next;
    env_events_kernel<<<nb, kEnvThreads, 0, ctx->stream>>>(p);
    env_block_scan_kernel<<<1, kEnvThreads, 0, ctx->stream>>>(p.block_a, nb);
    env_transitions_kernel<<<nb, kEnvThreads, 0, ctx->stream>>>(p);
    env_block_scan_kernel<<<1, kEnvThreads, 0, ctx->stream>>>(p.block_b, nb);
    env_apply_kernel<<<nb, kEnvThreads, 0, ctx->stream>>>(p, staged);
    env_commit_kernel<<<1, 1, 0, ctx->stream>>>(p.state, staged);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) MXL_FAIL(MXL_ERR_CUDA, "envelope launch failed: %s", cudaGetErrorString(e));
    ctx->launches += 6;
    return MXL_OK;
}

}  // namespace k
}  // namespace mxl

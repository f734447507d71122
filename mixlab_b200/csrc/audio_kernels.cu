// audio_kernels.cu -- element-wise audio module kernels for sm_100a.
//
// All of these are streaming, HBM-bound passes: one 16-byte (float4) access per thread per line,
// fully coalesced, read-once inputs loaded through the non-coherent path without L1 allocation.
// Arithmetic follows the reference operation by operation: f64 exactly where the Rust widens,
// no FMA contraction (-fmad=false), round-to-nearest conversions.
#include <stdlib.h>

#include "dsp_math.cuh"
#include "kernels.h"
#include "osc_core.cuh"

namespace mxl {
namespace k {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float4 ldg_stream(const float* p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

inline unsigned blocks_for(uint64_t items, int threads = kThreads)
{
    return (unsigned)((items + threads - 1) / threads);
}

// ------------------------------------------------------------------------------------------
// Oscillator: oscillator.rs:73-89
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) oscillator_kernel(const __grid_constant__ OscBatch b)
{
    pdl_prologue();
    const OscInst& in = b.inst[blockIdx.y];
    uint64_t f0 = ((uint64_t)blockIdx.x * kThreads + threadIdx.x) * 4;
    if (f0 >= b.frames) return;
    const double freq = in.freq, sr = b.sample_rate, inv_sr = b.inv_sample_rate;
    const int wf = in.waveform;
    if (f0 + 4 <= b.frames) {
        // one u64 -> f64 conversion per thread (XU pipe); the neighbours are exact +1.0 steps below 2^53
        const double q0 = (double)(b.t0 + f0);
        const bool exact = (b.t0 + f0 + 3) < (1ull << 53);
        double n[4];
        if (exact) {                                   // one well-predicted branch: no conversions on this side
#pragma unroll
            for (int j = 0; j < 4; j++) n[j] = osc_phase(q0 + (double)j, sr, inv_sr, freq);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) n[j] = osc_phase((double)(b.t0 + f0 + j), sr, inv_sr, freq);
        }
        float s[4];
        osc_wave4(wf, n, s);
        if (in.mono) st4(in.mono + f0, make_float4(s[0], s[1], s[2], s[3]));
        if (in.stereo) {
            st4(in.stereo + 2 * f0, make_float4(s[0], s[0], s[1], s[1]));
            st4(in.stereo + 2 * f0 + 4, make_float4(s[2], s[2], s[3], s[3]));
        }
    } else {
        for (uint64_t f = f0; f < b.frames; f++) {
            float s = osc_wave(osc_phase((double)(b.t0 + f), sr, inv_sr, freq), wf);
            if (in.mono) in.mono[f] = s;
            if (in.stereo) { in.stereo[2 * f] = s; in.stereo[2 * f + 1] = s; }
        }
    }
}

// ------------------------------------------------------------------------------------------
// FmSine: fm_sine.rs:45-53
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float fm_sample(double seq, double sr, double inv_sr, double mid, double amp, float x)
{
    double t = div_by_const(seq, sr, inv_sr);
    double co = (mid + amp * (double)x) * kTwoPi;      // (a * 2.0) * PI == a * (2.0 * PI): the doubling is exact
    return (float)sin_f64(co * t);
}

__global__ void __launch_bounds__(kThreads) fm_sine_kernel(const __grid_constant__ FmBatch b)
{
    pdl_prologue();
    const FmInst& in = b.inst[blockIdx.y];
    uint64_t f0 = ((uint64_t)blockIdx.x * kThreads + threadIdx.x) * 4;
    if (f0 >= b.frames) return;
    if (f0 + 4 <= b.frames) {
        float4 x = in.in ? ldg_stream(in.in + f0) : make_float4(0.f, 0.f, 0.f, 0.f);
        const double q0 = (double)(b.t0 + f0);
        const bool exact = (b.t0 + f0 + 3) < (1ull << 53);
        const float xs[4] = {x.x, x.y, x.z, x.w};
        double seq[4], arg[4], y[4];
        if (exact) {
#pragma unroll
            for (int j = 0; j < 4; j++) seq[j] = q0 + (double)j;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) seq[j] = (double)(b.t0 + f0 + j);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {                                  // fm_sine.rs:45-47
            const double t = div_by_const(seq[j], b.sample_rate, b.inv_sample_rate);
            const double co = (in.freq_mid + in.freq_amp * (double)xs[j]) * kTwoPi;
            arg[j] = co * t;
        }
        sin_f64x4(arg, y);
        const float s0 = (float)y[0], s1 = (float)y[1], s2 = (float)y[2], s3 = (float)y[3];
        st4(in.out + 2 * f0, make_float4(s0, s0, s1, s1));
        st4(in.out + 2 * f0 + 4, make_float4(s2, s2, s3, s3));
    } else {
        for (uint64_t f = f0; f < b.frames; f++) {
            float x = in.in ? in.in[f] : 0.f;
            float s = fm_sample((double)(b.t0 + f), b.sample_rate, b.inv_sample_rate, in.freq_mid, in.freq_amp, x);
            in.out[2 * f] = s;
            in.out[2 * f + 1] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Mixer: mixer.rs:54-68.  One thread per float4 of output, channels summed in channel order so
// the f32 accumulation order is the reference's.  No inter-thread reduction exists on this bus:
// the sum runs over channels, which a thread walks serially with all its loads in flight.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float mix1(float x, double g) { return (float)((double)x * g); }

template <int UNROLL, int CU>
__global__ void __launch_bounds__(kThreads) mixer_kernel(const __grid_constant__ MixerLaunch p)
{
    pdl_prologue();
    const uint64_t n4 = p.len >> 2;
    // the UNROLL float4 of a thread are kThreads apart: every load/store instruction of a warp covers
    // 512 contiguous bytes
    const uint64_t i0 = (uint64_t)blockIdx.x * (kThreads * UNROLL) + threadIdx.x;
    constexpr uint64_t kStep = kThreads;
    if (i0 < n4) {
        float4 m[UNROLL], c[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const bool live = i0 + u * kStep < n4;
            if (p.accumulate && live) {
                m[u] = *reinterpret_cast<const float4*>(p.master + 4 * (i0 + u * kStep));
                c[u] = *reinterpret_cast<const float4*>(p.cue + 4 * (i0 + u * kStep));
            } else {
                m[u] = make_float4(0.f, 0.f, 0.f, 0.f);   // util::zero(master), util::zero(cue)
                c[u] = m[u];
            }
        }
        // CU channels at a time: their loads are all in flight before the first sum, the sums then run in
        // channel order (the reference's f32 accumulation order)
        for (int ch0 = 0; ch0 < p.channels; ch0 += CU) {
            float4 x[CU][UNROLL];
#pragma unroll
            for (int k = 0; k < CU; k++) {
                const float* src = ch0 + k < p.channels ? p.ch[ch0 + k].in : nullptr;
#pragma unroll
                for (int u = 0; u < UNROLL; u++)
                    x[k][u] = (src && i0 + u * kStep < n4) ? ldg_stream(src + 4 * (i0 + u * kStep)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int k = 0; k < CU; k++) {
                if (ch0 + k >= p.channels) break;
                const double g = p.ch[ch0 + k].gain;
                const bool cue = p.ch[ch0 + k].cue != 0;
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    m[u].x += mix1(x[k][u].x, g); m[u].y += mix1(x[k][u].y, g);
                    m[u].z += mix1(x[k][u].z, g); m[u].w += mix1(x[k][u].w, g);
                    if (cue) { c[u].x += x[k][u].x; c[u].y += x[k][u].y; c[u].z += x[k][u].z; c[u].w += x[k][u].w; }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            if (i0 + u * kStep < n4) {
                st4(p.master + 4 * (i0 + u * kStep), m[u]);
                st4(p.cue + 4 * (i0 + u * kStep), c[u]);
            }
        }
    }
    // scalar tail (len % 4 floats), handled by the first threads of block 0
    const uint64_t tail0 = n4 << 2;
    if (blockIdx.x == 0 && tail0 + threadIdx.x < p.len) {
        const uint64_t i = tail0 + threadIdx.x;
        float m = p.accumulate ? p.master[i] : 0.f, c = p.accumulate ? p.cue[i] : 0.f;
        for (int ch = 0; ch < p.channels; ch++) {
            float x = p.ch[ch].in ? p.ch[ch].in[i] : 0.f;
            m += mix1(x, p.ch[ch].gain);
            if (p.ch[ch].cue) c += x;
        }
        p.master[i] = m;
        p.cue[i] = c;
    }
}

// One tick of a WIDE bus (many channels, few samples): the kernel above leaves a handful of threads walking all the
// channels, 16 loads at a time.  Here a CTA owns 8 output vectors; its 256 threads first form the products of ALL
// channels in parallel (8 vectors x 32 channel lanes, every load of the CTA in flight at once) into shared memory,
// then one warp -- a lane per output float -- adds them up in channel order: the f32 accumulation order stays the
// reference's, only the products were formed early (each is its own rounding, mixer.rs:62).
constexpr int kMixWideVec = 8;
__global__ void __launch_bounds__(kThreads) mixer_wide_kernel(const __grid_constant__ MixerLaunch p)
{
    pdl_prologue();
    __shared__ float4 s_prod[kMixerMaxCh][kMixWideVec];     // (float)((double)x * gain)
    __shared__ float4 s_raw[kMixerMaxCh][kMixWideVec];      // x, for the cue bus
    const uint64_t n4 = p.len >> 2;
    const int v = threadIdx.x & (kMixWideVec - 1), cl = threadIdx.x / kMixWideVec;       // 8 vectors x 32 channel lanes
    const uint64_t i = (uint64_t)blockIdx.x * kMixWideVec + v;
    constexpr int kLanes = kThreads / kMixWideVec;
    constexpr int kPer = (kMixerMaxCh + kLanes - 1) / kLanes;
    float4 x[kPer];
#pragma unroll
    for (int k = 0; k < kPer; k++) {
        const int ch = cl + k * kLanes;
        const float* src = ch < p.channels ? p.ch[ch].in : nullptr;
        x[k] = (src && i < n4) ? ldg_stream(src + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < kPer; k++) {
        const int ch = cl + k * kLanes;
        if (ch < p.channels) {
            const double g = p.ch[ch].gain;
            s_prod[ch][v] = make_float4(mix1(x[k].x, g), mix1(x[k].y, g), mix1(x[k].z, g), mix1(x[k].w, g));
            s_raw[ch][v] = x[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < kMixWideVec * 4) {                    // one warp: lane = (vector, component)
        const int ov = threadIdx.x >> 2, oc = threadIdx.x & 3;
        const uint64_t o = ((uint64_t)blockIdx.x * kMixWideVec + ov) * 4 + oc;
        if (o < (n4 << 2)) {
            float m = p.accumulate ? p.master[o] : 0.f, c = p.accumulate ? p.cue[o] : 0.f;     // util::zero (mixer.rs:54-55)
            const float* prod = reinterpret_cast<const float*>(&s_prod[0][ov]) + oc;
            const float* raw = reinterpret_cast<const float*>(&s_raw[0][ov]) + oc;
            for (int ch = 0; ch < p.channels; ch++) {
                m += prod[ch * kMixWideVec * 4];
                if (p.ch[ch].cue) c += raw[ch * kMixWideVec * 4];
            }
            p.master[o] = m;
            p.cue[o] = c;
        }
    }
    // scalar tail (len % 4 floats), handled by the first threads of block 0
    const uint64_t tail0 = n4 << 2;
    if (blockIdx.x == 0 && tail0 + threadIdx.x < p.len) {
        const uint64_t t = tail0 + threadIdx.x;
        float m = p.accumulate ? p.master[t] : 0.f, c = p.accumulate ? p.cue[t] : 0.f;
        for (int ch = 0; ch < p.channels; ch++) {
            float xv = p.ch[ch].in ? p.ch[ch].in[t] : 0.f;
            m += mix1(xv, p.ch[ch].gain);
            if (p.ch[ch].cue) c += xv;
        }
        p.master[t] = m;
        p.cue[t] = c;
    }
}

// ------------------------------------------------------------------------------------------
// Amplifier: amplifier.rs:52-57,71-73.  One thread = 4 frames = 8 stereo floats + 4 control floats.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float amp1(float x, double mod_value, double depth, double amplitude)
{
    double dv = 1.0 - depth + depth * mod_value;      // depth(value, depth)
    return (float)((double)x * dv * amplitude);
}

__device__ __forceinline__ float2 ldg_stream2(const float* p)
{
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

// kAmpUnroll float4 (2 frames each) per thread, kThreads apart: every warp instruction covers 512
// contiguous bytes of the stereo lines (256 of the control line) and all of a thread's loads are in
// flight before the first f64 conversion (the conversions run on the 16-lane XU pipe and would
// otherwise sit between a thread's loads: 0.73 of the copy peak with one vector pair per thread).
constexpr int kAmpUnroll = 4;

__global__ void __launch_bounds__(kThreads) amplifier_kernel(const __grid_constant__ AmpBatch b)
{
    pdl_prologue();
    const AmpInst& in = b.inst[blockIdx.y];
    const uint64_t n4 = b.frames >> 1;                              // float4 = 2 stereo frames
    const uint64_t i0 = (uint64_t)blockIdx.x * (kThreads * kAmpUnroll) + threadIdx.x;
    const double d = in.mod_depth, a = in.amplitude;
    if (i0 < n4) {
        float4 x[kAmpUnroll];
        float2 m[kAmpUnroll];
#pragma unroll
        for (int u = 0; u < kAmpUnroll; u++) {
            const uint64_t v = i0 + (uint64_t)u * kThreads;
            x[u] = (in.in && v < n4) ? ldg_stream(in.in + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
            m[u] = (in.mod && v < n4) ? ldg_stream2(in.mod + 2 * v) : make_float2(1.f, 1.f);
        }
#pragma unroll
        for (int u = 0; u < kAmpUnroll; u++) {
            const uint64_t v = i0 + (uint64_t)u * kThreads;
            if (v >= n4) break;
            // without a control line the reference uses depth(1.0, d) = 1 - d + d * 1.0 as well (amplifier.rs:52-57)
            const double m0 = (double)m[u].x, m1 = (double)m[u].y;
            st4(in.out + 4 * v, make_float4(amp1(x[u].x, m0, d, a), amp1(x[u].y, m0, d, a),
                                            amp1(x[u].z, m1, d, a), amp1(x[u].w, m1, d, a)));
        }
    }
    if ((b.frames & 1) && blockIdx.x == 0 && threadIdx.x == 0) {    // odd frame count: last frame
        const uint64_t f = b.frames - 1;
        const double mv = in.mod ? (double)in.mod[f] : 1.0;
        in.out[2 * f] = amp1(in.in ? in.in[2 * f] : 0.f, mv, d, a);
        in.out[2 * f + 1] = amp1(in.in ? in.in[2 * f + 1] : 0.f, mv, d, a);
    }
}

// ------------------------------------------------------------------------------------------
// StereoPanner stereo_panner.rs:35-38 ; StereoSplitter stereo_splitter.rs:41-44 ; Trigger trigger.rs:38-45
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) panner_kernel(const __grid_constant__ PanBatch b)
{
    pdl_prologue();
    const PanInst& in = b.inst[blockIdx.y];
    uint64_t f0 = ((uint64_t)blockIdx.x * kThreads + threadIdx.x) * 4;
    if (f0 >= b.frames) return;
    if (f0 + 4 <= b.frames) {
        float4 l = in.left ? ldg_stream(in.left + f0) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 r = in.right ? ldg_stream(in.right + f0) : make_float4(0.f, 0.f, 0.f, 0.f);
        st4(in.out + 2 * f0, make_float4(l.x, r.x, l.y, r.y));
        st4(in.out + 2 * f0 + 4, make_float4(l.z, r.z, l.w, r.w));
    } else {
        for (uint64_t f = f0; f < b.frames; f++) {
            in.out[2 * f] = in.left ? in.left[f] : 0.f;
            in.out[2 * f + 1] = in.right ? in.right[f] : 0.f;
        }
    }
}

__global__ void __launch_bounds__(kThreads) splitter_kernel(const __grid_constant__ SplitBatch b)
{
    pdl_prologue();
    const SplitInst& in = b.inst[blockIdx.y];
    uint64_t f0 = ((uint64_t)blockIdx.x * kThreads + threadIdx.x) * 4;
    if (f0 >= b.frames) return;
    if (f0 + 4 <= b.frames) {
        float4 a = in.in ? ldg_stream(in.in + 2 * f0) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 c = in.in ? ldg_stream(in.in + 2 * f0 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        st4(in.left + f0, make_float4(a.x, a.z, c.x, c.z));
        st4(in.right + f0, make_float4(a.y, a.w, c.y, c.w));
    } else {
        for (uint64_t f = f0; f < b.frames; f++) {
            in.left[f] = in.in ? in.in[2 * f] : 0.f;
            in.right[f] = in.in ? in.in[2 * f + 1] : 0.f;
        }
    }
}

__global__ void __launch_bounds__(kThreads) fill_kernel(const __grid_constant__ FillBatch b)
{
    // (no programmatic-launch prologue here: a CTA of this kernel is one store per thread, and the wait cost it
    // 0.96 -> 0.58 of the copy peak on long lines)
    const FillInst& in = b.inst[blockIdx.y];
    uint64_t i0 = ((uint64_t)blockIdx.x * kThreads + threadIdx.x) * 4;
    if (i0 >= b.len) return;
    const float v = in.value;
    if (i0 + 4 <= b.len) st4(in.out + i0, make_float4(v, v, v, v));
    else for (uint64_t i = i0; i < b.len; i++) in.out[i] = v;
}

// ------------------------------------------------------------------------------------------
// Meter (new): per tick slot, per channel: peak |s|, sum of s^2 in f64, and the
// OutputDevice clip predicate `s < -1 || s > 1` (output_device.rs:192-194,202-204).
// Two shapes: a warp per (slot, instance) when the call has enough slots to fill the machine, else a block per
// slot (a live one-tick call: shortest latency).
// ------------------------------------------------------------------------------------------
constexpr int kMeterThreads = 128;
constexpr int kMeterWarps = kMeterThreads / 32;

// One block per (slot, instance); warp-shuffle tree, then one smem hop across warps.
__global__ void __launch_bounds__(kMeterThreads) meter_block_kernel(const __grid_constant__ MeterBatch b)
{
    pdl_prologue();
    const MeterInst& in = b.inst[blockIdx.y];
    const uint64_t slot = blockIdx.x;
    const uint64_t f_begin = slot * b.spt;
    uint64_t f_end = f_begin + b.spt;
    if (f_end > b.frames) f_end = b.frames;
    float pk0 = 0.f, pk1 = 0.f;
    double sq0 = 0.0, sq1 = 0.0;
    int clip = 0;
    if (in.in && ((f_begin & 1) == 0) && (reinterpret_cast<uintptr_t>(in.in) & 15) == 0) {
        // the slot starts 16-byte aligned: float4 = two frames, four vectors in flight per thread
        // (a tick of 800 frames is one round of 128 threads x 4 vectors, predicated)
        constexpr int kU = 4;
        const uint64_t v_begin = f_begin >> 1, v_end = f_end >> 1;
        for (uint64_t vb = v_begin; vb < v_end; vb += kU * kMeterThreads) {
            float4 s[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const uint64_t v = vb + (uint64_t)u * kMeterThreads + threadIdx.x;
                s[u] = ldg_stream(in.in + 4 * (v < v_end ? v : v_end - 1));     // see meter_warp_kernel
            }
#pragma unroll
            for (int u = 0; u < kU; u++)
                if (vb + (uint64_t)u * kMeterThreads + threadIdx.x >= v_end) s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < kU; u++) {
                pk0 = fmaxf(pk0, fmaxf(fabsf(s[u].x), fabsf(s[u].z)));   // fmaxf drops NaN like the oracle's `a > peak`
                pk1 = fmaxf(pk1, fmaxf(fabsf(s[u].y), fabsf(s[u].w)));
                sq0 += (double)s[u].x * (double)s[u].x;
                sq1 += (double)s[u].y * (double)s[u].y;
                sq0 += (double)s[u].z * (double)s[u].z;
                sq1 += (double)s[u].w * (double)s[u].w;
            }
        }
        if (((f_end - f_begin) & 1) && threadIdx.x == 0) {              // odd frame count: last frame
            const float2 t = ldg_stream2(in.in + 2 * (f_end - 1));
            pk0 = fmaxf(pk0, fabsf(t.x)); pk1 = fmaxf(pk1, fabsf(t.y));
            sq0 += (double)t.x * (double)t.x; sq1 += (double)t.y * (double)t.y;
        }
    } else if (in.in) {
        constexpr int kU = 4;                  // frames in flight per thread
        for (uint64_t base = f_begin; base < f_end; base += kU * kMeterThreads) {
            float2 s[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const uint64_t f = base + (uint64_t)u * kMeterThreads + threadIdx.x;
                s[u] = f < f_end ? ldg_stream2(in.in + 2 * f) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                pk0 = fmaxf(pk0, fabsf(s[u].x));
                pk1 = fmaxf(pk1, fabsf(s[u].y));
                sq0 += (double)s[u].x * (double)s[u].x;
                sq1 += (double)s[u].y * (double)s[u].y;
            }
        }
    }
    // output_device.rs:192-194: `s < -1.0 || s > 1.0` for any sample <=> the peak of |s| exceeds 1 (a NaN
    // sample fails both tests there and is dropped by fmaxf here)
    clip = (pk0 > 1.0f || pk1 > 1.0f) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        pk0 = fmaxf(pk0, __shfl_xor_sync(0xffffffffu, pk0, o));
        pk1 = fmaxf(pk1, __shfl_xor_sync(0xffffffffu, pk1, o));
        sq0 += __shfl_xor_sync(0xffffffffu, sq0, o);
        sq1 += __shfl_xor_sync(0xffffffffu, sq1, o);
        clip |= __shfl_xor_sync(0xffffffffu, clip, o);
    }
    __shared__ float s_pk[2][kMeterThreads / 32];
    __shared__ double s_sq[2][kMeterThreads / 32];
    __shared__ int s_clip[kMeterThreads / 32];
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_pk[0][w] = pk0; s_pk[1][w] = pk1; s_sq[0][w] = sq0; s_sq[1][w] = sq1; s_clip[w] = clip; }
    __syncthreads();
    if (threadIdx.x == 0) {
        MeterRecord r;
        r.peak[0] = r.peak[1] = 0.f; r.sumsq[0] = r.sumsq[1] = 0.0; r.clip = 0; r._pad = 0;
        for (int i = 0; i < kMeterThreads / 32; i++) {
            r.peak[0] = fmaxf(r.peak[0], s_pk[0][i]); r.peak[1] = fmaxf(r.peak[1], s_pk[1][i]);
            r.sumsq[0] += s_sq[0][i]; r.sumsq[1] += s_sq[1][i];
            r.clip |= s_clip[i];
        }
        in.out[slot] = r;
    }
}

// One WARP per (slot, instance): a tick of 800 stereo frames is 400 float4 = 12.5 per lane, so the
// shuffle tree is paid once per ~13 loads (a block per slot paid it, a shared-memory hop and a barrier
// once per ~3 loads and was issue-bound at 0.79 of the copy peak).  No shared memory, no barrier.
__global__ void __launch_bounds__(kMeterThreads) meter_warp_kernel(const __grid_constant__ MeterBatch b, uint32_t n_slots)
{
    pdl_prologue();
    const MeterInst& in = b.inst[blockIdx.y];
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t slot = (uint64_t)blockIdx.x * kMeterWarps + (threadIdx.x >> 5);
    if (slot >= n_slots) return;
    const uint64_t f_begin = slot * b.spt;
    uint64_t f_end = f_begin + b.spt;
    if (f_end > b.frames) f_end = b.frames;
    float pk0 = 0.f, pk1 = 0.f;
    double sq0 = 0.0, sq1 = 0.0;
    if (in.in && ((f_begin & 1) == 0) && (reinterpret_cast<uintptr_t>(in.in) & 15) == 0) {
        // the slot starts 16-byte aligned: float4 = two frames, eight vectors in flight per lane (a warp has no
        // other warp of its slot to hide the latency behind), each warp instruction covering 512 contiguous bytes
        constexpr int kU = 8;
        const uint64_t v_begin = f_begin >> 1, v_end = f_end >> 1;
        for (uint64_t vb = v_begin; vb < v_end; vb += kU * 32) {
            float4 s[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) {
                // unconditional loads (index clamped, value masked below): a predicated load ends up scheduled
                // behind the arithmetic of the one before it, which leaves one or two loads in flight per warp
                const uint64_t v = vb + (uint64_t)u * 32 + lane;
                s[u] = ldg_stream(in.in + 4 * (v < v_end ? v : v_end - 1));
            }
#pragma unroll
            for (int u = 0; u < kU; u++)
                if (vb + (uint64_t)u * 32 + lane >= v_end) s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < kU; u++) {
                pk0 = fmaxf(pk0, fmaxf(fabsf(s[u].x), fabsf(s[u].z)));   // fmaxf drops NaN like the oracle's `a > peak`
                pk1 = fmaxf(pk1, fmaxf(fabsf(s[u].y), fabsf(s[u].w)));
                sq0 += (double)s[u].x * (double)s[u].x;
                sq1 += (double)s[u].y * (double)s[u].y;
                sq0 += (double)s[u].z * (double)s[u].z;
                sq1 += (double)s[u].w * (double)s[u].w;
            }
        }
        if (((f_end - f_begin) & 1) && lane == 0) {                     // odd frame count: last frame
            const float2 t = ldg_stream2(in.in + 2 * (f_end - 1));
            pk0 = fmaxf(pk0, fabsf(t.x)); pk1 = fmaxf(pk1, fabsf(t.y));
            sq0 += (double)t.x * (double)t.x; sq1 += (double)t.y * (double)t.y;
        }
    } else if (in.in) {
        constexpr int kU = 4;                  // frames in flight per lane
        for (uint64_t base = f_begin; base < f_end; base += kU * 32) {
            float2 s[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const uint64_t f = base + (uint64_t)u * 32 + lane;
                s[u] = f < f_end ? ldg_stream2(in.in + 2 * f) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                pk0 = fmaxf(pk0, fabsf(s[u].x));
                pk1 = fmaxf(pk1, fabsf(s[u].y));
                sq0 += (double)s[u].x * (double)s[u].x;
                sq1 += (double)s[u].y * (double)s[u].y;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        pk0 = fmaxf(pk0, __shfl_xor_sync(0xffffffffu, pk0, o));
        pk1 = fmaxf(pk1, __shfl_xor_sync(0xffffffffu, pk1, o));
        sq0 += __shfl_xor_sync(0xffffffffu, sq0, o);
        sq1 += __shfl_xor_sync(0xffffffffu, sq1, o);
    }
    if (lane == 0) {
        MeterRecord r;
        r.peak[0] = pk0; r.peak[1] = pk1; r.sumsq[0] = sq0; r.sumsq[1] = sq1; r._pad = 0;
        // output_device.rs:192-194: `s < -1.0 || s > 1.0` for any sample <=> the peak of |s| exceeds 1 (a NaN
        // sample fails both tests there and is dropped by fmaxf here)
        r.clip = (pk0 > 1.0f || pk1 > 1.0f) ? 1 : 0;
        in.out[slot] = r;
    }
}

// ------------------------------------------------------------------------------------------
// OutputDevice (output_device.rs:177-206): left / right of the stereo line into two channels of the
// device's interleaved scratch buffer, clip = any routed sample outside [-1, 1].  Left is stored first,
// so right wins when both sides map to one channel, as in the reference's loop.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) route_kernel(const __grid_constant__ RouteLaunch p)
{
    pdl_prologue();
    const uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    bool clip = false;
    if (i < p.frames) {
        const float2 s = p.in ? ldg_stream2(p.in + 2 * i) : make_float2(0.f, 0.f);
        float* o = p.scratch + i * p.channels;
        if (p.left >= 0) { clip |= s.x < -1.0f || s.x > 1.0f; o[p.left] = s.x; }
        if (p.right >= 0) { clip |= s.y < -1.0f || s.y > 1.0f; o[p.right] = s.y; }
    }
    if (__any_sync(0xffffffffu, clip) && (threadIdx.x & 31) == 0) atomicOr(p.clip, 1);
}

// ------------------------------------------------------------------------------------------
// PCM pack: src/video/encode.rs:184-195 (clamp, *32767, `as i16`); unpack: stream_input.rs:167-173
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ short pack1(float s)
{
    s = s > 1.0f ? 1.0f : (s < -1.0f ? -1.0f : s);        // NaN falls through both tests, as in Rust
    float v = s * 32767.0f;
    // Rust `as i16`: NaN -> 0, saturating, toward zero.  cvt.rzi.s16.f32 saturates; NaN needs the guard.
    return (v != v) ? (short)0 : (short)__float2int_rz(fminf(fmaxf(v, -32768.0f), 32767.0f));
}

constexpr int kPcmUnroll = 4;

__global__ void __launch_bounds__(kThreads) pcm_pack_kernel(const float* __restrict__ in, short* __restrict__ out, uint64_t len)
{
    pdl_prologue();
    const uint64_t n4 = len >> 2;
    const uint64_t i0 = (uint64_t)blockIdx.x * (kThreads * kPcmUnroll) + threadIdx.x;
    float4 x[kPcmUnroll];
#pragma unroll
    for (int u = 0; u < kPcmUnroll; u++) {
        const uint64_t v = i0 + (uint64_t)u * kThreads;
        x[u] = v < n4 ? ldg_stream(in + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kPcmUnroll; u++) {
        const uint64_t v = i0 + (uint64_t)u * kThreads;
        if (v < n4) *reinterpret_cast<short4*>(out + 4 * v) = make_short4(pack1(x[u].x), pack1(x[u].y), pack1(x[u].z), pack1(x[u].w));
    }
    const uint64_t tail = (n4 << 2) + threadIdx.x;
    if (blockIdx.x == 0 && tail < len) out[tail] = pack1(in[tail]);
}

__global__ void __launch_bounds__(kThreads) pcm_unpack_kernel(const short* __restrict__ in, float* __restrict__ out, uint64_t len)
{
    pdl_prologue();
    const uint64_t n4 = len >> 2;
    const uint64_t i0 = (uint64_t)blockIdx.x * (kThreads * kPcmUnroll) + threadIdx.x;
    short4 x[kPcmUnroll];
#pragma unroll
    for (int u = 0; u < kPcmUnroll; u++) {
        const uint64_t v = i0 + (uint64_t)u * kThreads;
        x[u] = v < n4 ? *reinterpret_cast<const short4*>(in + 4 * v) : make_short4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < kPcmUnroll; u++) {
        const uint64_t v = i0 + (uint64_t)u * kThreads;
        if (v < n4) st4(out + 4 * v, make_float4((float)x[u].x / 32768.0f, (float)x[u].y / 32768.0f,
                                                 (float)x[u].z / 32768.0f, (float)x[u].w / 32768.0f));
    }
    const uint64_t tail = (n4 << 2) + threadIdx.x;
    if (blockIdx.x == 0 && tail < len) out[tail] = (float)in[tail] / 32768.0f;
}

__global__ void __launch_bounds__(kThreads) fill_bytes_kernel(uint4* dst, size_t n16, uint32_t word)
{
    size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    size_t stride = (size_t)gridDim.x * kThreads;
    for (; i < n16; i += stride) dst[i] = make_uint4(word, word, word, word);
}

int check_launch(mxl_ctx* ctx, const char* name)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", name, cudaGetErrorString(e));
        return MXL_ERR_CUDA;
    }
    ctx->launches++;
    return MXL_OK;
}

}  // namespace

#define MXL_REQUIRE_DEVICE(ctx)                                                  \
    do {                                                                         \
        if (!(ctx) || !(ctx)->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context"); \
        MXL_TRY((ctx)->activate());                                              \
    } while (0)

int launch_oscillator(mxl_ctx* ctx, const OscBatch& b)
{
    MXL_REQUIRE_DEVICE(ctx);
    if (b.n <= 0 || b.frames == 0) return MXL_OK;
    dim3 grid(blocks_for((b.frames + 3) / 4), b.n);
    MXL_TIMED(ctx, "oscillator_kernel");
    launch_chained(ctx, oscillator_kernel, grid, dim3(kThreads), 0, b);
    return check_launch(ctx, "oscillator_kernel");
}

int launch_fm_sine(mxl_ctx* ctx, const FmBatch& b)
{
    MXL_REQUIRE_DEVICE(ctx);
    if (b.n <= 0 || b.frames == 0) return MXL_OK;
    dim3 grid(blocks_for((b.frames + 3) / 4), b.n);
    MXL_TIMED(ctx, "fm_sine_kernel");
    launch_chained(ctx, fm_sine_kernel, grid, dim3(kThreads), 0, b);
    return check_launch(ctx, "fm_sine_kernel");
}

int launch_mixer(mxl_ctx* ctx, const MixerLaunch& p)
{
    MXL_REQUIRE_DEVICE(ctx);
    if (p.len == 0) return MXL_OK;
    const uint64_t n4 = p.len >> 2;
    // Long lines: two float4 per thread, kThreads apart, four channels in flight (measured on B200, C = 2..10,
    // 805 MB..2.4 GB per launch: 0.93-0.98 of the copy peak).  Short lines (one tick of a wide bus) do not
    // fill the machine with threads, so each thread keeps 16 channels in flight instead.
    int unroll = 2;
    if (const char* e = getenv("MXL_MIXER_UNROLL")) unroll = atoi(e);
    const uint64_t machine = (uint64_t)(ctx->sm_count > 0 ? ctx->sm_count : 148) * 2048;
    static const bool wide_ok = getenv("MXL_MIXER_NO_WIDE") == nullptr;
    if (wide_ok && p.channels >= 16 && n4 * (uint64_t)p.channels < machine) {
        // one tick of a wide bus: all products in parallel, then the ordered sum (mixer_wide_kernel)
        const unsigned g = (unsigned)((n4 + kMixWideVec - 1) / kMixWideVec);
        MXL_TIMED(ctx, "mixer_kernel");
        launch_chained(ctx, mixer_wide_kernel, dim3(g ? g : 1), dim3(kThreads), 0, p);
    } else if (n4 < machine && p.channels > 4) {
        unsigned g = blocks_for(n4);
        MXL_TIMED(ctx, "mixer_kernel");
        launch_chained(ctx, mixer_kernel<1, 16>, dim3(g ? g : 1), dim3(kThreads), 0, p);
    } else if (unroll >= 2) {
        unsigned g = blocks_for((n4 + 1) / 2);
        MXL_TIMED(ctx, "mixer_kernel");
        launch_chained(ctx, mixer_kernel<2, 4>, dim3(g ? g : 1), dim3(kThreads), 0, p);
    } else {
        unsigned g = blocks_for(n4);
        MXL_TIMED(ctx, "mixer_kernel");
        launch_chained(ctx, mixer_kernel<1, 4>, dim3(g ? g : 1), dim3(kThreads), 0, p);
    }
    return check_launch(ctx, "mixer_kernel");
}

int launch_amplifier(mxl_ctx* ctx, const AmpBatch& b)
{
    MXL_REQUIRE_DEVICE(ctx);
    if (b.n <= 0 || b.frames == 0) return MXL_OK;
    const uint64_t n4 = b.frames >> 1;
    unsigned gx = blocks_for((n4 + kAmpUnroll - 1) / kAmpUnroll);
    dim3 grid(gx ? gx : 1, b.n);
    MXL_TIMED(ctx, "amplifier_kernel");
    launch_chained(ctx, amplifier_kernel, grid, dim3(kThreads), 0, b);
    return check_launch(ctx, "amplifier_kernel");
}

int launch_panner(mxl_ctx* ctx, const PanBatch& b)
{
    MXL_REQUIRE_DEVICE(ctx);
    if (b.n <= 0 || b.frames == 0) return MXL_OK;
    dim3 grid(blocks_for((b.frames + 3) / 4), b.n);
    MXL_TIMED(ctx, "panner_kernel");
    launch_chained(ctx, panner_kernel, grid, dim3(kThreads), 0, b);
    return check_launch(ctx, "panner_kernel");
}

int launch_splitter(mxl_ctx* ctx, const SplitBatch& b)
{
    MXL_REQUIRE_DEVICE(ctx);
    if (b.n <= 0 || b.frames == 0) return MXL_OK;
    dim3 grid(blocks_for((b.frames + 3) / 4), b.n);
    MXL_TIMED(ctx, "splitter_kernel");
    launch_chained(ctx, splitter_kernel, grid, dim3(kThreads), 0, b);
    return check_launch(ctx, "splitter_kernel");
}

int launch_fill(mxl_ctx* ctx, const FillBatch& b)
{
    MXL_REQUIRE_DEVICE(ctx);
    if (b.n <= 0 || b.len == 0) return MXL_OK;
    dim3 grid(blocks_for((b.len + 3) / 4), b.n);
    MXL_TIMED(ctx, "fill_kernel");
    fill_kernel<<<grid, kThreads, 0, ctx->stream>>>(b);
    return check_launch(ctx, "fill_kernel");
}

int launch_meter(mxl_ctx* ctx, const MeterBatch& b, uint32_t n_slots)
{
    MXL_REQUIRE_DEVICE(ctx);
    if (b.n <= 0 || n_slots == 0) return MXL_OK;
    MXL_TIMED(ctx, "meter_kernel");
    const uint64_t machine = (uint64_t)(ctx->sm_count > 0 ? ctx->sm_count : 148) * 16;     // warps that fill every SM's schedulers 4 deep
    // Even ticks always take the warp kernel: its summation order (lane-strided vectors, then the xor tree) is the
    // one the fused voice kernel reproduces, so a graph gives the same meter bits fused or staged, one tick or many.
    if ((uint64_t)n_slots * b.n >= machine || (b.spt & 1u) == 0) {
        dim3 grid((n_slots + kMeterWarps - 1) / kMeterWarps, b.n);
        launch_chained(ctx, meter_warp_kernel, grid, dim3(kMeterThreads), 0, b, n_slots);
    } else {
        dim3 grid(n_slots, b.n);
        launch_chained(ctx, meter_block_kernel, grid, dim3(kMeterThreads), 0, b);
    }
    return check_launch(ctx, "meter_kernel");
}

int launch_route(mxl_ctx* ctx, const RouteLaunch& p)
{
    MXL_REQUIRE_DEVICE(ctx);
    if (p.frames == 0) return MXL_OK;
    const unsigned g = blocks_for(p.frames);
    MXL_TIMED(ctx, "route_kernel");
    launch_chained(ctx, route_kernel, dim3(g), dim3(kThreads), 0, p);
    return check_launch(ctx, "route_kernel");
}

int launch_pcm_pack(mxl_ctx* ctx, const float* in, int16_t* out, uint64_t len)
{
    MXL_REQUIRE_DEVICE(ctx);
    if (len == 0) return MXL_OK;
    const unsigned g = blocks_for(((len >> 2) + kPcmUnroll - 1) / kPcmUnroll);
    MXL_TIMED(ctx, "pcm_pack_kernel");
    launch_chained(ctx, pcm_pack_kernel, dim3(g ? g : 1), dim3(kThreads), 0, in, reinterpret_cast<short*>(out), len);
    return check_launch(ctx, "pcm_pack_kernel");
}

int launch_pcm_unpack(mxl_ctx* ctx, const int16_t* in, float* out, uint64_t len)
{
    MXL_REQUIRE_DEVICE(ctx);
    if (len == 0) return MXL_OK;
    const unsigned g = blocks_for(((len >> 2) + kPcmUnroll - 1) / kPcmUnroll);
    MXL_TIMED(ctx, "pcm_unpack_kernel");
    launch_chained(ctx, pcm_unpack_kernel, dim3(g ? g : 1), dim3(kThreads), 0, reinterpret_cast<const short*>(in), out, len);
    return check_launch(ctx, "pcm_unpack_kernel");
}

int launch_fill_bytes(mxl_ctx* ctx, void* dst, size_t bytes, uint8_t value)
{
    MXL_REQUIRE_DEVICE(ctx);
    if (bytes == 0) return MXL_OK;
    uint32_t word = 0x01010101u * value;
    size_t n16 = bytes / 16;
    unsigned blocks = ctx->sm_count > 0 ? ctx->sm_count * 8 : 1184;
    MXL_TIMED(ctx, "fill_bytes_kernel");
    fill_bytes_kernel<<<blocks, kThreads, 0, ctx->stream>>>(reinterpret_cast<uint4*>(dst), n16, word);
    MXL_TRY(check_launch(ctx, "fill_bytes_kernel"));
    if (bytes % 16) MXL_CUDA(cudaMemsetAsync((uint8_t*)dst + n16 * 16, value, bytes % 16, ctx->stream));
    return MXL_OK;
}

}  // namespace k
}  // namespace mxl

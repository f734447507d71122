// fused_voice.cu -- Oscillator -> EqThree -> StereoPanner -> Mixer [-> Meter] as TWO launches.
//
// The reference runs these five modules one after another over freshly allocated buffers
// (src/engine.rs:464-507); the staged back end made that five dependent launches whose lines round-trip
// through HBM (464*S bytes per tick for BASELINE config 2, of which 16*S -- master and cue -- are compulsory).
//
//   fused_voice_kernel   grid (time tiles, voices).  A CTA generates its oscillator's samples (oscillator.rs:65-92;
//      osc_core.cuh) straight into the swizzled EqThree tile in shared memory -- the oscillator line never exists --
//      folding the zero pass of the time-parallel EqThree scheme (eq_stream.cu) into the same loop, then scans, re-runs
//      its chunks exactly (eq_three.rs:66-86 in the reference's operation order), and stores the EqThree line together
//      with the voice's mixer products (f64(y) * gain) as f32 (mixer.rs:59-62), one scratch line per distinct gain.
//   fused_mix_kernel     one CTA per tick (or per slice of frames without a meter), launched behind the first with
//      programmatic dependent launch: every output vector adds the channels' products IN CHANNEL ORDER (mixer.rs:57-68:
//      the f32 accumulation order is the reference's), the panner's interleave (stereo_panner.rs:35-38) is the addressing
//      of that walk; the tick's master samples stay in shared memory and one warp reduces them to the meter record in
//      meter_warp_kernel's order (same bits as the staged meter).  Its inputs are L2-resident.
//
// An earlier version did both in ONE launch (a thread-block cluster per time tile, a CTA per voice, the mix after a
// cluster barrier, through L2 and then through distributed shared memory).  Measured with per-phase clocks
// (mxl_ctx_fused_profile; profiles/r2_fused_phases*.jsonl): the mix phase cost a CTA 11-14 k cycles per tick (4 dependent
// L2 round trips, or 48 KB per tick through the ~20 B/clk DSMEM port, on 8 of 10 CTAs) plus 4-6 k cycles at the barrier
// waiting for the slowest voice (a sine costs 5x a saw), against 8 k for generating and 5 k for filtering; the product
// tiles it needed in shared memory cost a resident CTA per SM.  As its own launch the mix is ~3 us of the whole machine.
//
// Also measured and dropped (r2): CHAINED tiles -- no halo chunks recomputed ahead of a tile's own (at 16-sample chunks the
// halo is 73 of a tile's 256 chunks, 1.4 x the oscillator work); instead a tile hands the aggregates of its last three warps
// to the next tile of its voice through global memory right after its first scan barrier.  Bit-identical results, but
// slower: 7.0-7.3 M against 7.7 M ticks/s at 128 ticks per call, 12.8 M against 15.2 M at 1024 -- the extra block barrier
// and the neighbour's hand-off sit on every CTA's critical path, and the kernel is bound by that path (FP64 pipe 25-40 %
// busy), not by the amount of oscillator arithmetic.
//
// Lines nobody observes (Oscillator outputs, StereoPanner output) are not written.  State, numerics and results are those
// of the staged kernels: same oscillator code, same EqThree scheme (chunk 0 from the stored state, state_out after the
// last chunk), same channel order.  Roofline: FP64 pipe (~22 ops per sine sample + ~46 per EqThree sample + 1 per channel
// product against 64 lanes/clk/SM), not HBM.
#include <stdlib.h>

#include <algorithm>

#include "eq_stream.cuh"
#include "osc_core.cuh"

namespace mxl {
namespace k {

namespace {

using namespace eqs;

__device__ __forceinline__ float mix1(float x, double g) { return (float)((double)x * g); }

__device__ __forceinline__ float2 ldg2(const float* p)
{
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

// shared memory of the voice kernel: [EqThree tile][EqShared]
template <int LC>
__host__ __device__ constexpr size_t tile_bytes() { return (size_t)kT * LC * sizeof(float); }

template <int LC>
__global__ void __launch_bounds__(kT, 3) fused_voice_kernel(const __grid_constant__ FusedVoiceBatch b)
{
    constexpr int VPR = LC / 4;
    // Steady state (b.late_wait): the kernel before this one is the group's mix kernel of the previous call, which still
    // READS the product / EqThree lines this kernel will overwrite, and which itself only started after the previous voice
    // kernel had completed (so the EqThree states this kernel reads are final).  Everything up to the first global store
    // -- generating, the scan, the exact pass, all in shared memory and registers -- runs beside that mix kernel; the
    // dependency wait sits in front of the stores.
    const bool late_wait = b.late_wait != 0;
    if (!late_wait) pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(16) unsigned char fv_smem[];
    float4* tile = reinterpret_cast<float4*>(fv_smem);                              // [256][VPR], swizzled
    EqShared<LC>* sh = reinterpret_cast<EqShared<LC>*>(fv_smem + tile_bytes<LC>());
    const FusedVoice& vc = b.voice[blockIdx.y];
    const EqStreamConsts& q = b.eq;
    const int tid = threadIdx.x;
    unsigned long long* prof = b.prof ? b.prof + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kFusedProfStamps : nullptr;
#define MXL_STAMP(i) do { if (prof && tid == (int)q.halo) prof[i] = clock64(); } while (0)   /* the tile's first owner thread */
    MXL_STAMP(0);
    // Tables (device memory) -> shared memory with cp.async: one round of overlapping copies (as kernel parameters
    // they cost a constant-bank miss per 64-byte line, in program order), landing while the first samples are generated.
    load_tables_async<LC>(q.tab, sh, tid);
    MXL_STAMP(1);
    const int halo = (int)q.halo;
    const int U = (int)b.owned;
    const int64_t c0 = (int64_t)blockIdx.x * U - halo;        // chunk of thread 0
    const int64_t c = c0 + tid;
    const bool active = c >= 0 && c < (int64_t)b.n_chunks && tid < halo + U;
    const double* st = vc.state;
    RowIo<LC> row{tile, tid};

    // ---- A1. oscillator -> tile row, zero pass folded in: v = K + sum_j s_j V_j ----
    // Eight samples per trip, two groups of four sines interleaved.  The first trip's samples are generated before
    // the tables are waited for.
    double v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = 0.0;
    {
        // (t + i) as f64: one conversion per chunk, the samples are exact +1.0 steps below 2^53 (checked by the host)
        double seq0 = active ? (double)(b.t0 + (uint64_t)c * LC) : 0.0;
        const double freq = vc.freq, sr = b.sample_rate, inv_sr = b.inv_sample_rate;
        const int wf = vc.waveform;
        bool tables_in = false;
#pragma unroll 1
        for (int vv = 0; vv < VPR; vv += 2, seq0 += 8.0) {
            float s[8];
            if (active) {
                double n[8];
#pragma unroll
                for (int j = 0; j < 8; j++) n[j] = osc_phase(seq0 + (double)j, sr, inv_sr, freq);
                osc_wave8(wf, n, s);
                row.store(vv, EqF4{s[0], s[1], s[2], s[3]});
                row.store(vv + 1, EqF4{s[4], s[5], s[6], s[7]});
            }
            if (!tables_in) {                              // uniform: first trip only
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncthreads();
                tables_in = true;
                if (active) {
#pragma unroll
                    for (int e = 0; e < 8; e++) v[e] = sh->K[e];
                }
            }
            if (active) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const double sd = (double)s[j];
#pragma unroll
                    for (int e = 0; e < 8; e++) v[e] = fma(sd, sh->V[vv * 4 + j][e], v[e]);
                }
            }
        }
        if (active && c == 0) {                            // v_0 = A p_init + z_0
            const double pl[4] = {st[0], st[1], st[2], st[3]}, ph[4] = {st[4], st[5], st[6], st[7]};
            double yl[4], yh[4];
            tri_apply(sh->pow_lo[0], pl, yl);
            tri_apply(sh->pow_hi[0], ph, yh);
#pragma unroll
            for (int e = 0; e < 4; e++) { v[e] += yl[e]; v[4 + e] += yh[e]; }
        }
    }
    // A non-finite pole in the stored state never leaves the reference's cascades (every later output is NaN);
    // the scan forgets by construction, so every chunk is told.  (The oscillator itself is finite: the host only
    // fuses finite frequencies.)
    bool bad_init;
    {
        const double p0[8] = {st[0], st[1], st[2], st[3], st[4], st[5], st[6], st[7]};
        bad_init = any_non_finite(p0);
    }

    // ---- A2. scan (eq_stream.cuh): two block barriers, the rows are visible after the first ----
    MXL_STAMP(2);
    double S[8];
    scan_start_states<LC>(q, sh, v, S, tid);
    if (bad_init) {
#pragma unroll
        for (int e = 0; e < 8; e++) S[e] = __longlong_as_double(0x7ff8000000000000ll);
    }

    // ---- A3. exact re-run of the owned chunks ----
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
    // observed oscillator lines leave from the tile before the exact pass overwrites it
    if (vc.osc_mono || vc.osc_stereo) {
        if (late_wait) pdl_wait();
        const int total = U * VPR;
        for (int idx = tid; idx < total; idx += kT) {
            const int r = halo + idx / VPR, vv = idx % VPR;
            const int64_t ci = c0 + r;
            if (ci < 0 || ci >= (int64_t)b.n_chunks) continue;
            const uint64_t g = (uint64_t)ci * LC + vv * 4;
            const float4 y = tile[slot_of<VPR>(r, vv)];
            const float ys[4] = {y.x, y.y, y.z, y.w};
            for (int j = 0; j < 4; j++) {
                if (g + j >= b.frames) break;
                if (vc.osc_mono) vc.osc_mono[g + j] = ys[j];
                if (vc.osc_stereo) { vc.osc_stereo[2 * (g + j)] = ys[j]; vc.osc_stereo[2 * (g + j) + 1] = ys[j]; }
            }
        }
    }
    __syncthreads();                                       // every history / tile read precedes any overwrite
    MXL_STAMP(3);
    if (owner) {
        const EqGains g{q.c_lo, q.c_hi, vc.g_lo, vc.g_mid, vc.g_hi};
        if (count == LC) {
            eq_run_chunk_skewed<LC, true>(p, hist, row, g);
        } else {                                           // ragged end of the call: one thread, sequential form
            float* mine = reinterpret_cast<float*>(tile);
            for (uint32_t j = 0; j < count; j++) {
                float* cell = mine + slot_of<VPR>(tid, j >> 2) * 4 + (j & 3);
                *cell = eq_step_seq(p, hist, *cell, g);
            }
        }
        if (c + 1 == (int64_t)b.n_chunks) {                // state after this call
            double* so = vc.state_out;
            so[0] = p.l0; so[1] = p.l1; so[2] = p.l2; so[3] = p.l3;
            so[4] = p.h0; so[5] = p.h1; so[6] = p.h2; so[7] = p.h3;
            so[8] = hist[0]; so[9] = hist[1]; so[10] = hist[2];
        }
    }
    __syncthreads();
    MXL_STAMP(4);

    // ---- A4. coalesced store of the owned rows: the EqThree line and this voice's mixer products ----
    if (late_wait) pdl_wait();                             // (a second wait returns at once)
    {
        float* dst = vc.eq_out;
        bool vec_ok = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
        for (int pi = 0; pi < vc.n_products; pi++) vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(vc.product_out[pi]) & 15) == 0;
        const int total = U * VPR;
        for (int idx = tid; idx < total; idx += kT) {
            const int r = halo + idx / VPR, vv = idx % VPR;
            const int64_t ci = c0 + r;
            if (ci < 0 || ci >= (int64_t)b.n_chunks) continue;
            const uint64_t g = (uint64_t)ci * LC + vv * 4;
            const float4 y = tile[slot_of<VPR>(r, vv)];
            if (g + 4 <= b.frames && vec_ok) {
                *reinterpret_cast<float4*>(dst + g) = y;
                for (int pi = 0; pi < vc.n_products; pi++) {
                    const double gain = vc.product_gain[pi];                        // mixer.rs:59-62
                    *reinterpret_cast<float4*>(vc.product_out[pi] + g) = make_float4(mix1(y.x, gain), mix1(y.y, gain), mix1(y.z, gain), mix1(y.w, gain));
                }
            } else {
                const float ys[4] = {y.x, y.y, y.z, y.w};
                for (int j = 0; j < 4; j++) {
                    if (g + j >= b.frames) break;
                    dst[g + j] = ys[j];
                    for (int pi = 0; pi < vc.n_products; pi++) vc.product_out[pi][g + j] = mix1(ys[j], vc.product_gain[pi]);
                }
            }
        }
    }
    MXL_STAMP(5);
#undef MXL_STAMP
}

// ------------------------------------------------------------------------------------------
// fused_mix_kernel: Mixer (+ StereoPanner addressing, + Meter) over the voices' product lines
// ------------------------------------------------------------------------------------------
constexpr int kMixThreads = 512; // one frame pair per thread: a tick of 800 frames is one round of loads
constexpr int kMixCU = 10;       // channels whose loads are in flight together

// master / cue of the frame pair at f (even): the channels' products added IN CHANNEL ORDER (mixer.rs:57-68), EqThree
// samples for the cue bus; the panner outputs of observed channels are written on the way.
__device__ __forceinline__ void mix_pair(const FusedChan* __restrict__ chan, int n_channels, uint64_t f, float4& m, float4& cu)
{
    m = make_float4(0.f, 0.f, 0.f, 0.f);                    // util::zero(master), util::zero(cue) (mixer.rs:54-55)
    cu = m;
    for (int ch0 = 0; ch0 < n_channels; ch0 += kMixCU) {
        float2 L[kMixCU], R[kMixCU], RL[kMixCU], RR[kMixCU];
#pragma unroll
        for (int k = 0; k < kMixCU; k++) {
            RL[k] = make_float2(0.f, 0.f);
            RR[k] = RL[k];
            L[k] = RL[k];
            R[k] = RL[k];
            if (ch0 + k < n_channels) {
                const FusedChan& c = chan[ch0 + k];
                L[k] = make_float2(c.zero_product, c.zero_product);
                R[k] = L[k];
                if (c.left) L[k] = ldg2(c.left + f);
                if (c.right) R[k] = (c.right == c.left) ? L[k] : ldg2(c.right + f);
                if (c.left_raw) RL[k] = ldg2(c.left_raw + f);
                if (c.right_raw) RR[k] = (c.right_raw == c.left_raw) ? RL[k] : ldg2(c.right_raw + f);
            }
        }
#pragma unroll
        for (int k = 0; k < kMixCU; k++) {
            if (ch0 + k >= n_channels) break;
            const FusedChan& c = chan[ch0 + k];
            m.x += L[k].x; m.y += R[k].x; m.z += L[k].y; m.w += R[k].y;
            if (c.cue) { cu.x += RL[k].x; cu.y += RR[k].x; cu.z += RL[k].y; cu.w += RR[k].y; }
            if (c.pan_out) *reinterpret_cast<float4*>(c.pan_out + 2 * f) = make_float4(RL[k].x, RR[k].x, RL[k].y, RR[k].y);
        }
    }
}

// The same when every channel's panner takes ONE voice on both sides (BASELINE configs 2 and 4): half the loads, half the
// registers (two CTAs per SM), left and right sums share their addends.
__device__ __forceinline__ void mix_pair_mono(const FusedChan* __restrict__ chan, int n_channels, uint64_t f, float4& m, float4& cu)
{
    float2 ms = make_float2(0.f, 0.f), cs = ms;
    for (int ch0 = 0; ch0 < n_channels; ch0 += kMixCU) {
        float2 P[kMixCU], W[kMixCU];
#pragma unroll
        for (int k = 0; k < kMixCU; k++) {
            P[k] = make_float2(0.f, 0.f);
            W[k] = P[k];
            if (ch0 + k < n_channels) {
                const FusedChan& c = chan[ch0 + k];
                P[k] = make_float2(c.zero_product, c.zero_product);
                if (c.left) P[k] = ldg2(c.left + f);
                if (c.left_raw) W[k] = ldg2(c.left_raw + f);
            }
        }
#pragma unroll
        for (int k = 0; k < kMixCU; k++) {
            if (ch0 + k >= n_channels) break;
            const FusedChan& c = chan[ch0 + k];
            ms.x += P[k].x; ms.y += P[k].y;
            if (c.cue) { cs.x += W[k].x; cs.y += W[k].y; }
            if (c.pan_out) *reinterpret_cast<float4*>(c.pan_out + 2 * f) = make_float4(W[k].x, W[k].x, W[k].y, W[k].y);
        }
    }
    m = make_float4(ms.x, ms.x, ms.y, ms.y);                // identical addends in identical order: left == right, bit for bit
    cu = make_float4(cs.x, cs.x, cs.y, cs.y);
}

// the odd last frame of a call (one frame, scalar)
__device__ __forceinline__ void mix_frame(const FusedChan* __restrict__ chan, int n_channels, uint64_t f, float2& m, float2& cu)
{
    m = make_float2(0.f, 0.f);
    cu = m;
    for (int ch = 0; ch < n_channels; ch++) {
        const FusedChan& c = chan[ch];
        m.x += c.left ? __ldg(c.left + f) : c.zero_product;
        m.y += c.right ? __ldg(c.right + f) : c.zero_product;
        const float l = c.left_raw ? __ldg(c.left_raw + f) : 0.f, r = c.right_raw ? __ldg(c.right_raw + f) : 0.f;
        if (c.cue) { cu.x += l; cu.y += r; }
        if (c.pan_out) { c.pan_out[2 * f] = l; c.pan_out[2 * f + 1] = r; }
    }
}

// blockIdx.x = segment of b.spt frames (a tick when the meter rides along).  Dynamic shared memory: the channel table
// and, with a meter, the segment's master samples.
template <bool MONO>
__global__ void __launch_bounds__(kMixThreads, MONO ? 2 : 1) fused_mix_kernel(const __grid_constant__ FusedMixBatch b)
{
    extern __shared__ __align__(16) unsigned char fm_smem[];
    FusedChan* s_chan = reinterpret_cast<FusedChan*>(fm_smem);
    float4* seg = reinterpret_cast<float4*>(s_chan + kFusedMaxChans);
    const int tid = threadIdx.x;
    {   // the channel table (device memory, written by the host before the voice kernel was launched) -> shared memory:
        // does not depend on the voice kernel, so it runs before the grid dependency wait
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(b.chan);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(s_chan);
        const int words = b.n_channels * (int)(sizeof(FusedChan) / 8);
        for (int i = tid; i < words; i += kMixThreads) dst[i] = src[i];
    }
    __syncthreads();
    pdl_prologue();                                        // the voices' lines are complete and visible from here on
    const uint64_t fb = (uint64_t)blockIdx.x * b.spt;
    uint64_t fe = fb + b.spt;
    if (fe > b.frames) fe = b.frames;
    if (fb >= fe) return;
    const uint32_t npairs = (uint32_t)((fe - fb) >> 1);
    for (uint32_t pi = tid; pi < npairs; pi += kMixThreads) {
        float4 m, cu;
        const uint64_t f = fb + 2ull * pi;
        if (MONO) mix_pair_mono(s_chan, b.n_channels, f, m, cu);
        else mix_pair(s_chan, b.n_channels, f, m, cu);
        *reinterpret_cast<float4*>(b.master + 2 * f) = m;
        *reinterpret_cast<float4*>(b.cue + 2 * f) = cu;
        if (b.meter) seg[pi] = m;
    }
    if (((fe - fb) & 1) && tid == 0) {                     // odd frame count: last frame
        float2 m, cu;
        mix_frame(s_chan, b.n_channels, fe - 1, m, cu);
        b.master[2 * (fe - 1)] = m.x; b.master[2 * (fe - 1) + 1] = m.y;
        b.cue[2 * (fe - 1)] = cu.x; b.cue[2 * (fe - 1) + 1] = cu.y;
        if (b.meter) seg[npairs] = make_float4(m.x, m.y, 0.f, 0.f);
    }
    if (!b.meter) return;
    __syncthreads();
    if (tid < 32) {                                        // meter_warp_kernel's order: lane-strided vectors, xor tree
        float pk0 = 0.f, pk1 = 0.f;
        double sq0 = 0.0, sq1 = 0.0;
        for (uint32_t pi = tid; pi < npairs; pi += 32) {
            const float4 s = seg[pi];
            pk0 = fmaxf(pk0, fmaxf(fabsf(s.x), fabsf(s.z)));
            pk1 = fmaxf(pk1, fmaxf(fabsf(s.y), fabsf(s.w)));
            sq0 += (double)s.x * (double)s.x;
            sq1 += (double)s.y * (double)s.y;
            sq0 += (double)s.z * (double)s.z;
            sq1 += (double)s.w * (double)s.w;
        }
        if (((fe - fb) & 1) && tid == 0) {
            const float4 t = seg[npairs];
            pk0 = fmaxf(pk0, fabsf(t.x)); pk1 = fmaxf(pk1, fabsf(t.y));
            sq0 += (double)t.x * (double)t.x; sq1 += (double)t.y * (double)t.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pk0 = fmaxf(pk0, __shfl_xor_sync(0xffffffffu, pk0, o));
            pk1 = fmaxf(pk1, __shfl_xor_sync(0xffffffffu, pk1, o));
            sq0 += __shfl_xor_sync(0xffffffffu, sq0, o);
            sq1 += __shfl_xor_sync(0xffffffffu, sq1, o);
        }
        if (tid == 0) {
            MeterRecord r;
            r.peak[0] = pk0; r.peak[1] = pk1; r.sumsq[0] = sq0; r.sumsq[1] = sq1; r._pad = 0;
            r.clip = (pk0 > 1.0f || pk1 > 1.0f) ? 1 : 0;   // output_device.rs:192-194
            b.meter[blockIdx.x] = r;
        }
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
template <int LC>
int launch_voice_lc(mxl_ctx* ctx, FusedVoiceBatch& b)
{
    const size_t smem = tile_bytes<LC>() + sizeof(EqShared<LC>);
    const uint32_t bit = 1u << (8 + LC / 16);
    if (smem > 48 * 1024 && !(ctx->eq_stream_smem_set & bit)) {
        MXL_CUDA(cudaFuncSetAttribute(fused_voice_kernel<LC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->eq_stream_smem_set |= bit;
    }
    const uint32_t tiles = (b.n_chunks + b.owned - 1) / b.owned;
    b.prof = nullptr;
    // late dependency wait: only behind this group's own mix kernel (the host's steady-state path sets group_tag)
    b.late_wait = (b.late_wait && ctx->launches == ctx->fused_mix_launch_no && ctx->fused_group_now && ctx->fused_mix_group == ctx->fused_group_now &&
                   !ctx->fused_prof_cap && pdl_enabled(ctx)) ? 1 : 0;
    if (ctx->fused_prof_cap) {                             // diagnostics on: phase clocks of this launch
        const uint32_t ctas = tiles * (uint32_t)b.n_voices;
        if (ctas <= ctx->fused_prof_cap) { b.prof = ctx->fused_prof; ctx->fused_prof_ctas = ctas; }
    }
    MXL_TIMED(ctx, "fused_voice_kernel");
    launch_chained(ctx, fused_voice_kernel<LC>, dim3(tiles, b.n_voices), dim3(kT), smem, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) MXL_FAIL(MXL_ERR_CUDA, "launch of fused_voice_kernel<%d> failed: %s", LC, cudaGetErrorString(e));
    ctx->launches++;
    return MXL_OK;
}

}  // namespace

int launch_fused_voice(mxl_ctx* ctx, FusedVoiceBatch& b)
{
    if (!ctx || !ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_TRY(ctx->activate());
    if (b.frames == 0) return MXL_OK;
    const EqStreamConsts& q = b.eq;
    if (b.n_voices < 1 || b.n_voices > kFusedMaxVoices) MXL_FAIL(MXL_ERR_INVALID, "fused_voice_kernel: %d voices", b.n_voices);
    if (q.halo == 0 || q.halo > (uint32_t)kT / 2 || q.lev_lo > (uint32_t)kEqPlanLevels || q.lev_hi > (uint32_t)kEqPlanLevels ||
        q.back_lo > 3 || q.back_hi > 3 || !q.tab || b.owned == 0 || b.owned + q.halo > (uint32_t)kT)
        MXL_FAIL(MXL_ERR_INVALID, "fused_voice_kernel: bad plan (chunk %u, halo %u, owned %u)", q.chunk, q.halo, b.owned);
    switch (q.chunk) {
    case 16: return launch_voice_lc<16>(ctx, b);
    case 32: return launch_voice_lc<32>(ctx, b);
    case 64: return launch_voice_lc<64>(ctx, b);
    default: MXL_FAIL(MXL_ERR_INVALID, "fused_voice_kernel: unsupported chunk length %u", q.chunk);
    }
}

int launch_fused_mix(mxl_ctx* ctx, const FusedMixBatch& b)
{
    if (!ctx || !ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_TRY(ctx->activate());
    if (b.frames == 0) return MXL_OK;
    if (b.n_channels < 1 || b.n_channels > kFusedMaxChans || b.spt == 0 || (b.spt & 1u) || !b.chan)
        MXL_FAIL(MXL_ERR_INVALID, "fused_mix_kernel: %d channels, segments of %u frames", b.n_channels, b.spt);
    const size_t smem = (size_t)kFusedMaxChans * sizeof(FusedChan) + (b.meter ? (size_t)(b.spt / 2 + 1) * sizeof(float4) : 0);
    if (smem > 48 * 1024) MXL_FAIL(MXL_ERR_INVALID, "fused_mix_kernel: a tick of %u frames does not fit its shared-memory segment", b.spt);
    const unsigned grid = (unsigned)((b.frames + b.spt - 1) / b.spt);
    MXL_TIMED(ctx, "fused_mix_kernel");
    if (b.mono) launch_chained(ctx, fused_mix_kernel<true>, dim3(grid), dim3(kMixThreads), smem, b);
    else launch_chained(ctx, fused_mix_kernel<false>, dim3(grid), dim3(kMixThreads), smem, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) MXL_FAIL(MXL_ERR_CUDA, "launch of fused_mix_kernel failed: %s", cudaGetErrorString(e));
    ctx->launches++;
    ctx->fused_mix_launch_no = ctx->launches;
    ctx->fused_mix_group = ctx->fused_group_now;
    return MXL_OK;
}

}  // namespace k
}  // namespace mxl

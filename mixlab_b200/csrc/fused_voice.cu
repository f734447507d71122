// fused_voice.cu -- Oscillator -> EqThree -> StereoPanner -> Mixer [-> Meter] as ONE launch.
//
// The reference runs these five modules one after another over freshly allocated buffers
// (src/engine.rs:464-507); the staged back end made that five dependent launches whose lines round-trip
// through HBM (464*S bytes per tick for BASELINE config 2, of which 16*S -- master and cue -- are compulsory).
// Here the whole group is one kernel:
//
//   grid (time tiles, voices), one thread-block CLUSTER per time tile, a CTA per voice;
//   A. each CTA generates its oscillator's samples (oscillator.rs:65-92; osc_core.cuh) straight into the
//      swizzled EqThree tile in shared memory -- the oscillator line never exists -- folding the zero pass of
//      the time-parallel EqThree scheme (eq_stream.cu) into the same loop, then scans, re-runs its chunks
//      exactly (eq_three.rs:66-86 in the reference's operation order) and stores the EqThree line;
//   B. after the cluster barrier (release / acquire at cluster scope) the CTAs of the cluster share the tile's
//      mixer sum: every output vector walks the channels IN ORDER (mixer.rs:57-68: the f32 accumulation order
//      is the reference's), the panner's interleave (stereo_panner.rs:35-38) is the addressing of that walk;
//      when the tile is a whole number of ticks each tick is mixed by one CTA, kept in shared memory and
//      reduced to its meter record by one warp in meter_warp_kernel's order (same bits as the staged meter).
//
// Lines nobody observes (Oscillator outputs, StereoPanner output) are not written; the EqThree lines are (they
// carry the samples from phase A to phase B through L2).  State, numerics and results are those of the staged
// kernels: same oscillator code, same EqThree scheme (chunk 0 from the stored state, state_out after the last
// chunk), same channel order.  Roofline: FP64 pipe (~22 ops per sine sample + ~46 per EqThree sample + 1 per
// channel product against 64 lanes/clk/SM), not HBM.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "eq_stream.cuh"
#include "osc_core.cuh"

namespace cg = cooperative_groups;

namespace mxl {
namespace k {

namespace {

using namespace eqs;

__device__ __forceinline__ float2 ldcg2(const float* p)
{
    float2 v;
    asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float mix1(float x, double g) { return (float)((double)x * g); }

constexpr int kMixCU = 8;        // channels whose loads are in flight together

// master / cue of the two frames at f (even), channels walked in order (mixer.rs:57-68); the panner outputs of
// observed channels are written on the way.
__device__ __forceinline__ void mix_pair(const FusedChan* __restrict__ chan, int n_channels, uint64_t f, float4& m, float4& cu)
{
    m = make_float4(0.f, 0.f, 0.f, 0.f);                    // util::zero(master), util::zero(cue) (mixer.rs:54-55)
    cu = m;
    for (int ch0 = 0; ch0 < n_channels; ch0 += kMixCU) {
        float2 L[kMixCU], R[kMixCU];
#pragma unroll
        for (int k = 0; k < kMixCU; k++) {
            L[k] = make_float2(0.f, 0.f);
            R[k] = L[k];
            if (ch0 + k < n_channels) {
                const float* l = chan[ch0 + k].left;
                const float* r = chan[ch0 + k].right;
                if (l) L[k] = ldcg2(l + f);
                if (r) R[k] = (r == l) ? L[k] : ldcg2(r + f);
            }
        }
#pragma unroll
        for (int k = 0; k < kMixCU; k++) {
            if (ch0 + k >= n_channels) break;
            const FusedChan& c = chan[ch0 + k];
            const double g = c.gain;
            float lx = mix1(L[k].x, g), ly = mix1(L[k].y, g), rx, ry;
            if (c.left == c.right) { rx = lx; ry = ly; }
            else { rx = mix1(R[k].x, g); ry = mix1(R[k].y, g); }
            m.x += lx; m.y += rx; m.z += ly; m.w += ry;
            if (c.cue) { cu.x += L[k].x; cu.y += R[k].x; cu.z += L[k].y; cu.w += R[k].y; }
            if (c.pan_out) *reinterpret_cast<float4*>(c.pan_out + 2 * f) = make_float4(L[k].x, R[k].x, L[k].y, R[k].y);
        }
    }
}

// the odd last frame of a call (one frame, scalar)
__device__ __forceinline__ void mix_frame(const FusedChan* __restrict__ chan, int n_channels, uint64_t f, float2& m, float2& cu)
{
    m = make_float2(0.f, 0.f);
    cu = m;
    for (int ch = 0; ch < n_channels; ch++) {
        const FusedChan& c = chan[ch];
        const float l = c.left ? __ldcg(c.left + f) : 0.f, r = c.right ? __ldcg(c.right + f) : 0.f;
        m.x += mix1(l, c.gain); m.y += mix1(r, c.gain);
        if (c.cue) { cu.x += l; cu.y += r; }
        if (c.pan_out) { c.pan_out[2 * f] = l; c.pan_out[2 * f + 1] = r; }
    }
}

template <int LC>
__global__ void __launch_bounds__(kT) fused_voice_mix_kernel(const __grid_constant__ FusedBatch b)
{
    constexpr int VPR = LC / 4;
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char fv_smem[];
    float4* tile = reinterpret_cast<float4*>(fv_smem);                              // [256][VPR], swizzled
    EqShared<LC>* sh = reinterpret_cast<EqShared<LC>*>(fv_smem + (size_t)kT * LC * sizeof(float));
    FusedChan* s_chan = reinterpret_cast<FusedChan*>(sh + 1);                       // [n_channels]
    const FusedVoice vc = b.voice[blockIdx.y];
    const EqStreamConsts& q = b.eq;
    const int tid = threadIdx.x;
    // Tables (device memory) and the channel table (kernel parameters) -> shared memory, a word per thread: one
    // round of overlapping loads instead of a constant-bank miss per table line in program order.
    load_tables<LC>(q.tab, sh, tid);
    {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(b.chan);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(s_chan);
        const int words = b.n_channels * (int)(sizeof(FusedChan) / 8);
        for (int i = tid; i < words; i += kT) dst[i] = src[i];
    }
    __syncthreads();
    const int halo = (int)q.halo;
    const int U = (int)b.owned;
    const int64_t c0 = (int64_t)blockIdx.x * U - halo;        // chunk of thread 0
    const int64_t c = c0 + tid;
    const bool active = c >= 0 && c < (int64_t)b.n_chunks && tid < halo + U;
    const double* st = vc.state;
    RowIo<LC> row{tile, tid};

    // ---- A1. oscillator -> tile row, zero pass folded in: v = K + sum_j s_j V_j ----
    double v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = 0.0;
    if (active) {
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = sh->K[e];
        // (t + i) as f64: one conversion per chunk, the samples are exact +1.0 steps below 2^53 (checked by the host)
        const double base = (double)(b.t0 + (uint64_t)c * LC);
        const double freq = vc.freq, sr = b.sample_rate, inv_sr = b.inv_sample_rate;
        const int wf = vc.waveform;
        // a real loop (four samples per trip): unrolled it is VPR copies of the sine code and the warps of the CTA
        // stall on instruction fetch
#pragma unroll 1
        for (int vv = 0; vv < VPR; vv++) {
            double n[4];
#pragma unroll
            for (int j = 0; j < 4; j++) n[j] = osc_phase(base + (double)(vv * 4 + j), sr, inv_sr, freq);
            float s[4];
            osc_wave4(wf, n, s);
            row.store(vv, EqF4{s[0], s[1], s[2], s[3]});
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const double sd = (double)s[j];
#pragma unroll
                for (int e = 0; e < 8; e++) v[e] = fma(sd, sh->V[vv * 4 + j][e], v[e]);
            }
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
    // A non-finite pole in the stored state never leaves the reference's cascades (every later output is NaN);
    // the scan forgets by construction, so every chunk is told.  (The oscillator itself is finite: the host only
    // fuses finite frequencies.)
    bool bad_init;
    {
        const double p0[8] = {st[0], st[1], st[2], st[3], st[4], st[5], st[6], st[7]};
        bad_init = any_non_finite(p0);
    }

    // ---- A2. scan (eq_stream.cuh): two block barriers, the rows are visible after the first ----
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
    if (owner) {
        const EqGains g{q.c_lo, q.c_hi, vc.g_lo, vc.g_mid, vc.g_hi};
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
            double* so = vc.state_out;
            so[0] = p.l0; so[1] = p.l1; so[2] = p.l2; so[3] = p.l3;
            so[4] = p.h0; so[5] = p.h1; so[6] = p.h2; so[7] = p.h3;
            so[8] = hist[0]; so[9] = hist[1]; so[10] = hist[2];
        }
    }
    __syncthreads();

    // ---- A4. coalesced store of the owned rows: the EqThree line ----
    {
        float* dst = vc.eq_out;
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

    // ---- B. every voice of this time tile is in L2: mix it, CTAs of the cluster sharing the work ----
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();                                        // barrier.cluster arrive.release / wait.acquire
    const uint32_t rank = cluster.block_rank(), n_rank = cluster.num_blocks();
    const uint64_t f_begin = (uint64_t)blockIdx.x * U * LC;
    uint64_t f_end = f_begin + (uint64_t)U * LC;
    if (f_end > b.frames) f_end = b.frames;
    if (f_begin >= f_end) return;

    if (b.meter) {
        // the tile is a whole number of ticks: one tick per CTA at a time; the tick's master samples stay in
        // shared memory (the tile is free now) for the meter warp
        float4* seg = tile;
        const uint32_t spt = b.spt;
        const uint32_t n_slots = (uint32_t)((f_end - f_begin + spt - 1) / spt);
        for (uint32_t sidx = rank; sidx < n_slots; sidx += n_rank) {
            const uint64_t fb = f_begin + (uint64_t)sidx * spt;
            uint64_t fe = fb + spt;
            if (fe > f_end) fe = f_end;
            const uint32_t npairs = (uint32_t)((fe - fb) >> 1);
            __syncthreads();                               // the previous tick's meter warp is done with seg
            for (uint32_t pi = tid; pi < npairs; pi += kT) {
                float4 m, cu;
                const uint64_t f = fb + 2ull * pi;
                mix_pair(s_chan, b.n_channels, f, m, cu);
                *reinterpret_cast<float4*>(b.master + 2 * f) = m;
                *reinterpret_cast<float4*>(b.cue + 2 * f) = cu;
                seg[pi] = m;
            }
            if (((fe - fb) & 1) && tid == 0) {             // odd frame count: last frame
                float2 m, cu;
                mix_frame(s_chan, b.n_channels, fe - 1, m, cu);
                b.master[2 * (fe - 1)] = m.x; b.master[2 * (fe - 1) + 1] = m.y;
                b.cue[2 * (fe - 1)] = cu.x; b.cue[2 * (fe - 1) + 1] = cu.y;
                seg[npairs] = make_float4(m.x, m.y, 0.f, 0.f);
            }
            __syncthreads();
            if (tid < 32) {                                // meter_warp_kernel's order: lane-strided vectors, xor tree
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
                    b.meter[fb / spt] = r;
                }
            }
        }
    } else {
        // no meter in the group: the tile's frame pairs split evenly over the cluster
        const uint64_t pairs = (f_end - f_begin) >> 1;
        const uint64_t per = (pairs + n_rank - 1) / n_rank;
        const uint64_t p_lo = (uint64_t)rank * per;
        uint64_t p_hi = p_lo + per;
        if (p_hi > pairs) p_hi = pairs;
        for (uint64_t pi = p_lo + tid; pi < p_hi; pi += kT) {
            float4 m, cu;
            const uint64_t f = f_begin + 2ull * pi;
            mix_pair(s_chan, b.n_channels, f, m, cu);
            *reinterpret_cast<float4*>(b.master + 2 * f) = m;
            *reinterpret_cast<float4*>(b.cue + 2 * f) = cu;
        }
        if (((f_end - f_begin) & 1) && rank == 0 && tid == 0) {
            float2 m, cu;
            mix_frame(s_chan, b.n_channels, f_end - 1, m, cu);
            b.master[2 * (f_end - 1)] = m.x; b.master[2 * (f_end - 1) + 1] = m.y;
            b.cue[2 * (f_end - 1)] = cu.x; b.cue[2 * (f_end - 1) + 1] = cu.y;
        }
    }
}

template <int LC>
size_t smem_bytes() { return (size_t)kT * LC * sizeof(float) + sizeof(EqShared<LC>) + (size_t)kFusedMaxChans * sizeof(FusedChan); }

template <int LC>
int configure(mxl_ctx* ctx)
{
    const uint32_t bit = 1u << (8 + LC / 16);
    if (ctx->eq_stream_smem_set & bit) return MXL_OK;
    MXL_CUDA(cudaFuncSetAttribute(fused_voice_mix_kernel<LC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<LC>()));
    MXL_CUDA(cudaFuncSetAttribute(fused_voice_mix_kernel<LC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    ctx->eq_stream_smem_set |= bit;
    return MXL_OK;
}

template <int LC>
void fill_config(const mxl_ctx* ctx, cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, dim3 grid, int n_voices, bool pdl)
{
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = grid; cfg->blockDim = dim3(kT); cfg->dynamicSmemBytes = smem_bytes<LC>(); cfg->stream = ctx->stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = (unsigned)n_voices; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg->attrs = attr;
    cfg->numAttrs = pdl ? 2 : 1;
}

template <int LC>
int launch_lc(mxl_ctx* ctx, const FusedBatch& b)
{
    MXL_TRY(configure<LC>(ctx));
    const uint32_t tiles = (b.n_chunks + b.owned - 1) / b.owned;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    fill_config<LC>(ctx, &cfg, attr, dim3(tiles, b.n_voices), b.n_voices, pdl_enabled(ctx));
    MXL_TIMED(ctx, "fused_voice_mix_kernel");
    cudaError_t e = cudaLaunchKernelEx(&cfg, fused_voice_mix_kernel<LC>, b);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) MXL_FAIL(MXL_ERR_CUDA, "launch of fused_voice_mix_kernel<%d> (cluster of %d) failed: %s", LC, b.n_voices, cudaGetErrorString(e));
    ctx->launches++;
    return MXL_OK;
}

template <int LC>
int supported_lc(mxl_ctx* ctx, int n_voices)
{
    if (configure<LC>(ctx) != MXL_OK) return 0;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    fill_config<LC>(ctx, &cfg, attr, dim3(1, n_voices), n_voices, false);
    int clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&clusters, fused_voice_mix_kernel<LC>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return clusters;
}

}  // namespace

uint32_t fused_owned_chunks(const EqStreamConsts& eq, uint32_t spt, bool* meter_ok)
{
    const uint32_t room = (uint32_t)kT - eq.halo;
    *meter_ok = false;
    if (spt == 0 || (spt & 1)) return room;                           // odd ticks: vector pairs would straddle ticks
    // chunks per whole number of ticks: multiples of spt / gcd(spt, LC)
    uint32_t a = spt, c = eq.chunk;
    while (c) { const uint32_t t = a % c; a = c; c = t; }
    const uint32_t step = spt / a;
    if (step == 0 || step > room) return room;
    if ((size_t)(spt / 2 + 1) * sizeof(float4) > (size_t)kT * eq.chunk * sizeof(float)) return room;   // a tick must fit the tile's shared memory
    *meter_ok = true;
    return room / step * step;
}

int fused_voice_mix_supported(mxl_ctx* ctx, uint32_t chunk, int n_voices)
{
    if (!ctx || !ctx->has_device() || n_voices < 1 || n_voices > kFusedMaxVoices) return 0;
    const uint32_t key = (chunk << 8) | (uint32_t)n_voices;
    auto it = ctx->fused_clusters.find(key);
    if (it != ctx->fused_clusters.end()) return it->second;
    if (ctx->activate() != MXL_OK) return 0;
    int clusters = 0;
    switch (chunk) {
    case 32: clusters = supported_lc<32>(ctx, n_voices); break;
    case 64: clusters = supported_lc<64>(ctx, n_voices); break;
    default: break;
    }
    if (getenv("MXL_DEBUG")) fprintf(stderr, "[mxl] fused_voice_mix_kernel<%u>: cluster of %d CTAs -> %d clusters resident\n", chunk, n_voices, clusters);
    return ctx->fused_clusters[key] = clusters;
}

int launch_fused_voice_mix(mxl_ctx* ctx, const FusedBatch& b)
{
    if (!ctx || !ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_TRY(ctx->activate());
    if (b.frames == 0) return MXL_OK;
    const EqStreamConsts& q = b.eq;
    if (b.n_voices < 1 || b.n_voices > kFusedMaxVoices || b.n_channels < 1 || b.n_channels > kFusedMaxChans)
        MXL_FAIL(MXL_ERR_INVALID, "fused_voice_mix_kernel: %d voices / %d channels", b.n_voices, b.n_channels);
    if (q.halo == 0 || q.halo > (uint32_t)kT / 2 || q.lev_lo > (uint32_t)kEqPlanLevels || q.lev_hi > (uint32_t)kEqPlanLevels ||
        q.back_lo > 3 || q.back_hi > 3 || !q.tab || b.owned == 0 || b.owned + q.halo > (uint32_t)kT)
        MXL_FAIL(MXL_ERR_INVALID, "fused_voice_mix_kernel: bad plan (chunk %u, halo %u, owned %u)", q.chunk, q.halo, b.owned);
    switch (q.chunk) {
    case 32: return launch_lc<32>(ctx, b);
    case 64: return launch_lc<64>(ctx, b);
    default: MXL_FAIL(MXL_ERR_INVALID, "fused_voice_mix_kernel: unsupported chunk length %u", q.chunk);
    }
}

}  // namespace k
}  // namespace mxl

// envelope.cu -- Envelope (src/module/envelope.rs:16-58,91-120) as a single streaming pass.
//
// The reference walks a 3-state machine sample by sample.  Its transitions depend only on whether
// the machine is "on" (TriggerOn) or not (Initial / TriggerOff): a gate sample == 1.0 switches
// not-on to on, a gate sample == 0.0 switches on to not-on, every other sample is inert.  Hence
//   (1) the class after sample i is the class of the last EVENT (sample that is exactly 1.0 or 0.0)
//       at or before i -- "last event wins", a max-scan over keys (index+1)<<1 | class;
//   (2) an event is a TRANSITION iff its class differs from the class just before it;
//   (3) the machine state at sample i is fixed by the last transition p <= i (TriggerOn{on: p} or
//       TriggerOff{off: p, ..}) and, for an off, the transition q before it:
//       off_amplitude = amplitude(TriggerOn{on: q}, p) -- a "latest two" scan over transition keys.
// Both scans have trivial combine steps (max / latest-two), so they run as decoupled look-back
// scans inside ONE kernel: a CTA owns a tile of 2048 consecutive samples (16 per thread, read once
// with four float4 loads), scans inside the tile with warp shuffles, publishes its tile aggregate,
// and looks back over the preceding tiles 32 at a time (ballot + shuffle, no loops over lanes)
// until the carry is decided: the nearest tile with any event decides scan (1), two transitions or
// a tile whose inclusive value is known decide scan (3).  Every thread then starts the REFERENCE
// state machine from its exact incoming state and walks its 16 samples, evaluating amplitude()
// with the reference's f64 expressions (bit-exact: only IEEE +,-,*,/).
// Line traffic is the algorithmic 8 B/sample; tile descriptors add 24 B per 2048 samples.  Tiles
// are handed out by an atomic ticket, so a tile only ever waits for tiles that already run.
#include <stdlib.h>

#include "dsp_math.cuh"
#include "kernels.h"

namespace mxl {
namespace k {

namespace {

constexpr int kEnvThreads = 128;      // 256 threads: 0.126 ms, 128: 0.117 (0.110 at 16 CTAs per SM), 64: 0.120: more, smaller CTAs cover the look-backs
constexpr int kEnvPerThread = 16;     // per 2^25 samples: 8 per thread 0.169 ms, 16: 0.117, 32: 0.155
constexpr int kEnvTileSamples = kEnvThreads * kEnvPerThread;
constexpr int kEnvWarps = kEnvThreads / 32;

// event / transition key of sample idx (index inside the call): 0 = none; later samples have larger keys
__device__ __forceinline__ uint32_t key_of(uint32_t idx, bool on) { return ((idx + 1u) << 1) | (on ? 1u : 0u); }
__device__ __forceinline__ uint32_t key_idx(uint32_t key) { return (key >> 1) - 1u; }
__device__ __forceinline__ bool key_on(uint32_t key) { return (key & 1u) != 0; }

// latest two transitions, a later than b (0 = none); `x` earlier in time than `y`
struct Top2 { uint32_t a, b; };
__device__ __forceinline__ Top2 top2_combine(Top2 x, Top2 y)
{
    if (y.a == 0u) return x;
    if (y.b != 0u) return y;
    return Top2{y.a, x.a};
}

struct EnvParams { double sr, inv_sr, attack_ms, inv_attack, inv_decay, sustain, inv_release; };

// envelope.rs:16-18 -- the quotient by the sample rate correctly rounded (dsp_math.cuh)
__device__ __forceinline__ double duration_ms(uint64_t first, uint64_t last, const EnvParams& p)
{
    return div_by_const((double)(last - first), p.sr, p.inv_sr) * 1000.0;
}

// envelope.rs:38-51, TriggerOn arm (1.0 / x_ms is loop-invariant and formed once on the host)
__device__ __forceinline__ double amp_on(const EnvParams& p, uint64_t on, uint64_t t)
{
    const double ms_since_on = duration_ms(on, t, p);
    if (ms_since_on < p.attack_ms) return p.inv_attack * ms_since_on;
    const double ms_since_decay_started = ms_since_on - p.attack_ms;
    const double decay_amplitude = 1.0 - clamp01(p.inv_decay * ms_since_decay_started);
    return p.sustain + ((1.0 - p.sustain) * decay_amplitude);
}

// envelope.rs:52-57, TriggerOff arm
__device__ __forceinline__ double amp_off(const EnvParams& p, uint64_t off, double off_amplitude, uint64_t t)
{
    const double ms_since_off = duration_ms(off, t, p);
    const double release_amplitude = 1.0 - clamp01(p.inv_release * ms_since_off);
    return off_amplitude * release_amplitude;
}

// Tile descriptors: 64-bit words (tag << 32 | value), tag = epoch << 2 | status, written and read
// whole, so a word is either stale (other epoch), an aggregate or an inclusive value -- no fences.
constexpr uint32_t kAgg = 1u, kIncl = 2u;

__device__ __forceinline__ unsigned long long ld_word(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_word(unsigned long long* p, uint32_t tag, uint32_t value)
{
    const unsigned long long v = ((unsigned long long)tag << 32) | value;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// spins until the word carries this launch's epoch; returns status, value through *value
__device__ __forceinline__ uint32_t wait_word(const unsigned long long* p, uint32_t epoch, uint32_t* value)
{
    unsigned long long v;
    do { v = ld_word(p); } while ((uint32_t)(v >> 34) != epoch);
    *value = (uint32_t)v;
    return (uint32_t)(v >> 32) & 3u;
}

__global__ void __launch_bounds__(kEnvThreads, 16) envelope_kernel(const __grid_constant__ EnvBatch b)
{
    pdl_prologue();
    const EnvInst& in = b.inst[blockIdx.y];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_ev[kEnvWarps];                  // per-warp last-event key
    __shared__ Top2 s_tr[kEnvWarps];                      // per-warp latest two transitions
    __shared__ uint32_t s_ev_in;                          // last-event key before the tile (0 = none in this call)
    __shared__ Top2 s_tr_in;                              // latest two transitions before the tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = (uint32_t)(atomicAdd(in.ticket, 1ull) - in.ticket_base);
    __syncthreads();
    const uint32_t tile = s_tile;
    EnvTile* const desc = in.tiles + tile;
    const EnvParams p{b.sample_rate, b.inv_sample_rate, in.attack_ms, in.inv_attack, in.inv_decay, in.sustain, in.inv_release};
    const EnvState st0 = *in.state;                        // machine state before the call

    // ---- my samples ----
    const uint64_t base = (uint64_t)tile * kEnvTileSamples + (uint64_t)tid * kEnvPerThread;
    float x[kEnvPerThread];
    if (in.in && base + kEnvPerThread <= b.frames && (reinterpret_cast<uintptr_t>(in.in) & 15) == 0) {
#pragma unroll
        for (int v = 0; v < kEnvPerThread / 4; v++) {
            const float4 a = *reinterpret_cast<const float4*>(in.in + base + 4 * v);
            x[4 * v] = a.x; x[4 * v + 1] = a.y; x[4 * v + 2] = a.z; x[4 * v + 3] = a.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kEnvPerThread; j++)            // disconnected input = zeros (io.rs:8-9); past the end = inert
            x[j] = base + j < b.frames ? (in.in ? in.in[base + j] : 0.0f) : 2.0f;
    }

    // ---- scan 1: last event.  envelope.rs:101,106: exact float ==, so -0.0 counts as 0.0 ----
    uint32_t on_mask = 0, off_mask = 0;                    // bit j: sample j is exactly 1.0 / 0.0
#pragma unroll
    for (int j = 0; j < kEnvPerThread; j++) {
        on_mask |= (x[j] == 1.0f ? 1u : 0u) << j;
        off_mask |= (x[j] == 0.0f ? 1u : 0u) << j;
    }
    uint32_t ev = 0;
    if (on_mask | off_mask) {
        const int j = 31 - __clz(on_mask | off_mask);
        ev = key_of((uint32_t)(base + j), (on_mask >> j) & 1u);
    }
    uint32_t ev_incl = ev;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) ev_incl = max(ev_incl, __shfl_up_sync(0xffffffffu, ev_incl, d));
    uint32_t ev_excl = __shfl_up_sync(0xffffffffu, ev_incl, 1);
    if (lane == 0) ev_excl = 0;
    if (lane == 31) s_ev[warp] = ev_incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t agg = 0;
#pragma unroll
        for (int w = 0; w < kEnvWarps; w++) agg = max(agg, s_ev[w]);
        uint32_t carry = 0;                                // last event before the tile, inside this call
        if (tile != 0) {
            // a tile with an event decides everything after it: its aggregate is already inclusive
            if (lane == 0) st_word(&desc->ev, (b.epoch << 2) | (agg ? kIncl : kAgg), agg);
            int64_t j = (int64_t)tile - 1;
            for (;;) {
                const int64_t mj = j - lane;
                uint32_t status = kIncl, val = 0;          // before the call: inclusive "none"
                if (mj >= 0) status = wait_word(&in.tiles[mj].ev, b.epoch, &val);
                const uint32_t decided = __ballot_sync(0xffffffffu, status == kIncl);
                if (decided) {                             // nearest decided tile; the ones nearer had no event
                    carry = __shfl_sync(0xffffffffu, val, __ffs(decided) - 1);
                    break;
                }
                j -= 32;
            }
            if (agg == 0 && lane == 0) st_word(&desc->ev, (b.epoch << 2) | kIncl, carry);
        } else if (lane == 0) {
            st_word(&desc->ev, (b.epoch << 2) | kIncl, agg);
        }
        if (lane == 0) s_ev_in = carry;
    }
    __syncthreads();
    // class just before my first sample
    uint32_t before = s_ev_in;
    for (int w = 0; w < warp; w++) before = max(before, s_ev[w]);
    before = max(before, ev_excl);
    bool cls = before ? key_on(before) : (st0.state == 1);

    // ---- scan 3: latest two transitions ----
    Top2 tr{0u, 0u};
    if ((cls ? off_mask : on_mask) != 0u) {                // an event of the other class: at least one transition
#pragma unroll
        for (int j = 0; j < kEnvPerThread; j++) {
            const bool on = (on_mask >> j) & 1u, off = (off_mask >> j) & 1u;
            if ((on && !cls) || (off && cls)) {
                tr.b = tr.a;
                tr.a = key_of((uint32_t)(base + j), on);
                cls = on;
            }
        }
    }
    Top2 tr_incl = tr;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Top2 o;
        o.a = __shfl_up_sync(0xffffffffu, tr_incl.a, d);
        o.b = __shfl_up_sync(0xffffffffu, tr_incl.b, d);
        if (lane >= d) tr_incl = top2_combine(o, tr_incl);
    }
    Top2 tr_excl;
    tr_excl.a = __shfl_up_sync(0xffffffffu, tr_incl.a, 1);
    tr_excl.b = __shfl_up_sync(0xffffffffu, tr_incl.b, 1);
    if (lane == 0) tr_excl = Top2{0u, 0u};
    if (lane == 31) s_tr[warp] = tr_incl;
    __syncthreads();
    if (warp == 0) {
        Top2 agg{0u, 0u};
#pragma unroll
        for (int w = 0; w < kEnvWarps; w++) agg = top2_combine(agg, s_tr[w]);
        Top2 carry{0u, 0u};
        if (tile != 0) {
            const bool full = agg.b != 0u;                 // two transitions of its own: inclusive as it stands
            if (lane == 0) {
                st_word(&desc->tr_a, (b.epoch << 2) | (full ? kIncl : kAgg), agg.a);
                st_word(&desc->tr_b, (b.epoch << 2) | (full ? kIncl : kAgg), agg.b);
            }
            int64_t j = (int64_t)tile - 1;
            for (;;) {
                const int64_t mj = j - lane;
                uint32_t status = kIncl, va = 0, vb = 0;   // before the call: inclusive "none"
                if (mj >= 0) {
                    uint32_t sa, sb;
                    do {                                   // the two words of a descriptor change status one after the other
                        sa = wait_word(&in.tiles[mj].tr_a, b.epoch, &va);
                        sb = wait_word(&in.tiles[mj].tr_b, b.epoch, &vb);
                    } while (sa != sb);
                    status = sa;
                }
                // walk from the nearest tile back: collect transitions until two are known or a tile is inclusive
                const uint32_t has = __ballot_sync(0xffffffffu, va != 0u);
                const uint32_t incl = __ballot_sync(0xffffffffu, status == kIncl);
                const int stop = incl ? __ffs(incl) - 1 : 32;          // nearest inclusive tile
                const uint32_t reach = stop >= 31 ? 0xffffffffu : ((2u << stop) - 1u);   // lanes 0..stop
                uint32_t cand = has & reach;
                while (cand && carry.b == 0u) {
                    const int l = __ffs(cand) - 1;
                    const uint32_t la = __shfl_sync(0xffffffffu, va, l), lb = __shfl_sync(0xffffffffu, vb, l);
                    if (carry.a == 0u) { carry.a = la; carry.b = lb; }
                    else carry.b = la;
                    cand &= cand - 1u;
                }
                if (carry.b != 0u || incl) break;
                j -= 32;
            }
            if (!full && lane == 0) {
                const Top2 inc = top2_combine(carry, agg);
                st_word(&desc->tr_a, (b.epoch << 2) | kIncl, inc.a);
                st_word(&desc->tr_b, (b.epoch << 2) | kIncl, inc.b);
            }
        } else if (lane == 0) {
            st_word(&desc->tr_a, (b.epoch << 2) | kIncl, agg.a);
            st_word(&desc->tr_b, (b.epoch << 2) | kIncl, agg.b);
        }
        if (lane == 0) s_tr_in = carry;
    }
    __syncthreads();
    Top2 tin = s_tr_in;
    for (int w = 0; w < warp; w++) tin = top2_combine(tin, s_tr[w]);
    tin = top2_combine(tin, tr_excl);

    // ---- machine state before my first sample ----
    EnvState s = st0;
    if (tin.a != 0u) {
        s.seq = b.t0 + key_idx(tin.a);
        if (key_on(tin.a)) {
            s.state = 1;
        } else {
            // envelope.rs:108-112: off_amplitude = amplitude(TriggerOn{on}, off); the class before an off
            // transition is "on": the transition before it, or the incoming TriggerOn of the call
            const uint64_t on = tin.b != 0u ? b.t0 + key_idx(tin.b) : st0.seq;
            s.state = 2;
            s.off_amplitude = amp_on(p, on, s.seq);
        }
    }

    // ---- outputs ----
    const uint64_t seq0 = b.t0 + base;
    if (tr.a == 0u && base + kEnvPerThread <= b.frames && seq0 + kEnvPerThread - s.seq < (1ull << 53) &&
        (reinterpret_cast<uintptr_t>(in.out) & 15) == 0) {
        // no transition among my samples: the machine keeps state s; amplitude() (envelope.rs:33-58) per
        // sample with the elapsed sample count stepped in f64 (exact below 2^53)
        float y[kEnvPerThread];
        if (s.state == 1) {
            const double d0 = (double)(seq0 - s.seq), rest = 1.0 - p.sustain;
#pragma unroll
            for (int j = 0; j < kEnvPerThread; j++) {
                const double ms = div_by_const(d0 + (double)j, p.sr, p.inv_sr) * 1000.0;
                const double decay_amplitude = 1.0 - clamp01(p.inv_decay * (ms - p.attack_ms));
                y[j] = (float)(ms < p.attack_ms ? p.inv_attack * ms : p.sustain + (rest * decay_amplitude));
            }
        } else if (s.state == 2) {
            const double d0 = (double)(seq0 - s.seq);
#pragma unroll
            for (int j = 0; j < kEnvPerThread; j++) {
                const double ms = div_by_const(d0 + (double)j, p.sr, p.inv_sr) * 1000.0;
                y[j] = (float)(s.off_amplitude * (1.0 - clamp01(p.inv_release * ms)));
            }
        } else {
#pragma unroll
            for (int j = 0; j < kEnvPerThread; j++) y[j] = 0.f;
        }
#pragma unroll
        for (int v = 0; v < kEnvPerThread / 4; v++)
            *reinterpret_cast<float4*>(in.out + base + 4 * v) = make_float4(y[4 * v], y[4 * v + 1], y[4 * v + 2], y[4 * v + 3]);
        if (base + kEnvPerThread == b.frames) *in.state_out = s;
    } else {
        // the reference state machine, sample by sample (envelope.rs:96-117)
#pragma unroll 1
        for (int j = 0; j < kEnvPerThread; j++) {
            const uint64_t i = base + j;
            if (i >= b.frames) break;
            const uint64_t seq = b.t0 + i;
            if (s.state != 1) {
                if ((on_mask >> j) & 1u) { s.state = 1; s.seq = seq; }
            } else if ((off_mask >> j) & 1u) {
                s.off_amplitude = amp_on(p, s.seq, seq);
                s.state = 2; s.seq = seq;
            }
            double a = 0.0;                                // Initial
            if (s.state == 1) a = amp_on(p, s.seq, seq);
            else if (s.state == 2) a = amp_off(p, s.seq, s.off_amplitude, seq);
            in.out[i] = (float)a;
            if (i + 1 == b.frames) *in.state_out = s;
        }
    }
}

}  // namespace

uint32_t envelope_tiles(uint64_t frames) { return (uint32_t)((frames + kEnvTileSamples - 1) / kEnvTileSamples); }

int launch_envelope(mxl_ctx* ctx, const EnvBatch& b)
{
    if (!ctx || !ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_TRY(ctx->activate());
    if (b.n <= 0 || b.frames == 0) return MXL_OK;
    if (b.frames >= 0x7ffffff0ull) MXL_FAIL(MXL_ERR_LENGTH, "Envelope: call longer than 2^31 samples");
    dim3 grid(envelope_tiles(b.frames), b.n);
    MXL_TIMED(ctx, "envelope_kernel");
    launch_chained(ctx, envelope_kernel, grid, dim3(kEnvThreads), 0, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) MXL_FAIL(MXL_ERR_CUDA, "envelope launch failed: %s", cudaGetErrorString(e));
    ctx->launches++;
    return MXL_OK;
}

}  // namespace k
}  // namespace mxl

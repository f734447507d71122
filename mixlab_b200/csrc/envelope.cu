// envelope.cu -- Envelope (src/module/envelope.rs:16-58,91-120) as a single streaming pass.
//
// The reference walks a 3-state machine sample by sample.  Its transitions depend only on whether
// the machine is "on" (TriggerOn) or not (Initial / TriggerOff): a gate sample == 1.0 switches
// not-on to on, a gate sample == 0.0 switches on to not-on, every other sample is inert.  Hence
//   (1) the class after sample i is the class of the last EVENT (sample that is exactly 1.0 or 0.0)
//       at or before i;
//   (2) an event is a TRANSITION iff its class differs from the class just before it, i.e. from the
//       class of the event before it (or, for the first event of the call, from the stored state);
//   (3) the machine state at sample i is fixed by the last transition p <= i (TriggerOn{on: p} or
//       TriggerOff{off: p, ..}) and, for an off, the transition q before it:
//       off_amplitude = amplitude(TriggerOn{on: q}, p).
// All of that follows from ONE associative summary of a run of samples, Seg = {first event, last
// event, the latest two transitions among the events after the first}: joining two runs adds at most
// the boundary transition (first event of the later run against the last event of the earlier one).
// The kernel is therefore one decoupled look-back scan over Seg.  A CTA owns a tile of 2048 consecutive
// samples, 16 per thread:
//   1. the samples are read once, coalesced (lane l takes quad 32 v + l of its warp's 128 quads), their
//      event masks change hands so that every thread holds the masks of its 16 consecutive samples;
//   2. the thread's summary (bit scans), the warp's (ballots + three shuffles, no scan), the tile's;
//   3. warp 0 publishes the tile's Seg and looks back over the preceding tiles, nearest first and only as
//      far as they are already published, until the carry is decided -- two transitions known (nothing
//      earlier can matter) or a tile whose published value already covers everything before it;
//   4. every thread starts the REFERENCE state machine from its exact incoming state: no transition among
//      its samples (the common case) -> amplitude() straight-line with the reference's f64 expressions,
//      specialised to the one arm all sixteen samples take; otherwise the warp takes the thread's samples
//      one per lane.  Outputs cross a shared-memory tile so that they, too, leave coalesced.
// Bit-exact: only IEEE +,-,*,/ in the reference's order.  Line traffic is the algorithmic 8 B/sample; tile
// descriptors add 32 B per 2048 samples.
//
// History: the first single-pass version (r1) ran (1) and (3) as two scans with two look-backs behind an atomic
// ticket -- four dependent global round trips per CTA and five block barriers: 0.110 ms per 2^25 samples.  One
// scan, ballots, the nearest-first walk, coalesced accesses: see DESIGN.md 4.3 for what each step bought.
// Tiles are taken in blockIdx order, which relies on CTAs being dispatched in index order (as single-pass scans
// generally do): a CTA only ever waits for tiles with lower indices.
#include <stdlib.h>

#include <algorithm>

#include "dsp_math.cuh"
#include "kernels.h"

namespace mxl {
namespace k {

namespace {

constexpr int kEnvThreads = 128;      // 256 threads: 0.126 ms, 128: 0.117 (0.110 at 16 CTAs per SM), 64: 0.120: more, smaller CTAs cover the look-backs
constexpr int kEnvWarps = kEnvThreads / 32;

// event / transition key of sample idx (index inside the call): 0 = none; later samples have larger keys
__device__ __forceinline__ uint32_t key_of(uint32_t idx, uint32_t on) { return ((idx + 1u) << 1) | on; }
__device__ __forceinline__ uint32_t key_idx(uint32_t key) { return (key >> 1) - 1u; }
__device__ __forceinline__ bool key_on(uint32_t key) { return (key & 1u) != 0; }

// Summary of a run of samples: F / L = its first / last event; a, b = the latest two transitions among the events
// after F (a later than b; 0 = none) -- whether F itself is a transition depends on what came before the run.
struct Seg { uint32_t F, L, a, b; };

// x earlier in time than y.  Transitions in time order: x.b, x.a, boundary, y.b, y.a -- keep the latest two.
__device__ __forceinline__ Seg seg_combine(const Seg x, const Seg y)
{
    const bool xh = x.L != 0u, yh = y.L != 0u;
    const uint32_t boundary = (xh && ((x.L ^ y.F) & 1u)) ? y.F : 0u;   // y's first event changes the class x left (y.F = 0: none)
    Seg r;
    r.F = xh ? x.F : y.F;
    r.L = yh ? y.L : x.L;
    const uint32_t third = boundary ? boundary : x.a;                  // the latest transition before y's own
    const uint32_t fourth = boundary ? x.a : x.b;
    r.a = y.a ? y.a : third;
    r.b = y.a ? (y.b ? y.b : third) : fourth;
    return r;
}
// A Seg with two transitions is SATURATED: joined with anything later it stays saturated, and nothing earlier can
// change its L, a, b -- only its F, which a saturated Seg's users never read.

__device__ __forceinline__ Seg seg_shfl(const Seg s, int l)
{
    Seg o;
    o.F = __shfl_sync(0xffffffffu, s.F, l); o.L = __shfl_sync(0xffffffffu, s.L, l);
    o.a = __shfl_sync(0xffffffffu, s.a, l); o.b = __shfl_sync(0xffffffffu, s.b, l);
    return o;
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

// Tile descriptors: four 64-bit words (tag << 32 | value), tag = epoch << 2 | status, each written and read whole, so a
// word is either stale (other epoch), part of the tile's own summary (kAgg) or of its inclusive one (kIncl); a reader
// retries until the four tags agree -- no fences.
constexpr uint32_t kAgg = 1u, kIncl = 2u;

__device__ __forceinline__ void st_seg(EnvTile* d, uint32_t tag, const Seg s)
{
    const unsigned long long t = (unsigned long long)tag << 32;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" :: "l"(&d->w[0]), "l"(t | s.F), "l"(t | s.L) : "memory");
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" :: "l"(&d->w[2]), "l"(t | s.a), "l"(t | s.b) : "memory");
}

// one look at a descriptor: true once its four words carry this launch's epoch and one status
__device__ __forceinline__ bool try_seg(const EnvTile* d, uint32_t epoch, Seg* s, uint32_t* status)
{
    unsigned long long w0, w1, w2, w3;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(&d->w[0]) : "memory");
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "l"(&d->w[2]) : "memory");
    const uint32_t t0 = (uint32_t)(w0 >> 32), t1 = (uint32_t)(w1 >> 32), t2 = (uint32_t)(w2 >> 32), t3 = (uint32_t)(w3 >> 32);
    if ((t0 >> 2) != epoch || t0 != t1 || t0 != t2 || t0 != t3) return false;
    s->F = (uint32_t)w0; s->L = (uint32_t)w1; s->a = (uint32_t)w2; s->b = (uint32_t)w3;
    *status = t0 & 3u;
    return true;
}

__device__ __forceinline__ int top_bit(uint32_t m) { return 31 - __clz(m); }

constexpr int kEnvPerThread = 16;     // per 2^25 samples (r1 kernel): 8 per thread 0.169 ms, 16: 0.117, 32: 0.155; r2: 32 no better than 16
constexpr int kEnvTileSamples = kEnvThreads * kEnvPerThread;

// A warp's 512 samples as 128 quads (float4).  A THREAD owns 16 consecutive samples = quads 4t .. 4t+3, but a warp-wide
// 16-byte access is only coalesced when LANE l touches quad 32v + l: with the thread layout every load / store instruction
// spreads over sixteen 128-byte lines (16 L1 wavefronts instead of 4, and each 32-byte sector crosses to L2 twice, half
// filled: 365 MB of L1->L2 write traffic for 134 MB of output).  So samples are loaded in the lane layout and only their
// 8-bit event masks change hands (4 shuffles), and outputs cross a 2 KB shared-memory tile per warp.  quad_slot keeps both
// the thread-layout writes and the lane-layout reads of that tile free of bank conflicts.
__device__ __forceinline__ int quad_slot(int q) { return q ^ ((q >> 3) & 3); }

// a thread's outputs, four at a time: f(j) -> sample j, into the warp's staging tile
template <class F>
__device__ __forceinline__ void emit(float4* stage, int lane, F f)
{
#pragma unroll
    for (int v = 0; v < kEnvPerThread / 4; v++) {
        float4 y;
        y.x = f(4 * v); y.y = f(4 * v + 1); y.z = f(4 * v + 2); y.w = f(4 * v + 3);
        stage[4 * lane + (v ^ ((lane >> 1) & 3))] = y;    // = quad_slot(4 lane + v): (4 lane + v) >> 3 == lane >> 1
    }
}

// What a thread keeps of its sixteen samples and of its warp: everything the outputs need besides the carry.
struct EnvThread {
    uint32_t masks;        // bit j: sample j == 1.0; bit 16 + j: sample j == 0.0
    uint32_t tmf;          // bits 0..15: Tm (my events, after my first, whose class differs from the event before);
                           // bit 16: an event exists in the lanes before me (e_has); bit 17: its class (e_cls)
    uint32_t e_a, e_b;     // latest two warp-internal transitions in the lanes before me
    uint32_t Fw;           // the warp's first event
};

// Steps 1 and 2 for the warp whose 512 samples start at warp_base: my masks, my summary, the warp's summary (*warp_seg, in
// every lane).
__device__ __forceinline__ EnvThread env_summarize(const EnvInst& in, const EnvBatch& b, uint64_t warp_base, Seg* warp_seg, int lane)
{
    const uint64_t base = warp_base + (uint64_t)lane * kEnvPerThread;
    // envelope.rs:101,106: exact float ==, so -0.0 counts as 0.0
    uint32_t on_mask = 0, off_mask = 0;                    // bit j: MY sample j is exactly 1.0 / 0.0
    if (in.in && warp_base + 32 * kEnvPerThread <= b.frames && (reinterpret_cast<uintptr_t>(in.in) & 15) == 0) {
        // lane layout: quad 32 v + lane; the quads' masks (on | off << 4, one byte per v) then go to their owners
        uint32_t pk = 0;
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const float4 a = *reinterpret_cast<const float4*>(in.in + warp_base + 4 * (32 * v + lane));
            const uint32_t on4 = (a.x == 1.0f ? 1u : 0u) | (a.y == 1.0f ? 2u : 0u) | (a.z == 1.0f ? 4u : 0u) | (a.w == 1.0f ? 8u : 0u);
            const uint32_t off4 = (a.x == 0.0f ? 1u : 0u) | (a.y == 0.0f ? 2u : 0u) | (a.z == 0.0f ? 4u : 0u) | (a.w == 0.0f ? 8u : 0u);
            pk |= (on4 | (off4 << 4)) << (8 * v);
        }
        const int src0 = 4 * (lane & 7), sh = 8 * (lane >> 3);      // my quads 4 lane + i sit in lane (4 lane + i) % 32, byte lane / 8
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t byte = (__shfl_sync(0xffffffffu, pk, src0 + i) >> sh) & 0xffu;
            on_mask |= (byte & 15u) << (4 * i);
            off_mask |= (byte >> 4) << (4 * i);
        }
    } else {
#pragma unroll
        for (int j = 0; j < kEnvPerThread; j++) {          // disconnected input = zeros (io.rs:8-9); past the end = inert
            const float x = base + j < b.frames ? (in.in ? in.in[base + j] : 0.0f) : 2.0f;
            on_mask |= (x == 1.0f ? 1u : 0u) << j;
            off_mask |= (x == 0.0f ? 1u : 0u) << j;
        }
    }
    const uint32_t E = on_mask | off_mask;
    uint32_t Tm = 0;
    Seg me{0u, 0u, 0u, 0u};
    if (E) {
        // class of the nearest event at or below each bit, by doubling (a bit is filled from the nearest event below it
        // because nearer events reach it in earlier, shorter steps)
        uint32_t val = on_mask, known = E;
#pragma unroll
        for (int s = 1; s < kEnvPerThread; s <<= 1) {
            val |= (val << s) & ~known;
            known |= known << s;
        }
        Tm = E & (known << 1) & ((val << 1) ^ on_mask) & ((1u << kEnvPerThread) - 1u);
        const int f0 = __ffs(E) - 1, l0 = top_bit(E);
        me.F = key_of((uint32_t)base + f0, (on_mask >> f0) & 1u);
        me.L = key_of((uint32_t)base + l0, (on_mask >> l0) & 1u);
        if (Tm) {
            const int h = top_bit(Tm);
            me.a = key_of((uint32_t)base + h, (on_mask >> h) & 1u);
            const uint32_t rest = Tm & ~(1u << h);
            if (rest) {
                const int h2 = top_bit(rest);
                me.b = key_of((uint32_t)base + h2, (on_mask >> h2) & 1u);
            }
        }
    }
    // Inside the warp: ballots instead of a scan.  Every lane learns the class of the last event before it from two
    // ballots, hence whether its first event is a transition; the latest two transitions before a lane are then found
    // with two more ballots (lanes with >= 1 / >= 2 transitions) and three shuffles from lanes picked by bit scans.  The
    // warp's first event stays apart: whether it is a transition depends on what came before the warp.
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t H = __ballot_sync(0xffffffffu, E != 0u);
    const uint32_t CL = __ballot_sync(0xffffffffu, E != 0u && (me.L & 1u));
    const uint32_t Hb = H & lt;                            // lanes before me with events
    const bool e_has = Hb != 0u;
    const bool e_cls = e_has && ((CL >> top_bit(Hb | 1u)) & 1u);     // class of the last event before me in the warp
    uint32_t wa = me.a, wb = me.b;                          // my latest two warp-internal transitions
    if (e_has && E != 0u && ((me.F & 1u) != 0u) != e_cls) {  // my first event changes the class the lanes before me left
        if (wa == 0u) wa = me.F;
        else if (wb == 0u) wb = me.F;
    }
    const uint32_t T1 = __ballot_sync(0xffffffffu, wa != 0u), T2 = __ballot_sync(0xffffffffu, wb != 0u);
    EnvThread th;
    {   // among the lanes before me
        const uint32_t m1 = T1 & lt;
        const int q1 = top_bit(m1 | 1u);
        const uint32_t m2 = m1 & ~(1u << q1);
        const int q2 = top_bit(m2 | 1u);
        const uint32_t a1 = __shfl_sync(0xffffffffu, wa, q1), b1 = __shfl_sync(0xffffffffu, wb, q1), a2 = __shfl_sync(0xffffffffu, wa, q2);
        th.e_a = m1 ? a1 : 0u;
        th.e_b = m1 ? (((T2 >> q1) & 1u) ? b1 : (m2 ? a2 : 0u)) : 0u;
    }
    uint32_t g_a, g_b;
    {   // in the whole warp
        const int q1 = top_bit(T1 | 1u);
        const uint32_t m2 = T1 & ~(1u << q1);
        const int q2 = top_bit(m2 | 1u);
        const uint32_t a1 = __shfl_sync(0xffffffffu, wa, q1), b1 = __shfl_sync(0xffffffffu, wb, q1), a2 = __shfl_sync(0xffffffffu, wa, q2);
        g_a = T1 ? a1 : 0u;
        g_b = T1 ? (((T2 >> q1) & 1u) ? b1 : (m2 ? a2 : 0u)) : 0u;
    }
    th.Fw = __shfl_sync(0xffffffffu, me.F, __ffs(H | 0x80000000u) - 1);     // the warp's first / last event
    const uint32_t Lw = __shfl_sync(0xffffffffu, me.L, top_bit(H | 1u));
    *warp_seg = H ? Seg{th.Fw, Lw, g_a, g_b} : Seg{0u, 0u, 0u, 0u};
    th.masks = on_mask | (off_mask << 16);
    th.tmf = Tm | (e_has ? 1u << 16 : 0u) | (e_cls ? 1u << 17 : 0u);
    return th;
}

// Step 3, first half: a tile's own summary, published.  The call's first tile, or a saturated one: inclusive as it stands.
__device__ __forceinline__ void env_publish(const EnvInst& in, const EnvBatch& b, uint32_t tile, const Seg agg, int lane)
{
    if (lane == 0) st_seg(in.tiles + tile, (b.epoch << 2) | ((tile == 0 || agg.b != 0u) ? kIncl : kAgg), agg);
}

// Step 3, second half (one warp): the summary of everything in the call before `tile`, as far as anybody will read it.
// Lane l looks at tile j - l.  Only the NEAREST tiles matter (until two transitions are known), so the walk never waits for
// a far tile: it folds the contiguous run of published descriptors from lane 0 on, and looks again only while that run has
// not decided the carry.  A tile that was not inclusive on its own then publishes its inclusive value.
__device__ __forceinline__ Seg env_look_back(const EnvInst& in, const EnvBatch& b, uint32_t tile, const Seg agg, int lane)
{
    Seg carry{0u, 0u, 0u, 0u};
    if (tile == 0) return carry;
    int64_t j = (int64_t)tile - 1;
    bool decided = false;
    while (!decided) {
        const int64_t mj = j - lane;
        bool ready = mj < 0;                               // before the call: inclusive "nothing"
        uint32_t status = kIncl;
        Seg d{0u, 0u, 0u, 0u};
        int folded = 0;                                    // lanes 0 .. folded-1 are in `carry`
        for (;;) {
            if (!ready) ready = try_seg(in.tiles + mj, b.epoch, &d, &status);
            const uint32_t rdy = __ballot_sync(0xffffffffu, ready);
            const int n_ready = rdy == 0xffffffffu ? 32 : __ffs(~rdy) - 1;       // published tiles, counted from the nearest
            if (n_ready > folded) {
                const uint32_t window = (n_ready == 32 ? 0xffffffffu : ((1u << n_ready) - 1u)) & ~((1u << folded) - 1u);
                const uint32_t has = __ballot_sync(0xffffffffu, ready && d.L != 0u) & window;
                const uint32_t inc = __ballot_sync(0xffffffffu, ready && status == kIncl) & window;
                const int stop = inc ? __ffs(inc) - 1 : 32;                       // nearest inclusive tile
                const uint32_t reach = stop >= 31 ? 0xffffffffu : ((2u << stop) - 1u);   // lanes 0..stop
                uint32_t cand = has & reach;
                while (cand && carry.b == 0u) {            // from the nearest tile back: join until saturated
                    const int l = __ffs(cand) - 1;
                    carry = seg_combine(seg_shfl(d, l), carry);
                    cand &= cand - 1u;
                }
                if (carry.b != 0u || inc) { decided = true; break; }
                folded = n_ready;
                if (folded == 32) break;                   // the whole batch had nothing decisive: 32 tiles further back
            }
        }
        j -= 32;
    }
    if (agg.b == 0u && lane == 0) st_seg(in.tiles + tile, (b.epoch << 2) | kIncl, seg_combine(carry, agg));
    return carry;
}

// Step 4: outputs of a tile.  seg = the tile's per-warp summaries, carry = everything in the call before the tile.
// Step 4: outputs of the warp whose samples start at warp_base.  P = everything in the call before the warp; stage = the
// warp's 128 output quads.
__device__ __forceinline__ void env_emit(const EnvInst& in, const EnvBatch& b, const EnvState st0, uint64_t warp_base, const EnvThread th,
                                         const Seg P, float4* stage, int lane)
{
    const EnvParams p{b.sample_rate, b.inv_sample_rate, in.attack_ms, in.inv_attack, in.inv_decay, in.sustain, in.inv_release};
    const uint64_t base = warp_base + (uint64_t)lane * kEnvPerThread;
    const uint32_t on_mask = th.masks & 0xffffu, off_mask = th.masks >> 16, E = on_mask | off_mask, Tm = th.tmf & 0xffffu;
    const bool e_has = (th.tmf >> 16) & 1u, e_cls = (th.tmf >> 17) & 1u;

    // ---- everything in this call before my first sample ----
    const bool c0 = st0.state == 1;                        // class the call starts in
    const bool cls_w = P.L ? key_on(P.L) : c0;             // class just before my warp's first sample
    // latest two transitions before my warp: P's own, then P's first event if it changed the stored class
    const uint32_t first_tr = (P.F != 0u && key_on(P.F) != c0) ? P.F : 0u;
    const uint32_t pw_a = P.a ? P.a : first_tr;
    const uint32_t pw_b = P.a ? (P.b ? P.b : first_tr) : 0u;
    // ... then, inside the warp, its first event (if before me and a transition) and the warp-internal ones before me
    const uint32_t fw_tr = (e_has && key_on(th.Fw) != cls_w) ? th.Fw : 0u;
    const uint32_t third = fw_tr ? fw_tr : pw_a, fourth = fw_tr ? pw_a : pw_b;
    const uint32_t tin_a = th.e_a ? th.e_a : third;
    const uint32_t tin_b = th.e_a ? (th.e_b ? th.e_b : third) : fourth;
    const bool cls = e_has ? e_cls : cls_w;                // class just before my first sample
    const uint32_t first_on = E ? (on_mask >> (__ffs(E) - 1)) & 1u : 0u;
    // a transition among my own samples: one of Tm, or my first event against cls
    const bool own_transition = Tm != 0u || (E != 0u && (first_on != 0u) != cls);

    // ---- machine state before my first sample ----
    EnvState s = st0;
    if (tin_a != 0u) {
        s.seq = b.t0 + key_idx(tin_a);
        if (key_on(tin_a)) {
            s.state = 1;
        } else {
            // envelope.rs:108-112: off_amplitude = amplitude(TriggerOn{on}, off); the class before an off
            // transition is "on": the transition before it, or the incoming TriggerOn of the call
            const uint64_t on = tin_b != 0u ? b.t0 + key_idx(tin_b) : st0.seq;
            s.state = 2;
            s.off_amplitude = amp_on(p, on, s.seq);
        }
    }

    // ---- outputs ----
    const uint64_t seq0 = b.t0 + base;
    const bool fast = !own_transition && base + kEnvPerThread <= b.frames && seq0 + kEnvPerThread - s.seq < (1ull << 53) &&
                      (reinterpret_cast<uintptr_t>(in.out) & 15) == 0;
    if (fast) {
        // no transition among my samples: the machine keeps state s; amplitude() (envelope.rs:33-58) per
        // sample with the elapsed sample count stepped in f64 (exact below 2^53).  ms grows with the sample
        // index (every step of its evaluation is monotone), so the branch amplitude() takes at my first and
        // last sample is the branch of all sixteen: the other arm's arithmetic is skipped, the results are the same bits.
        if (s.state == 1) {
            const double d0 = (double)(seq0 - s.seq), rest = 1.0 - p.sustain;
            const double ms_first = div_by_const(d0, p.sr, p.inv_sr) * 1000.0;
            const double ms_last = div_by_const(d0 + (double)(kEnvPerThread - 1), p.sr, p.inv_sr) * 1000.0;
            auto ms_of = [&](int j) { return div_by_const(d0 + (double)j, p.sr, p.inv_sr) * 1000.0; };
            if (ms_last < p.attack_ms) {                   // attack
                emit(stage, lane, [&](int j) { return (float)(p.inv_attack * ms_of(j)); });
            } else if (!(ms_first < p.attack_ms)) {        // decay / sustain
                // x = inv_decay * (ms - attack) grows with the sample index when inv_decay >= 0: clamp(x) is 1 for all
                // sixteen, or x itself for all sixteen, when it is so at the ends
                const double x_first = p.inv_decay * (ms_first - p.attack_ms), x_last = p.inv_decay * (ms_last - p.attack_ms);
                if (p.inv_decay >= 0.0 && x_first > 1.0) {
                    const float c = (float)(p.sustain + (rest * (1.0 - 1.0)));
                    emit(stage, lane, [&](int) { return c; });
                } else if (p.inv_decay >= 0.0 && x_first >= 0.0 && x_last <= 1.0) {
                    emit(stage, lane, [&](int j) { return (float)(p.sustain + (rest * (1.0 - (p.inv_decay * (ms_of(j) - p.attack_ms))))); });
                } else {
                    emit(stage, lane, [&](int j) { return (float)(p.sustain + (rest * (1.0 - clamp01(p.inv_decay * (ms_of(j) - p.attack_ms))))); });
                }
            } else {                                       // the attack ends among my samples
                emit(stage, lane, [&](int j) {
                    const double ms = ms_of(j);
                    const double decay_amplitude = 1.0 - clamp01(p.inv_decay * (ms - p.attack_ms));
                    return (float)(ms < p.attack_ms ? p.inv_attack * ms : p.sustain + (rest * decay_amplitude));
                });
            }
        } else if (s.state == 2) {
            const double d0 = (double)(seq0 - s.seq);
            auto ms_of = [&](int j) { return div_by_const(d0 + (double)j, p.sr, p.inv_sr) * 1000.0; };
            const double x_first = p.inv_release * ms_of(0), x_last = p.inv_release * ms_of(kEnvPerThread - 1);
            if (p.inv_release >= 0.0 && x_first > 1.0) {
                const float c = (float)(s.off_amplitude * (1.0 - 1.0));            // released: the clamp holds from here on
                emit(stage, lane, [&](int) { return c; });
            } else if (p.inv_release >= 0.0 && x_first >= 0.0 && x_last <= 1.0) {  // the clamp is the identity for all sixteen
                emit(stage, lane, [&](int j) { return (float)(s.off_amplitude * (1.0 - (p.inv_release * ms_of(j)))); });
            } else {
                emit(stage, lane, [&](int j) { return (float)(s.off_amplitude * (1.0 - clamp01(p.inv_release * ms_of(j)))); });
            }
        } else {
            emit(stage, lane, [&](int) { return 0.f; });
        }
        if (base + kEnvPerThread == b.frames) *in.state_out = s;
    }
    {   // the staged quads of the fast threads leave in the lane layout (quad 32 v + lane)
        const uint32_t fast_lanes = __ballot_sync(0xffffffffu, fast);
        if (fast_lanes) {
            __syncwarp();
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const int q = 32 * v + lane;
                // quad_slot(q) = 32 v + (lane ^ ((lane >> 3) & 3)): (32 v + lane) >> 3 == 4 v + (lane >> 3)
                if ((fast_lanes >> (q >> 2)) & 1u)
                    *reinterpret_cast<float4*>(in.out + warp_base + 4 * q) = stage[32 * v + (lane ^ ((lane >> 3) & 3))];
            }
            __syncwarp();                                  // the tile is reused by the next tile / launch phase
        }
    }
    // Threads with a transition among their samples (or at the ragged end of the call): the reference state machine
    // (envelope.rs:96-117), one SAMPLE per lane -- the warp takes such a thread's samples together instead of leaving one
    // lane to walk them while 31 wait.  The machine state at sample k follows from the thread's own transitions at or
    // below k (bit scans): none -> the incoming state; the last one an "on" -> TriggerOn{that sample}; an "off" ->
    // TriggerOff{that sample, amplitude(TriggerOn{the transition before it, or the incoming on}, that sample)}.
    const uint32_t own = Tm | ((E != 0u && (first_on != 0u) != cls) ? (E & (0u - E)) : 0u);   // my transitions, given cls
    uint32_t slow = __ballot_sync(0xffffffffu, !fast && base < b.frames);
    while (slow) {
        const int l = __ffs(slow) - 1;
        slow &= slow - 1u;
        const uint32_t M = __shfl_sync(0xffffffffu, own, l), onm = __shfl_sync(0xffffffffu, on_mask, l);
        const int st = __shfl_sync(0xffffffffu, s.state, l);
        const uint64_t sseq = __shfl_sync(0xffffffffu, (unsigned long long)s.seq, l);
        const double soff = __shfl_sync(0xffffffffu, s.off_amplitude, l);
        const uint64_t base_l = base + (uint64_t)((int64_t)(l - lane) * kEnvPerThread);
        for (int k = lane; k < kEnvPerThread; k += 32) {
            const uint64_t i = base_l + k;
            if (i >= b.frames) continue;
            const uint64_t seq = b.t0 + i;
            const uint32_t Mk = M & ((2u << k) - 1u);      // transitions at or below sample k
            int state = st;
            uint64_t start = sseq;
            double offa = soff;
            if (Mk) {
                const int t = top_bit(Mk);
                start = b.t0 + base_l + t;
                if ((onm >> t) & 1u) {
                    state = 1;
                } else {
                    const uint32_t before = Mk & ~(1u << t);
                    const uint64_t on = before ? b.t0 + base_l + top_bit(before) : sseq;
                    state = 2;
                    offa = amp_on(p, on, start);           // envelope.rs:108-112
                }
            }
            double a = 0.0;                                // Initial
            if (state == 1) a = amp_on(p, start, seq);
            else if (state == 2) a = amp_off(p, start, offa, seq);
            in.out[i] = (float)a;
            if (i + 1 == b.frames) {
                EnvState so;
                so.state = state; so._pad = 0; so.seq = start; so.off_amplitude = offa;
                *in.state_out = so;
            }
        }
    }
}

// One tile of 2048 samples per CTA, taken in blockIdx order (relies on CTAs being dispatched in index order, as single-pass
// scans generally do): the warps' summaries meet in shared memory, warp 0 publishes and looks back, two block barriers.
template <int MINB>
__global__ void __launch_bounds__(kEnvThreads, MINB) envelope_kernel(const __grid_constant__ EnvBatch b)
{
    pdl_prologue();
    const EnvInst& in = b.inst[blockIdx.y];
    __shared__ Seg s_seg[kEnvWarps];                      // per-warp summaries
    __shared__ Seg s_in;                                  // summary of everything in this call before the tile
    __shared__ float4 s_stage[kEnvWarps * 128];           // output quads on their way from the thread layout to the lane layout
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const uint64_t warp_base = (uint64_t)tile * kEnvTileSamples + (uint64_t)warp * (32 * kEnvPerThread);
    Seg mine;
    const EnvThread th = env_summarize(in, b, warp_base, &mine, lane);
    if (lane == 0) s_seg[warp] = mine;
    const EnvState st0 = *in.state;                        // machine state before the call
    __syncthreads();
    if (warp == 0) {
        Seg agg = s_seg[0];
#pragma unroll
        for (int w = 1; w < kEnvWarps; w++) agg = seg_combine(agg, s_seg[w]);
        env_publish(in, b, tile, agg, lane);
        const Seg carry = env_look_back(in, b, tile, agg, lane);
        if (lane == 0) s_in = carry;
    }
    __syncthreads();
    Seg P = s_in;                                          // everything before my warp
    for (int w = 0; w < warp; w++) P = seg_combine(P, s_seg[w]);
    env_emit(in, b, st0, warp_base, th, P, s_stage + warp * 128, lane);
}

}  // namespace

// look-back descriptors a call of `frames` samples needs
uint32_t envelope_tiles(uint64_t frames) { return (uint32_t)((frames + kEnvTileSamples - 1) / kEnvTileSamples); }

// Two r2 variants measured the same or worse than this kernel and were removed.  (1) Pipelined: resident CTAs taking
// tiles by ticket and running one tile behind their look-backs, so that a look-back never meets an unpublished neighbour:
// 0.082-0.088 against 0.084-0.085 ms per 2^25 samples -- the barrier samples are the warps of a CTA waiting for each
// other's loads, not for the look-back.  (2) Warp grain: a tile = one warp's 512 samples, every warp publishing and looking
// back for itself, no block barrier, no shared-memory exchange: 0.118-0.125 ms (as in r1: four times the descriptors and
// four times the warps that poll cost more than the two barriers).

int launch_envelope(mxl_ctx* ctx, const EnvBatch& b)
{
    if (!ctx || !ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_TRY(ctx->activate());
    if (b.n <= 0 || b.frames == 0) return MXL_OK;
    if (b.frames >= 0x7ffffff0ull) MXL_FAIL(MXL_ERR_LENGTH, "Envelope: call longer than 2^31 samples");
    static const int minb = getenv("MXL_ENV_MINB") ? atoi(getenv("MXL_ENV_MINB")) : 12;
    const dim3 grid((unsigned)((b.frames + kEnvTileSamples - 1) / kEnvTileSamples), b.n), block(kEnvThreads);
    if (!ctx->env_carveout_set) {                          // 8 KB of staging per CTA: let 12-16 CTAs of an SM have it
        cudaFuncSetAttribute(envelope_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, 60);
        cudaFuncSetAttribute(envelope_kernel<10>, cudaFuncAttributePreferredSharedMemoryCarveout, 60);
        cudaFuncSetAttribute(envelope_kernel<12>, cudaFuncAttributePreferredSharedMemoryCarveout, 60);
        cudaFuncSetAttribute(envelope_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, 60);
        ctx->env_carveout_set = true;
    }
    MXL_TIMED(ctx, "envelope_kernel");
    if (minb == 8) launch_chained(ctx, envelope_kernel<8>, grid, block, 0, b);
    else if (minb == 10) launch_chained(ctx, envelope_kernel<10>, grid, block, 0, b);
    else if (minb == 16) launch_chained(ctx, envelope_kernel<16>, grid, block, 0, b);
    else launch_chained(ctx, envelope_kernel<12>, grid, block, 0, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) MXL_FAIL(MXL_ERR_CUDA, "envelope launch failed: %s", cudaGetErrorString(e));
    ctx->launches++;
    return MXL_OK;
}

}  // namespace k
}  // namespace mxl

// kernels.h -- launch-parameter structs and launchers of the sm_100a kernels.
//
// Every kernel serves a BATCH of module instances of one kind (blockIdx.y = instance): the graph
// executor issues one launch per (dependency level, kind) instead of one per module, and the
// single-module ABI path is the same launch with n = 1.  Instance tables travel as
// __grid_constant__ kernel parameters: they are snapshotted at launch, so the host may rewrite
// its copy immediately, and they are read through the constant bank as warp-uniform loads.
#pragma once

#include "common.h"

namespace mxl {
namespace k {

constexpr int kMaxBatch = 32;        // instances per launch (keeps every table under 4 KB)
constexpr int kMixerMaxCh = 160;     // channels per mixer launch (more => accumulate passes)

// ---- Oscillator (src/module/oscillator.rs:65-92) ----
struct OscInst { float* mono; float* stereo; double freq; int32_t waveform; int32_t _pad; };
struct OscBatch {
    uint64_t t0, frames;
    double sample_rate, inv_sample_rate;
    int32_t n, _pad;
    OscInst inst[kMaxBatch];
};
int launch_oscillator(mxl_ctx* ctx, const OscBatch& b);

// ---- FmSine (src/module/fm_sine.rs:37-56) ----
struct FmInst { const float* in; float* out; double freq_mid, freq_amp; };
struct FmBatch {
    uint64_t t0, frames;
    double sample_rate, inv_sample_rate;
    int32_t n, _pad;
    FmInst inst[kMaxBatch];
};
int launch_fm_sine(mxl_ctx* ctx, const FmBatch& b);

// ---- Mixer (src/module/mixer.rs:46-71): one instance per launch, channel table inline ----
struct MixChan { const float* in; double gain; int32_t cue; int32_t _pad; };
struct MixerLaunch {
    float* master; float* cue;
    uint64_t len;                 // f32 per line (2*frames)
    int32_t channels;
    int32_t accumulate;           // continue a previous pass (channel counts above kMixerMaxCh)
    MixChan ch[kMixerMaxCh];
};
int launch_mixer(mxl_ctx* ctx, const MixerLaunch& p);

// ---- Amplifier (src/module/amplifier.rs:38-73) ----
struct AmpInst { const float* in; const float* mod; float* out; double amplitude, mod_depth; };
struct AmpBatch { uint64_t frames; int32_t n, _pad; AmpInst inst[kMaxBatch]; };
int launch_amplifier(mxl_ctx* ctx, const AmpBatch& b);

// ---- StereoPanner / StereoSplitter / Trigger ----
struct PanInst { const float* left; const float* right; float* out; };
struct PanBatch { uint64_t frames; int32_t n, _pad; PanInst inst[kMaxBatch]; };
int launch_panner(mxl_ctx* ctx, const PanBatch& b);
struct SplitInst { const float* in; float* left; float* right; };
struct SplitBatch { uint64_t frames; int32_t n, _pad; SplitInst inst[kMaxBatch]; };
int launch_splitter(mxl_ctx* ctx, const SplitBatch& b);
struct FillInst { float* out; float value; int32_t _pad; };
struct FillBatch { uint64_t len; int32_t n, _pad; FillInst inst[kMaxBatch]; };
int launch_fill(mxl_ctx* ctx, const FillBatch& b);

// ---- EqThree (src/module/eq_three.rs:58-89,106-125) ----
constexpr int kEqMaxCarry = 8;       // J: previous chunks whose end states are summed into a start state
struct EqInst {
    const float* in; float* out;
    const double* state;             // device: lo poles[4], hi poles[4], history[3] before the call
    double* state_out;               // same layout, after the call (other half of a double buffer)
    double* zend;                    // device scratch: per chunk 8 doubles (zero-state end poles)
    uint32_t* poison;                // device: first chunk with a non-finite carry (~0 = none)
    double g_lo, g_mid, g_hi;
};
struct EqBatch {
    uint64_t frames;
    uint32_t chunk;                  // Lc samples per chunk
    uint32_t n_chunks;
    uint32_t carry_terms;            // J <= kEqMaxCarry
    int32_t n;
    double c_lo, c_hi;               // LowPass.freq (eq_three.rs:117-119)
    // lower-triangular powers A^j = M^(j*Lc) of the homogeneous 4-pole step, row-major packed
    // (10 entries each), j = 0..J, for the lo and hi filters
    double pow_lo[kEqMaxCarry + 1][10];
    double pow_hi[kEqMaxCarry + 1][10];
    EqInst inst[kMaxBatch];
};
int launch_eq_three(mxl_ctx* ctx, const EqBatch& b);

// Time-parallel single launch with dot-product zero pass and skewed exact pass (eq_stream.cu).
constexpr int kEqStreamThreads = 256;
struct EqStreamInst {
    const float* in; float* out;
    const double* state; double* state_out;      // lo poles[4], hi poles[4], history[3]; double-buffered
    uint32_t* poison;                            // device: {first chunk with a non-finite carry (~0 = none), CTAs finished}
    double g_lo, g_mid, g_hi;
};
// Tables of the time-parallel scheme for one chunk length (host: EqStreamPlan, eq_plan.h).  They live in DEVICE memory,
// one copy per chunk length and context, and every CTA brings them into shared memory with one round of coalesced
// loads.  (As kernel parameters they sat in the constant bank: a cold bank serves one 64-byte line per ~700-cycle miss,
// in program order -- 32 to 64 serialised misses at the head of every CTA, 10 us of a live one-tick call.)
struct EqDevTables {
    double V[64][8];                             // end-state response to a unit input at sample j
    double K[8];                                 // zero-input (VSA) end state of a chunk
    double pow_lo[8][10];                        // A^(2^d), packed lower-triangular
    double pow_hi[8][10];
    double lane_pow[2][10][32];                  // A^(lane+1), [cascade][entry][lane] (EqStreamPlan::lane_pow)
};
// Scalars of the scheme, shared by eq_stream_kernel and the fused voice kernel (kernel parameters).
struct EqStreamConsts {
    uint32_t chunk;                              // LC: 16, 32 or 64
    uint32_t halo;                               // chunks recomputed ahead of a CTA's own range
    uint32_t lev_lo, lev_hi;                     // scan levels per cascade (<= 8); levels >= 5 cross warps
    uint32_t back_lo, back_hi;                   // previous warps a warp's start states still hear (<= 3)
    const EqDevTables* tab;                      // device
    const double (*host_V)[8];                   // HOST copies of V and K (the plan's, alive as long as the context): the
    const double* host_K;                        // long-call variant of eq_stream_kernel takes them as kernel parameters
    double c_lo, c_hi;
};
struct EqStreamBatch {
    uint64_t frames;
    uint32_t n_chunks;
    int32_t n;
    EqStreamConsts eq;
    EqStreamInst inst[kMaxBatch];
};
int launch_eq_stream(mxl_ctx* ctx, const EqStreamBatch& b);

// ---- Fused voice group: Oscillator -> EqThree -> StereoPanner -> Mixer [-> Meter] in TWO launches (fused_voice.cu) ----
// The graph executor (graph.cu) replaces the five stages of such a sub-graph.  A "voice" is an Oscillator feeding an
// EqThree; a mixer channel takes a StereoPanner whose sides are voices (or disconnected).
//   fused_voice_kernel   grid (time tiles, voices): a CTA generates its voice's samples straight into the EqThree tile
//                        (the oscillator line never exists), filters them (the eq_stream scheme), stores the EqThree line
//                        and the voice's mixer PRODUCTS (f64(y) * gain) as f32, one line per distinct channel gain;
//   fused_mix_kernel     a CTA per tick: adds the channels' products IN CHANNEL ORDER (mixer.rs:57-68) -- the panner's
//                        interleave is the addressing of that walk -- writes master and cue and reduces the tick to its
//                        meter record in meter_warp_kernel's order.  Launched behind the first with programmatic
//                        dependent launch; its inputs are L2-resident.
struct MeterRecord;
constexpr int kFusedMaxVoices = 16;    // (a voice is 128 bytes of kernel parameters; launches get slower with their size)
constexpr int kFusedMaxChans = 64;
constexpr int kFusedMaxProducts = 2;   // distinct channel gains one voice is multiplied by (more: the group stays staged)
struct FusedVoice {
    double freq; int32_t waveform; int32_t n_products;   // Oscillator params (waveform Off = EqThree input disconnected)
    const double* state; double* state_out;        // EqThree state: lo poles[4], hi poles[4], history[3]; double-buffered
    double g_lo, g_mid, g_hi;
    float* eq_out;                                 // EqThree output line: always written (cue bus and observers read it)
    float* osc_mono; float* osc_stereo;            // Oscillator output lines: written only when something observes them
    double product_gain[kFusedMaxProducts];        // fader * 10^(gain_dB / 20) of the channels this voice feeds (mixer.rs:59)
    float* product_out[kFusedMaxProducts];         // scratch lines: (f64(y) * gain) as f32 (mixer.rs:62)
};
struct FusedVoiceBatch {
    uint64_t t0, frames;
    double sample_rate, inv_sample_rate;
    uint32_t n_chunks;
    uint32_t owned;                                // chunks a tile owns (256 - halo)
    int32_t n_voices;
    int32_t late_wait;                             // the kernel launched before this one is this group's mix kernel of the previous
                                                   // call: generate / scan / filter before waiting for it (fused_voice.cu)
    unsigned long long* prof;                      // diagnostics: kFusedProfStamps clock64 stamps per CTA, or nullptr
    EqStreamConsts eq;
    FusedVoice voice[kFusedMaxVoices];
};
struct FusedChan {
    const float* left; const float* right;         // product lines of the voices feeding the panner's L / R; nullptr = disconnected
    const float* left_raw; const float* right_raw; // their EqThree lines, set only when the cue bus or an observed panner line needs them
    float* pan_out;                                // StereoPanner output line: written only when observed
    float zero_product;                            // (f64(0.0) * gain) as f32: what a disconnected side adds (+-0, NaN for a non-finite gain)
    int32_t cue;
};
struct FusedMixBatch {
    uint64_t frames;
    uint32_t spt;                                  // frames per CTA: a tick when the meter rides along, else any even number
    int16_t n_channels;
    int16_t mono;                                  // every channel's panner takes the same voice on both sides (or none)
    float* master; float* cue;
    MeterRecord* meter;                            // nullptr = no meter in the group
    const FusedChan* chan;                         // device: [n_channels] (uploaded when it changes; as kernel parameters the
                                                   // table cost a constant-bank miss per 64-byte line at the head of every CTA)
};
int launch_fused_voice(mxl_ctx* ctx, FusedVoiceBatch& b);
int launch_fused_mix(mxl_ctx* ctx, const FusedMixBatch& b);
constexpr int kFusedProfStamps = 8;

// ---- Envelope (src/module/envelope.rs:91-120) ----
struct EnvState { int32_t state; int32_t _pad; uint64_t seq; double off_amplitude; };
// look-back descriptor of one tile: four words (epoch << 2 | status) << 32 | key -- first event, last event, latest two
// transitions after the first event (envelope.cu)
struct alignas(16) EnvTile { unsigned long long w[4]; };
struct EnvInst {
    const float* in; float* out;
    const EnvState* state;           // machine state before the call
    EnvState* state_out;             // after the call (other half of a double buffer)
    EnvTile* tiles;                  // device, one per tile of the call; flags carry the launch epoch
    double attack_ms, inv_attack, inv_decay, sustain, inv_release;
};
struct EnvBatch {
    uint64_t t0, frames;
    double sample_rate, inv_sample_rate;
    uint32_t epoch;                  // launch number of this context: stale tile flags never match
    int32_t n;
    EnvInst inst[kMaxBatch];
};
int launch_envelope(mxl_ctx* ctx, const EnvBatch& b);
uint32_t envelope_tiles(uint64_t frames);

// ---- Meter (new): one record per tick slot ----
struct MeterRecord { float peak[2]; int32_t clip; int32_t _pad; double sumsq[2]; };
struct MeterInst { const float* in; MeterRecord* out; };
struct MeterBatch { uint64_t frames; uint32_t spt; int32_t n; MeterInst inst[kMaxBatch]; };
int launch_meter(mxl_ctx* ctx, const MeterBatch& b, uint32_t n_slots);

// ---- OutputDevice channel routing + clip (src/module/output_device.rs:177-206) ----
struct RouteLaunch {
    const float* in;                 // stereo line, nullptr = disconnected (zeros)
    float* scratch;                  // [frames][channels], device
    uint64_t frames;
    uint32_t channels;
    int32_t left, right;             // output channel of each side, -1 = None
    int32_t* clip;                   // device flag, set to 1 when a routed sample lies outside [-1, 1]
};
int launch_route(mxl_ctx* ctx, const RouteLaunch& p);

// ---- PCM (src/video/encode.rs:184-195 ; src/module/stream_input.rs:167-173) ----
int launch_pcm_pack(mxl_ctx* ctx, const float* in, int16_t* out, uint64_t len);
int launch_pcm_unpack(mxl_ctx* ctx, const int16_t* in, float* out, uint64_t len);

// ---- Video ----
struct FadeJob {                     // one output frame
    const uint8_t* a; const uint8_t* b;   // nullptr = layer missing (blank), video_mixer.rs:180-188
    uint8_t* out;
    uint32_t fade; uint32_t _pad;
};
// All jobs share one layout.  jobs_dev is a device array of n_jobs FadeJob.
int launch_crossfade(mxl_ctx* ctx, const mxl_frame_layout& lay, const FadeJob* jobs_dev, uint32_t n_jobs);
// Same launch with the job table (a HOST array of at most kFadeInlineJobs entries) carried in the kernel parameters
constexpr int kFadeInlineJobs = 8;
struct FadeJobsInline { FadeJob job[kFadeInlineJobs]; };
int launch_crossfade_inline(mxl_ctx* ctx, const mxl_frame_layout& lay, const FadeJob* jobs_host, uint32_t n_jobs);
int launch_blank(mxl_ctx* ctx, const mxl_frame_layout& lay, uint8_t* frame);

// Tiled letterbox scaler: all planes of a batch of frames in one launch (video_kernels.cu)
struct ScaleJob { const uint8_t* src; uint8_t* dst; };            // frame base pointers
struct ScalePlane {
    uint32_t src_off, src_w, src_h, src_stride;                   // plane inside the source frame
    uint32_t dst_off, dst_w, dst_h, dst_stride;                   // scaled rectangle inside the destination frame
    const int32_t* xpos; const int16_t* xcoef;                    // device tables: first tap and 4 x 14-bit weights per output column
    const int32_t* ypos; const int16_t* ycoef;                    // ... per output row
    uint32_t tiles_x, tiles_y, tile_base;
    uint32_t tiles_x_magic;                                       // ceil(2^32 / tiles_x): t / tiles_x == umulhi(t, magic) for tiles_x > 1
};
struct ScaleLaunch {
    ScalePlane pl[3];
    uint32_t total_tiles;
    uint32_t region_pitch, region_rows;                           // staging bounds over all tiles (pitch % 16 == 0)
    uint32_t tile_h;                                              // output rows per tile: 64, 32, 8 or 2
    const ScaleJob* jobs;                                         // device
};
int launch_scale_tiled(mxl_ctx* ctx, const ScaleLaunch& L, uint32_t n_jobs);
constexpr size_t kScaleMaxSmem = 200 * 1024;                      // per CTA; beyond it a shorter tile is used
size_t scale_smem_bytes(uint32_t region_rows, uint32_t region_pitch);
uint32_t scale_tile_width();

// Crossfade + yuv420p -> RGBA8 in one pass
struct ComposeRgbaJob { const uint8_t* a; const uint8_t* b; uint8_t* rgba; uint32_t fade; uint32_t _pad; };
int launch_compose_rgba(mxl_ctx* ctx, const mxl_frame_layout& lay, const ComposeRgbaJob* jobs_dev, uint32_t n_jobs);

// RGBA8 -> yuv420p (new, self-specified BT.601 integer form: oracle/mixlab_oracle.h orc_rgba_to_yuv420p)
struct RgbaToYuvJob { const uint8_t* rgba; uint8_t* yuv; };
int launch_rgba_to_yuv(mxl_ctx* ctx, const mxl_frame_layout& lay, const RgbaToYuvJob* jobs_dev, uint32_t n_jobs);

// L2 flush helper
int launch_fill_bytes(mxl_ctx* ctx, void* dst, size_t bytes, uint8_t value);

}  // namespace k
}  // namespace mxl

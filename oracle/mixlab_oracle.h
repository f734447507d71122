/*
 * mixlab_oracle.h -- CPU ORACLE for the mixlab tick hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's (haileys/mixlab @ d73346d) per-tick module
 * arithmetic and of Engine::run_tick's buffer routing.  It exists so that tests/, the
 * smoke() check and bench.py's cpu_baseline / --impl reference legs have something to
 * compare the CUDA path against and to time on host cores.  Nothing in the product
 * (mixlab_b200/, include/) may include, link or call this file.
 *
 * Parity pinning: EqThree is pinned bit-exactly against the reference's only golden
 * vector (fixtures/module/eq_three/chronos{,-eq}.f32.raw, src/module/eq_three.rs:150-167;
 * copies under tests/golden/eq_three/).  Every other function is pinned by source text
 * only (the reference has no tests for them); functions marked UNPINNED restate nothing
 * in the reference and are defined by this file.
 *
 * The reference itself (Rust, nightly packed_simd, un-vendored FFmpeg/x264/fdk-aac) cannot
 * be built in this environment (no cargo/rustc), so there is no oracle/_ref.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (Rust never contracts a*b+c into an FMA).
 * All `file:line` citations are relative to /root/reference.
 */
#ifndef MIXLAB_ORACLE_H
#define MIXLAB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- scalars --------------------------------------------------------------------------- */

/* protocol/src/lib.rs:469-471  Decibel::to_linear = powf(10, dB/20) in f64 */
double orc_db_to_linear(double db);

/* ---- audio modules (all length-agnostic, like the reference's run_tick loops) ------------ */

/* src/module/mixer.rs:46-71.  inputs[ch]==NULL means Disconnected (static zero buffer,
 * src/engine/io.rs:8-9,36-52).  len = number of f32 in each stereo line (2*frames). */
void orc_mixer(const float *const *inputs, const double *gain_db, const double *fader,
               const uint8_t *cue, int channels, float *master, float *cue_out, size_t len);

/* src/module/amplifier.rs:38-73.  mod==NULL means control input disconnected (value 1.0). */
void orc_amplifier(const float *input, const float *mod, double amplitude, double mod_depth,
                   float *output, size_t len);

/* src/module/eq_three.rs:8-11,17-27,106-125 */
typedef struct {
    double lo_coef, hi_coef;      /* LowPass.freq after set_freq (eq_three.rs:117-119) */
    double lo_poles[4], hi_poles[4];
    double history[3];
} orc_eq_three;
void orc_eq_three_create(orc_eq_three *eq, double sample_rate);
/* src/module/eq_three.rs:58-89 */
void orc_eq_three_run(orc_eq_three *eq, double gain_lo_db, double gain_mid_db, double gain_hi_db,
                      const float *input, float *output, size_t n);

/* protocol/src/lib.rs:233-241 (declaration order of the Rust enum) */
enum { ORC_WAVE_ON = 0, ORC_WAVE_OFF = 1, ORC_WAVE_SINE = 2, ORC_WAVE_SQUARE = 3,
       ORC_WAVE_TRIANGLE = 4, ORC_WAVE_SAW = 5 };
/* src/module/oscillator.rs:15-37,65-92 */
void orc_oscillator(uint64_t t, double sample_rate, double freq, int waveform,
                    float *mono, float *stereo, size_t n);

/* src/module/envelope.rs:9-58,91-120 */
enum { ORC_ENV_INITIAL = 0, ORC_ENV_ON = 1, ORC_ENV_OFF = 2 };
typedef struct {
    int state;
    uint64_t seq;             /* `on` or `off` sample index */
    double off_amplitude;
} orc_envelope;
void orc_envelope_create(orc_envelope *env);
void orc_envelope_run(orc_envelope *env, uint64_t t, double sample_rate, double attack_ms,
                      double decay_ms, double sustain_amplitude, double release_ms,
                      const float *input, float *output, size_t n);

/* src/module/fm_sine.rs:37-56.  n = frames; output has 2n floats */
void orc_fm_sine(uint64_t t, double sample_rate, double freq_lo, double freq_hi,
                 const float *input, float *output, size_t n);

/* src/module/stereo_panner.rs:30-41 ; stereo_splitter.rs:33-47 ; trigger.rs:35-48 */
void orc_stereo_panner(const float *left, const float *right, float *output, size_t n);
void orc_stereo_splitter(const float *input, float *left, float *right, size_t n);
void orc_trigger(int open, float *output, size_t n);

/* src/video/encode.rs:184-195 (f32 -> i16 pack) ; src/module/stream_input.rs:167-173 (unpack) */
void orc_pcm_pack_i16(const float *samples, int16_t *pcm, size_t len);
void orc_pcm_unpack_i16(const int16_t *pcm, float *samples, size_t len);

/* Meter stand-ins: src/module/plotter.rs:37-56 (de-interleave tap) and
 * src/module/output_device.rs:188-208 (clip detect `s < -1 || s > 1`). */
void orc_plotter_tap(const float *stereo, float *left, float *right, size_t n);
int  orc_clip_detect(const float *stereo, size_t len);
/* UNPINNED (no reference counterpart): per-channel peak |s| and sum of squares in f64,
 * accumulated in sample order. */
void orc_meter(const float *stereo, size_t n, float peak[2], double sumsq[2], int *clip);

/* ---- video ------------------------------------------------------------------------------ */

/* yuv420p frame in the layout av_frame_get_buffer(frame, 0) produces for the FFmpeg the
 * reference pins (ffmpeg-dev 0.3.8, third party, not vendored): every linesize is a multiple
 * of 32 (the reference asserts exactly this, src/module/video_mixer.rs:196-201).
 * codec/src/ffmpeg/frame.rs:76-138 */
typedef struct {
    uint32_t width, height;       /* luma size */
    uint32_t stride[3];           /* bytes per row, multiple of 32 */
    uint32_t plane_h[3];          /* rows per plane */
    size_t offset[3];             /* byte offset of each plane inside data */
    size_t size;                  /* total bytes */
} orc_frame_layout;
void orc_frame_layout_yuv420p(uint32_t width, uint32_t height, orc_frame_layout *out);
/* AvFrame::blank: Y=0x00, U=V=0x80 (frame.rs:128-134).  The reference memsets
 * stride*(h-1)+w bytes per plane and leaves the last row's padding uninitialised; the oracle
 * (and the product) define the whole plane as blank-valued. */
void orc_frame_blank(const orc_frame_layout *lay, uint8_t *data);

/* src/module/video_mixer.rs:168  crossfade = (fader * 255.0) as u8  (saturating, NaN -> 0) */
uint8_t orc_fader_to_u8(double fader);
/* src/module/video_mixer.rs:170-237: per plane, per row, fade_line over ceil(w/32)*32 bytes.
 * a / b == NULL means "layer missing": the pointer aliases the blank output plane
 * (video_mixer.rs:180-188).  out must hold a blank frame on entry, as in the reference. */
void orc_video_crossfade(const orc_frame_layout *lay, const uint8_t *a, const uint8_t *b,
                         uint8_t fade, uint8_t *out);

/* src/module/video_mixer.rs:261-297 unify_picture_settings */
void orc_unify_picture(uint32_t aw, uint32_t ah, uint32_t bw, uint32_t bh, uint32_t *w, uint32_t *h);
/* src/video/encode.rs:354-374 DynamicScaler geometry (exact Ratio<usize> arithmetic) */
typedef struct { uint32_t scaled_w, scaled_h, letterbox_x, letterbox_y; } orc_scale_geometry;
void orc_scale_geometry_yuv420p(uint32_t in_w, uint32_t in_h, uint32_t out_w, uint32_t out_h,
                                orc_scale_geometry *g);

/* UNPINNED: yuv420p -> RGBA8, BT.601 limited range, 16.16-free integer form:
 *   C=Y-16, D=U-128, E=V-128
 *   R=clip((298*C+409*E+128)>>8) G=clip((298*C-100*D-208*E+128)>>8) B=clip((298*C+516*D+128)>>8) A=255
 * chroma sampled nearest (x>>1, y>>1).  rgba stride = 4*width. */
void orc_yuv420p_to_rgba(const orc_frame_layout *lay, const uint8_t *yuv, uint8_t *rgba);

/* UNPINNED (the other direction; the reference never converts colour, video_mixer.rs:282-283): RGBA8 -> yuv420p,
 * BT.601 limited range, the integer form that inverts the one above to +-2 levels:
 *   Y = ((66*R + 129*G + 25*B + 128) >> 8) + 16            per pixel
 *   U = ((-38*R - 74*G + 112*B + 128) >> 8) + 128          per 2x2 block, from the block's rounded mean colour
 *   V = ((112*R - 94*G - 18*B + 128) >> 8) + 128           (mean = (sum of the block's existing pixels * (4/n) + 2) >> 2,
 *                                                            i.e. a block cut by an odd edge repeats its last column / row)
 * alpha is ignored; >> is an arithmetic shift.  Bytes of the frame outside the picture (stride padding) are left alone. */
void orc_rgba_to_yuv420p(const orc_frame_layout *lay, const uint8_t *rgba, uint8_t *yuv);

/* UNPINNED: bicubic (Keys a=-0.6 as swscale's SWS_BICUBIC default, 4 taps, edge clamp,
 * separable, horizontal then vertical, each pass rounded to u8 via 14-bit fixed point)
 * resample of one plane.  Stands in for the third-party sws_scale call
 * (codec/src/ffmpeg/scale.rs:23-27,68). */
void orc_bicubic_plane(const uint8_t *src, uint32_t sw, uint32_t sh, uint32_t sstride,
                       uint8_t *dst, uint32_t dw, uint32_t dh, uint32_t dstride);

/* The same arithmetic as orc_bicubic_plane (bit-identical, tests/test_oracle_golden.py) arranged the way a CPU scaler
 * is written: tap tables built once per geometry and cached, edge columns clamped outside the inner loop, the
 * vertical pass row-contiguous so the compiler vectorises it.  Used where the CPU arm is TIMED (bench.py's session
 * variant), so that the baseline is not handicapped by the clamps in the definition's inner loops. */
void orc_bicubic_plane_fast(const uint8_t *src, uint32_t sw, uint32_t sh, uint32_t sstride,
                            uint8_t *dst, uint32_t dw, uint32_t dh, uint32_t dstride);
/* DynamicScaler::scale (src/video/encode.rs:338-397): identity copy when the sizes agree, else the three planes
 * resampled (orc_bicubic_plane_fast) into the letterboxed sub-frame of a blank `out` frame. */
void orc_letterbox_scale(const orc_frame_layout *lay_in, const uint8_t *in, const orc_frame_layout *lay_out, uint8_t *out);

/* ---- engine walker ---------------------------------------------------------------------- */

enum { ORC_LINE_MONO = 0, ORC_LINE_STEREO = 1, ORC_LINE_VIDEO = 2 };
enum { ORC_MOD_AMPLIFIER = 0, ORC_MOD_ENVELOPE, ORC_MOD_EQ_THREE, ORC_MOD_FM_SINE, ORC_MOD_MIXER,
       ORC_MOD_OSCILLATOR, ORC_MOD_PLOTTER, ORC_MOD_STEREO_PANNER, ORC_MOD_STEREO_SPLITTER,
       ORC_MOD_TRIGGER, ORC_MOD_METER, ORC_MOD_SOURCE_STEREO, ORC_MOD_SOURCE_MONO };

#define ORC_MAX_PORTS 256

typedef struct orc_graph orc_graph;

orc_graph *orc_graph_create(double sample_rate, uint32_t samples_per_tick);
void orc_graph_destroy(orc_graph *g);
/* params: kind-specific doubles, see mixlab_oracle.c:orc_graph_add.  Returns module id. */
int orc_graph_add(orc_graph *g, int kind, const double *params, int n_params);
/* workspace.rs:97-114: 0 ok, -1 NoInput, -2 NoOutput, -3 TypeMismatch */
int orc_graph_connect(orc_graph *g, int in_module, int in_index, int out_module, int out_index);
/* external source data for ORC_MOD_SOURCE_*: a ring of `frames` frames read at t % frames */
void orc_graph_set_source(orc_graph *g, int module, const float *data, size_t frames);
/* One Engine::run_tick (engine.rs:400-510): terminal set, DFS topsort, per-output zeroed
 * allocation, serial dispatch.  If capture_module>=0 the given output line of that module is
 * copied to capture (len floats of that line type). */
void orc_graph_run_tick(orc_graph *g, uint64_t tick, int capture_module, int capture_output,
                        float *capture);
/* run order of the last tick, for tests */
int orc_graph_last_order(const orc_graph *g, int *order, int cap);
/* meter values recorded by ORC_MOD_METER at the last tick */
void orc_graph_meter(const orc_graph *g, int module, float peak[2], double sumsq[2], int *clip);

/* ---- audio sample-rate converter: UNPINNED, this is its definition (the reference has only the TODO,
 * src/icecast/mod.rs:94-97).  Rational ratio L / M = out_rate / in_rate reduced; output frame m of the stream sits at input
 * position m * M / L: n0 = floor(m * M / L), phase = (m * M) mod L, and
 *     y[m] = (float) sum_{k = 0}^{31} c[phase][k] * (double) x[n0 - 15 + k]      accumulated with fma(), k ascending, from 0.0
 *     c[phase][k] = h(k - 15 - phase / L) / (sum over k of the same),  h(x) = w * sinc(w * x) * bh(x / 16),
 *     w = 0.92 * min(1, L / M),  sinc(u) = sin(pi u) / (pi u),  bh(u) = 0.35875 + 0.48829 cos(pi u) + 0.14128 cos(2 pi u) + 0.01168 cos(3 pi u)
 * x before the start of the stream is 0; i16 input is first converted like StreamInput does (s / 32768.0f).  After N input
 * frames, ceil((N - 16) * L / M) output frames are determined.  The oracle keeps the whole stream. */
typedef struct orc_resampler orc_resampler;
orc_resampler *orc_resampler_create(uint32_t in_rate, uint32_t out_rate, uint32_t channels);
void orc_resampler_destroy(orc_resampler *r);
/* appends in_frames interleaved f32 frames; writes the newly determined output frames to out (cap_frames); returns their count */
size_t orc_resampler_push(orc_resampler *r, const float *in, size_t in_frames, float *out, size_t cap_frames);
/* c[phase][k], for tests */
double orc_resampler_coef(const orc_resampler *r, uint32_t phase, uint32_t k);
void orc_resampler_ratio(const orc_resampler *r, uint32_t *L, uint32_t *M);

/* ---- one live session tick after tick, as the reference's engine thread and its two neighbours do it per tick
 * (timed CPU arm of bench.py's session variant; the per-function restatements above are what it calls):
 *   StreamInput x2: convert_sample of the tick's i16 (stream_input.rs:110-112,167-173)
 *   Engine::run_tick over the audio graph (engine.rs:400-510), master bus captured
 *   VideoMixer: AvFrame::blank + crossfade of the two stored 1080p layers (video_mixer.rs:150-239); the sources
 *     deliver a new frame every `ticks_per_frame` ticks and the mixer re-uses the stored ones in between (92-143)
 *   Monitor: DynamicScaler to the monitor's size (encode.rs:279-287,338-397) and AudioCtx::send_audio's pack (184-195)
 * layers_a / layers_b hold n_layers frames each, used round-robin.  Outputs of the LAST tick are left in
 * monitor_out (lay_mon.size bytes) and pcm_out (2 * spt i16). */
typedef struct {
    orc_graph *graph; int master_module;
    orc_frame_layout lay, lay_mon;
    const uint8_t *layers_a, *layers_b; uint32_t n_layers, ticks_per_frame;
    const int16_t *pcm_in; size_t pcm_in_samples;      /* ring of interleaved i16 the two StreamInputs read */
    uint8_t fade;
    uint8_t *composite, *monitor_out; int16_t *pcm_out; float *scratch;   /* caller-owned work buffers */
} orc_session;
void orc_session_run(orc_session *s, uint64_t tick0, uint32_t n_ticks);

#ifdef __cplusplus
}
#endif
#endif

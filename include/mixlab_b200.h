/*
 * mixlab_b200.h -- C ABI of the B200 (sm_100a) back end for Mixlab's per-tick module-graph path.
 *
 * This is the drop-in boundary: the entry points a Rust `impl ModuleT` shim
 * binds with `extern "C"` to replace the arithmetic of the reference's src/module/ sources and the
 * buffer routing of Engine::run_tick.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *  - Every function returning `int` returns MXL_OK (0) or a negative mxl_status.  The message
 *    of the last failure on the calling thread is available from mxl_last_error().  Nothing ever
 *    unwinds across this boundary (the reference's own FFI convention: negative c_int sentinels,
 *    codec/src/ffmpeg.rs:25-26, codec/src/ffmpeg/ioctx.rs:136-152).
 *  - Where the reference panics (line-type mismatch src/engine/io.rs:40-41,49-50; params-variant
 *    mismatch src/engine/module.rs:108) this ABI returns MXL_ERR_LINE_TYPE / MXL_ERR_PARAMS.
 *  - A context is single-threaded like the engine thread that owns the modules
 *    (src/engine.rs:78-93): one CUDA stream per context, no internal locking.
 *  - Line buffers are device-resident and owned by the back end.  Host pointers appear only in
 *    upload/download calls.  A NULL input line means InputRef::Disconnected
 *    (src/engine/io.rs:19-61).
 *  - `t` is the absolute sample index of the first sample of the call (src/engine.rs:490).
 *    run_tick is length-agnostic exactly like the reference loops: a line of n_ticks*S frames
 *    processed in one call gives the same samples as n_ticks calls of S frames.
 *  - There is no CPU fallback.  A context created with device = MXL_DEVICE_NONE can only be
 *    used for host-side logic (graph planning, connect type checks, picture geometry); every
 *    compute entry point returns MXL_ERR_NO_DEVICE on it.
 *
 * Citations `file:line` are relative to the reference repository (haileys/mixlab @ d73346d).
 */
#ifndef MIXLAB_B200_H
#define MIXLAB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MXL_API __attribute__((visibility("default")))

typedef struct mxl_ctx mxl_ctx;
typedef struct mxl_line mxl_line;
typedef struct mxl_frame mxl_frame;
typedef struct mxl_module mxl_module;
typedef struct mxl_graph mxl_graph;

typedef enum mxl_status {
    MXL_OK = 0,
    MXL_ERR_INVALID = -1,        /* NULL handle, bad index, bad size */
    MXL_ERR_LINE_TYPE = -2,      /* io.rs:40-41 "expected mono input, got stereo" */
    MXL_ERR_PARAMS = -3,         /* module.rs:108 "module params mismatch" */
    MXL_ERR_CUDA = -4,
    MXL_ERR_NO_DEVICE = -5,
    MXL_ERR_NO_INPUT = -6,       /* workspace.rs:100 ConnectError::NoInput */
    MXL_ERR_NO_OUTPUT = -7,      /* workspace.rs:105 ConnectError::NoOutput */
    MXL_ERR_TYPE_MISMATCH = -8,  /* workspace.rs:112 ConnectError::TypeMismatch */
    MXL_ERR_OOM = -9,
    MXL_ERR_UNSUPPORTED = -10,
    MXL_ERR_LENGTH = -11         /* line shorter than the call needs (Rust: slice index panic) */
} mxl_status;

#define MXL_DEVICE_NONE (-1)

/* protocol/src/lib.rs:176-181 */
typedef enum mxl_line_type { MXL_LINE_MONO = 0, MXL_LINE_STEREO = 1, MXL_LINE_VIDEO = 2 } mxl_line_type;

/* src/module/mod.rs:28-49 enumerate_modules!, same order.  Kinds marked (io) are I/O edges that
 * stay in the reference; they are listed so the numbering matches protocol ModuleParams.
 * METER and SOURCE_* have no reference counterpart. */
typedef enum mxl_module_kind {
    MXL_MOD_AMPLIFIER = 0,
    MXL_MOD_ENVELOPE = 1,
    MXL_MOD_EQ_THREE = 2,
    MXL_MOD_FM_SINE = 3,
    MXL_MOD_MIXER = 4,
    MXL_MOD_MONITOR = 5,          /* run_tick + the codec thread's feed up to the encoder calls provided
                                   * (monitor.rs:112-140,235-247; EncodeStream, src/video/encode.rs:34-107,184-221);
                                   * x264 / fdk-aac and the websocket stay in the host application */
    MXL_MOD_OSCILLATOR = 6,
    MXL_MOD_OUTPUT_DEVICE = 7,    /* run_tick provided (channel routing into the device's interleaved buffer + clip,
                                   * output_device.rs:177-206); cpal and its callback stay in the host application */
    MXL_MOD_PLOTTER = 8,
    MXL_MOD_STEREO_PANNER = 9,
    MXL_MOD_STEREO_SPLITTER = 10,
    MXL_MOD_STREAM_INPUT = 11,    /* run_tick provided (queue assembly + gating, stream_input.rs:72-147); the
                                   * RTMP / Icecast receivers that feed it stay in the host application and push
                                   * with mxl_stream_input_write_audio / _write_video */
    MXL_MOD_STREAM_OUTPUT = 12,   /* the Monitor feed at 1120 x 700 while the host reports the RTMP connection Live
                                   * (mxl_stream_output_set_live); the publisher stays in the host application */
    MXL_MOD_TRIGGER = 13,
    MXL_MOD_VIDEO_MIXER = 14,
    MXL_MOD_MEDIA_SOURCE = 15,    /* (io) not provided */
    MXL_MOD_METER = 32,           /* new: per-tick peak / sum-of-squares / clip */
    MXL_MOD_SOURCE_MONO = 33,     /* new: host-fed line, stands where StreamInput's output was */
    MXL_MOD_SOURCE_STEREO = 34,
    MXL_MOD_SOURCE_VIDEO = 35,
    MXL_MOD_PCM_SINK = 36         /* new: f32 -> i16 pack of a stereo line (src/video/encode.rs:184-195) */
} mxl_module_kind;

/* ---- parameter PODs: mirrors of protocol/src/lib.rs structs, f64 fields stay f64 ----------- */

typedef struct mxl_amplifier_params { double amplitude; double mod_depth; } mxl_amplifier_params;      /* lib.rs:298-302 */
typedef struct mxl_envelope_params {                                                                   /* lib.rs:310-327 */
    double attack_ms, decay_ms, sustain_amplitude, release_ms;
} mxl_envelope_params;
typedef struct mxl_eq_three_params { double gain_lo_db, gain_mid_db, gain_hi_db; } mxl_eq_three_params; /* lib.rs:285-290 */
typedef struct mxl_fm_sine_params { double freq_lo, freq_hi; } mxl_fm_sine_params;                     /* lib.rs:292-296 */
typedef struct mxl_mixer_channel_params { double gain_db; double fader; int32_t cue; int32_t _pad; } mxl_mixer_channel_params; /* lib.rs:342-347 */
typedef struct mxl_mixer_params { const mxl_mixer_channel_params *channels; uint32_t n_channels; } mxl_mixer_params;          /* lib.rs:329-332 */
/* lib.rs:233-241, declaration order */
typedef enum mxl_waveform { MXL_WAVE_ON = 0, MXL_WAVE_OFF = 1, MXL_WAVE_SINE = 2, MXL_WAVE_SQUARE = 3,
                            MXL_WAVE_TRIANGLE = 4, MXL_WAVE_SAW = 5 } mxl_waveform;
typedef struct mxl_oscillator_params { double freq; int32_t waveform; int32_t _pad; } mxl_oscillator_params; /* lib.rs:243-247 */
typedef enum mxl_gate_state { MXL_GATE_OPEN = 0, MXL_GATE_CLOSED = 1 } mxl_gate_state;                   /* lib.rs:304-308 */
typedef struct mxl_trigger_params { int32_t gate; } mxl_trigger_params;
/* lib.rs:405-420; a / b = -1 encodes Option::None */
typedef struct mxl_video_mixer_params { int32_t a; int32_t b; double fader; } mxl_video_mixer_params;
#define MXL_VIDEO_MIXER_CHANNELS 4                                                                     /* lib.rs:403 */

/* ---- context --------------------------------------------------------------------------------- */

/* sample_rate / samples_per_tick are the reference's compile-time SAMPLE_RATE / SAMPLES_PER_TICK
 * (src/engine.rs:52-55) made runtime parameters. */
MXL_API mxl_ctx *mxl_ctx_create(int device, uint32_t sample_rate, uint32_t samples_per_tick);
/* Same, but launches on a caller-owned cudaStream_t (so the caller can time with its own events). */
MXL_API mxl_ctx *mxl_ctx_create_on_stream(int device, uint32_t sample_rate, uint32_t samples_per_tick,
                                          void *cuda_stream);
MXL_API int mxl_ctx_destroy(mxl_ctx *ctx);
MXL_API int mxl_ctx_synchronize(mxl_ctx *ctx);
/* Released frame buffers are parked per size class for reuse (the reference mallocs a frame per tick,
 * src/module/video_mixer.rs:150-160 via AvFrame::blank).  Frees parked buffers until at most keep_bytes remain
 * (0 = everything; call it after a source changed resolution) and, when new_cap_bytes != 0, sets the cap beyond which a
 * released buffer goes straight back to the driver (default 8 GiB).  freed_out (may be NULL) = bytes returned. */
MXL_API int mxl_ctx_trim_frame_pool(mxl_ctx *ctx, uint64_t keep_bytes, uint64_t new_cap_bytes, uint64_t *freed_out);
/* Copy/compute overlap for host-fed sessions.  When enabled, the *_async upload calls run on an
 * upload stream and the *_async download calls on a download stream, ordered against the compute
 * stream by events (upload waits for the last run, a run waits for the uploads and downloads enqueued
 * before it, a download waits for the last run), so step k+1's uploads overlap step k's downloads on
 * the full-duplex bus.  Synchronous calls and mxl_ctx_synchronize still wait for everything. */
MXL_API int mxl_ctx_set_copy_overlap(mxl_ctx *ctx, int enabled);
/* Marks "all downloads enqueued so far" (slot 0..3) / blocks the host until that mark is reached. */
MXL_API int mxl_ctx_download_fence(mxl_ctx *ctx, uint32_t slot);
MXL_API int mxl_ctx_wait_fence(mxl_ctx *ctx, uint32_t slot);
/* Bytes moved by the upload / download entry points of this context so far. */
MXL_API uint64_t mxl_ctx_h2d_bytes(const mxl_ctx *ctx);
MXL_API uint64_t mxl_ctx_d2h_bytes(const mxl_ctx *ctx);
MXL_API void *mxl_ctx_stream(mxl_ctx *ctx);
MXL_API uint32_t mxl_ctx_sample_rate(const mxl_ctx *ctx);
MXL_API uint32_t mxl_ctx_samples_per_tick(const mxl_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
MXL_API uint64_t mxl_ctx_launch_count(const mxl_ctx *ctx);
/* Device-timed interval on the context's stream: begin/end record CUDA events, elapsed_ms
 * synchronises on the end event. */
MXL_API int mxl_ctx_timer_begin(mxl_ctx *ctx);
MXL_API int mxl_ctx_timer_end(mxl_ctx *ctx);
MXL_API int mxl_ctx_timer_elapsed_ms(mxl_ctx *ctx, float *ms);
/* Overwrites a scratch buffer larger than L2 (bench hygiene between timed iterations). */
MXL_API int mxl_ctx_flush_l2(mxl_ctx *ctx);
/* Free / total device memory of the context's GPU (cudaMemGetInfo), after synchronising the context. */
MXL_API int mxl_ctx_device_memory(mxl_ctx *ctx, uint64_t *free_bytes, uint64_t *total_bytes);
/* Per-kernel device timing: while enabled, every kernel launch of the context is bracketed by a CUDA event
 * pair on its launching stream (host preparation and table copies stay outside).  mxl_ctx_kernel_times
 * synchronises, folds the pairs recorded since the last call into one entry per kernel and returns the
 * number of entries written.  (The reference's EngineStat, src/engine/timing.rs:46-60, at kernel grain.) */
typedef struct mxl_kernel_time { char name[48]; uint32_t launches; float total_ms; } mxl_kernel_time;
MXL_API int mxl_ctx_set_kernel_timing(mxl_ctx *ctx, int enabled);
MXL_API int mxl_ctx_kernel_times(mxl_ctx *ctx, mxl_kernel_time *out, uint32_t cap);
/* Diagnostics of the fused voice kernel: with max_ctas > 0 every later launch of at most that many CTAs records 8 SM-clock
 * stamps per CTA (start, table copies issued, samples generated, scanned, filtered, stored; two spare).  The call first
 * copies the stamps of the LAST such launch to stamps_out (8 per CTA, voice-major; returns the CTA count), then applies
 * max_ctas (0 = off).  Synchronises. */
MXL_API int mxl_ctx_fused_profile(mxl_ctx *ctx, uint32_t max_ctas, uint64_t *stamps_out, uint32_t cap_ctas);
MXL_API const char *mxl_last_error(void);
MXL_API const char *mxl_version(void);
/* Decibel::to_linear, protocol/src/lib.rs:469-471 (host scalar; exposed for tests) */
MXL_API double mxl_db_to_linear(double db);

/* Host-fed sessions on a multi-socket box: moves the CALLING thread (the engine thread of this context) onto the CPUs of
 * the NUMA node the context's GPU hangs off (/sys/bus/pci/devices/<bdf>/numa_node) and makes that node the preferred
 * one for its later allocations, so that pinned staging buffers (mxl_host_alloc) are first-touched next to the GPU.
 * node_out = the node, or -1 when the platform reports none (nothing was changed); cpus_out = CPUs in the new mask. */
MXL_API int mxl_ctx_bind_host_to_gpu_node(mxl_ctx *ctx, int32_t *node_out, int32_t *cpus_out);
/* Pinned host memory for upload/download staging. */
MXL_API void *mxl_host_alloc(size_t bytes);
MXL_API int mxl_host_free(void *p);

/* ---- audio lines: Output::{Mono,Stereo} of src/engine/io.rs:64-77, device-resident ------------- */

/* `frames` samples per channel; zero-filled like `vec![0.0; N]` (io.rs:73-74). */
MXL_API mxl_line *mxl_line_alloc(mxl_ctx *ctx, int line_type, uint64_t frames);
MXL_API int mxl_line_free(mxl_line *line);
MXL_API int mxl_line_type_of(const mxl_line *line);
MXL_API uint64_t mxl_line_frames(const mxl_line *line);
/* number of f32 in the line (frames for mono, 2*frames for stereo) */
MXL_API uint64_t mxl_line_len(const mxl_line *line);
MXL_API void *mxl_line_device_ptr(mxl_line *line);
MXL_API int mxl_line_zero(mxl_line *line);
/* host <-> device, ordered on the context stream; both return after the copy has completed */
MXL_API int mxl_line_upload(mxl_line *line, const float *host, uint64_t n_floats);
MXL_API int mxl_line_download(const mxl_line *line, float *host, uint64_t n_floats);
/* asynchronous variants for pinned host memory (mxl_host_alloc); complete at mxl_ctx_synchronize */
MXL_API int mxl_line_upload_async(mxl_line *line, const float *host, uint64_t n_floats);
MXL_API int mxl_line_download_async(const mxl_line *line, float *host, uint64_t n_floats);

/* ---- video frames: AvFrame<Video> in yuv420p, codec/src/ffmpeg/frame.rs:76-138 ---------------- */

typedef struct mxl_frame_layout {
    uint32_t width, height;     /* luma size */
    uint32_t stride[3];         /* bytes per row, each a multiple of 32 (video_mixer.rs:196-201) */
    uint32_t plane_h[3];
    uint64_t offset[3];         /* byte offset of each plane in the frame buffer */
    uint64_t size;              /* bytes */
} mxl_frame_layout;

/* Host-side geometry; valid without a device. */
MXL_API int mxl_frame_layout_yuv420p(uint32_t width, uint32_t height, mxl_frame_layout *out);
/* video_mixer.rs:276-297 unify_picture_settings */
MXL_API int mxl_unify_picture_settings(uint32_t aw, uint32_t ah, uint32_t bw, uint32_t bh,
                                       uint32_t *w, uint32_t *h);
/* src/video/encode.rs:354-374 DynamicScaler letterbox geometry */
typedef struct mxl_scale_geometry { uint32_t scaled_w, scaled_h, letterbox_x, letterbox_y; } mxl_scale_geometry;
MXL_API int mxl_scale_geometry_yuv420p(uint32_t in_w, uint32_t in_h, uint32_t out_w, uint32_t out_h,
                                       mxl_scale_geometry *out);
/* video_mixer.rs:168 `(fader * 255.0) as u8` */
MXL_API uint8_t mxl_fader_to_u8(double fader);

/* Frames are reference-counted like av_frame_clone / av_frame_free (frame.rs:351-361). */
MXL_API mxl_frame *mxl_frame_alloc(mxl_ctx *ctx, uint32_t width, uint32_t height);
/* AvFrame::blank: Y=0x00, U=V=0x80 (frame.rs:128-134), filled by a kernel */
MXL_API mxl_frame *mxl_frame_blank(mxl_ctx *ctx, uint32_t width, uint32_t height);
MXL_API mxl_frame *mxl_frame_retain(mxl_frame *frame);
MXL_API int mxl_frame_release(mxl_frame *frame);
MXL_API int mxl_frame_get_layout(const mxl_frame *frame, mxl_frame_layout *out);
MXL_API void *mxl_frame_device_ptr(mxl_frame *frame);
/* plane-wise copy with caller strides (what AvFrame::frame_data() exposes, frame.rs:188-197) */
MXL_API int mxl_frame_upload(mxl_frame *frame, const uint8_t *const planes[3], const uint32_t strides[3]);
MXL_API int mxl_frame_download(const mxl_frame *frame, uint8_t *const planes[3], const uint32_t strides[3]);
/* whole buffer in mxl_frame_layout order, padding included */
MXL_API int mxl_frame_upload_raw(mxl_frame *frame, const uint8_t *host, uint64_t size);
MXL_API int mxl_frame_download_raw(const mxl_frame *frame, uint8_t *host, uint64_t size);
MXL_API int mxl_frame_upload_raw_async(mxl_frame *frame, const uint8_t *host, uint64_t size);
/* n frames of one size in ONE device allocation, adjacent in memory (freed with the last of them): what an
 * ingest that hands over several frames per engine call allocates, so that mxl_frames_upload_raw_async /
 * _download_raw_async move every run of adjacent frames as a single copy (host = n * bytes_each contiguous
 * bytes).  Frames from mxl_frame_alloc work too, one copy each. */
MXL_API int mxl_frames_alloc_batch(mxl_ctx *ctx, uint32_t width, uint32_t height, uint32_t n, mxl_frame **frames_out);
MXL_API int mxl_frames_upload_raw_async(mxl_frame *const *frames, uint32_t n, const uint8_t *host, uint64_t bytes_each);
MXL_API int mxl_frames_download_raw_async(mxl_frame *const *frames, uint32_t n, uint8_t *host, uint64_t bytes_each);
MXL_API int mxl_frame_download_raw_async(const mxl_frame *frame, uint8_t *host, uint64_t size);

/* Video line = Output::Video(Option<VideoFrame>) per tick (io.rs:8-17,64-69): `ticks` slots. */
MXL_API mxl_line *mxl_video_line_alloc(mxl_ctx *ctx, uint32_t ticks);
/* frame may be NULL (None).  duration_hint / tick_offset are MediaDuration rationals
 * (src/video.rs, util/src/time.rs:77-105).  The line retains the frame. */
MXL_API int mxl_video_line_set(mxl_line *line, uint32_t slot, mxl_frame *frame, int64_t duration_num,
                               int64_t duration_den, int64_t offset_num, int64_t offset_den);
/* borrowed pointer, NULL if the slot is empty */
MXL_API mxl_frame *mxl_video_line_get(const mxl_line *line, uint32_t slot);
/* duration_hint and tick_offset of a slot's VideoFrame (io.rs:11-17) as {numerator, denominator};
 * MXL_ERR_INVALID if the slot is empty */
MXL_API int mxl_video_line_get_timing(const mxl_line *line, uint32_t slot, int64_t duration[2], int64_t offset[2]);
MXL_API int mxl_video_line_clear(mxl_line *line);

/* ---- modules: trait ModuleT, src/module/mod.rs:7-19 ------------------------------------------- */

/* create(params, ctx).  `params` points at the POD of that kind (NULL for kinds with `()`). */
MXL_API mxl_module *mxl_module_create(mxl_ctx *ctx, int kind, const void *params);
MXL_API void mxl_module_destroy(mxl_module *m);
MXL_API int mxl_module_kind_of(const mxl_module *m);
/* update(new_params); kind-checked against the module (module.rs:104-110) */
MXL_API int mxl_module_update(mxl_module *m, int kind, const void *params);
/* params(): copies the POD out.  For the mixer, `channels_out` receives up to `cap` channels and
 * the return value is the channel count. */
MXL_API int mxl_module_params(const mxl_module *m, void *params_out);
MXL_API int mxl_mixer_params_get(const mxl_module *m, mxl_mixer_channel_params *channels_out, uint32_t cap);
/* inputs() / outputs(): Terminal(label, LineType), protocol/src/lib.rs:160-174 */
MXL_API uint32_t mxl_module_n_inputs(const mxl_module *m);
MXL_API uint32_t mxl_module_n_outputs(const mxl_module *m);
MXL_API int mxl_module_input_type(const mxl_module *m, uint32_t index);
MXL_API int mxl_module_output_type(const mxl_module *m, uint32_t index);
MXL_API const char *mxl_module_input_label(const mxl_module *m, uint32_t index);   /* NULL = unlabeled */
MXL_API const char *mxl_module_output_label(const mxl_module *m, uint32_t index);
/* run_tick(t, inputs, outputs).  The number of frames processed is taken from the lines, as the
 * reference takes it from slice lengths.  Asynchronous on the context stream. */
MXL_API int mxl_module_run_tick(mxl_module *m, uint64_t t, const mxl_line *const *inputs, uint32_t n_inputs,
                                mxl_line *const *outputs, uint32_t n_outputs);

/* The reference's own run_tick, host slices in and out: what the UNMODIFIED engine loop hands a module
 * (src/engine.rs:461-494).  One mxl_host_ref per terminal mirrors InputRef::{Disconnected, Mono(&[f32]),
 * Stereo(&[f32]), Video(Option<&VideoFrame>)} and OutputRef::{Mono(&mut [f32]), Stereo(&mut [f32]),
 * Video(&mut Option<VideoFrame>)} (src/engine/io.rs:19-34,79-98).  Audio slices are ordinary host memory:
 * inputs are copied to device lines cached in the module, the module runs, outputs are copied back, and
 * the call returns when the output slices are complete (it synchronises the context).  The number of frames
 * is taken from the slice lengths, as in the reference.  A video terminal carries a frame handle
 * (Arc<AvFrame> there, a retained mxl_frame here: upload once with mxl_frame_upload*, pass it to as many
 * ticks and modules as hold it); a video output is written by the call -- NULL = None, else a frame the
 * caller owns one reference of (mxl_frame_download, mxl_frame_release).  A `type` that differs from the
 * terminal's is MXL_ERR_TYPE_MISMATCH (the reference panics, io.rs:40-41,49-50,58-59).
 * This is the literal drop-in (one bus round trip per module per tick); mxl_graph_run_ticks is the fast path. */
typedef struct mxl_host_ref {
    int32_t type;               /* mxl_line_type of the terminal */
    int32_t connected;          /* inputs: 0 = InputRef::Disconnected (the other fields are ignored) */
    float *samples;             /* Mono / Stereo: first sample of the slice */
    uint64_t len;               /* f32 in the slice: frames (mono), 2 * frames (stereo) */
    mxl_frame *frame;           /* Video: Some(frame) / NULL = None */
    int64_t duration_num, duration_den;   /* video::Frame.duration_hint (src/video.rs), seconds as a ratio */
    int64_t offset_num, offset_den;       /* engine::VideoFrame.tick_offset (io.rs:11-17) */
} mxl_host_ref;
MXL_API int mxl_module_run_tick_host(mxl_module *m, uint64_t t, const mxl_host_ref *inputs, uint32_t n_inputs,
                                     mxl_host_ref *outputs, uint32_t n_outputs);

/* State read-back (synchronises).  EqThree: lo poles[4], hi poles[4], history[3]. */
MXL_API int mxl_eq_three_state(mxl_module *m, double state[11]);
/* Envelope: state (0 Initial, 1 TriggerOn, 2 TriggerOff), seq, off_amplitude */
MXL_API int mxl_envelope_state(mxl_module *m, int32_t *state, uint64_t *seq, double *off_amplitude);
/* Meter: values of tick slot `slot` of the last call */
MXL_API int mxl_meter_read(mxl_module *m, uint32_t slot, float peak[2], double sumsq[2], int32_t *clip);
/* All tick slots of the last call in one copy (synchronises).  Returns the number of records. */
typedef struct mxl_meter_record { float peak[2]; int32_t clip; int32_t _pad; double sumsq[2]; } mxl_meter_record;
MXL_API int mxl_meter_download(mxl_module *m, mxl_meter_record *records, uint32_t cap);
/* Same copy without the synchronisation (records should be pinned memory). */
MXL_API int mxl_meter_download_async(mxl_module *m, mxl_meter_record *records, uint32_t cap);
/* Plotter (plotter.rs:37-56): de-interleaved tap of the most recent tick whose count % 6 == 0
 * within the last call.  Returns the number of frames written (0 = no indication). */
MXL_API int mxl_plotter_read(mxl_module *m, float *left, float *right, uint32_t cap_frames);
/* Source modules: the line (audio or video) that the module presents as its output. */
MXL_API int mxl_source_set_line(mxl_module *m, mxl_line *line);
/* PCM sink: packed i16 of the last call (src/video/encode.rs:184-195) */
MXL_API int mxl_pcm_sink_download(mxl_module *m, int16_t *host, uint64_t n_samples);

/* ---- StreamInput (src/module/stream_input.rs): the module just before the path ------------------
 * outputs: Video "Video", Stereo "Audio" (stream_input.rs:44-47); params: none here (protocol and mountpoint
 * select the receiver, which stays in the host application).  The receiver side pushes what
 * SourceSend::write_audio / write_video push (src/source.rs:156-190): Frame { source_id, source_time, data }.
 * Audio data is interleaved i16 as the decoder delivers it: it crosses the bus at 2 B/sample and is converted
 * by the device (convert_sample, stream_input.rs:167-173).  run_tick then does what the reference does per tick
 * (72-147): fills the tick's stereo line from the queued frames (several frames, or part of one), zero-fills on
 * underrun, re-bases the source clock when the source id changes, and holds a video frame back until it is
 * due (tick_offset <= tick duration).  A call over n ticks (video line of n slots, stereo line of n * S frames)
 * is n such ticks with one upload and one conversion launch.
 * Queues hold 65536 frames like the reference's ring buffers (source.rs:97-98); a full queue is MXL_ERR_LENGTH
 * (write_* returns Err(()) there).  `frame` is retained by the queue. */
MXL_API int mxl_stream_input_write_audio(mxl_module *m, uint64_t source_id, int64_t time_num, int64_t time_den,
                                         const int16_t *samples, uint64_t n_samples);
MXL_API int mxl_stream_input_write_video(mxl_module *m, uint64_t source_id, int64_t time_num, int64_t time_den,
                                         mxl_frame *frame, int64_t duration_num, int64_t duration_den);
/* Frames waiting in the two queues (a partly consumed audio frame / a held-back video frame count as one). */
MXL_API int mxl_stream_input_pending(const mxl_module *m, uint32_t *audio_frames, uint32_t *video_frames);

/* ---- Monitor (src/module/monitor.rs) + EncodeStream (src/video/encode.rs): the module just after the path ----
 * inputs: Video "Video", Stereo "Audio" (monitor.rs:97-100); no outputs.  run_tick does what Monitor::run_tick
 * (112-140) and the codec thread's loop body (235-247) do with a tick, up to the two encoder calls:
 *   timestamp = tick time - time of the module's first tick (epoch);
 *   AudioCtx::send_audio (encode.rs:184-221): samples clamped to [-1, 1], * 32767, `as i16` -- on the device, the
 *     whole call in one launch, downloaded at 2 B/sample -- appended to the PCM buffer; whenever it holds MORE than
 *     2 * 1024 samples, one fragment of exactly 2 * 1024 leaves it (decode_timestamp = running audio clock,
 *     duration = 1024 / sample_rate): what aac::Encoder::encode is called with;
 *   EncodeStream::send_video (61-76): frame_timestamp = timestamp + tick_offset; a frame that ends before the video
 *     clock is dropped, else its duration becomes end - video clock;  EncodeStream::barrier (78-84): a gap up to the
 *     tick's timestamp is filled with the blank frame;  encode_video (86-100): pts = round_to_base(start),
 *     duration = round_to_base(end) - pts in `time_base` units; VideoCtx::send_frame (279-287): the picture is
 *     letterbox-scaled to the monitor's size (DynamicScaler) -- on the device, one launch per geometry per call:
 *     what AvcEncoder::send_frame is called with.
 * The reference drops a tick when its 2-deep channel to the codec thread is full (monitor.rs:162-172): timing
 * dependent, not mirrored -- every tick is processed.
 * mxl_monitor_recv_audio / _video pop the two feeds in order; each returns 1 if it wrote an entry, 0 if none is
 * pending, negative on error.  recv_audio waits for the download that carries the fragment. */
typedef struct mxl_monitor_params { uint32_t width, height; int64_t time_base; } mxl_monitor_params;   /* 560 x 350, SAMPLE_RATE (monitor.rs:21-22,193-196) */
typedef struct mxl_audio_fragment {
    int64_t decode_num, decode_den;        /* AudioSegment.decode_timestamp */
    int64_t duration_num, duration_den;    /* AudioSegment.duration = 1024 / sample_rate */
    uint32_t n_samples, _pad;              /* interleaved i16 written: 2 * 1024 */
} mxl_audio_fragment;
typedef struct mxl_video_job {
    int64_t pts;                           /* frame.set_presentation_timestamp(frame_start_in_base) */
    int64_t duration;                      /* duration_in_base */
    int64_t time_base;
    int32_t blank, _pad;                   /* 1 = VideoCtx::blank_frame() filling a gap */
    mxl_frame *frame;                      /* monitor-sized picture, one reference owned by the caller */
} mxl_video_job;
MXL_API int mxl_monitor_recv_audio(mxl_module *m, mxl_audio_fragment *info, int16_t *pcm, uint32_t cap_samples);
MXL_API int mxl_monitor_recv_video(mxl_module *m, mxl_video_job *out);
/* ---- OutputDevice (src/module/output_device.rs) ---------------------------------------------------
 * input: Stereo (unlabeled).  params: the output channel each side is routed to (-1 = None) and the channel count
 * of the opened device stream (stream.config.channels; 0 = no stream: run_tick queues nothing).  update() zeroes the
 * scratch buffer when an assignment changes and drops assignments beyond the channel count (152-168).  run_tick
 * (173-206) writes left then right into the interleaved scratch buffer, sets the clip flag when a routed sample lies
 * outside [-1, 1] and queues frames * channels samples for the device callback; the queue holds 65536 samples like
 * the reference's ring buffer (128) and takes only what fits.  mxl_output_device_read pops what the cpal callback
 * would pop (waits for the download); mxl_output_device_clip returns the flag of the last run_tick (synchronises;
 * the Instant-based temporal warnings, 208-230, stay with the host). */
typedef struct mxl_output_device_params { int32_t left, right; uint32_t channels, _pad; } mxl_output_device_params;
MXL_API int64_t mxl_output_device_read(mxl_module *m, float *out, uint64_t cap_samples);
MXL_API int mxl_output_device_clip(mxl_module *m, int32_t *clip);

/* StreamOutput (src/module/stream_output.rs): the same two feeds (LiveOutput::tick, 369-381) at 1120 x 700 (13-14),
 * read with mxl_monitor_recv_*.  The RTMP connection state machine lives in the host application; it reports
 * Connection::Live with live = 1 (a new EncodeStream; the next tick is the epoch, 126,328-366) and anything else
 * with live = 0 (run_tick sends nothing, 112-151).  Created Offline. */
MXL_API int mxl_stream_output_set_live(mxl_module *m, int live);

/* Stand-alone PCM converters on raw device memory of the context (N2/N3 rows of SURVEY §8f):
 * i16 -> f32 `sample / 32768.0` (stream_input.rs:167-173) and the pack above. */
MXL_API int mxl_pcm_unpack_i16(mxl_ctx *ctx, const int16_t *host_pcm, uint64_t n_samples, mxl_line *dst);
MXL_API int mxl_pcm_pack_i16(mxl_ctx *ctx, const mxl_line *src, int16_t *host_pcm, uint64_t n_samples);
/* The same without allocation or synchronisation per call (device staging is a ring owned by the context):
 * host_pcm should be pinned (mxl_host_alloc) and stay valid until the context has been synchronised or a
 * download fence has passed.  With copy overlap enabled the bus transfer runs on the copy streams. */
MXL_API int mxl_pcm_unpack_i16_async(mxl_ctx *ctx, const int16_t *host_pcm, uint64_t n_samples, mxl_line *dst);
MXL_API int mxl_pcm_pack_i16_async(mxl_ctx *ctx, const mxl_line *src, int16_t *host_pcm, uint64_t n_samples);

/* ---- audio sample-rate converter (north_star "resample"; NEW and self-specified -- the reference has only the TODO:
 * Icecast ingest drops every stream that is not at the engine's rate, src/icecast/mod.rs:94-97, RTMP ingest panics,
 * src/rtmp/mod.rs:229-232).  What a receiver calls between its decoder and mxl_stream_input_write_audio / a source line:
 * interleaved i16 (or an f32 line) at the source's rate in, an f32 line at out_rate out.  Polyphase windowed sinc for
 * the exact ratio (44.1 k -> 48 k: 160 / 147), 32 taps per output frame accumulated in f64 with fma; the exact definition
 * is oracle/mixlab_oracle.h (orc_resampler_*), parity unpinned.  The converter is a pure function of the stream pushed so
 * far: any split into calls gives the same bits.  After N input frames, ceil((N - 16) * L / M) output frames are
 * determined; each push resizes `out` to the frames it produced and returns that count (>= 0) or a negative status. */
typedef struct mxl_resampler mxl_resampler;
MXL_API mxl_resampler *mxl_resampler_create(mxl_ctx *ctx, uint32_t in_rate, uint32_t out_rate, uint32_t channels);
MXL_API void mxl_resampler_destroy(mxl_resampler *r);
MXL_API int mxl_resampler_reset(mxl_resampler *r);
MXL_API uint64_t mxl_resampler_output_frames(const mxl_resampler *r, uint64_t in_frames);
MXL_API int64_t mxl_resampler_push_i16(mxl_resampler *r, const int16_t *host_pcm, uint64_t in_frames, mxl_line *out);
MXL_API int64_t mxl_resampler_push_line(mxl_resampler *r, const mxl_line *in, mxl_line *out);

/* yuv420p -> RGBA8 of a frame (self-specified BT.601 integer form; the reference
 * never converts colour, video_mixer.rs:282-283).  rgba_host receives width*height*4 bytes. */
MXL_API int mxl_frame_to_rgba(const mxl_frame *frame, uint8_t *rgba_host);
/* Letterboxed bicubic rescale of `src` into a new frame of (out_w,out_h): DynamicScaler::scale
 * (src/video/encode.rs:338-397).  Returns a retained `src` when sizes are equal (342-345). */
MXL_API mxl_frame *mxl_frame_scale(mxl_frame *src, uint32_t out_w, uint32_t out_h);
/* The same for n frames that share one source size: all planes of all frames in ONE launch of the
 * tiled scaler (what VideoMixer does with the frames of a multi-tick call).  dst[i] receives a new
 * frame each (a retained src[i] when the sizes already agree). */
MXL_API int mxl_frames_scale(mxl_ctx *ctx, mxl_frame *const *src, mxl_frame **dst, uint32_t n,
                             uint32_t out_w, uint32_t out_h);

/* Device-resident RGBA8 pictures (width*height*4 bytes each, rows packed), the output of the
 * compositor's colour-converting path (BASELINE config 3).  Not a reference type: the reference
 * keeps everything yuv420p (video_mixer.rs:282-283). */
typedef struct mxl_rgba mxl_rgba;
MXL_API mxl_rgba *mxl_rgba_alloc(mxl_ctx *ctx, uint32_t width, uint32_t height, uint32_t n_pictures);
MXL_API int mxl_rgba_free(mxl_rgba *pics);
MXL_API void *mxl_rgba_device_ptr(mxl_rgba *pics, uint32_t index);
MXL_API int mxl_rgba_download(const mxl_rgba *pics, uint32_t first, uint32_t count, uint8_t *host);
MXL_API int mxl_rgba_download_async(const mxl_rgba *pics, uint32_t first, uint32_t count, uint8_t *host);
/* VideoMixer's crossfade (video_mixer.rs:150-239: out = (a*f + b*(255-f)) / 255 per byte, a missing
 * layer = blank) of n layer pairs, written directly as RGBA8 of the blend -- one pass, one launch:
 * the layers are read once and no yuv420p composite is stored.  a[i] / b[i] may be NULL.  Bit-exact
 * to mxl_frame_to_rgba of the VideoMixer output for the same inputs. */
MXL_API int mxl_video_compose_rgba(mxl_ctx *ctx, mxl_frame *const *a, mxl_frame *const *b, uint32_t n,
                                   double fader, mxl_rgba *out, uint32_t first_picture);
/* yuv420p -> RGBA8 of n frames of one size, one launch (the compose path with a single layer). */
MXL_API int mxl_frames_to_rgba(mxl_ctx *ctx, mxl_frame *const *frames, uint32_t n, mxl_rgba *out,
                               uint32_t first_picture);

/* The other direction (north_star "YUV<->RGB"; NEW, self-specified like the conversion above -- the reference keeps
 * everything yuv420p, video_mixer.rs:282-283): RGBA8 pictures -> yuv420p frames of the same size, BT.601 limited range,
 * Y per pixel, U / V from the rounded mean colour of each 2x2 block (an odd edge repeats its last column / row), alpha
 * ignored; one launch for n pictures.  Exact integer definition: oracle/mixlab_oracle.h (orc_rgba_to_yuv420p).  What an
 * RGB source (a screen capture, a rendered title) needs before it can enter VideoMixer.  Stride padding of the frames
 * is left untouched. */
MXL_API int mxl_rgba_upload(mxl_rgba *pics, uint32_t first, uint32_t count, const uint8_t *host);
MXL_API int mxl_rgba_to_frames(mxl_ctx *ctx, const mxl_rgba *pics, uint32_t first_picture, uint32_t n, mxl_frame *const *frames);

/* ---- optional shared-source mode (NEW: no reference counterpart; one receiver per mountpoint there,
 * src/source.rs:93-95).  Sessions on different GPUs share nothing on the tick path; when several graphs fan
 * out from ONE ingest, the ingest GPU broadcasts the source line / frame to the others over NVLink with NCCL
 * (bound at run time; these calls fail with MXL_ERR_INVALID where libnccl.so.2 is absent).  One process per GPU:
 * rank 0 makes the id, every rank passes it to mxl_ctx_comm_init.  Broadcasts are asynchronous on the context
 * stream, ordered with the kernels that produce / consume the buffer. */
#define MXL_COMM_ID_BYTES 128
MXL_API int mxl_comm_unique_id(uint8_t id_out[MXL_COMM_ID_BYTES]);
MXL_API int mxl_ctx_comm_init(mxl_ctx *ctx, const uint8_t id[MXL_COMM_ID_BYTES], int rank, int world);
MXL_API int mxl_ctx_comm_destroy(mxl_ctx *ctx);
MXL_API int mxl_line_broadcast(mxl_line *line, int root);
MXL_API int mxl_frame_broadcast(mxl_frame *frame, int root);

/* ---- graph: Workspace + Engine::run_tick, src/engine/workspace.rs, src/engine.rs:400-510 ------ */

MXL_API mxl_graph *mxl_graph_create(mxl_ctx *ctx);
MXL_API void mxl_graph_destroy(mxl_graph *g);
/* The graph takes ownership of the module.  Returns the ModuleId (>= 0) or a negative status. */
MXL_API int mxl_graph_add_module(mxl_graph *g, mxl_module *m);
MXL_API int mxl_graph_remove_module(mxl_graph *g, int module_id);
MXL_API mxl_module *mxl_graph_module(mxl_graph *g, int module_id);
/* Workspace::connect (workspace.rs:97-114) / disconnect (116-118) */
MXL_API int mxl_graph_connect(mxl_graph *g, int in_module, uint32_t in_index, int out_module, uint32_t out_index);
MXL_API int mxl_graph_disconnect(mxl_graph *g, int in_module, uint32_t in_index);
/* Run order the next run will use (terminal set + DFS, engine.rs:408-457).  Returns the count. */
MXL_API int mxl_graph_plan(mxl_graph *g, int *order_out, uint32_t cap);
/* Runs ticks tick0 .. tick0+n_ticks-1 (t = tick * samples_per_tick, engine.rs:490) as one batch:
 * every audio line holds n_ticks*S frames, every video line n_ticks slots. */
MXL_API int mxl_graph_run_ticks(mxl_graph *g, uint64_t tick0, uint32_t n_ticks);
/* Audio and video sub-graphs share no line, so by default their stages run on two streams of the
 * context concurrently (joined before the call returns to the stream's order).  0 = one stream. */
MXL_API int mxl_graph_set_stream_split(mxl_graph *g, int enabled);
/* Output line of a module after the last run (borrowed; valid until the next run or edit).
 * Fused voice groups: a sub-graph Oscillator -> EqThree -> StereoPanner -> Mixer [-> Meter] whose interior lines no
 * module outside it consumes runs as ONE launch (the reference runs the five modules one after another over fresh
 * buffers, src/engine.rs:464-507).  Oscillator and StereoPanner lines inside such a group are written only when
 * the host observes them: asking for one here (or mxl_graph_pin_output) makes it observed from the next run on; asked
 * after a run that did not write it, this call returns NULL and says so.  Mixer, Meter and EqThree outputs are always
 * written.  Results are those of the staged path (same kernels' arithmetic, same channel order). */
MXL_API mxl_line *mxl_graph_output(mxl_graph *g, int module_id, uint32_t out_index);
MXL_API int mxl_graph_pin_output(mxl_graph *g, int module_id, uint32_t out_index);
/* 0 = every module runs as its own stage (default 1). */
MXL_API int mxl_graph_set_fusion(mxl_graph *g, int enabled);
/* mxl_stage_info.kind of a fused voice group's stage (not a module kind) */
#define MXL_STAGE_FUSED_VOICE_MIX 1000
/* Per-launch device timings, shaped like EngineStat (src/engine/timing.rs:46-60,86-94). */
MXL_API int mxl_graph_set_profiling(mxl_graph *g, int enabled);
typedef struct mxl_stage_info {
    int32_t kind;              /* mxl_module_kind of the batched launch */
    int32_t n_modules;         /* module instances served by the launch */
    int32_t n_launches;        /* kernels launched by this stage in the last run */
    float last_ms;             /* device time of the last run (profiling on), else -1 */
    uint64_t algorithmic_bytes;/* API-level line bytes read+written by the stage in the last run */
    float host_us;             /* host time the engine thread spent enqueueing the stage in the last run
                                * (what a one-tick live call is bound by); always measured */
    float _pad;
} mxl_stage_info;
/* The same timings folded into PerformanceInfo.accounts (protocol/src/lib.rs:32-56; EngineStat::report,
 * src/engine/timing.rs:45-60): one account per module that ran in the last call -- its stage's device time divided
 * evenly among the stage's modules and by the ticks of the call, in microseconds per tick (`last`) -- and, first, the
 * Engine account (module_id = -1): the host time of the last mxl_graph_run_ticks that no stage accounts for
 * (planning, line bookkeeping, fork / join), per tick.  Needs mxl_graph_set_profiling(g, 1) before the run for the
 * module accounts (else they are -1).  Returns the number of accounts written. */
typedef struct mxl_perf_account { int32_t module_id; int32_t kind; float last_us; float host_us; } mxl_perf_account;
MXL_API int mxl_graph_performance(mxl_graph *g, mxl_perf_account *out, uint32_t cap);
MXL_API int mxl_graph_stage_count(mxl_graph *g);
MXL_API int mxl_graph_stage_info(mxl_graph *g, uint32_t stage, mxl_stage_info *out);

#ifdef __cplusplus
}
#endif
#endif

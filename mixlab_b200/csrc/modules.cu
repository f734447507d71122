// modules.cu -- host side of the modules (see modules.h).
#include "modules.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <deque>

namespace mxl {

// protocol/src/lib.rs:469-471  Decibel::to_linear
static double db_to_linear(double db) { return pow(10.0, db / 20.0); }

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(mxl_ctx* ctx, size_t bytes)
    {
        if (bytes <= cap) return MXL_OK;
        MXL_TRY(ctx->activate());
        if (p) { MXL_CUDA(cudaFree(p)); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 4;
        MXL_CUDA(cudaMalloc(&p, want));
        cap = want;
        return MXL_OK;
    }
    void release(mxl_ctx* ctx)
    {
        if (p) { if (ctx && ctx->has_device()) ctx->activate(); cudaFree(p); p = nullptr; cap = 0; }
    }
};

// ================================================================================================
// module classes
// ================================================================================================

struct Amplifier : mxl_module {                       // src/module/amplifier.rs
    mxl_amplifier_params p{};
    Amplifier(const mxl_amplifier_params* in)
    {
        kind = MXL_MOD_AMPLIFIER;
        if (in) p = *in;
        inputs = {labeled(MXL_LINE_STEREO, "Input"), labeled(MXL_LINE_MONO, "Control")};   // amplifier.rs:21-24
        outputs = {unlabeled(MXL_LINE_STEREO)};
    }
    int update(const void* np) override { p = *(const mxl_amplifier_params*)np; return MXL_OK; }
    int get_params(void* out) const override { *(mxl_amplifier_params*)out = p; return MXL_OK; }
};

struct Envelope : mxl_module {                        // src/module/envelope.rs
    mxl_envelope_params p{25.0, 500.0, 0.8, 200.0};   // protocol lib.rs:318-327
    DevBuf state, scratch;
    bool state_init = false;
    int cur = 0;                                      // which half of the state double buffer is current
    size_t tiles_cap = 0;                             // look-back descriptors allocated (scratch)
    Envelope(const mxl_envelope_params* in)
    {
        kind = MXL_MOD_ENVELOPE;
        if (in) p = *in;
        inputs = {unlabeled(MXL_LINE_MONO)};
        outputs = {unlabeled(MXL_LINE_MONO)};
    }
    ~Envelope() override { state.release(ctx); scratch.release(ctx); }
    int update(const void* np) override { p = *(const mxl_envelope_params*)np; return MXL_OK; }
    int get_params(void* out) const override { *(mxl_envelope_params*)out = p; return MXL_OK; }
    int ensure_state()
    {
        if (state_init) return MXL_OK;
        MXL_TRY(state.ensure(ctx, 2 * sizeof(k::EnvState)));
        MXL_CUDA(cudaMemsetAsync(state.p, 0, 2 * sizeof(k::EnvState), ctx->stream));   // EnvelopeState::Initial
        state_init = true;
        return MXL_OK;
    }
};

struct EqThree : mxl_module {                         // src/module/eq_three.rs
    mxl_eq_three_params p{};
    DevBuf state, zend;
    bool state_init = false;
    int cur = 0;                                      // which half of the state double buffer is current
    EqThree(const mxl_eq_three_params* in)
    {
        kind = MXL_MOD_EQ_THREE;
        if (in) p = *in;
        inputs = {unlabeled(MXL_LINE_MONO)};
        outputs = {unlabeled(MXL_LINE_MONO)};
    }
    ~EqThree() override { state.release(ctx); zend.release(ctx); }
    int update(const void* np) override { p = *(const mxl_eq_three_params*)np; return MXL_OK; }   // state survives (eq_three.rs:53-56)
    int get_params(void* out) const override { *(mxl_eq_three_params*)out = p; return MXL_OK; }
    int ensure_state()
    {
        if (state_init) return MXL_OK;
        // [2][11] doubles, then the poison words of eq_stream_kernel: {first non-finite chunk (none = ~0), CTAs done}
        MXL_TRY(state.ensure(ctx, 2 * 11 * sizeof(double) + 2 * sizeof(uint32_t)));
        MXL_CUDA(cudaMemsetAsync(state.p, 0, 2 * 11 * sizeof(double) + 2 * sizeof(uint32_t), ctx->stream));   // poles, history = 0 (eq_three.rs:34-40,113)
        MXL_CUDA(cudaMemsetAsync((char*)state.p + 2 * 11 * sizeof(double), 0xFF, sizeof(uint32_t), ctx->stream));
        state_init = true;
        return MXL_OK;
    }
    double* state_ptr(int which) { return (double*)state.p + 11 * which; }
    uint32_t* poison_ptr() { return (uint32_t*)((double*)state.p + 22); }
};

struct FmSine : mxl_module {                          // src/module/fm_sine.rs
    mxl_fm_sine_params p{};
    FmSine(const mxl_fm_sine_params* in)
    {
        kind = MXL_MOD_FM_SINE;
        if (in) p = *in;
        inputs = {unlabeled(MXL_LINE_MONO)};
        outputs = {unlabeled(MXL_LINE_STEREO)};
    }
    int update(const void* np) override { p = *(const mxl_fm_sine_params*)np; return MXL_OK; }
    int get_params(void* out) const override { *(mxl_fm_sine_params*)out = p; return MXL_OK; }
};

struct Mixer : mxl_module {                           // src/module/mixer.rs
    std::vector<mxl_mixer_channel_params> channels;
    std::vector<double> channel_gain;                 // fader * gain.to_linear() (mixer.rs:59), formed when the params arrive:
                                                      // the reference recomputes the same product (a powf) every tick
    DevBuf fused_products;                            // fused voice group: the channels' product lines (scratch, one per voice and gain)
    DevBuf fused_chan;                                // ... and the mix kernel's channel table on the device,
    std::vector<k::FusedChan> fused_chan_host;        //     uploaded when it differs from this copy
    // launch parameters of the group's last run: rebuilt only when something that feeds them changed (ctx->change_epoch,
    // the call length); a steady-state call patches t0 and the EqThree state pointers and launches
    struct FusedCache { uint64_t epoch = 0, frames = 0; const void* group = nullptr; k::FusedVoiceBatch vb; k::FusedMixBatch mb; };
    FusedCache* fused_cache = nullptr;
    ~Mixer() override { fused_products.release(ctx); fused_chan.release(ctx); delete fused_cache; }
    Mixer(const mxl_mixer_params* in)
    {
        kind = MXL_MOD_MIXER;
        set(in);
    }
    void set(const mxl_mixer_params* in)
    {
        channels.clear();
        if (in && in->channels) channels.assign(in->channels, in->channels + in->n_channels);
        channel_gain.resize(channels.size());
        for (size_t i = 0; i < channels.size(); i++) channel_gain[i] = channels[i].fader * db_to_linear(channels[i].gain_db);
        inputs.clear();
        for (size_t i = 0; i < channels.size(); i++) inputs.push_back(labeled(MXL_LINE_STEREO, std::to_string(i + 1)));   // mixer.rs:23-25
        outputs = {labeled(MXL_LINE_STEREO, "Master"), labeled(MXL_LINE_STEREO, "Cue")};                                  // mixer.rs:26-29
    }
    int update(const void* np) override { set((const mxl_mixer_params*)np); return MXL_OK; }   // mixer.rs:40-44 re-creates
    int get_params(void* out) const override
    {
        mxl_mixer_params* o = (mxl_mixer_params*)out;
        o->channels = channels.data();
        o->n_channels = (uint32_t)channels.size();
        return MXL_OK;
    }
};

struct Oscillator : mxl_module {                      // src/module/oscillator.rs
    mxl_oscillator_params p{100.0, MXL_WAVE_SINE, 0};
    Oscillator(const mxl_oscillator_params* in)
    {
        kind = MXL_MOD_OSCILLATOR;
        if (in) p = *in;
        outputs = {labeled(MXL_LINE_MONO, "Mono"), labeled(MXL_LINE_STEREO, "Stereo")};   // oscillator.rs:48-51
    }
    int update(const void* np) override { p = *(const mxl_oscillator_params*)np; return MXL_OK; }
    int get_params(void* out) const override { *(mxl_oscillator_params*)out = p; return MXL_OK; }
};

struct Plotter : mxl_module {                         // src/module/plotter.rs
    uint64_t count = 0;
    DevBuf tap;                                       // left[S] then right[S]
    uint32_t tap_frames = 0;
    Plotter()
    {
        kind = MXL_MOD_PLOTTER;
        inputs = {unlabeled(MXL_LINE_STEREO)};
    }
    ~Plotter() override { tap.release(ctx); }
    int update(const void*) override { return MXL_OK; }
    int get_params(void*) const override { return MXL_OK; }
};

struct StereoPanner : mxl_module {                    // src/module/stereo_panner.rs
    StereoPanner()
    {
        kind = MXL_MOD_STEREO_PANNER;
        inputs = {labeled(MXL_LINE_MONO, "L"), labeled(MXL_LINE_MONO, "R")};
        outputs = {unlabeled(MXL_LINE_STEREO)};
    }
    int update(const void*) override { return MXL_OK; }
    int get_params(void*) const override { return MXL_OK; }
};

struct StereoSplitter : mxl_module {                  // src/module/stereo_splitter.rs
    StereoSplitter()
    {
        kind = MXL_MOD_STEREO_SPLITTER;
        inputs = {unlabeled(MXL_LINE_STEREO)};
        outputs = {labeled(MXL_LINE_MONO, "L"), labeled(MXL_LINE_MONO, "R")};
    }
    int update(const void*) override { return MXL_OK; }
    int get_params(void*) const override { return MXL_OK; }
};

struct Trigger : mxl_module {                         // src/module/trigger.rs
    mxl_trigger_params p{MXL_GATE_CLOSED};
    Trigger(const mxl_trigger_params* in)
    {
        kind = MXL_MOD_TRIGGER;
        if (in) p = *in;
        outputs = {unlabeled(MXL_LINE_MONO)};
    }
    int update(const void* np) override { p = *(const mxl_trigger_params*)np; return MXL_OK; }
    int get_params(void* out) const override { *(mxl_trigger_params*)out = p; return MXL_OK; }
};

struct Meter : mxl_module {                           // new
    DevBuf records;
    uint32_t n_slots = 0;
    Meter()
    {
        kind = MXL_MOD_METER;
        inputs = {unlabeled(MXL_LINE_STEREO)};
    }
    ~Meter() override { records.release(ctx); }
    int update(const void*) override { return MXL_OK; }
    int get_params(void*) const override { return MXL_OK; }
};

struct Source : mxl_module {                          // new: host-fed line (stands in for StreamInput's output)
    mxl_line* line = nullptr;                         // borrowed
    Source(int k, int type)
    {
        kind = k;
        outputs = {unlabeled(type)};
    }
    int update(const void*) override { return MXL_OK; }
    int get_params(void*) const override { return MXL_OK; }
};

struct PcmSink : mxl_module {                         // new: f32 -> i16 pack (src/video/encode.rs:184-195)
    DevBuf pcm;
    uint64_t n_samples = 0;
    PcmSink()
    {
        kind = MXL_MOD_PCM_SINK;
        inputs = {unlabeled(MXL_LINE_STEREO)};
    }
    ~PcmSink() override { pcm.release(ctx); }
    int update(const void*) override { return MXL_OK; }
    int get_params(void*) const override { return MXL_OK; }
};

// ---- StreamInput: src/module/stream_input.rs -------------------------------------------------------
// The receiver (RTMP / Icecast) stays in the host application and pushes Frame { source_id, source_time, data }
// (src/source.rs:63-70) through mxl_stream_input_write_*; this class is the module's run_tick: the queue
// assembly and gating on the host (integer / rational bookkeeping, a few frames per tick), the sample
// conversion on the device.
struct StreamInput : mxl_module {
    static constexpr size_t kRingCapacity = 65536;                 // RingBuffer::new(65536), source.rs:97-98
    struct AudioFrame { uint64_t source_id = 0; Rational source_time; std::vector<int16_t> data; size_t head = 0; };
    struct VideoFrame { uint64_t source_id = 0; Rational source_time; mxl_frame* frame = nullptr; Rational duration_hint; };
    std::deque<AudioFrame> audio_rx;                               // SourceRecv::audio_rx
    std::deque<VideoFrame> video_rx;                               // SourceRecv::video_rx
    bool has_audio_frame = false, has_video_frame = false;         // self.audio_frame / self.video_frame (stream_input.rs:17-18)
    AudioFrame audio_frame;
    VideoFrame video_frame;
    bool has_source = false;                                       // Option<SourceTiming> (21-25)
    uint64_t source_id = 0;
    Rational epoch;
    // i16 of one call on its way to the device: two pinned buffers used in turn, each guarded by an event
    int16_t* pinned[2] = {nullptr, nullptr};
    size_t pinned_cap[2] = {0, 0};
    cudaEvent_t pinned_ev[2] = {nullptr, nullptr};
    int turn = 0;
    DevBuf pcm;

    StreamInput()
    {
        kind = MXL_MOD_STREAM_INPUT;
        outputs = {labeled(MXL_LINE_VIDEO, "Video"), labeled(MXL_LINE_STEREO, "Audio")};   // stream_input.rs:44-47
    }
    ~StreamInput() override
    {
        if (ctx && ctx->has_device()) { ctx->activate(); cudaStreamSynchronize(ctx->stream); }
        for (auto& v : video_rx) frame_release(v.frame);
        if (has_video_frame) frame_release(video_frame.frame);
        for (int i = 0; i < 2; i++) {
            if (pinned[i]) cudaFreeHost(pinned[i]);
            if (pinned_ev[i]) cudaEventDestroy(pinned_ev[i]);
        }
        pcm.release(ctx);
    }
    int update(const void*) override { return MXL_OK; }            // protocol / mountpoint select the receiver (host side)
    int get_params(void*) const override { return MXL_OK; }
    int staging(size_t samples, int16_t** out);
    int run(uint64_t t0, const IoSet& io, uint64_t* bytes);
};

// ---- Monitor: src/module/monitor.rs + EncodeStream (src/video/encode.rs) ------------------------------
struct Monitor : mxl_module {
    static constexpr uint64_t kFragmentSamples = 2 * 1024;         // AUDIO_CHANNELS * SAMPLES_PER_CHANNEL_PER_FRAGMENT (encode.rs:20-22,197)
    mxl_monitor_params p{560, 350, 0};                             // monitor.rs:21-22; time_base 0 = the context's sample rate
    bool has_epoch = false;                                        // self.epoch (monitor.rs:119)
    Rational epoch;
    Rational audio_timestamp, video_timestamp;                     // EncodeStream::new: MediaTime::new(0, 1)
    uint64_t pcm_len = 0;                                          // AudioCtx::pcm_buff.len()
    // the packed i16 stream, one pinned chunk per call, in order; fragments are consecutive windows of it
    struct Chunk { int16_t* host = nullptr; size_t cap = 0, n = 0; cudaEvent_t ev = nullptr; bool in_flight = false; };
    std::deque<Chunk> chunks;
    size_t chunk_head = 0;                                         // samples of chunks.front() already handed out
    std::vector<Chunk> free_chunks;
    struct Fragment { Rational decode_timestamp, duration; };
    std::deque<Fragment> audio_segments;
    struct Job { int64_t pts, duration; bool blank; mxl_frame* frame; };
    std::deque<Job> video_jobs;
    mxl_frame* blank = nullptr;                                    // VideoCtx::blank_frame (encode.rs:275)
    DevBuf pcm;

    // StreamOutput (src/module/stream_output.rs) feeds the same EncodeStream from LiveOutput::tick (369-381) once its
    // RTMP connection is Live: 1120 x 700 (13-14,347-351), epoch = the tick the connection completed (126,328-366).
    // The connection itself stays in the host application, which reports it with mxl_stream_output_set_live.
    bool live = true;
    Monitor(int k, const mxl_monitor_params* in)
    {
        kind = k;
        if (k == MXL_MOD_STREAM_OUTPUT) { p.width = 1120; p.height = 700; live = false; }   // Connection::Offline (stream_output.rs:42)
        if (in) p = *in;
        inputs = {labeled(MXL_LINE_VIDEO, "Video"), labeled(MXL_LINE_STEREO, "Audio")};   // monitor.rs:97-100; stream_output.rs:43-46
    }
    void reset_stream();
    ~Monitor() override
    {
        if (ctx && ctx->has_device()) { ctx->activate(); cudaStreamSynchronize(ctx->stream); }
        for (auto& j : video_jobs) frame_release(j.frame);
        frame_release(blank);
        for (auto& c : chunks) { if (c.host) cudaFreeHost(c.host); if (c.ev) cudaEventDestroy(c.ev); }
        for (auto& c : free_chunks) { if (c.host) cudaFreeHost(c.host); if (c.ev) cudaEventDestroy(c.ev); }
        pcm.release(ctx);
    }
    int update(const void*) override { return MXL_OK; }            // Params = () (monitor.rs:105-111)
    int get_params(void* out) const override { if (out) *(mxl_monitor_params*)out = p; return MXL_OK; }
    int64_t time_base() const { return p.time_base ? p.time_base : (int64_t)ctx->sample_rate; }
    int chunk_for(size_t samples, Chunk* out);
    void encode_video(Rational duration, mxl_frame* frame, bool is_blank);
    int run(uint64_t t0, const IoSet& io, uint64_t* bytes);
};

// ---- OutputDevice: src/module/output_device.rs ---------------------------------------------------------
struct OutputDevice : mxl_module {
    static constexpr uint64_t kRingCapacity = 65536;               // RingBuffer::<f32>::new(65536) (output_device.rs:128)
    mxl_output_device_params p{-1, -1, 0, 0};
    DevBuf scratch;                                                // self.scratch, on the device
    uint64_t scratch_len = 0;
    DevBuf clip_flag;
    int32_t* clip_host = nullptr;                                  // pinned
    // samples queued for the device callback, one pinned chunk per call
    struct Chunk { float* host = nullptr; size_t cap = 0, n = 0; cudaEvent_t ev = nullptr; bool in_flight = false; };
    std::deque<Chunk> chunks;
    std::vector<Chunk> spare;
    size_t head = 0;
    uint64_t pending = 0;

    OutputDevice(const mxl_output_device_params* in)
    {
        kind = MXL_MOD_OUTPUT_DEVICE;
        inputs = {unlabeled(MXL_LINE_STEREO)};                     // output_device.rs:84
        if (in) { p = *in; filter(); }
    }
    ~OutputDevice() override
    {
        if (ctx && ctx->has_device()) { ctx->activate(); cudaStreamSynchronize(ctx->stream); }
        for (auto& c : chunks) { if (c.host) cudaFreeHost(c.host); if (c.ev) cudaEventDestroy(c.ev); }
        for (auto& c : spare) { if (c.host) cudaFreeHost(c.host); if (c.ev) cudaEventDestroy(c.ev); }
        if (clip_host) cudaFreeHost(clip_host);
        scratch.release(ctx);
        clip_flag.release(ctx);
    }
    void filter()                                                  // 161-167: assignments must be within range
    {
        if (p.left >= 0 && (uint32_t)p.left >= p.channels) p.left = -1;
        if (p.right >= 0 && (uint32_t)p.right >= p.channels) p.right = -1;
    }
    int update(const void* params) override
    {
        if (!params) MXL_FAIL(MXL_ERR_PARAMS, "OutputDevice: NULL params");
        const mxl_output_device_params n = *(const mxl_output_device_params*)params;
        const bool reopened = n.channels != p.channels;            // another device stream (100-149): a new, empty ring
        if (n.channels) {                                          // `if let Some(stream)` (152)
            // 156-160: zero the scratch buffer when an assignment changes, so that no left-over data keeps playing
            if ((p.left != n.left || p.right != n.right || reopened) && scratch.p && ctx && ctx->has_device()) {
                MXL_TRY(ctx->activate());
                MXL_CUDA(cudaMemsetAsync(scratch.p, 0, scratch_len * sizeof(float), ctx->stream));
            }
            p.left = n.left; p.right = n.right;
        }
        p.channels = n.channels;
        filter();
        if (reopened) {
            if (ctx && ctx->has_device() && !chunks.empty()) { ctx->activate(); cudaStreamSynchronize(ctx->stream); }
            for (auto& c : chunks) { c.in_flight = false; spare.push_back(c); }
            chunks.clear(); head = 0; pending = 0;
        }
        return MXL_OK;
    }
    int get_params(void* out) const override { if (out) *(mxl_output_device_params*)out = p; return MXL_OK; }
    int run(const IoSet& io, uint64_t* bytes);
};

// ---- VideoMixer: src/module/video_mixer.rs ---------------------------------------------------------
struct PictureSettings { uint32_t w = 0, h = 0; bool operator==(const PictureSettings& o) const { return w == o.w && h == o.h; } bool operator!=(const PictureSettings& o) const { return !(*this == o); } };

struct VideoMixer : mxl_module {
    mxl_video_mixer_params p{-1, -1, 1.0};            // protocol lib.rs:412-420
    struct Channel {
        bool has_stored = false;
        Rational active_until;
        mxl_frame* frame = nullptr;                   // retained; scaled to the scaler's output
        bool has_scaler = false;
        PictureSettings scaler_out;
    } ch[MXL_VIDEO_MIXER_CHANNELS];
    DevBuf jobs;
    VideoMixer(const mxl_video_mixer_params* in)
    {
        kind = MXL_MOD_VIDEO_MIXER;
        if (in) p = *in;
        for (int i = 0; i < MXL_VIDEO_MIXER_CHANNELS; i++) inputs.push_back(labeled(MXL_LINE_VIDEO, std::to_string(i + 1)));   // video_mixer.rs:29-31
        outputs = {labeled(MXL_LINE_VIDEO, "Output"), labeled(MXL_LINE_VIDEO, "A"), labeled(MXL_LINE_VIDEO, "B")};            // video_mixer.rs:32-36
    }
    ~VideoMixer() override
    {
        for (auto& c : ch) frame_release(c.frame);
        jobs.release(ctx);
    }
    int update(const void* np) override { p = *(const mxl_video_mixer_params*)np; return MXL_OK; }
    int get_params(void* out) const override { *(mxl_video_mixer_params*)out = p; return MXL_OK; }
    void clear_stored(Channel& c) { frame_release(c.frame); c.frame = nullptr; c.has_stored = false; }
    int rescale(Channel& c, const PictureSettings& target);
    int run(uint64_t t0, const IoSet& io, uint64_t* bytes);
};

}  // namespace mxl

mxl_module::~mxl_module()
{
    for (mxl_line* l : host_in) mxl::line_free(l);
    for (mxl_line* l : host_out) mxl::line_free(l);
}

const char* mxl_module::kind_name() const
{
    switch (kind) {
    case MXL_MOD_AMPLIFIER: return "Amplifier";
    case MXL_MOD_ENVELOPE: return "Envelope";
    case MXL_MOD_EQ_THREE: return "EqThree";
    case MXL_MOD_FM_SINE: return "FmSine";
    case MXL_MOD_MIXER: return "Mixer";
    case MXL_MOD_OSCILLATOR: return "Oscillator";
    case MXL_MOD_PLOTTER: return "Plotter";
    case MXL_MOD_STEREO_PANNER: return "StereoPanner";
    case MXL_MOD_STEREO_SPLITTER: return "StereoSplitter";
    case MXL_MOD_TRIGGER: return "Trigger";
    case MXL_MOD_VIDEO_MIXER: return "VideoMixer";
    case MXL_MOD_METER: return "Meter";
    case MXL_MOD_SOURCE_MONO: return "SourceMono";
    case MXL_MOD_SOURCE_STEREO: return "SourceStereo";
    case MXL_MOD_SOURCE_VIDEO: return "SourceVideo";
    case MXL_MOD_STREAM_INPUT: return "StreamInput";
    case MXL_MOD_OUTPUT_DEVICE: return "OutputDevice";
    case MXL_MOD_MONITOR: return "Monitor";
    case MXL_MOD_STREAM_OUTPUT: return "StreamOutput";
    case MXL_MOD_PCM_SINK: return "PcmSink";
    default: return "?";
    }
}

namespace mxl {

mxl_module* module_create(mxl_ctx* ctx, int kind, const void* params)
{
    if (!ctx) { set_error("mxl_module_create: NULL context"); return nullptr; }
    mxl_module* m = nullptr;
    switch (kind) {
    case MXL_MOD_AMPLIFIER: m = new Amplifier((const mxl_amplifier_params*)params); break;
    case MXL_MOD_ENVELOPE: m = new Envelope((const mxl_envelope_params*)params); break;
    case MXL_MOD_EQ_THREE: m = new EqThree((const mxl_eq_three_params*)params); break;
    case MXL_MOD_FM_SINE: m = new FmSine((const mxl_fm_sine_params*)params); break;
    case MXL_MOD_MIXER: m = new Mixer((const mxl_mixer_params*)params); break;
    case MXL_MOD_OSCILLATOR: m = new Oscillator((const mxl_oscillator_params*)params); break;
    case MXL_MOD_PLOTTER: m = new Plotter(); break;
    case MXL_MOD_STEREO_PANNER: m = new StereoPanner(); break;
    case MXL_MOD_STEREO_SPLITTER: m = new StereoSplitter(); break;
    case MXL_MOD_TRIGGER: m = new Trigger((const mxl_trigger_params*)params); break;
    case MXL_MOD_VIDEO_MIXER: m = new VideoMixer((const mxl_video_mixer_params*)params); break;
    case MXL_MOD_METER: m = new Meter(); break;
    case MXL_MOD_SOURCE_MONO: m = new Source(kind, MXL_LINE_MONO); break;
    case MXL_MOD_SOURCE_STEREO: m = new Source(kind, MXL_LINE_STEREO); break;
    case MXL_MOD_SOURCE_VIDEO: m = new Source(kind, MXL_LINE_VIDEO); break;
    case MXL_MOD_PCM_SINK: m = new PcmSink(); break;
    case MXL_MOD_STREAM_INPUT: m = new StreamInput(); break;
    case MXL_MOD_MONITOR: case MXL_MOD_STREAM_OUTPUT: m = new Monitor(kind, (const mxl_monitor_params*)params); break;
    case MXL_MOD_OUTPUT_DEVICE: m = new OutputDevice((const mxl_output_device_params*)params); break;
    case MXL_MOD_MEDIA_SOURCE:
        set_error("module kind %d is an I/O edge that stays in the host application (out of scope of the tick hot path)", kind);
        return nullptr;
    default:
        set_error("unknown module kind %d", kind);
        return nullptr;
    }
    m->ctx = ctx;
    return m;
}

// ================================================================================================
// batched dispatch
// ================================================================================================

#define NEED_IO(io, nin, nout, name)                                                                     \
    do {                                                                                                 \
        if ((io).n_in != (nin) || (io).n_out != (nout))                                                  \
            MXL_FAIL(MXL_ERR_INVALID, "%s: expected %u inputs / %u outputs, got %u / %u", name, (unsigned)(nin), (unsigned)(nout), (io).n_in, (io).n_out); \
    } while (0)

static int need_len(const mxl_line* l, uint64_t floats, const char* what)
{
    if (l && l->len() < floats)
        MXL_FAIL(MXL_ERR_LENGTH, "%s: line holds %llu samples, the call needs %llu", what, (unsigned long long)l->len(), (unsigned long long)floats);
    return MXL_OK;
}

// --- Oscillator ---
static int run_oscillators(mxl_ctx* ctx, mxl_module* const* mods, int n, uint64_t t, const IoSet* io, uint64_t* bytes)
{
    int i = 0;
    while (i < n) {
        k::OscBatch b{};
        b.t0 = t; b.sample_rate = (double)ctx->sample_rate; b.inv_sample_rate = 1.0 / b.sample_rate;
        uint64_t frames = 0;
        int cnt = 0;
        for (; i < n && cnt < k::kMaxBatch; i++) {
            NEED_IO(io[i], 0, 2, "Oscillator");
            MXL_TRY(expect_output(io[i].out[0], MXL_LINE_MONO, "Oscillator.Mono"));
            MXL_TRY(expect_output(io[i].out[1], MXL_LINE_STEREO, "Oscillator.Stereo"));
            const uint64_t f = io[i].out[0]->frames;                      // let len = mono.len()  (oscillator.rs:71)
            MXL_TRY(need_len(io[i].out[1], 2 * f, "Oscillator.Stereo"));
            if (cnt && f != frames) break;
            frames = f;
            const Oscillator* o = (const Oscillator*)mods[i];
            b.inst[cnt++] = k::OscInst{io[i].out[0]->dev, io[i].out[1]->dev, o->p.freq, o->p.waveform, 0};
            if (bytes) *bytes += 12 * f;
        }
        b.frames = frames; b.n = cnt;
        MXL_TRY(k::launch_oscillator(ctx, b));
    }
    return MXL_OK;
}

// --- FmSine ---
static int run_fm_sines(mxl_ctx* ctx, mxl_module* const* mods, int n, uint64_t t, const IoSet* io, uint64_t* bytes)
{
    int i = 0;
    while (i < n) {
        k::FmBatch b{};
        b.t0 = t; b.sample_rate = (double)ctx->sample_rate; b.inv_sample_rate = 1.0 / b.sample_rate;
        uint64_t frames = 0;
        int cnt = 0;
        for (; i < n && cnt < k::kMaxBatch; i++) {
            NEED_IO(io[i], 1, 1, "FmSine");
            MXL_TRY(expect_input(io[i].in[0], MXL_LINE_MONO, "FmSine"));
            MXL_TRY(expect_output(io[i].out[0], MXL_LINE_STEREO, "FmSine"));
            const uint64_t f = io[i].out[0]->frames;                      // len = output.len() / CHANNELS (fm_sine.rs:40)
            MXL_TRY(need_len(io[i].in[0], f, "FmSine input"));
            if (cnt && f != frames) break;
            frames = f;
            const FmSine* m = (const FmSine*)mods[i];
            const double freq_amp = (m->p.freq_hi - m->p.freq_lo) / 2.0;   // fm_sine.rs:42-43
            const double freq_mid = m->p.freq_lo + freq_amp;
            b.inst[cnt++] = k::FmInst{io[i].in[0] ? io[i].in[0]->dev : nullptr, io[i].out[0]->dev, freq_mid, freq_amp};
            if (bytes) *bytes += (io[i].in[0] ? 4 * f : 0) + 8 * f;
        }
        b.frames = frames; b.n = cnt;
        MXL_TRY(k::launch_fm_sine(ctx, b));
    }
    return MXL_OK;
}

// --- Mixer ---
static int run_mixers(mxl_ctx* ctx, mxl_module* const* mods, int n, const IoSet* io, uint64_t* bytes)
{
    for (int i = 0; i < n; i++) {
        const Mixer* m = (const Mixer*)mods[i];
        const uint32_t C = (uint32_t)m->channels.size();
        NEED_IO(io[i], C, 2, "Mixer");
        MXL_TRY(expect_output(io[i].out[0], MXL_LINE_STEREO, "Mixer.Master"));
        MXL_TRY(expect_output(io[i].out[1], MXL_LINE_STEREO, "Mixer.Cue"));
        const uint64_t len = io[i].out[0]->len();                         // let len = master.len()  (mixer.rs:52)
        MXL_TRY(need_len(io[i].out[1], len, "Mixer.Cue"));
        for (uint32_t c = 0; c < C; c++) {
            MXL_TRY(expect_input(io[i].in[c], MXL_LINE_STEREO, "Mixer input"));
            MXL_TRY(need_len(io[i].in[c], len, "Mixer input"));
        }
        uint32_t done = 0;
        do {
            k::MixerLaunch p{};
            p.master = io[i].out[0]->dev; p.cue = io[i].out[1]->dev; p.len = len;
            p.accumulate = done > 0;
            const uint32_t take = std::min<uint32_t>(C - done, k::kMixerMaxCh);
            for (uint32_t c = 0; c < take; c++) {
                const mxl_mixer_channel_params& cp = m->channels[done + c];
                p.ch[c].in = io[i].in[done + c] ? io[i].in[done + c]->dev : nullptr;
                p.ch[c].gain = m->channel_gain[done + c];                 // mixer.rs:59
                p.ch[c].cue = cp.cue ? 1 : 0;
                if (bytes && io[i].in[done + c]) *bytes += 4 * len;
            }
            p.channels = (int32_t)take;
            MXL_TRY(k::launch_mixer(ctx, p));
            done += take;
        } while (done < C);
        if (bytes) *bytes += 8 * len;
    }
    return MXL_OK;
}

// --- Amplifier ---
static int run_amplifiers(mxl_ctx* ctx, mxl_module* const* mods, int n, const IoSet* io, uint64_t* bytes)
{
    int i = 0;
    while (i < n) {
        k::AmpBatch b{};
        uint64_t frames = 0;
        int cnt = 0;
        for (; i < n && cnt < k::kMaxBatch; i++) {
            NEED_IO(io[i], 2, 1, "Amplifier");
            MXL_TRY(expect_input(io[i].in[0], MXL_LINE_STEREO, "Amplifier.Input"));
            MXL_TRY(expect_input(io[i].in[1], MXL_LINE_MONO, "Amplifier.Control"));
            MXL_TRY(expect_output(io[i].out[0], MXL_LINE_STEREO, "Amplifier"));
            // let len = input.len() (amplifier.rs:50); a disconnected input is the engine's zero buffer,
            // which has the length of the tick's lines
            const uint64_t f = io[i].in[0] ? io[i].in[0]->frames : io[i].out[0]->frames;
            MXL_TRY(need_len(io[i].out[0], 2 * f, "Amplifier output"));
            MXL_TRY(need_len(io[i].in[1], f, "Amplifier.Control"));
            if (cnt && f != frames) break;
            frames = f;
            const Amplifier* m = (const Amplifier*)mods[i];
            b.inst[cnt++] = k::AmpInst{io[i].in[0] ? io[i].in[0]->dev : nullptr, io[i].in[1] ? io[i].in[1]->dev : nullptr,
                                       io[i].out[0]->dev, m->p.amplitude, m->p.mod_depth};
            if (bytes) *bytes += (io[i].in[0] ? 8 * f : 0) + (io[i].in[1] ? 4 * f : 0) + 8 * f;
        }
        b.frames = frames; b.n = cnt;
        MXL_TRY(k::launch_amplifier(ctx, b));
    }
    return MXL_OK;
}

// --- StereoPanner / StereoSplitter / Trigger ---
static int run_panners(mxl_ctx* ctx, int n, const IoSet* io, uint64_t* bytes)
{
    int i = 0;
    while (i < n) {
        k::PanBatch b{};
        uint64_t frames = 0;
        int cnt = 0;
        for (; i < n && cnt < k::kMaxBatch; i++) {
            NEED_IO(io[i], 2, 1, "StereoPanner");
            MXL_TRY(expect_input(io[i].in[0], MXL_LINE_MONO, "StereoPanner.L"));
            MXL_TRY(expect_input(io[i].in[1], MXL_LINE_MONO, "StereoPanner.R"));
            MXL_TRY(expect_output(io[i].out[0], MXL_LINE_STEREO, "StereoPanner"));
            const uint64_t f = io[i].in[0] ? io[i].in[0]->frames : io[i].out[0]->frames;   // 0..left.len() (stereo_panner.rs:35)
            MXL_TRY(need_len(io[i].in[1], f, "StereoPanner.R"));
            MXL_TRY(need_len(io[i].out[0], 2 * f, "StereoPanner output"));
            if (cnt && f != frames) break;
            frames = f;
            b.inst[cnt++] = k::PanInst{io[i].in[0] ? io[i].in[0]->dev : nullptr, io[i].in[1] ? io[i].in[1]->dev : nullptr, io[i].out[0]->dev};
            if (bytes) *bytes += (io[i].in[0] ? 4 * f : 0) + (io[i].in[1] ? 4 * f : 0) + 8 * f;
        }
        b.frames = frames; b.n = cnt;
        MXL_TRY(k::launch_panner(ctx, b));
    }
    return MXL_OK;
}

static int run_splitters(mxl_ctx* ctx, int n, const IoSet* io, uint64_t* bytes)
{
    int i = 0;
    while (i < n) {
        k::SplitBatch b{};
        uint64_t frames = 0;
        int cnt = 0;
        for (; i < n && cnt < k::kMaxBatch; i++) {
            NEED_IO(io[i], 1, 2, "StereoSplitter");
            MXL_TRY(expect_input(io[i].in[0], MXL_LINE_STEREO, "StereoSplitter"));
            MXL_TRY(expect_output(io[i].out[0], MXL_LINE_MONO, "StereoSplitter.L"));
            MXL_TRY(expect_output(io[i].out[1], MXL_LINE_MONO, "StereoSplitter.R"));
            const uint64_t f = io[i].out[0]->frames;                      // 0..left.len() (stereo_splitter.rs:41)
            MXL_TRY(need_len(io[i].out[1], f, "StereoSplitter.R"));
            MXL_TRY(need_len(io[i].in[0], 2 * f, "StereoSplitter input"));
            if (cnt && f != frames) break;
            frames = f;
            b.inst[cnt++] = k::SplitInst{io[i].in[0] ? io[i].in[0]->dev : nullptr, io[i].out[0]->dev, io[i].out[1]->dev};
            if (bytes) *bytes += (io[i].in[0] ? 8 * f : 0) + 8 * f;
        }
        b.frames = frames; b.n = cnt;
        MXL_TRY(k::launch_splitter(ctx, b));
    }
    return MXL_OK;
}

static int run_triggers(mxl_ctx* ctx, mxl_module* const* mods, int n, const IoSet* io, uint64_t* bytes)
{
    int i = 0;
    while (i < n) {
        k::FillBatch b{};
        uint64_t len = 0;
        int cnt = 0;
        for (; i < n && cnt < k::kMaxBatch; i++) {
            NEED_IO(io[i], 0, 1, "Trigger");
            MXL_TRY(expect_output(io[i].out[0], MXL_LINE_MONO, "Trigger"));
            const uint64_t f = io[i].out[0]->frames;
            if (cnt && f != len) break;
            len = f;
            const Trigger* m = (const Trigger*)mods[i];
            b.inst[cnt++] = k::FillInst{io[i].out[0]->dev, m->p.gate == MXL_GATE_OPEN ? 1.0f : 0.0f, 0};   // trigger.rs:38-41
            if (bytes) *bytes += 4 * f;
        }
        b.len = len; b.n = cnt;
        MXL_TRY(k::launch_fill(ctx, b));
    }
    return MXL_OK;
}

// --- EqThree ---
// Homogeneous Lc-sample transition A = M^Lc of one 4-pole cascade and its powers A^0..A^J, packed
// lower-triangular (10 doubles each).  Plan layout: [J, pow_lo[(J+1)*10], pow_hi[(J+1)*10]].
static void cascade_power(double c, uint32_t steps, double A[4][4])
{
    for (int col = 0; col < 4; col++) {
        double p[4] = {0, 0, 0, 0};
        p[col] = 1.0;
        for (uint32_t s = 0; s < steps; s++) {
            p[0] += c * (0.0 - p[0]);
            p[1] += c * (p[0] - p[1]);
            p[2] += c * (p[1] - p[2]);
            p[3] += c * (p[2] - p[3]);
        }
        for (int r = 0; r < 4; r++) A[r][col] = p[r];
    }
}

static void matmul4(const double X[4][4], const double Y[4][4], double Z[4][4])
{
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            double s = 0;
            for (int q = 0; q < 4; q++) s += X[r][q] * Y[q][c];
            Z[r][c] = s;
        }
}

static double absmax4(const double X[4][4])
{
    double m = 0;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) m = std::max(m, fabs(X[r][c]));
    return m;
}

static void pack_tri(const double X[4][4], double* out)
{
    int q = 0;
    for (int r = 0; r < 4; r++) for (int c = 0; c <= r; c++) out[q++] = X[r][c];
}

struct EqCoefs { double c_lo, c_hi; };
static EqCoefs eq_coefs(const mxl_ctx* ctx)
{
    // LowPass::set_freq (eq_three.rs:117-119), FREQ_LO = 420, FREQ_HI = 2700 (eq_three.rs:8-9)
    const double pi = 3.14159265358979323846264338327950288;
    return EqCoefs{2.0 * sin(pi * 420.0 / (double)ctx->sample_rate), 2.0 * sin(pi * 2700.0 / (double)ctx->sample_rate)};
}

// Returns the plan for chunk length `chunk`; if the cascade does not forget within kEqMaxCarry
// chunks of that length (|A^J| >= 2^-75), *chunk is doubled until it does.
static const std::vector<double>& eq_plan_for(mxl_ctx* ctx, uint32_t* chunk)
{
    const EqCoefs co = eq_coefs(ctx);
    for (;;) {
        auto it = ctx->eq_plans.find(*chunk);
        if (it != ctx->eq_plans.end()) {
            if (!it->second.empty()) return it->second;
            *chunk *= 2;                      // remembered as "too short"
            continue;
        }
        double Al[4][4], Ah[4][4];
        cascade_power(co.c_lo, *chunk, Al);
        cascade_power(co.c_hi, *chunk, Ah);
        double Pl[k::kEqMaxCarry + 1][4][4], Ph[k::kEqMaxCarry + 1][4][4];
        memset(Pl, 0, sizeof Pl); memset(Ph, 0, sizeof Ph);
        for (int d = 0; d < 4; d++) { Pl[0][d][d] = 1.0; Ph[0][d][d] = 1.0; }
        int J = -1;
        const double tiny = ldexp(1.0, -75);
        for (int j = 1; j <= k::kEqMaxCarry; j++) {
            matmul4(Pl[j - 1], Al, Pl[j]);
            matmul4(Ph[j - 1], Ah, Ph[j]);
            if (J < 0 && absmax4(Pl[j]) < tiny && absmax4(Ph[j]) < tiny) J = j;
        }
        if (J < 0 && *chunk < (1u << 24)) {
            ctx->eq_plans[*chunk] = std::vector<double>();
            *chunk *= 2;
            continue;
        }
        if (J < 0) J = k::kEqMaxCarry;
        std::vector<double> plan(1 + 2 * (k::kEqMaxCarry + 1) * 10, 0.0);
        plan[0] = (double)J;
        for (int j = 0; j <= k::kEqMaxCarry; j++) {
            pack_tri(Pl[j], &plan[1 + j * 10]);
            pack_tri(Ph[j], &plan[1 + (k::kEqMaxCarry + 1) * 10 + j * 10]);
        }
        return ctx->eq_plans[*chunk] = plan;
    }
}

static uint32_t eq_pick_chunk(const mxl_ctx* ctx, uint64_t frames, int n_inst)
{
    if (const char* e = getenv("MXL_EQ_CHUNK")) {
        long v = atol(e);
        if (v >= 4) return (uint32_t)(v / 4 * 4);
    }
    const uint64_t total = frames * (uint64_t)n_inst;
    const uint64_t target_threads = (uint64_t)(ctx->sm_count > 0 ? ctx->sm_count : 148) * 768;
    uint64_t want = (total + target_threads - 1) / target_threads;
    uint32_t chunk = 256;
    while (chunk < want && chunk < 8192) chunk *= 2;
    return chunk;
}

// Kernel-parameter scalars of a plan (eq_plan.h); its tables live in device memory (9 KB per plan, uploaded once).
static int fill_eq_consts(mxl_ctx* ctx, const mxl::EqStreamPlan& plan, k::EqStreamConsts* q)
{
    q->chunk = plan.lc;
    q->halo = plan.halo;
    q->lev_lo = plan.lev_lo;
    q->lev_hi = plan.lev_hi;
    q->back_lo = plan.back_lo;
    q->back_hi = plan.back_hi;
    void*& tab = ctx->eq_stream_tables[plan.lc];
    if (!tab) {
        static_assert(sizeof(plan.V) == sizeof(k::EqDevTables::V) && sizeof(plan.pow_lo) == sizeof(k::EqDevTables::pow_lo) &&
                      sizeof(plan.lane_pow) == sizeof(k::EqDevTables::lane_pow), "plan layout");
        k::EqDevTables* host = new k::EqDevTables();
        memcpy(host->V, plan.V, sizeof host->V);
        memcpy(host->K, plan.K, sizeof host->K);
        memcpy(host->pow_lo, plan.pow_lo, sizeof host->pow_lo);
        memcpy(host->pow_hi, plan.pow_hi, sizeof host->pow_hi);
        memcpy(host->lane_pow, plan.lane_pow, sizeof host->lane_pow);
        cudaError_t e = cudaMalloc(&tab, sizeof(k::EqDevTables));
        if (e == cudaSuccess) e = cudaMemcpy(tab, host, sizeof(k::EqDevTables), cudaMemcpyHostToDevice);   // synchronous: `host` dies here
        delete host;
        if (e != cudaSuccess) { tab = nullptr; MXL_FAIL(MXL_ERR_CUDA, "EqThree tables: %s", cudaGetErrorString(e)); }
    }
    q->tab = (const k::EqDevTables*)tab;
    q->host_V = plan.V; q->host_K = plan.K;
    q->c_lo = plan.c_lo; q->c_hi = plan.c_hi;
    return MXL_OK;
}

static const mxl::EqStreamPlan* eq_stream_plan_for(mxl_ctx* ctx, uint32_t lc)
{
    auto it = ctx->eq_stream_plans.find(lc);
    if (it == ctx->eq_stream_plans.end())
        it = ctx->eq_stream_plans.emplace(lc, mxl::eq_stream_plan(ctx->sample_rate, lc, k::kEqStreamThreads / 2)).first;
    return it->second.ok ? &it->second : nullptr;
}

// Single-launch path (eq_stream_kernel).  Returns 1 if it ran, 0 if no usable plan exists at this
// sample rate (the cascades would need more than 128 chunks to forget), < 0 on error.
static int run_eq_stream(mxl_ctx* ctx, mxl_module* const* mods, const IoSet* io, int first, int cnt, uint64_t frames, uint64_t* bytes)
{
    // Chunk length: 64 samples per thread once that still fills every SM with a CTA (per-sample scan and
    // halo overhead halve), else 32 (twice the threads for short calls); 16 only if forced.  A length whose
    // cascade would not forget inside half a CTA at this sample rate is skipped.  MXL_EQ_STREAM_CHUNK forces one.
    uint32_t forced = 0;
    if (const char* e = getenv("MXL_EQ_STREAM_CHUNK")) forced = (uint32_t)atol(e);
    auto plan_for = [&](uint32_t lc) { return eq_stream_plan_for(ctx, lc); };
    const mxl::EqStreamPlan* plan = nullptr;
    if (forced) {
        plan = plan_for(forced);
    } else {
        const mxl::EqStreamPlan* p64 = plan_for(64);
        const mxl::EqStreamPlan* p32 = plan_for(32);
        if (p64) {
            const uint64_t ctas64 = (uint64_t)cnt * (((frames + 63) / 64 + (k::kEqStreamThreads - p64->halo) - 1) / (k::kEqStreamThreads - p64->halo));
            if (!p32 || ctas64 >= (uint64_t)(ctx->sm_count > 0 ? ctx->sm_count : 148)) plan = p64;
        }
        if (!plan) plan = p32 ? p32 : plan_for(16);
    }
    if (!plan) return 0;
    k::EqStreamBatch b{};
    b.frames = frames;
    b.n_chunks = (uint32_t)((frames + plan->lc - 1) / plan->lc);
    b.n = cnt;
    MXL_TRY(fill_eq_consts(ctx, *plan, &b.eq));
    for (int j = 0; j < cnt; j++) {
        EqThree* m = (EqThree*)mods[first + j];
        MXL_TRY(m->ensure_state());
        k::EqStreamInst& e = b.inst[j];
        e.in = io[first + j].in[0] ? io[first + j].in[0]->dev : nullptr;
        e.out = io[first + j].out[0]->dev;
        e.state = m->state_ptr(m->cur);
        e.state_out = m->state_ptr(m->cur ^ 1);
        e.poison = m->poison_ptr();
        e.g_lo = db_to_linear(m->p.gain_lo_db);            // eq_three.rs:62-64
        e.g_mid = db_to_linear(m->p.gain_mid_db);
        e.g_hi = db_to_linear(m->p.gain_hi_db);
        if (bytes) *bytes += (io[first + j].in[0] ? 4 * frames : 0) + 4 * frames;
    }
    MXL_TRY(k::launch_eq_stream(ctx, b));
    for (int j = 0; j < cnt; j++) ((EqThree*)mods[first + j])->cur ^= 1;
    return 1;
}

static int run_eq_threes(mxl_ctx* ctx, mxl_module* const* mods, int n, const IoSet* io, uint64_t* bytes)
{
    int i = 0;
    while (i < n) {
        // gather a batch of equal length
        int first = i, cnt = 0;
        uint64_t frames = 0;
        for (; i < n && cnt < k::kMaxBatch; i++, cnt++) {
            NEED_IO(io[i], 1, 1, "EqThree");
            MXL_TRY(expect_input(io[i].in[0], MXL_LINE_MONO, "EqThree"));
            MXL_TRY(expect_output(io[i].out[0], MXL_LINE_MONO, "EqThree"));
            // input.iter().zip(output.iter_mut()) (eq_three.rs:66): the shorter of the two
            uint64_t f = io[i].out[0]->frames;
            if (io[i].in[0] && io[i].in[0]->frames < f) f = io[i].in[0]->frames;
            if (cnt && f != frames) break;
            frames = f;
        }
        if (frames == 0) continue;
        if (frames >= (1ull << 36)) MXL_FAIL(MXL_ERR_LENGTH, "EqThree: call too long");
        if (!getenv("MXL_EQ_CHUNK")) {                     // MXL_EQ_CHUNK selects the two-launch path
            const int ran = run_eq_stream(ctx, mods, io, first, cnt, frames, bytes);
            if (ran < 0) return ran;
            if (ran == 1) continue;
        }
        uint32_t chunk = eq_pick_chunk(ctx, frames, cnt);
        const std::vector<double>& plan = eq_plan_for(ctx, &chunk);
        k::EqBatch b{};
        b.frames = frames;
        b.chunk = chunk;
        b.n_chunks = (uint32_t)((frames + chunk - 1) / chunk);
        b.carry_terms = (uint32_t)plan[0];
        b.n = cnt;
        const EqCoefs co = eq_coefs(ctx);
        b.c_lo = co.c_lo; b.c_hi = co.c_hi;
        memcpy(b.pow_lo, &plan[1], sizeof b.pow_lo);
        memcpy(b.pow_hi, &plan[1 + (k::kEqMaxCarry + 1) * 10], sizeof b.pow_hi);
        for (int j = 0; j < cnt; j++) {
            EqThree* m = (EqThree*)mods[first + j];
            MXL_TRY(m->ensure_state());
            MXL_TRY(m->zend.ensure(ctx, (size_t)b.n_chunks * 8 * sizeof(double)));
            k::EqInst& e = b.inst[j];
            e.in = io[first + j].in[0] ? io[first + j].in[0]->dev : nullptr;
            e.out = io[first + j].out[0]->dev;
            e.state = m->state_ptr(m->cur);
            e.state_out = m->state_ptr(m->cur ^ 1);
            e.zend = (double*)m->zend.p;
            e.poison = m->poison_ptr();
            e.g_lo = db_to_linear(m->p.gain_lo_db);        // eq_three.rs:62-64
            e.g_mid = db_to_linear(m->p.gain_mid_db);
            e.g_hi = db_to_linear(m->p.gain_hi_db);
            if (bytes) *bytes += (io[first + j].in[0] ? 4 * frames : 0) + 4 * frames;
        }
        MXL_TRY(k::launch_eq_three(ctx, b));
        for (int j = 0; j < cnt; j++) ((EqThree*)mods[first + j])->cur ^= 1;
    }
    return MXL_OK;
}

// --- Envelope ---
static int run_envelopes(mxl_ctx* ctx, mxl_module* const* mods, int n, uint64_t t, const IoSet* io, uint64_t* bytes)
{
    int i = 0;
    while (i < n) {
        k::EnvBatch b{};
        uint64_t frames = 0;
        int cnt = 0, first = i;
        for (; i < n && cnt < k::kMaxBatch; i++) {
            NEED_IO(io[i], 1, 1, "Envelope");
            MXL_TRY(expect_input(io[i].in[0], MXL_LINE_MONO, "Envelope"));
            MXL_TRY(expect_output(io[i].out[0], MXL_LINE_MONO, "Envelope"));
            const uint64_t f = io[i].in[0] ? io[i].in[0]->frames : io[i].out[0]->frames;   // let len = input.len() (envelope.rs:95)
            MXL_TRY(need_len(io[i].out[0], f, "Envelope output"));
            if (f >= 0x7ffffff0ull) MXL_FAIL(MXL_ERR_LENGTH, "Envelope: call longer than 2^31 samples");
            if (cnt && f != frames) break;
            frames = f;
            cnt++;
        }
        if (frames == 0) continue;
        const uint32_t nt = k::envelope_tiles(frames);
        for (int j = 0; j < cnt; j++) {
            Envelope* m = (Envelope*)mods[first + j];
            MXL_TRY(m->ensure_state());
            if (m->tiles_cap < nt) {                       // look-back descriptors; zeroed once: epoch 0 is never used
                const size_t cap = (size_t)nt + nt / 2 + 16;
                MXL_TRY(m->scratch.ensure(ctx, cap * sizeof(k::EnvTile)));
                MXL_CUDA(cudaMemsetAsync(m->scratch.p, 0, cap * sizeof(k::EnvTile), ctx->stream));
                m->tiles_cap = cap;
            }
            k::EnvInst& e = b.inst[j];
            e.in = io[first + j].in[0] ? io[first + j].in[0]->dev : nullptr;
            e.out = io[first + j].out[0]->dev;
            e.state = (const k::EnvState*)m->state.p + m->cur;
            e.state_out = (k::EnvState*)m->state.p + (m->cur ^ 1);
            e.tiles = (k::EnvTile*)m->scratch.p;
            e.attack_ms = m->p.attack_ms;
            e.inv_attack = 1.0 / m->p.attack_ms;           // envelope.rs:43,48,55: `1.0 / x_ms * ms`, left to right
            e.inv_decay = 1.0 / m->p.decay_ms;
            e.sustain = m->p.sustain_amplitude;
            e.inv_release = 1.0 / m->p.release_ms;
            if (bytes) *bytes += (io[first + j].in[0] ? 4 * frames : 0) + 4 * frames;
        }
        b.t0 = t; b.frames = frames; b.n = cnt;
        b.sample_rate = (double)ctx->sample_rate;
        b.inv_sample_rate = 1.0 / (double)ctx->sample_rate;
        b.epoch = ++ctx->env_epoch;
        if ((b.epoch << 2) == 0) b.epoch = ctx->env_epoch = 1;   // 30-bit epoch wrapped: flags of 2^30 launches ago are long gone
        MXL_TRY(k::launch_envelope(ctx, b));
        for (int j = 0; j < cnt; j++) ((Envelope*)mods[first + j])->cur ^= 1;
    }
    return MXL_OK;
}

// --- Meter ---
static int run_meters(mxl_ctx* ctx, mxl_module* const* mods, int n, const IoSet* io, uint64_t* bytes)
{
    int i = 0;
    while (i < n) {
        k::MeterBatch b{};
        uint64_t frames = 0;
        int cnt = 0;
        uint32_t slots = 0;
        for (; i < n && cnt < k::kMaxBatch; i++) {
            NEED_IO(io[i], 1, 0, "Meter");
            MXL_TRY(expect_input(io[i].in[0], MXL_LINE_STEREO, "Meter"));
            Meter* m = (Meter*)mods[i];
            const uint64_t f = io[i].in[0] ? io[i].in[0]->frames : ctx->spt;
            if (cnt && f != frames) break;
            frames = f;
            slots = (uint32_t)((f + ctx->spt - 1) / ctx->spt);
            MXL_TRY(m->records.ensure(ctx, (size_t)slots * sizeof(k::MeterRecord)));
            m->n_slots = slots;
            b.inst[cnt++] = k::MeterInst{io[i].in[0] ? io[i].in[0]->dev : nullptr, (k::MeterRecord*)m->records.p};
            if (bytes && io[i].in[0]) *bytes += 8 * f;
        }
        b.frames = frames; b.spt = ctx->spt; b.n = cnt;
        MXL_TRY(k::launch_meter(ctx, b, slots));
    }
    return MXL_OK;
}

// ================================================================================================
// Fused voice group (fused_voice.cu): Oscillator -> EqThree -> StereoPanner -> Mixer [-> Meter]
// ================================================================================================
bool fused_meter_supported(mxl_ctx* ctx)
{
    // the mix kernel gives a tick to one CTA: frame pairs must not straddle ticks, and the tick's master samples sit in
    // its shared memory for the meter warp
    return ctx->spt > 0 && (ctx->spt & 1u) == 0 && (size_t)(ctx->spt / 2 + 1) * 16 + (size_t)k::kFusedMaxChans * sizeof(k::FusedChan) <= 48 * 1024;
}

static int fused_max_products(const FusedGroup& g);

bool fused_group_params_ok(const FusedGroup& g)
{
    for (const FusedVoiceRef& v : g.voices) {
        if (!v.osc) continue;
        const Oscillator* o = (const Oscillator*)v.osc;
        // a finite frequency keeps every phase finite, so the oscillator never produces a NaN the cascades would
        // keep for ever (the staged EqThree kernel carries that case; the fused one does not)
        if (!(fabs(o->p.freq) <= 1e12)) return false;
    }
    if (!g.mixer || ((const Mixer*)g.mixer)->channels.size() != g.chans.size()) return false;
    if (fused_max_products(g) < 0) return false;            // a voice multiplied by too many distinct channel gains
    return true;
}

// largest number of distinct channel gains any voice of the group is multiplied by (= product tiles per CTA); -1 = too many
static int fused_max_products(const FusedGroup& g)
{
    const Mixer* mx = (const Mixer*)g.mixer;
    int most = 0;
    for (size_t v = 0; v < g.voices.size(); v++) {
        double seen[k::kFusedMaxProducts];
        int n = 0;
        for (size_t c = 0; c < g.chans.size(); c++)
            for (int side : {g.chans[c].left, g.chans[c].right}) {
                if (side != (int)v) continue;
                bool found = false;
                for (int i = 0; i < n; i++) found = found || memcmp(&seen[i], &mx->channel_gain[c], sizeof(double)) == 0;
                if (found) continue;
                if (n == k::kFusedMaxProducts) return -1;
                seen[n++] = mx->channel_gain[c];
            }
        most = std::max(most, n);
    }
    return most;
}

bool fused_group_supported(mxl_ctx* ctx, const FusedGroup& g)
{
    if (!ctx->has_device() || !g.mixer || g.voices.empty() || g.voices.size() > (size_t)k::kFusedMaxVoices ||
        g.chans.empty() || g.chans.size() > (size_t)k::kFusedMaxChans)
        return false;
    if (!fused_group_params_ok(g)) return false;
    return eq_stream_plan_for(ctx, 32) != nullptr;          // else: a sample rate the time-parallel scheme does not handle
}

static int fused_bytes(const FusedGroup& g, uint64_t frames, uint64_t* bytes);

int run_fused_group(mxl_ctx* ctx, const FusedGroup& g, uint64_t t, uint64_t* bytes)
{
    if (bytes) *bytes = 0;
    {   // steady state: same plan, same lines, same parameters, same call length as the last run
        Mixer* mxc = (Mixer*)g.mixer;
        Mixer::FusedCache* fc = mxc ? mxc->fused_cache : nullptr;
        if (fc && fc->epoch == ctx->change_epoch && fc->group == (const void*)&g && g.master && fc->frames == g.master->frames && fc->frames) {
            if (t + fc->frames + (1ull << 20) >= (1ull << 53)) MXL_FAIL(MXL_ERR_LENGTH, "fused voice group: sample index beyond 2^53");
            fc->vb.t0 = t;
            for (int i = 0; i < fc->vb.n_voices; i++) {
                EqThree* e = (EqThree*)g.voices[i].eq;
                fc->vb.voice[i].state = e->state_ptr(e->cur);
                fc->vb.voice[i].state_out = e->state_ptr(e->cur ^ 1);
            }
            if (bytes) fused_bytes(g, fc->frames, bytes);
            ctx->fused_group_now = (const void*)&g;
            fc->vb.late_wait = 1;                          // (the launcher checks that this group's mix kernel is what runs before it)
            MXL_TRY(k::launch_fused_voice(ctx, fc->vb));
            for (int i = 0; i < fc->vb.n_voices; i++) ((EqThree*)g.voices[i].eq)->cur ^= 1;
            return k::launch_fused_mix(ctx, fc->mb);
        }
    }
    MXL_TRY(expect_output(g.master, MXL_LINE_STEREO, "Mixer.Master"));
    MXL_TRY(expect_output(g.cue, MXL_LINE_STEREO, "Mixer.Cue"));
    const uint64_t frames = g.master->frames;
    if (frames == 0) return MXL_OK;
    MXL_TRY(need_len(g.cue, 2 * frames, "Mixer.Cue"));
    if (t + frames + (1ull << 20) >= (1ull << 53)) MXL_FAIL(MXL_ERR_LENGTH, "fused voice group: sample index beyond 2^53");
    Mixer* mx = (Mixer*)g.mixer;
    if (mx->channels.size() != g.chans.size()) MXL_FAIL(MXL_ERR_INVALID, "fused voice group: mixer has %zu channels, the plan %zu", mx->channels.size(), g.chans.size());
    if (fused_max_products(g) < 0) MXL_FAIL(MXL_ERR_INVALID, "fused voice group: a voice is multiplied by more than %d gains", k::kFusedMaxProducts);
    const int nv = (int)g.voices.size();

    // Chunk length of the voice kernel.  A thread walks its chunk serially, so the chunk sets the latency of a CTA, while
    // the halo (the ~1100 samples a tile recomputes ahead of its own) and the scan cost the same per tile whatever the
    // chunk: 32 samples by default; 16 while the call is so short that 32 would leave SMs without a CTA; 64 for long
    // calls (half the tiles, half the halo work) once that still fills the machine several CTAs deep.
    const uint64_t sms = (uint64_t)(ctx->sm_count > 0 ? ctx->sm_count : 148);
    auto tiles_for = [&](const mxl::EqStreamPlan* pl) -> uint64_t {
        const uint64_t own = k::kEqStreamThreads - pl->halo;
        return ((frames + pl->lc - 1) / pl->lc + own - 1) / own;
    };
    const mxl::EqStreamPlan* plan = eq_stream_plan_for(ctx, 32);
    if (!plan) MXL_FAIL(MXL_ERR_UNSUPPORTED, "fused voice group: no EqThree plan at %u Hz", ctx->sample_rate);
    const char* forced_env = getenv("MXL_FUSED_CHUNK");             // tests and tuning: 16, 32 or 64
    const uint32_t forced = forced_env ? (uint32_t)atol(forced_env) : 0u;
    const mxl::EqStreamPlan* p16 = eq_stream_plan_for(ctx, 16);
    const mxl::EqStreamPlan* p64 = eq_stream_plan_for(ctx, 64);
    if (p16 && (forced == 16 || (!forced && tiles_for(plan) * nv < 2 * sms))) plan = p16;
    else if (p64 && (forced == 64 || (!forced && tiles_for(p64) * nv >= 6 * sms))) plan = p64;

    k::FusedVoiceBatch vb{};
    MXL_TRY(fill_eq_consts(ctx, *plan, &vb.eq));
    vb.t0 = t; vb.frames = frames;
    vb.sample_rate = (double)ctx->sample_rate; vb.inv_sample_rate = 1.0 / vb.sample_rate;
    vb.n_chunks = (uint32_t)((frames + plan->lc - 1) / plan->lc);
    vb.owned = k::kEqStreamThreads - plan->halo;
    vb.n_voices = nv;
    for (int i = 0; i < nv; i++) {
        const FusedVoiceRef& v = g.voices[i];
        EqThree* e = (EqThree*)v.eq;
        MXL_TRY(e->ensure_state());
        MXL_TRY(expect_output(v.eq_out, MXL_LINE_MONO, "EqThree"));
        MXL_TRY(need_len(v.eq_out, frames, "EqThree output"));
        k::FusedVoice& fv = vb.voice[i];
        if (v.osc) {
            const Oscillator* o = (const Oscillator*)v.osc;
            fv.freq = o->p.freq; fv.waveform = o->p.waveform;
            if (bytes) *bytes += 12 * frames;
        } else {
            fv.freq = 0.0; fv.waveform = MXL_WAVE_OFF;
        }
        fv.state = e->state_ptr(e->cur);
        fv.state_out = e->state_ptr(e->cur ^ 1);
        fv.g_lo = db_to_linear(e->p.gain_lo_db);             // eq_three.rs:62-64
        fv.g_mid = db_to_linear(e->p.gain_mid_db);
        fv.g_hi = db_to_linear(e->p.gain_hi_db);
        fv.eq_out = v.eq_out->dev;
        if (v.osc_mono) { MXL_TRY(need_len(v.osc_mono, frames, "Oscillator.Mono")); fv.osc_mono = v.osc_mono->dev; }
        if (v.osc_stereo) { MXL_TRY(need_len(v.osc_stereo, 2 * frames, "Oscillator.Stereo")); fv.osc_stereo = v.osc_stereo->dev; }
        if (bytes) *bytes += (v.osc ? 4 * frames : 0) + 4 * frames;
    }
    // product lines: one per (voice, distinct channel gain), in the mixer's scratch buffer
    const uint64_t line_floats = (frames + 3) & ~3ull;                  // 16-byte aligned lines
    int n_lines = 0;
    auto product_of = [&](int voice, double gain) -> int {
        k::FusedVoice& fv = vb.voice[voice];
        for (int i = 0; i < fv.n_products; i++)
            if (memcmp(&fv.product_gain[i], &gain, sizeof(double)) == 0) return i;
        fv.product_gain[fv.n_products] = gain;
        fv.product_out[fv.n_products] = (float*)nullptr + (size_t)(n_lines++) * line_floats;   // offset for now, based below
        return fv.n_products++;
    };
    std::vector<int> side_product(2 * g.chans.size(), -1);
    for (size_t c = 0; c < g.chans.size(); c++) {
        const FusedChanRef& cr = g.chans[c];
        if (cr.left >= 0) side_product[2 * c] = product_of(cr.left, mx->channel_gain[c]);
        if (cr.right >= 0) side_product[2 * c + 1] = product_of(cr.right, mx->channel_gain[c]);
    }
    MXL_TRY(mx->fused_products.ensure(ctx, (size_t)std::max(n_lines, 1) * line_floats * sizeof(float)));
    float* scratch = (float*)mx->fused_products.p;
    for (int i = 0; i < nv; i++)
        for (int pi = 0; pi < vb.voice[i].n_products; pi++)
            vb.voice[i].product_out[pi] = scratch + (vb.voice[i].product_out[pi] - (float*)nullptr);

    k::FusedMixBatch mb{};
    mb.frames = frames;
    mb.n_channels = (int16_t)g.chans.size();
    mb.mono = 1;
    for (const FusedChanRef& cr : g.chans) if (cr.left != cr.right) mb.mono = 0;
    mb.master = g.master->dev; mb.cue = g.cue->dev;
    if (g.meter) {
        if (!fused_meter_supported(ctx)) MXL_FAIL(MXL_ERR_INVALID, "fused voice group: meter folded in but ticks of %u frames do not fit", ctx->spt);
        Meter* me = (Meter*)g.meter;
        const uint32_t slots = (uint32_t)((frames + ctx->spt - 1) / ctx->spt);
        MXL_TRY(me->records.ensure(ctx, (size_t)slots * sizeof(k::MeterRecord)));
        me->n_slots = slots;
        mb.meter = (k::MeterRecord*)me->records.p;
        mb.spt = ctx->spt;
        if (bytes) *bytes += 8 * frames;
    } else {
        mb.spt = 1024;                                       // frames per CTA: two frame pairs per thread
    }
    std::vector<k::FusedChan> chan(g.chans.size());
    for (size_t c = 0; c < g.chans.size(); c++) {
        const FusedChanRef& cr = g.chans[c];
        k::FusedChan& fc = chan[c];
        memset(&fc, 0, sizeof fc);
        if (cr.pan_out) { MXL_TRY(need_len(cr.pan_out, 2 * frames, "StereoPanner output")); fc.pan_out = cr.pan_out->dev; }
        fc.cue = mx->channels[c].cue ? 1 : 0;
        const bool need_raw = fc.cue || fc.pan_out;
        if (cr.left >= 0) { fc.left = vb.voice[cr.left].product_out[side_product[2 * c]]; if (need_raw) fc.left_raw = vb.voice[cr.left].eq_out; }
        if (cr.right >= 0) { fc.right = vb.voice[cr.right].product_out[side_product[2 * c + 1]]; if (need_raw) fc.right_raw = vb.voice[cr.right].eq_out; }
        fc.zero_product = (float)(0.0 * mx->channel_gain[c]);           // what a disconnected side adds (mixer.rs:62 on the zero buffer)
        if (bytes && cr.connected) *bytes += (cr.left >= 0 ? 4 * frames : 0) + (cr.right >= 0 ? 4 * frames : 0) + 8 * frames + 8 * frames;
    }
    if (bytes) *bytes += 16 * frames;
    // the channel table travels to the device only when it changed (line pointers and gains are stable call to call)
    if (mx->fused_chan_host.size() != chan.size() || memcmp(mx->fused_chan_host.data(), chan.data(), chan.size() * sizeof(k::FusedChan)) != 0) {
        MXL_TRY(mx->fused_chan.ensure(ctx, (size_t)k::kFusedMaxChans * sizeof(k::FusedChan)));
        // pageable source: staged by the runtime before the call returns; ordered on the stream before the launches below
        MXL_CUDA(cudaMemcpyAsync(mx->fused_chan.p, chan.data(), chan.size() * sizeof(k::FusedChan), cudaMemcpyHostToDevice, ctx->stream));
        mx->fused_chan_host = chan;
    }
    mb.chan = (const k::FusedChan*)mx->fused_chan.p;
    if (bytes) { *bytes = 0; fused_bytes(g, frames, bytes); }
    if (!mx->fused_cache) mx->fused_cache = new Mixer::FusedCache();
    // (the epoch is read after the preparations above: table uploads and scratch growth do not touch it, line growth
    // happened before the stage ran)
    mx->fused_cache->epoch = ctx->change_epoch;
    mx->fused_cache->frames = frames;
    mx->fused_cache->group = (const void*)&g;
    mx->fused_cache->vb = vb;
    mx->fused_cache->mb = mb;
    ctx->fused_group_now = (const void*)&g;
    vb.late_wait = 0;                                      // parameters were just rebuilt (uploads, line growth): wait first
    MXL_TRY(k::launch_fused_voice(ctx, vb));
    for (int i = 0; i < nv; i++) ((EqThree*)g.voices[i].eq)->cur ^= 1;
    return k::launch_fused_mix(ctx, mb);
}

// API-level line bytes of the modules the group replaces (inputs read + outputs written), as the staged
// stages count them
static int fused_bytes(const FusedGroup& g, uint64_t frames, uint64_t* bytes)
{
    uint64_t n = 0;
    for (const FusedVoiceRef& v : g.voices) n += (v.osc ? 12 * frames : 0) + (v.osc ? 4 * frames : 0) + 4 * frames;
    for (const FusedChanRef& cr : g.chans)
        if (cr.connected) n += (cr.left >= 0 ? 4 * frames : 0) + (cr.right >= 0 ? 4 * frames : 0) + 8 * frames + 8 * frames;
    n += 16 * frames;
    if (g.meter) n += 8 * frames;
    *bytes += n;
    return MXL_OK;
}

// --- Plotter: plotter.rs:37-56 ---
static int run_plotters(mxl_ctx* ctx, mxl_module* const* mods, int n, const IoSet* io, uint64_t* bytes)
{
    for (int i = 0; i < n; i++) {
        Plotter* m = (Plotter*)mods[i];
        NEED_IO(io[i], 1, 0, "Plotter");
        MXL_TRY(expect_input(io[i].in[0], MXL_LINE_STEREO, "Plotter"));
        const uint32_t S = ctx->spt;
        const uint64_t f = io[i].in[0] ? io[i].in[0]->frames : S;
        const uint64_t ticks = (f + S - 1) / S;          // one reference run_tick per tick of the call
        // the last tick k in [0, ticks) with (count + k + 1) % 6 == 0
        int64_t hit = -1;
        for (uint64_t kk = ticks; kk-- > 0;) {
            if ((m->count + kk + 1) % 6 == 0) { hit = (int64_t)kk; break; }
        }
        m->count += ticks;
        m->tap_frames = 0;
        if (hit >= 0 && io[i].in[0]) {                   // `&& inputs[0].connected()` (plotter.rs:40)
            const uint64_t begin = (uint64_t)hit * S;
            const uint64_t len = std::min<uint64_t>(S, f - begin);
            MXL_TRY(m->tap.ensure(ctx, 2 * (size_t)S * sizeof(float)));
            // the splitter kernel's 16-byte vector path needs an aligned start; an odd `begin`
            // (odd S, odd tick) starts mid-vector and goes through strided copies instead
            if ((2 * begin * sizeof(float)) % 16 == 0) {
                k::SplitBatch b{};
                b.frames = len; b.n = 1;
                b.inst[0] = k::SplitInst{io[i].in[0]->dev + 2 * begin, (float*)m->tap.p, (float*)m->tap.p + S};
                MXL_TRY(k::launch_splitter(ctx, b));
            } else {
                MXL_CUDA(cudaMemcpy2DAsync(m->tap.p, sizeof(float), io[i].in[0]->dev + 2 * begin, 2 * sizeof(float), sizeof(float), len, cudaMemcpyDeviceToDevice, ctx->stream));
                MXL_CUDA(cudaMemcpy2DAsync((float*)m->tap.p + S, sizeof(float), io[i].in[0]->dev + 2 * begin + 1, 2 * sizeof(float), sizeof(float), len, cudaMemcpyDeviceToDevice, ctx->stream));
            }
            m->tap_frames = (uint32_t)len;
            if (bytes) *bytes += 16 * len;
        }
    }
    return MXL_OK;
}

// --- PcmSink ---
static int run_pcm_sinks(mxl_ctx* ctx, mxl_module* const* mods, int n, const IoSet* io, uint64_t* bytes)
{
    for (int i = 0; i < n; i++) {
        PcmSink* m = (PcmSink*)mods[i];
        NEED_IO(io[i], 1, 0, "PcmSink");
        MXL_TRY(expect_input(io[i].in[0], MXL_LINE_STEREO, "PcmSink"));
        m->n_samples = 0;
        if (!io[i].in[0]) continue;
        const uint64_t len = io[i].in[0]->len();
        MXL_TRY(m->pcm.ensure(ctx, len * sizeof(int16_t)));
        MXL_TRY(k::launch_pcm_pack(ctx, io[i].in[0]->dev, (int16_t*)m->pcm.p, len));
        m->n_samples = len;
        if (bytes) *bytes += 6 * len;
    }
    return MXL_OK;
}

// ================================================================================================
// StreamInput::run_tick: src/module/stream_input.rs:72-147
// ================================================================================================
int StreamInput::staging(size_t samples, int16_t** out)
{
    turn ^= 1;
    const int i = turn;
    if (!pinned_ev[i]) MXL_CUDA(cudaEventCreateWithFlags(&pinned_ev[i], cudaEventDisableTiming));
    else MXL_CUDA(cudaEventSynchronize(pinned_ev[i]));             // the copy that last read this buffer has finished
    if (pinned_cap[i] < samples) {
        if (pinned[i]) { MXL_CUDA(cudaFreeHost(pinned[i])); pinned[i] = nullptr; pinned_cap[i] = 0; }
        const size_t want = samples + samples / 4 + 64;
        MXL_CUDA(cudaMallocHost(&pinned[i], want * sizeof(int16_t)));
        pinned_cap[i] = want;
    }
    *out = pinned[i];
    return MXL_OK;
}

int StreamInput::run(uint64_t t0, const IoSet& io, uint64_t* bytes)
{
    NEED_IO(io, 0, 2, "StreamInput");
    MXL_TRY(expect_output(io.out[0], MXL_LINE_VIDEO, "StreamInput.Video"));
    MXL_TRY(expect_output(io.out[1], MXL_LINE_STEREO, "StreamInput.Audio"));
    const uint64_t ticks = io.out[0]->slots.size();
    const uint64_t total = io.out[1]->len();                          // f32 of the stereo line over the whole call
    if (ticks == 0) { if (total) MXL_FAIL(MXL_ERR_LENGTH, "StreamInput: audio line without a video tick slot"); return MXL_OK; }
    if (total % ticks) MXL_FAIL(MXL_ERR_LENGTH, "StreamInput: %llu samples do not divide into %llu ticks", (unsigned long long)total, (unsigned long long)ticks);
    const uint64_t len_tick = total / ticks;                          // audio_out.len()
    int16_t* stage = nullptr;
    if (total) MXL_TRY(staging(total, &stage));
    const int64_t sr = (int64_t)ctx->sample_rate;

    for (uint64_t kk = 0; kk < ticks; kk++) {
        const Rational engine_time = Rational::make((int64_t)(t0 + kk * (len_tick / 2)), sr);       // 73
        const Rational tick_duration = Rational::make((int64_t)(len_tick / 2), sr);                  // 80
        // 82-86: the frame held back from an earlier tick, else the next one from the receiver
        bool have_video = false;
        VideoFrame vf;
        if (has_video_frame) { vf = video_frame; has_video_frame = false; have_video = true; }
        else if (!video_rx.empty()) { vf = video_rx.front(); video_rx.pop_front(); have_video = true; }
        const bool had_source = has_source;                                                           // 88
        const uint64_t existing_source_id = source_id;
        // 92-124: several input audio frames (or part of one) fill the tick
        int16_t* out = stage + kk * len_tick;
        uint64_t remaining = len_tick;
        while (remaining > 0) {
            AudioFrame fr;
            bool have = false;
            if (has_audio_frame) { fr = std::move(audio_frame); has_audio_frame = false; have = true; }
            else if (!audio_rx.empty()) { fr = std::move(audio_rx.front()); audio_rx.pop_front(); have = true; }
            if (!have) {                                                                              // 120-123 util::zero
                memset(out, 0, remaining * sizeof(int16_t));                                          // 0 / 32768.0 == 0.0
                break;
            }
            if (!had_source || existing_source_id != fr.source_id) {                                  // 100-106 source changed
                has_source = true;
                source_id = fr.source_id;
                epoch = engine_time - fr.source_time;                                                 // remove_epoch
            }
            const uint64_t avail = fr.data.size() - fr.head;
            const uint64_t len = std::min<uint64_t>(remaining, avail);                                // 108
            memcpy(out, fr.data.data() + fr.head, len * sizeof(int16_t));                             // 110-112 (converted on the device)
            out += len;                                                                               // 114
            remaining -= len;
            if (len < avail) {                                                                        // 116-119 drain + put back
                fr.head += len;
                audio_frame = std::move(fr);
                has_audio_frame = true;
            }
        }
        // 126-143
        if (have_video) {
            Rational tick_offset;                                                                     // zero
            if (has_source) {
                const Rational off = (vf.source_time + epoch) - engine_time;                          // add_epoch(..) - engine_time
                if (off >= Rational()) tick_offset = off;                                             // filter(>= zero).unwrap_or(zero)
            }
            if (tick_offset > tick_duration) {                                                        // not due for this tick, put it back
                video_frame = vf;
                has_video_frame = true;
                video_slot_set(io.out[0]->slots[kk], nullptr, Rational(), Rational());
            } else {
                video_slot_set(io.out[0]->slots[kk], vf.frame, vf.duration_hint, tick_offset);
                frame_release(vf.frame);                                                              // the line holds the reference now
            }
        } else {
            video_slot_set(io.out[0]->slots[kk], nullptr, Rational(), Rational());
        }
    }
    if (total) {
        MXL_TRY(pcm.ensure(ctx, total * sizeof(int16_t)));
        MXL_CUDA(cudaMemcpyAsync(pcm.p, stage, total * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
        MXL_CUDA(cudaEventRecord(pinned_ev[turn], ctx->stream));
        ctx->h2d_bytes += total * sizeof(int16_t);
        MXL_TRY(k::launch_pcm_unpack(ctx, (const int16_t*)pcm.p, io.out[1]->dev, total));
        if (bytes) *bytes += 6 * total;
    }
    return MXL_OK;
}

int stream_input_write_audio(mxl_module* m, uint64_t source_id, Rational time, const int16_t* samples, uint64_t n)
{
    if (!m || m->kind != MXL_MOD_STREAM_INPUT) MXL_FAIL(MXL_ERR_PARAMS, "not a StreamInput module");
    if (n && !samples) MXL_FAIL(MXL_ERR_INVALID, "NULL samples");
    StreamInput* s = (StreamInput*)m;
    if (s->audio_rx.size() >= StreamInput::kRingCapacity) MXL_FAIL(MXL_ERR_LENGTH, "StreamInput: audio queue full");   // push -> Err(()) (source.rs:168)
    StreamInput::AudioFrame f;
    f.source_id = source_id;
    f.source_time = time;
    f.data.assign(samples, samples + n);
    s->audio_rx.push_back(std::move(f));
    return MXL_OK;
}

int stream_input_write_video(mxl_module* m, uint64_t source_id, Rational time, mxl_frame* frame, Rational duration)
{
    if (!m || m->kind != MXL_MOD_STREAM_INPUT) MXL_FAIL(MXL_ERR_PARAMS, "not a StreamInput module");
    if (!frame) MXL_FAIL(MXL_ERR_INVALID, "NULL frame");
    if (frame->ctx != m->ctx) MXL_FAIL(MXL_ERR_INVALID, "frame belongs to another context");
    StreamInput* s = (StreamInput*)m;
    if (s->video_rx.size() >= StreamInput::kRingCapacity) MXL_FAIL(MXL_ERR_LENGTH, "StreamInput: video queue full");   // source.rs:186
    StreamInput::VideoFrame f;
    f.source_id = source_id;
    f.source_time = time;
    f.frame = frame_retain(frame);
    f.duration_hint = duration;
    s->video_rx.push_back(f);
    return MXL_OK;
}

int stream_input_pending(const mxl_module* m, uint32_t* audio_frames, uint32_t* video_frames)
{
    if (!m || m->kind != MXL_MOD_STREAM_INPUT) MXL_FAIL(MXL_ERR_PARAMS, "not a StreamInput module");
    const StreamInput* s = (const StreamInput*)m;
    if (audio_frames) *audio_frames = (uint32_t)(s->audio_rx.size() + (s->has_audio_frame ? 1 : 0));
    if (video_frames) *video_frames = (uint32_t)(s->video_rx.size() + (s->has_video_frame ? 1 : 0));
    return MXL_OK;
}

// ================================================================================================
// Monitor::run_tick (monitor.rs:112-140), the codec thread's loop body (235-247) and EncodeStream
// (src/video/encode.rs:46-100) up to the encoder calls
// ================================================================================================
static int64_t round_to_base(const Rational& r, int64_t base)     // MediaTime::round_to_base (util/src/time.rs:17-19)
{
    return (int64_t)(((__int128)r.num * base) / r.den);            // Ratio * base, to_integer(): toward zero
}

int Monitor::chunk_for(size_t samples, Chunk* out)
{
    Chunk c;
    for (size_t i = 0; i < free_chunks.size(); i++)
        if (free_chunks[i].cap >= samples) { c = free_chunks[i]; free_chunks.erase(free_chunks.begin() + i); break; }
    if (!c.host) {
        if (!free_chunks.empty()) {                                // too small: replace the oldest spare
            Chunk old = free_chunks.front();
            free_chunks.erase(free_chunks.begin());
            cudaFreeHost(old.host);
            c.ev = old.ev;
        }
        const size_t want = samples + samples / 4 + 64;
        MXL_CUDA(cudaMallocHost(&c.host, want * sizeof(int16_t)));
        c.cap = want;
        if (!c.ev) MXL_CUDA(cudaEventCreateWithFlags(&c.ev, cudaEventDisableTiming));
    }
    c.n = samples;
    c.in_flight = false;
    *out = c;
    return MXL_OK;
}

// EncodeStream::encode_video (encode.rs:86-100) up to video_ctx.send_frame's scaler (done for the whole call at once)
void Monitor::encode_video(Rational duration, mxl_frame* frame, bool is_blank)
{
    const int64_t tb = time_base();
    const Rational start = video_timestamp, end = video_timestamp + duration;
    video_timestamp = end;
    const int64_t start_in_base = round_to_base(start, tb), end_in_base = round_to_base(end, tb);
    video_jobs.push_back(Job{start_in_base, end_in_base - start_in_base, is_blank, frame_retain(frame)});
}

// a new LiveOutput (stream_output.rs:328-366): fresh EncodeStream, epoch taken from the next tick
void Monitor::reset_stream()
{
    has_epoch = false;
    audio_timestamp = Rational();
    video_timestamp = Rational();
    pcm_len = 0;
    // what the previous stream still holds is the previous publisher's: dropped with it
    audio_segments.clear();
    for (auto& j : video_jobs) frame_release(j.frame);
    video_jobs.clear();
    if (ctx && ctx->has_device() && !chunks.empty()) { ctx->activate(); cudaStreamSynchronize(ctx->stream); }
    for (auto& c : chunks) { c.in_flight = false; free_chunks.push_back(c); }
    chunks.clear();
    chunk_head = 0;
}

int stream_output_set_live(mxl_module* m, int live)
{
    if (!m || m->kind != MXL_MOD_STREAM_OUTPUT) MXL_FAIL(MXL_ERR_PARAMS, "not a StreamOutput module");
    Monitor* mo = (Monitor*)m;
    if ((live != 0) == mo->live) return MXL_OK;
    mo->live = live != 0;
    if (mo->live) mo->reset_stream();
    return MXL_OK;
}

int Monitor::run(uint64_t t0, const IoSet& io, uint64_t* bytes)
{
    NEED_IO(io, 2, 0, "Monitor");
    MXL_TRY(expect_input(io.in[0], MXL_LINE_VIDEO, "Monitor.Video"));
    MXL_TRY(expect_input(io.in[1], MXL_LINE_STEREO, "Monitor.Audio"));
    if (!live) return MXL_OK;                                          // Connection::Offline / Failed / Connecting: nothing is sent (stream_output.rs:112-151)
    const mxl_line* video = io.in[0];
    const mxl_line* audio = io.in[1];
    // ticks of the call: the video line's slots, else the audio line in ticks of SAMPLES_PER_TICK, else one
    // (a disconnected input is the engine's static buffer of one tick, io.rs:8-9)
    uint64_t ticks = 1;
    if (video) ticks = video->slots.size();
    else if (audio) ticks = std::max<uint64_t>(1, (audio->frames + ctx->spt - 1) / ctx->spt);
    if (ticks == 0) return MXL_OK;
    const uint64_t total = audio ? audio->len() : 2ull * ctx->spt * ticks;
    if (total % ticks) MXL_FAIL(MXL_ERR_LENGTH, "Monitor: %llu samples do not divide into %llu ticks", (unsigned long long)total, (unsigned long long)ticks);
    const uint64_t len_tick = total / ticks;                          // audio.len()
    const int64_t sr = (int64_t)ctx->sample_rate;
    if (!blank) {
        blank = mxl_frame_blank(ctx, p.width, p.height);
        if (!blank) return MXL_ERR_OOM;
    }

    // ---- audio: the whole call packed in one launch, downloaded into one pinned chunk ----
    if (total) {
        Chunk c;
        MXL_TRY(chunk_for(total, &c));
        if (audio) {
            MXL_TRY(pcm.ensure(ctx, total * sizeof(int16_t)));
            MXL_TRY(k::launch_pcm_pack(ctx, audio->dev, (int16_t*)pcm.p, total));
            MXL_CUDA(cudaMemcpyAsync(c.host, pcm.p, total * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
            MXL_CUDA(cudaEventRecord(c.ev, ctx->stream));
            c.in_flight = true;
            ctx->d2h_bytes += total * sizeof(int16_t);
            if (bytes) *bytes += 6 * total;
        } else {
            memset(c.host, 0, total * sizeof(int16_t));               // (0.0 * 32767) as i16
        }
        chunks.push_back(c);
    }

    const size_t first_job = video_jobs.size();
    for (uint64_t kk = 0; kk < ticks; kk++) {
        const Rational absolute = Rational::make((int64_t)(t0 + kk * (len_tick / 2)), sr);             // monitor.rs:118
        if (!has_epoch) { has_epoch = true; epoch = absolute; }                                         // 119 get_or_insert
        const Rational timestamp = absolute - epoch;                                                    // 120
        // encode.send_audio(&tick.audio) (monitor.rs:236; encode.rs:46-59,184-221): one fragment at most per tick
        pcm_len += len_tick;
        if (pcm_len > kFragmentSamples) {
            const Rational duration = Rational::make(1024, sr);
            audio_segments.push_back(Fragment{audio_timestamp, duration});
            audio_timestamp = audio_timestamp + duration;
            pcm_len -= kFragmentSamples;
        }
        // monitor.rs:238-243
        if (video && video->slots[kk].frame) {
            const VideoSlot& vs = video->slots[kk];
            const Rational frame_timestamp = timestamp + vs.tick_offset;
            const Rational end_timestamp = frame_timestamp + vs.duration_hint;                          // encode.rs:62
            if (end_timestamp >= video_timestamp)                                                       // 64-67: ends before the clock -> dropped
                encode_video(end_timestamp - video_timestamp, vs.frame, false);                         // 73-75
        }
        if (timestamp > video_timestamp) encode_video(timestamp - video_timestamp, blank, true);       // 245; encode.rs:78-84
    }

    // ---- VideoCtx::send_frame's scaler (encode.rs:279-287) for the jobs of this call, one launch per geometry ----
    std::vector<size_t> todo;
    for (size_t j = first_job; j < video_jobs.size(); j++) {
        const mxl_frame_layout& l = video_jobs[j].frame->layout;
        if (!video_jobs[j].blank && (l.width != p.width || l.height != p.height)) todo.push_back(j);
    }
    while (!todo.empty()) {
        const mxl_frame_layout l0 = video_jobs[todo[0]].frame->layout;
        std::vector<size_t> grp, rest;
        for (size_t j : todo) {
            const mxl_frame_layout& l = video_jobs[j].frame->layout;
            (l.width == l0.width && l.height == l0.height ? grp : rest).push_back(j);
        }
        std::vector<mxl_frame*> src(grp.size()), dst(grp.size(), nullptr);
        for (size_t i = 0; i < grp.size(); i++) src[i] = video_jobs[grp[i]].frame;
        MXL_TRY(frames_scale(ctx, src.data(), dst.data(), (uint32_t)grp.size(), p.width, p.height));
        for (size_t i = 0; i < grp.size(); i++) {
            frame_release(video_jobs[grp[i]].frame);
            video_jobs[grp[i]].frame = dst[i];
        }
        todo.swap(rest);
    }
    return MXL_OK;
}

int monitor_recv_audio(mxl_module* m, mxl_audio_fragment* info, int16_t* pcm_out, uint32_t cap)
{
    if (!m || (m->kind != MXL_MOD_MONITOR && m->kind != MXL_MOD_STREAM_OUTPUT)) MXL_FAIL(MXL_ERR_PARAMS, "not a Monitor / StreamOutput module");
    Monitor* mo = (Monitor*)m;
    if (mo->audio_segments.empty()) return 0;
    if (!info || !pcm_out || cap < Monitor::kFragmentSamples) MXL_FAIL(MXL_ERR_INVALID, "mxl_monitor_recv_audio: room for %llu samples needed", (unsigned long long)Monitor::kFragmentSamples);
    const Monitor::Fragment f = mo->audio_segments.front();
    mo->audio_segments.pop_front();
    uint64_t need = Monitor::kFragmentSamples;
    int16_t* out = pcm_out;
    while (need > 0) {
        if (mo->chunks.empty()) MXL_FAIL(MXL_ERR_INVALID, "Monitor: PCM stream shorter than its fragments (internal)");
        Monitor::Chunk& c = mo->chunks.front();
        if (c.in_flight) { MXL_TRY(m->ctx->activate()); MXL_CUDA(cudaEventSynchronize(c.ev)); c.in_flight = false; }
        const uint64_t take = std::min<uint64_t>(need, c.n - mo->chunk_head);
        memcpy(out, c.host + mo->chunk_head, take * sizeof(int16_t));
        out += take; need -= take; mo->chunk_head += take;
        if (mo->chunk_head == c.n) {
            mo->free_chunks.push_back(c);
            mo->chunks.pop_front();
            mo->chunk_head = 0;
        }
    }
    info->decode_num = f.decode_timestamp.num; info->decode_den = f.decode_timestamp.den;
    info->duration_num = f.duration.num; info->duration_den = f.duration.den;
    info->n_samples = (uint32_t)Monitor::kFragmentSamples; info->_pad = 0;
    return 1;
}

int monitor_recv_video(mxl_module* m, mxl_video_job* out)
{
    if (!m || (m->kind != MXL_MOD_MONITOR && m->kind != MXL_MOD_STREAM_OUTPUT)) MXL_FAIL(MXL_ERR_PARAMS, "not a Monitor / StreamOutput module");
    Monitor* mo = (Monitor*)m;
    if (mo->video_jobs.empty()) return 0;
    if (!out) MXL_FAIL(MXL_ERR_INVALID, "NULL job");
    const Monitor::Job j = mo->video_jobs.front();
    mo->video_jobs.pop_front();
    out->pts = j.pts; out->duration = j.duration; out->time_base = mo->time_base();
    out->blank = j.blank ? 1 : 0; out->_pad = 0;
    out->frame = j.frame;                                             // the queue's reference passes to the caller
    return 1;
}

// ================================================================================================
// OutputDevice::run_tick: src/module/output_device.rs:173-206
// ================================================================================================
int OutputDevice::run(const IoSet& io, uint64_t* bytes)
{
    NEED_IO(io, 1, 0, "OutputDevice");
    MXL_TRY(expect_input(io.in[0], MXL_LINE_STEREO, "OutputDevice"));
    if (!clip_host) {
        MXL_CUDA(cudaMallocHost(&clip_host, sizeof(int32_t)));
        MXL_TRY(clip_flag.ensure(ctx, sizeof(int32_t)));
    }
    MXL_CUDA(cudaMemsetAsync(clip_flag.p, 0, sizeof(int32_t), ctx->stream));                      // let mut clip = false (176)
    if (p.channels == 0) return MXL_OK;                                                            // no stream (178)
    const uint64_t frames = io.in[0] ? io.in[0]->frames : ctx->spt;                                // input.len() / CHANNELS (180)
    const uint64_t n = frames * p.channels;                                                        // scratch_len (181)
    if (n == 0) return MXL_OK;
    if (scratch_len < n) {                                                                         // 183-185 resize(.., 0.0)
        // unrouted channels only ever hold zeros (update() clears on every re-assignment), so a fresh zeroed
        // buffer equals the grown one
        MXL_TRY(scratch.ensure(ctx, n * sizeof(float)));
        MXL_CUDA(cudaMemsetAsync(scratch.p, 0, n * sizeof(float), ctx->stream));
        scratch_len = n;
    }
    k::RouteLaunch r{io.in[0] ? io.in[0]->dev : nullptr, (float*)scratch.p, frames, p.channels, p.left, p.right, (int32_t*)clip_flag.p};
    MXL_TRY(k::launch_route(ctx, r));                                                              // 187-205
    if (bytes) *bytes += (io.in[0] ? 8 * frames : 0) + 4 * frames * ((p.left >= 0) + (p.right >= 0 && p.right != p.left));
    // stream.tx.push_slice(&scratch[0..n]) (207): as much as the ring still takes
    const uint64_t take = std::min<uint64_t>(n, kRingCapacity - std::min<uint64_t>(pending, kRingCapacity));
    if (take) {
        Chunk c;
        for (size_t i = 0; i < spare.size(); i++)
            if (spare[i].cap >= take) { c = spare[i]; spare.erase(spare.begin() + i); break; }
        if (!c.host) {
            if (!spare.empty()) { Chunk old = spare.front(); spare.erase(spare.begin()); cudaFreeHost(old.host); c.ev = old.ev; }
            const size_t want = take + take / 4 + 64;
            MXL_CUDA(cudaMallocHost(&c.host, want * sizeof(float)));
            c.cap = want;
            if (!c.ev) MXL_CUDA(cudaEventCreateWithFlags(&c.ev, cudaEventDisableTiming));
        }
        c.n = take;
        MXL_CUDA(cudaMemcpyAsync(c.host, scratch.p, take * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        MXL_CUDA(cudaEventRecord(c.ev, ctx->stream));
        c.in_flight = true;
        ctx->d2h_bytes += take * sizeof(float);
        chunks.push_back(c);
        pending += take;
    }
    return MXL_OK;
}

int64_t output_device_read(mxl_module* m, float* out, uint64_t cap)
{
    if (!m || m->kind != MXL_MOD_OUTPUT_DEVICE) MXL_FAIL(MXL_ERR_PARAMS, "not an OutputDevice module");
    if (cap && !out) MXL_FAIL(MXL_ERR_INVALID, "NULL buffer");
    OutputDevice* d = (OutputDevice*)m;
    uint64_t got = 0;
    while (got < cap && !d->chunks.empty()) {
        OutputDevice::Chunk& c = d->chunks.front();
        if (c.in_flight) { MXL_TRY(m->ctx->activate()); MXL_CUDA(cudaEventSynchronize(c.ev)); c.in_flight = false; }
        const uint64_t take = std::min<uint64_t>(cap - got, c.n - d->head);
        memcpy(out + got, c.host + d->head, take * sizeof(float));
        got += take; d->head += take;
        if (d->head == c.n) { d->spare.push_back(c); d->chunks.pop_front(); d->head = 0; }
    }
    d->pending -= got;
    return (int64_t)got;
}

int output_device_clip(mxl_module* m, int32_t* clip)
{
    if (!m || m->kind != MXL_MOD_OUTPUT_DEVICE || !clip) MXL_FAIL(MXL_ERR_PARAMS, "not an OutputDevice module");
    OutputDevice* d = (OutputDevice*)m;
    *clip = 0;
    if (!d->clip_host) return MXL_OK;                                 // never ran
    MXL_TRY(m->ctx->activate());
    MXL_CUDA(cudaMemcpyAsync(d->clip_host, d->clip_flag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, m->ctx->stream));
    MXL_CUDA(cudaStreamSynchronize(m->ctx->stream));
    *clip = *d->clip_host;
    return MXL_OK;
}

// ================================================================================================
// VideoMixer: src/module/video_mixer.rs:70-250
// ================================================================================================

// 4-tap tables of the letterbox scaler (self-specified stand-in for swscale's SWS_BICUBIC)
static double cubic_weight(double x)
{
    const double a = -0.6;
    x = fabs(x);
    if (x <= 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
    if (x < 2.0) return ((a * x - 5.0 * a) * x + 8.0 * a) * x - 4.0 * a;
    return 0.0;
}

static void bicubic_table(uint32_t src_n, uint32_t dst_n, std::vector<int32_t>& pos, std::vector<int16_t>& coef)
{
    pos.resize(dst_n);
    coef.resize((size_t)dst_n * 4);
    for (uint32_t d = 0; d < dst_n; d++) {
        const int64_t num = (int64_t)(2 * (uint64_t)d + 1) * src_n - dst_n;
        const int64_t den = 2 * (int64_t)dst_n;
        const int64_t ix = num >= 0 ? num / den : -((-num + den - 1) / den);
        const double frac = (double)(num - ix * den) / (double)den;
        int w[4], sum = 0, best = 0;
        for (int q = 0; q < 4; q++) {
            w[q] = (int)lrint(cubic_weight(frac - (double)(q - 1)) * 16384.0);
            sum += w[q];
            if (w[q] > w[best]) best = q;
        }
        w[best] += 16384 - sum;
        pos[d] = (int32_t)ix - 1;
        for (int q = 0; q < 4; q++) coef[(size_t)d * 4 + q] = (int16_t)w[q];
    }
}

// Device copy of one axis' tap table, cached per (source length, destination length): the tables only
// depend on the two lengths, so a live session builds each of them once.
static int axis_table(mxl_ctx* ctx, uint32_t src_n, uint32_t dst_n, const int32_t** pos, const int16_t** coef, const std::vector<int32_t>** host_pos)
{
    const uint64_t key = ((uint64_t)src_n << 32) | dst_n;
    const size_t pos_bytes = (((size_t)dst_n * sizeof(int32_t)) + 7) & ~(size_t)7;
    void*& dev = ctx->scale_tables[key];
    if (!dev) {
        std::vector<int32_t> p;
        std::vector<int16_t> c;
        bicubic_table(src_n, dst_n, p, c);
        std::vector<uint8_t> blob(pos_bytes + c.size() * sizeof(int16_t), 0);
        memcpy(blob.data(), p.data(), p.size() * sizeof(int32_t));
        memcpy(blob.data() + pos_bytes, c.data(), c.size() * sizeof(int16_t));
        MXL_CUDA(cudaMalloc(&dev, blob.size()));
        // pageable source: the copy is staged before the call returns, `blob` may die
        MXL_CUDA(cudaMemcpyAsync(dev, blob.data(), blob.size(), cudaMemcpyHostToDevice, ctx->stream));
        ctx->scale_positions[key] = std::move(p);
    }
    *pos = (const int32_t*)dev;
    *coef = (const int16_t*)((const uint8_t*)dev + pos_bytes);
    *host_pos = &ctx->scale_positions[key];
    return MXL_OK;
}

// DynamicScaler::scale (src/video/encode.rs:338-397), split so that the frames of a multi-tick call can be
// scaled by one launch: scale_target() makes the destination frame of one source (encode.rs:382), scale_run()
// scales a batch of (source, destination) pairs that share source and target sizes -- all planes of all
// frames in ONE launch of the tiled kernel.
mxl_frame* scale_target(mxl_ctx* ctx, const mxl_frame_layout& sl, uint32_t out_w, uint32_t out_h)
{
    mxl_scale_geometry g;
    if (mxl_scale_geometry_yuv420p(sl.width, sl.height, out_w, out_h, &g) != MXL_OK) return nullptr;
    // The reference scales into AvFrame::blank.  The blank fill can be skipped only when the scaled picture
    // covers every byte of the target, stride padding included.
    mxl_frame_layout tl;
    frame_layout_yuv420p(out_w, out_h, &tl);
    const bool bars = g.scaled_w != out_w || g.scaled_h != out_h || tl.stride[0] != out_w || tl.stride[1] != (out_w + 1) / 2;
    return bars ? mxl_frame_blank(ctx, out_w, out_h) : frame_alloc(ctx, out_w, out_h);
}

int scale_run(mxl_ctx* ctx, const mxl_frame_layout& sl, uint32_t out_w, uint32_t out_h, const k::ScaleJob* jobs, uint32_t n)
{
    if (n == 0) return MXL_OK;
    mxl_scale_geometry g;
    MXL_TRY(mxl_scale_geometry_yuv420p(sl.width, sl.height, out_w, out_h, &g));
    if (g.scaled_w == 0 || g.scaled_h == 0) return MXL_OK;
    mxl_frame_layout dl;
    frame_layout_yuv420p(out_w, out_h, &dl);
    k::ScaleLaunch L{};
    const uint32_t tw = k::scale_tile_width();
    auto clampi = [](int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); };
    // tallest tile whose staged source rectangle fits comfortably (several CTAs per SM): 128 output rows when the batch
    // still fills the machine four CTAs deep (a tile's prologue -- geometry, tap-table slices, staging set-up -- is a
    // quarter of the kernel's instructions at 64 rows), else 64, 32, or 8 / 2 for strong down-scales
    static const uint32_t th_max = getenv("MXL_SCALE_TH") ? (uint32_t)atoi(getenv("MXL_SCALE_TH")) : 128u;
    const uint64_t sms = (uint64_t)(ctx->sm_count > 0 ? ctx->sm_count : 148);
    for (uint32_t th : {128u, 64u, 32u, 8u, 2u}) {
        if (th > th_max && th > 2u) continue;
        uint32_t tile_base = 0, max_span = 16, max_rows = 1;
        for (int p = 0; p < 3; p++) {
            // subframe addressing (codec/src/ffmpeg/frame.rs:219-281): offsets are chroma-aligned already
            const uint32_t sh = p ? 1 : 0;
            k::ScalePlane& P = L.pl[p];
            P.src_off = (uint32_t)sl.offset[p];
            P.src_w = p ? (sl.width + 1) / 2 : sl.width;
            P.src_h = sl.plane_h[p];
            P.src_stride = sl.stride[p];
            P.dst_w = g.scaled_w >> sh;
            P.dst_h = g.scaled_h >> sh;
            P.dst_stride = dl.stride[p];
            P.dst_off = (uint32_t)(dl.offset[p] + (uint64_t)(g.letterbox_y >> sh) * dl.stride[p] + (g.letterbox_x >> sh));
            P.tile_base = tile_base;
            if (P.dst_w == 0 || P.dst_h == 0) { P.tiles_x = P.tiles_y = 0; continue; }
            const std::vector<int32_t>*hx, *hy;
            MXL_TRY(axis_table(ctx, P.src_w, P.dst_w, &P.xpos, &P.xcoef, &hx));
            MXL_TRY(axis_table(ctx, P.src_h, P.dst_h, &P.ypos, &P.ycoef, &hy));
            P.tiles_x = (P.dst_w + tw - 1) / tw;
            P.tiles_y = (P.dst_h + th - 1) / th;
            P.tiles_x_magic = (uint32_t)((0x100000000ull + P.tiles_x - 1) / P.tiles_x);     // unused when tiles_x == 1
            tile_base += P.tiles_x * P.tiles_y;
            for (uint32_t t = 0; t < P.tiles_x; t++) {      // staged bytes per row: from the 16-floored (unclamped) first tap
                const uint32_t x0 = t * tw, x1 = std::min(x0 + tw, P.dst_w) - 1;     // to the word after the last tap set
                const int lo = (*hx)[x0] & ~15, hi = (*hx)[x1] + 3;
                max_span = std::max<uint32_t>(max_span, (uint32_t)((hi - lo + 5 + 15) & ~15));
            }
            for (uint32_t t = 0; t < P.tiles_y; t++) {
                const uint32_t y0 = t * th, y1 = std::min(y0 + th, P.dst_h) - 1;
                const int lo = clampi((*hy)[y0], (int)P.src_h - 1), hi = clampi((*hy)[y1] + 3, (int)P.src_h - 1);
                max_rows = std::max<uint32_t>(max_rows, (uint32_t)(hi - lo + 1));
            }
        }
        L.total_tiles = tile_base;
        L.region_pitch = max_span;
        L.region_rows = max_rows;
        L.tile_h = th;
        if (th == 128 && ((uint64_t)tile_base * n < 20 * sms || k::scale_smem_bytes(max_rows, max_span) > 40u * 1024u)) continue;
        if (k::scale_smem_bytes(max_rows, max_span) <= (th == 64 ? 72u * 1024u : k::kScaleMaxSmem)) break;
    }
    // job tables rotate through a ring: a table must stay intact until the launch that reads it has run,
    // and several batches may be queued before the stream gets to the first
    const size_t need = (size_t)n * sizeof(k::ScaleJob);
    if (ctx->scale_jobs_cap < need * 2 || ctx->scale_jobs_used + need > ctx->scale_jobs_cap) {
        MXL_CUDA(cudaStreamSynchronize(ctx->stream));                                  // every queued reader is done
        if (ctx->scale_jobs_cap < need * 2) {
            if (ctx->scale_jobs) cudaFree(ctx->scale_jobs);
            ctx->scale_jobs = nullptr;
            const size_t cap = std::max<size_t>(64 * 1024, need * 4);
            MXL_CUDA(cudaMalloc(&ctx->scale_jobs, cap));
            ctx->scale_jobs_cap = cap;
        }
        ctx->scale_jobs_used = 0;
    }
    k::ScaleJob* dev_jobs = (k::ScaleJob*)((uint8_t*)ctx->scale_jobs + ctx->scale_jobs_used);
    ctx->scale_jobs_used += (need + 15) & ~(size_t)15;
    MXL_CUDA(cudaMemcpyAsync(dev_jobs, jobs, need, cudaMemcpyHostToDevice, ctx->stream));
    L.jobs = dev_jobs;
    return k::launch_scale_tiled(ctx, L, n);
}

// dst[i] receives a new frame (or a retained src[i] when the sizes already agree, encode.rs:342-345).
int frames_scale(mxl_ctx* ctx, mxl_frame* const* src, mxl_frame** dst, uint32_t n, uint32_t out_w, uint32_t out_h)
{
    if (n == 0) return MXL_OK;
    const mxl_frame_layout& sl = src[0]->layout;
    for (uint32_t i = 0; i < n; i++) {
        dst[i] = nullptr;
        if (!src[i] || src[i]->layout.width != sl.width || src[i]->layout.height != sl.height)
            MXL_FAIL(MXL_ERR_INVALID, "frames_scale: frames of one batch must share a size");
    }
    if (sl.width == out_w && sl.height == out_h) {
        for (uint32_t i = 0; i < n; i++) dst[i] = frame_retain(src[i]);
        return MXL_OK;
    }
    std::vector<k::ScaleJob> jobs(n);
    int st = MXL_OK;
    for (uint32_t i = 0; i < n && st == MXL_OK; i++) {
        dst[i] = scale_target(ctx, sl, out_w, out_h);
        if (!dst[i]) st = MXL_ERR_OOM;
        else jobs[i] = k::ScaleJob{src[i]->dev, dst[i]->dev};
    }
    if (st == MXL_OK) st = scale_run(ctx, sl, out_w, out_h, jobs.data(), n);
    if (st != MXL_OK)
        for (uint32_t i = 0; i < n; i++) { if (dst[i]) frame_release(dst[i]); dst[i] = nullptr; }
    return st;
}

mxl_frame* frame_scale(mxl_frame* src, uint32_t out_w, uint32_t out_h)
{
    if (!src) { set_error("frame_scale: NULL frame"); return nullptr; }
    mxl_frame* dst = nullptr;
    if (frames_scale(src->ctx, &src, &dst, 1, out_w, out_h) != MXL_OK) {
        if (!*last_error()) set_error("frame_scale: CUDA failure");
        return nullptr;
    }
    return dst;
}

// Channel::rescale (video_mixer.rs:262-274)
int VideoMixer::rescale(Channel& c, const PictureSettings& target)
{
    if (!c.has_scaler || c.scaler_out != target) {
        c.has_scaler = true;
        c.scaler_out = target;
        if (c.has_stored) {
            mxl_frame* scaled = frame_scale(c.frame, target.w, target.h);
            if (!scaled) return MXL_ERR_CUDA;
            frame_release(c.frame);
            c.frame = scaled;
        }
    }
    return MXL_OK;
}

int VideoMixer::run(uint64_t t0, const IoSet& io, uint64_t* bytes)
{
    NEED_IO(io, MXL_VIDEO_MIXER_CHANNELS, 3, "VideoMixer");
    for (int i = 0; i < MXL_VIDEO_MIXER_CHANNELS; i++) MXL_TRY(expect_input(io.in[i], MXL_LINE_VIDEO, "VideoMixer input"));
    for (int i = 0; i < 3; i++) MXL_TRY(expect_output(io.out[i], MXL_LINE_VIDEO, "VideoMixer output"));
    const uint64_t ticks = io.out[0]->slots.size();
    for (int i = 0; i < MXL_VIDEO_MIXER_CHANNELS; i++)
        if (io.in[i] && io.in[i]->slots.size() < ticks) MXL_FAIL(MXL_ERR_LENGTH, "VideoMixer input %d has fewer tick slots than the output", i);
    if (io.out[1]->slots.size() < ticks || io.out[2]->slots.size() < ticks) MXL_FAIL(MXL_ERR_LENGTH, "VideoMixer A/B outputs have fewer tick slots than Output");
    if (ticks > 65535) MXL_FAIL(MXL_ERR_LENGTH, "VideoMixer: at most 65535 ticks per call");

    std::vector<k::FadeJob> jobs;
    std::vector<mxl_frame_layout> job_layouts;
    struct ScaleGroup { mxl_frame_layout sl; uint32_t w, h; std::vector<k::ScaleJob> jobs; };
    std::vector<ScaleGroup> scale_groups;                 // input frames waiting for the scaler, by geometry
    // The crossfade of all ticks is ONE launch at the end of the call, so every frame a job points
    // at must outlive later ticks' expiry/replacement (a released buffer goes back to the pool and
    // could be handed out again as a later tick's output).
    std::vector<mxl_frame*> keepalive;
    struct KeepGuard { std::vector<mxl_frame*>& v; ~KeepGuard() { for (mxl_frame* f : v) frame_release(f); } } keep_guard{keepalive};
    jobs.reserve(ticks);
    const Rational tick_duration = Rational::make((int64_t)ctx->spt, (int64_t)ctx->sample_rate);   // 1/TICKS_PER_SECOND (video_mixer.rs:244)
    // Input frames are scaled by ONE launch per geometry, normally at the end of the call.  A stored frame that has
    // to be scaled AGAIN inside the call (another channel's picture grew the unified target, video_mixer.rs:262-274)
    // may still be waiting for that launch: the pending groups run first, as the reference's synchronous scaler would
    // have, then the immediate rescale reads finished pixels.
    auto flush_scales = [&]() -> int {
        for (auto& gq : scale_groups) MXL_TRY(scale_run(ctx, gq.sl, gq.w, gq.h, gq.jobs.data(), (uint32_t)gq.jobs.size()));
        scale_groups.clear();
        return MXL_OK;
    };

    for (uint64_t kk = 0; kk < ticks; kk++) {
        const uint64_t t = t0 + kk * ctx->spt;
        auto in_slot = [&](int idx) -> const VideoSlot* {
            if (idx < 0 || idx >= MXL_VIDEO_MIXER_CHANNELS || !io.in[idx]) return nullptr;
            const VideoSlot& s = io.in[idx]->slots[kk];
            return s.frame ? &s : nullptr;
        };
        // send channel specific outputs (video_mixer.rs:80-90)
        {
            const VideoSlot* a = in_slot(p.a);
            const VideoSlot* b = in_slot(p.b);
            video_slot_set(io.out[1]->slots[kk], a ? a->frame : nullptr, a ? a->duration_hint : Rational(), a ? a->tick_offset : Rational());
            video_slot_set(io.out[2]->slots[kk], b ? b->frame : nullptr, b ? b->duration_hint : Rational(), b ? b->tick_offset : Rational());
        }
        const Rational now = Rational::make((int64_t)t, (int64_t)ctx->sample_rate);                // video_mixer.rs:92
        // expire stored frames (94-101)
        for (auto& c : ch)
            if (c.has_stored && now >= c.active_until) clear_stored(c);
        // compatible output picture settings (104-119): fold1 over live-or-stored frames
        bool have_target = false;
        PictureSettings target;
        for (int idx = 0; idx < MXL_VIDEO_MIXER_CHANNELS; idx++) {
            const VideoSlot* s = in_slot(idx);
            const mxl_frame* f = s ? s->frame : (ch[idx].has_stored ? ch[idx].frame : nullptr);
            if (!f) continue;
            PictureSettings ps{f->layout.width, f->layout.height};
            if (!have_target) { target = ps; have_target = true; }
            else { uint32_t w, h; mxl_unify_picture_settings(target.w, target.h, ps.w, ps.h, &w, &h); target = PictureSettings{w, h}; }
        }
        if (!have_target) {                                 // no inputs and no stored pictures (113-119)
            video_slot_set(io.out[0]->slots[kk], nullptr, Rational(), Rational());
            continue;
        }
        // receive new input frames (122-148)
        for (int idx = 0; idx < MXL_VIDEO_MIXER_CHANNELS; idx++) {
            Channel& c = ch[idx];
            if (const VideoSlot* s = in_slot(idx)) {
                clear_stored(c);
                MXL_TRY(rescale(c, target));
                // Channel::scale (video_mixer.rs:276-279): identity when the sizes agree, else into a new
                // frame whose pixels are produced by ONE scaler launch per geometry at the end of the call
                mxl_frame* scaled;
                const mxl_frame_layout& sl = s->frame->layout;
                if (sl.width == target.w && sl.height == target.h) {
                    scaled = frame_retain(s->frame);
                } else {
                    scaled = scale_target(ctx, sl, target.w, target.h);
                    if (!scaled) return MXL_ERR_OOM;
                    ScaleGroup* grp = nullptr;
                    for (auto& gq : scale_groups)
                        if (gq.sl.width == sl.width && gq.sl.height == sl.height && gq.w == target.w && gq.h == target.h) grp = &gq;
                    if (!grp) { scale_groups.push_back(ScaleGroup{sl, target.w, target.h, {}}); grp = &scale_groups.back(); }
                    grp->jobs.push_back(k::ScaleJob{s->frame->dev, scaled->dev});
                    // source and destination stay out of the frame pool until the call ends: the launch that
                    // reads / writes them is still to come, whatever later ticks do with this channel
                    keepalive.push_back(frame_retain(s->frame));
                    keepalive.push_back(frame_retain(scaled));
                }
                c.has_stored = true;
                c.frame = scaled;
                c.active_until = now + s->tick_offset + s->duration_hint;                          // 140
            } else {
                if (c.has_stored && (!c.has_scaler || c.scaler_out != target)) MXL_TRY(flush_scales());
                MXL_TRY(rescale(c, target));
            }
        }
        // compose output frame (150-239)
        mxl_frame* outf = frame_alloc(ctx, target.w, target.h);
        if (!outf) return MXL_ERR_OOM;
        const Channel* ca = (p.a >= 0 && p.a < MXL_VIDEO_MIXER_CHANNELS && ch[p.a].has_stored) ? &ch[p.a] : nullptr;
        const Channel* cb = (p.b >= 0 && p.b < MXL_VIDEO_MIXER_CHANNELS && ch[p.b].has_stored) ? &ch[p.b] : nullptr;
        k::FadeJob job{};
        job.fade = mxl_fader_to_u8(p.fader);                                                       // 168
        // a layer whose weight is 0 (fader at an end stop) is handed to the kernel as missing and so never read:
        // (a*255 + b*0) / 255 == a exactly, whatever stands in for b
        if (job.fade == 0) ca = nullptr;
        if (job.fade == 255) cb = nullptr;
        job.a = ca ? ca->frame->dev : nullptr;
        job.b = cb ? cb->frame->dev : nullptr;
        if (ca) keepalive.push_back(frame_retain(ca->frame));
        if (cb) keepalive.push_back(frame_retain(cb->frame));
        job.out = outf->dev;
        jobs.push_back(job);
        job_layouts.push_back(outf->layout);
        if (bytes) *bytes += (ca ? outf->layout.size : 0) + (cb ? outf->layout.size : 0) + outf->layout.size;
        video_slot_set(io.out[0]->slots[kk], outf, tick_duration, Rational());                     // 241-247
        frame_release(outf);                                // the line holds the reference now
    }

    // the scaler first (its outputs are crossfade inputs): one launch per geometry
    MXL_TRY(flush_scales());
    // one batched launch per run of equal layouts (normally exactly one)
    if (!jobs.empty()) {
        // a call of a few ticks (the live engine thread: one) carries its job table in the kernel parameters
        const bool inline_jobs = jobs.size() <= (size_t)k::kFadeInlineJobs;
        if (!inline_jobs) {
            MXL_TRY(this->jobs.ensure(ctx, jobs.size() * sizeof(k::FadeJob)));
            MXL_CUDA(cudaMemcpyAsync(this->jobs.p, jobs.data(), jobs.size() * sizeof(k::FadeJob), cudaMemcpyHostToDevice, ctx->stream));
        }
        size_t begin = 0;
        while (begin < jobs.size()) {
            size_t end = begin + 1;
            while (end < jobs.size() && job_layouts[end].width == job_layouts[begin].width && job_layouts[end].height == job_layouts[begin].height) end++;
            if (inline_jobs) MXL_TRY(k::launch_crossfade_inline(ctx, job_layouts[begin], jobs.data() + begin, (uint32_t)(end - begin)));
            else MXL_TRY(k::launch_crossfade(ctx, job_layouts[begin], (const k::FadeJob*)this->jobs.p + begin, (uint32_t)(end - begin)));
            begin = end;
        }
    }
    return MXL_OK;
}

// ================================================================================================
// dispatch
// ================================================================================================
int run_batch(mxl_ctx* ctx, int kind, mxl_module* const* mods, int n, uint64_t t, const IoSet* io, uint64_t* bytes)
{
    if (!ctx || !ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context; there is no CPU fallback");
    if (bytes) *bytes = 0;
    switch (kind) {
    case MXL_MOD_OSCILLATOR: return run_oscillators(ctx, mods, n, t, io, bytes);
    case MXL_MOD_FM_SINE: return run_fm_sines(ctx, mods, n, t, io, bytes);
    case MXL_MOD_MIXER: return run_mixers(ctx, mods, n, io, bytes);
    case MXL_MOD_AMPLIFIER: return run_amplifiers(ctx, mods, n, io, bytes);
    case MXL_MOD_STEREO_PANNER: return run_panners(ctx, n, io, bytes);
    case MXL_MOD_STEREO_SPLITTER: return run_splitters(ctx, n, io, bytes);
    case MXL_MOD_TRIGGER: return run_triggers(ctx, mods, n, io, bytes);
    case MXL_MOD_EQ_THREE: return run_eq_threes(ctx, mods, n, io, bytes);
    case MXL_MOD_ENVELOPE: return run_envelopes(ctx, mods, n, t, io, bytes);
    case MXL_MOD_METER: return run_meters(ctx, mods, n, io, bytes);
    case MXL_MOD_PLOTTER: return run_plotters(ctx, mods, n, io, bytes);
    case MXL_MOD_PCM_SINK: return run_pcm_sinks(ctx, mods, n, io, bytes);
    case MXL_MOD_OUTPUT_DEVICE:
        for (int i = 0; i < n; i++) {
            uint64_t b = 0;
            MXL_TRY(((OutputDevice*)mods[i])->run(io[i], &b));
            if (bytes) *bytes += b;
        }
        return MXL_OK;
    case MXL_MOD_MONITOR: case MXL_MOD_STREAM_OUTPUT:
        for (int i = 0; i < n; i++) {
            uint64_t b = 0;
            MXL_TRY(((Monitor*)mods[i])->run(t, io[i], &b));
            if (bytes) *bytes += b;
        }
        return MXL_OK;
    case MXL_MOD_STREAM_INPUT:
        for (int i = 0; i < n; i++) {
            uint64_t b = 0;
            MXL_TRY(((StreamInput*)mods[i])->run(t, io[i], &b));
            if (bytes) *bytes += b;
        }
        return MXL_OK;
    case MXL_MOD_VIDEO_MIXER:
        for (int i = 0; i < n; i++) {
            uint64_t b = 0;
            MXL_TRY(((VideoMixer*)mods[i])->run(t, io[i], &b));
            if (bytes) *bytes += b;
        }
        return MXL_OK;
    case MXL_MOD_SOURCE_MONO: case MXL_MOD_SOURCE_STEREO: case MXL_MOD_SOURCE_VIDEO:
        return MXL_OK;       // the source's line IS its output
    default:
        MXL_FAIL(MXL_ERR_UNSUPPORTED, "run_tick: module kind %d has no device implementation", kind);
    }
}

// ================================================================================================
// accessors
// ================================================================================================
int eq_three_state(mxl_module* m, double state[11])
{
    if (!m || m->kind != MXL_MOD_EQ_THREE || !state) MXL_FAIL(MXL_ERR_PARAMS, "not an EqThree module");
    EqThree* e = (EqThree*)m;
    MXL_TRY(e->ensure_state());
    MXL_CUDA(cudaMemcpyAsync(state, e->state_ptr(e->cur), 11 * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
    MXL_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return MXL_OK;
}

int envelope_state(mxl_module* m, int32_t* state, uint64_t* seq, double* off_amplitude)
{
    if (!m || m->kind != MXL_MOD_ENVELOPE) MXL_FAIL(MXL_ERR_PARAMS, "not an Envelope module");
    Envelope* e = (Envelope*)m;
    MXL_TRY(e->ensure_state());
    k::EnvState s;
    MXL_CUDA(cudaMemcpyAsync(&s, (const k::EnvState*)e->state.p + e->cur, sizeof s, cudaMemcpyDeviceToHost, m->ctx->stream));
    MXL_CUDA(cudaStreamSynchronize(m->ctx->stream));
    if (state) *state = s.state;
    if (seq) *seq = s.seq;
    if (off_amplitude) *off_amplitude = s.off_amplitude;
    return MXL_OK;
}

int meter_read(mxl_module* m, uint32_t slot, float peak[2], double sumsq[2], int32_t* clip)
{
    if (!m || m->kind != MXL_MOD_METER) MXL_FAIL(MXL_ERR_PARAMS, "not a Meter module");
    Meter* me = (Meter*)m;
    if (slot >= me->n_slots) MXL_FAIL(MXL_ERR_LENGTH, "meter slot %u out of %u", slot, me->n_slots);
    k::MeterRecord r;
    MXL_TRY(m->ctx->activate());
    MXL_CUDA(cudaMemcpyAsync(&r, (k::MeterRecord*)me->records.p + slot, sizeof r, cudaMemcpyDeviceToHost, m->ctx->stream));
    MXL_CUDA(cudaStreamSynchronize(m->ctx->stream));
    if (peak) { peak[0] = r.peak[0]; peak[1] = r.peak[1]; }
    if (sumsq) { sumsq[0] = r.sumsq[0]; sumsq[1] = r.sumsq[1]; }
    if (clip) *clip = r.clip;
    return MXL_OK;
}

int meter_download_async(mxl_module* m, mxl_meter_record* records, uint32_t cap)
{
    if (!m || m->kind != MXL_MOD_METER) MXL_FAIL(MXL_ERR_PARAMS, "not a Meter module");
    Meter* me = (Meter*)m;
    const uint32_t n = std::min(cap, me->n_slots);
    if (n == 0) return 0;
    if (!records) MXL_FAIL(MXL_ERR_INVALID, "NULL records");
    MXL_TRY(m->ctx->activate());
    cudaStream_t st;
    MXL_TRY(m->ctx->download_stream(&st));
    MXL_CUDA(cudaMemcpyAsync(records, me->records.p, (size_t)n * sizeof(k::MeterRecord), cudaMemcpyDeviceToHost, st));
    m->ctx->d2h_bytes += (size_t)n * sizeof(k::MeterRecord);
    return (int)n;
}

int meter_download(mxl_module* m, mxl_meter_record* records, uint32_t cap)
{
    static_assert(sizeof(mxl_meter_record) == sizeof(k::MeterRecord), "ABI record mirrors the kernel's");
    if (!m || m->kind != MXL_MOD_METER) MXL_FAIL(MXL_ERR_PARAMS, "not a Meter module");
    Meter* me = (Meter*)m;
    const uint32_t n = std::min(cap, me->n_slots);
    if (n == 0) return 0;
    if (!records) MXL_FAIL(MXL_ERR_INVALID, "NULL records");
    const int got = meter_download_async(m, records, cap);
    if (got < 0) return got;
    MXL_TRY(mxl_ctx_synchronize(m->ctx));
    return (int)n;
}

int plotter_read(mxl_module* m, float* left, float* right, uint32_t cap)
{
    if (!m || m->kind != MXL_MOD_PLOTTER) MXL_FAIL(MXL_ERR_PARAMS, "not a Plotter module");
    Plotter* pl = (Plotter*)m;
    const uint32_t nfr = std::min(cap, pl->tap_frames);
    if (nfr == 0) return 0;
    if (m->ctx->activate() != MXL_OK) return MXL_ERR_CUDA;
    if (cudaMemcpyAsync(left, pl->tap.p, nfr * sizeof(float), cudaMemcpyDeviceToHost, m->ctx->stream) != cudaSuccess ||
        cudaMemcpyAsync(right, (float*)pl->tap.p + m->ctx->spt, nfr * sizeof(float), cudaMemcpyDeviceToHost, m->ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(m->ctx->stream) != cudaSuccess)
        MXL_FAIL(MXL_ERR_CUDA, "plotter read-back failed");
    return (int)nfr;
}

int source_set_line(mxl_module* m, mxl_line* line)
{
    if (!m || (m->kind != MXL_MOD_SOURCE_MONO && m->kind != MXL_MOD_SOURCE_STEREO && m->kind != MXL_MOD_SOURCE_VIDEO))
        MXL_FAIL(MXL_ERR_PARAMS, "not a Source module");
    if (line && line->type != m->outputs[0].type) MXL_FAIL(MXL_ERR_LINE_TYPE, "source line type does not match the module's output");
    ((Source*)m)->line = line;
    return MXL_OK;
}

mxl_line* source_line(mxl_module* m)
{
    if (!m || (m->kind != MXL_MOD_SOURCE_MONO && m->kind != MXL_MOD_SOURCE_STEREO && m->kind != MXL_MOD_SOURCE_VIDEO)) return nullptr;
    return ((Source*)m)->line;
}

int pcm_sink_download(mxl_module* m, int16_t* host, uint64_t n)
{
    if (!m || m->kind != MXL_MOD_PCM_SINK) MXL_FAIL(MXL_ERR_PARAMS, "not a PcmSink module");
    PcmSink* s = (PcmSink*)m;
    if (n > s->n_samples) MXL_FAIL(MXL_ERR_LENGTH, "%llu samples requested, sink holds %llu", (unsigned long long)n, (unsigned long long)s->n_samples);
    if (n == 0) return MXL_OK;
    MXL_TRY(m->ctx->activate());
    MXL_CUDA(cudaMemcpyAsync(host, s->pcm.p, n * sizeof(int16_t), cudaMemcpyDeviceToHost, m->ctx->stream));
    MXL_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return MXL_OK;
}

int mixer_params_get(const mxl_module* m, mxl_mixer_channel_params* out, uint32_t cap)
{
    if (!m || m->kind != MXL_MOD_MIXER) MXL_FAIL(MXL_ERR_PARAMS, "not a Mixer module");
    const Mixer* mx = (const Mixer*)m;
    const uint32_t nch = (uint32_t)mx->channels.size();
    for (uint32_t i = 0; i < nch && i < cap && out; i++) out[i] = mx->channels[i];
    return (int)nch;
}

}  // namespace mxl

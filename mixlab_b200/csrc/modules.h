// modules.h -- host side of the modules: the C++ mirror of `trait ModuleT` (src/module/mod.rs:7-19)
// and of ModuleHost (src/engine/module.rs:62-119).  Arithmetic lives in the kernels; these classes
// hold params, terminals and device-resident state, validate lines the way InputRef/OutputRef
// expect_* do, and turn a set of same-kind modules into one batched launch.
#pragma once

#include <string>
#include <vector>

#include "common.h"
#include "kernels.h"

namespace mxl {

struct Terminal {            // protocol/src/lib.rs:160-174  Terminal(Option<String>, LineType)
    int type;
    bool labeled;
    std::string label;
};

inline Terminal labeled(int type, const std::string& l) { return Terminal{type, true, l}; }
inline Terminal unlabeled(int type) { return Terminal{type, false, ""}; }

// inputs / outputs of one module for one call; in[i] == nullptr is InputRef::Disconnected
struct IoSet {
    const mxl_line* const* in;
    uint32_t n_in;
    mxl_line* const* out;
    uint32_t n_out;
};

}  // namespace mxl

struct mxl_module {
    mxl_ctx* ctx = nullptr;
    int kind = -1;
    std::vector<mxl::Terminal> inputs, outputs;
    // device lines behind the host slices of mxl_module_run_tick_host, one per terminal, kept across ticks
    std::vector<mxl_line*> host_in, host_out;

    virtual ~mxl_module();
    virtual int update(const void* params) = 0;          // ModuleT::update
    virtual int get_params(void* out) const = 0;         // ModuleT::params
    const char* kind_name() const;
};

namespace mxl {

mxl_module* module_create(mxl_ctx* ctx, int kind, const void* params);

// Runs `n` modules of one kind for the samples starting at absolute index t (ModuleT::run_tick for
// each of them, engine.rs:490-494), batching same-length instances into single launches.
// `bytes_out`, if given, receives the API-level line bytes read + written.
int run_batch(mxl_ctx* ctx, int kind, mxl_module* const* mods, int n, uint64_t t, const IoSet* io,
              uint64_t* bytes_out);

// A sub-graph Oscillator -> EqThree -> StereoPanner -> Mixer [-> Meter] that the graph executor runs as ONE launch
// (fused_voice.cu) instead of five stages.  Built by graph.cu's plan from the resolved connections.
struct FusedVoiceRef {
    mxl_module* osc;              // nullptr = the EqThree's input is disconnected (zeros)
    mxl_module* eq;
    mxl_line* eq_out;             // always written: it carries the voice from the filter to the mixer sum
    mxl_line* osc_mono;           // nullptr = nobody observes the line: it is not written
    mxl_line* osc_stereo;
};
struct FusedChanRef {
    bool connected;               // false = the mixer input is disconnected (no panner)
    int left, right;              // voice index feeding the panner's L / R, -1 = disconnected
    mxl_line* pan_out;            // nullptr = not observed, not written
};
struct FusedGroup {
    std::vector<FusedVoiceRef> voices;
    std::vector<FusedChanRef> chans;        // one per mixer channel, in channel order
    mxl_module* mixer = nullptr;
    mxl_line* master = nullptr;
    mxl_line* cue = nullptr;
    mxl_module* meter = nullptr;            // Meter on the master bus folded into the launch, or nullptr
    std::vector<int> members;               // ModuleIds of every module of the group (graph.cu's bookkeeping)
};
// Can this group run fused on this context (a cluster of that many CTAs fits, a plan exists at this sample rate)?
bool fused_group_supported(mxl_ctx* ctx, const FusedGroup& g);
// Do the modules' CURRENT parameters allow it (checked before every run: update() may have changed them)?
bool fused_group_params_ok(const FusedGroup& g);
// Can a Meter on the master bus be folded in (a time tile can be made a whole number of ticks)?
bool fused_meter_supported(mxl_ctx* ctx);
int run_fused_group(mxl_ctx* ctx, const FusedGroup& g, uint64_t t, uint64_t* bytes_out);

// kind-specific accessors used by the ABI
int eq_three_state(mxl_module* m, double state[11]);
int envelope_state(mxl_module* m, int32_t* state, uint64_t* seq, double* off_amplitude);
int meter_read(mxl_module* m, uint32_t slot, float peak[2], double sumsq[2], int32_t* clip);
int meter_download(mxl_module* m, mxl_meter_record* records, uint32_t cap);
int meter_download_async(mxl_module* m, mxl_meter_record* records, uint32_t cap);
int plotter_read(mxl_module* m, float* left, float* right, uint32_t cap);
int source_set_line(mxl_module* m, mxl_line* line);
mxl_line* source_line(mxl_module* m);
int pcm_sink_download(mxl_module* m, int16_t* host, uint64_t n);
int stream_input_write_audio(mxl_module* m, uint64_t source_id, Rational time, const int16_t* samples, uint64_t n);
int stream_input_write_video(mxl_module* m, uint64_t source_id, Rational time, mxl_frame* frame, Rational duration);
int stream_input_pending(const mxl_module* m, uint32_t* audio_frames, uint32_t* video_frames);
int monitor_recv_audio(mxl_module* m, mxl_audio_fragment* info, int16_t* pcm, uint32_t cap);
int monitor_recv_video(mxl_module* m, mxl_video_job* out);
int stream_output_set_live(mxl_module* m, int live);
int64_t output_device_read(mxl_module* m, float* out, uint64_t cap);
int output_device_clip(mxl_module* m, int32_t* clip);
int mixer_params_get(const mxl_module* m, mxl_mixer_channel_params* out, uint32_t cap);

mxl_frame* frame_scale(mxl_frame* src, uint32_t out_w, uint32_t out_h);
mxl_frame* scale_target(mxl_ctx* ctx, const mxl_frame_layout& sl, uint32_t out_w, uint32_t out_h);
int frames_scale(mxl_ctx* ctx, mxl_frame* const* src, mxl_frame** dst, uint32_t n, uint32_t out_w, uint32_t out_h);

}  // namespace mxl

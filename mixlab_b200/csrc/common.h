// common.h -- internal types of libmixlab_b200 (context, lines, frames, error plumbing).
#pragma once

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <utility>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mixlab_b200.h"
#include "eq_plan.h"

namespace mxl {

// ---- errors: never unwind across the C ABI ------------------------------------------------------
void set_error(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
const char* last_error();

#define MXL_FAIL(code, ...)            \
    do {                               \
        ::mxl::set_error(__VA_ARGS__); \
        return (code);                 \
    } while (0)

#define MXL_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            ::mxl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                             __LINE__);                                                         \
            return MXL_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define MXL_TRY(expr)              \
    do {                           \
        int _s = (expr);           \
        if (_s != MXL_OK) return _s; \
    } while (0)

// Rational64 as used by MediaTime / MediaDuration (util/src/time.rs:9-10,77-78): always reduced,
// denominator positive.
struct Rational {
    int64_t num = 0, den = 1;
    static Rational make(int64_t n, int64_t d);
    Rational operator+(const Rational& o) const;
    Rational operator-(const Rational& o) const;
    bool operator>=(const Rational& o) const;
    bool operator>(const Rational& o) const;
    bool operator==(const Rational& o) const { return num == o.num && den == o.den; }
};

}  // namespace mxl

struct mxl_ctx {
    int device = MXL_DEVICE_NONE;
    uint32_t sample_rate = 0;
    uint32_t spt = 0;                 // SAMPLES_PER_TICK
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t launches = 0;
    uint64_t fused_mix_launch_no = 0; // value of `launches` right after the last fused_mix_kernel launch ...
    const void* fused_mix_group = nullptr;   // ... and the group it belonged to (fused_voice_kernel's late dependency wait)
    const void* fused_group_now = nullptr;   // group whose launches are being issued (set by run_fused_group)
    uint64_t change_epoch = 1;        // bumped by whatever may invalidate cached launch parameters: a module update, a line
                                      // (re)allocation, a new graph plan
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    int sm_count = 0;

    // free yuv420p frame buffers by byte size: frames come and go every tick (AvFrame::blank per
    // tick in the reference, video_mixer.rs:151), cudaMalloc/cudaFree must not.
    std::unordered_map<uint64_t, std::vector<uint8_t*>> frame_pool;
    uint64_t frame_pool_bytes = 0;                    // bytes parked in frame_pool
    uint64_t frame_pool_cap = 8ull << 30;             // beyond this a released frame goes back to the driver (mxl_ctx_trim_frame_pool)
    // EqThree chunk plans by chunk length (see modules.cu: eq_plan_for)
    std::map<uint32_t, std::vector<double>> eq_plans;
    // eq_stream_kernel plans by chunk length (eq_plan.h); ok == false = unusable at this sample rate
    std::map<uint32_t, mxl::EqStreamPlan> eq_stream_plans;
    uint32_t eq_stream_smem_set = 0;  // bit LC/16: opt-in shared memory size configured
    uint32_t scale_smem[5] = {0, 0, 0, 0, 0};   // dynamic shared memory the scale_tiled_kernel variants have been configured for
    // scaler tap tables by (source length, destination length): device [pos int32 x n][coef int16 x 4n]
    std::map<uint64_t, void*> scale_tables;
    std::map<uint64_t, std::vector<int32_t>> scale_positions;   // host copies of the first-tap columns (tile bounds)
    void* scale_jobs = nullptr;       // device staging for ScaleJob arrays
    size_t scale_jobs_cap = 0, scale_jobs_used = 0;   // bytes; tables rotate through the buffer
    void* pcm_ring = nullptr;         // device staging ring of the asynchronous i16 converters (abi.cu)
    size_t pcm_ring_cap = 0, pcm_ring_used = 0;
    void* comm = nullptr;             // ncclComm_t of the optional shared-source mode (comm.cu)
    int comm_rank = 0, comm_world = 0;
    bool kernel_timing = false;
    bool pdl_hold = false;            // set by the graph executor while audio stages share the GPU with a batched compositor
    struct KernelEvents { const char* name; cudaEvent_t a, b; };
    std::vector<KernelEvents> kernel_events;      // pairs recorded since the last read
    std::vector<cudaEvent_t> kernel_event_pool;
    bool env_carveout_set = false;    // envelope_kernel's shared-memory carve-out preference has been set
    uint32_t env_epoch = 0;           // Envelope launches so far (tags the look-back flags of a launch)
    std::map<uint32_t, void*> eq_stream_tables;   // device copies of EqStreamPlan::lane_pow by chunk length
    unsigned long long* fused_prof = nullptr;     // diagnostics: per-CTA phase clocks of the last fused voice launch (mxl_ctx_fused_profile)
    uint32_t fused_prof_ctas = 0, fused_prof_cap = 0;
    std::map<uint32_t, int> fused_clusters;       // (chunk << 8 | voices) -> resident clusters of the fused voice kernel (0 = cannot launch)

    // Copy/compute overlap (mxl_ctx_set_copy_overlap): async uploads go to stream_in, async downloads
    // to stream_out, ordered against the compute stream with events:
    //   upload   waits for the last compute            (WAR on input lines/frames)
    //   compute  waits for the uploads and downloads enqueued so far (RAW on inputs, WAR on recycled outputs)
    //   download waits for the last compute            (RAW on outputs)
    bool overlap = false;
    cudaStream_t stream_in = nullptr, stream_out = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr, ev_compute = nullptr;
    cudaEvent_t fences[4] = {nullptr, nullptr, nullptr, nullptr};
    bool in_dirty = false, out_dirty = false, in_must_wait = false, out_must_wait = false;
    uint64_t h2d_bytes = 0, d2h_bytes = 0;
    // Audio and video sub-graphs never share a line (no module of the path has terminals of both
    // kinds), so the graph executor runs the small latency-bound audio kernels on a second,
    // higher-priority stream while the bandwidth-bound compositor owns the main stream.
    cudaStream_t stream_aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

    int upload_stream(cudaStream_t* s);     // stream an async upload must be issued on
    int download_stream(cudaStream_t* s);
    int compute_begin();                    // called before kernels of a run are enqueued
    int compute_end();

    bool has_device() const { return device >= 0; }
    int activate() const;            // cudaSetDevice
};

// One device allocation shared by a batch of frames (mxl_frames_alloc_batch): consecutive frames are
// adjacent in memory, so a whole batch moves over PCIe as one copy.  Freed with its last frame.
struct FrameSlab {
    uint8_t* base = nullptr;
    std::atomic<int> refs{0};
};

struct mxl_frame {
    mxl_ctx* ctx = nullptr;
    FrameSlab* slab = nullptr;        // non-null: `dev` points into a slab and never enters the frame pool
    mxl_frame_layout layout{};
    uint8_t* dev = nullptr;
    std::atomic<int> refs{1};
};

namespace mxl {
// Optional per-kernel timing (mxl_ctx_set_kernel_timing): every launcher records a CUDA event pair
// directly around its kernel launch on the launching stream -- host preparation, table copies and the
// work of other stages stay outside the pair.  bench.py's roofline numbers come from here.
struct KernelTimer {
    mxl_ctx* ctx;
    int slot = -1;
    KernelTimer(mxl_ctx* c, const char* name);
    ~KernelTimer();
};
}  // namespace mxl
#define MXL_TIMED(ctx, name) ::mxl::KernelTimer mxl_kernel_timer_((ctx), (name))

// Programmatic dependent launch for the chains of small audio kernels (Oscillator -> EqThree -> Panner -> Mixer ->
// Meter): a kernel launched with the attribute may be scheduled while its predecessor on the stream still runs; it
// parks in griddepcontrol.wait (pdl_prologue, the first thing every such kernel does) until the predecessor has
// completed and flushed, so only the launch latency and CTA scheduling overlap -- no kernel reads or writes early.
// Off while per-kernel event timing is on (the event records sit between the launches), with MXL_NO_PDL set, and
// while the audio stages of a many-tick call run beside the compositor on the other stream: kernels parked in
// griddepcontrol.wait hold registers and shared memory the bandwidth-bound crossfade wants (measured: 664 k ->
// 616 k ticks/s at 128 ticks per call with it on; audio alone 3.36 M -> 4.05 M, a live one-tick call 24.8 -> 17.2 us).
#ifdef __CUDACC__
namespace mxl {
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue()
{
    pdl_wait();
    pdl_launch_dependents();
}

bool pdl_enabled(const mxl_ctx* ctx);

template <typename... KArgs, typename... Args>
cudaError_t launch_chained(const mxl_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled(ctx) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
}  // namespace mxl
#endif

struct VideoSlot {
    mxl_frame* frame = nullptr;       // retained
    mxl::Rational duration_hint;      // video::Frame.duration_hint (src/video.rs)
    mxl::Rational tick_offset;        // engine::VideoFrame.tick_offset (io.rs:11-17)
};

struct mxl_line {
    mxl_ctx* ctx = nullptr;
    int type = MXL_LINE_MONO;
    uint64_t frames = 0;              // audio: samples per channel; video: tick slots
    uint64_t capacity = 0;            // audio: frames the allocation can hold
    float* dev = nullptr;             // audio payload
    std::vector<VideoSlot> slots;     // video payload

    uint64_t len() const { return type == MXL_LINE_STEREO ? frames * 2 : (type == MXL_LINE_MONO ? frames : 0); }
};

namespace mxl {

// Checked accessors mirroring InputRef::expect_* / OutputRef::expect_* (io.rs:36-61,100-126).
// `line == nullptr` is InputRef::Disconnected.
int expect_input(const mxl_line* line, int type, const char* what);
int expect_output(const mxl_line* line, int type, const char* what);

mxl_line* line_alloc(mxl_ctx* ctx, int type, uint64_t frames);
void line_free(mxl_line* line);
int line_resize(mxl_line* line, uint64_t frames);   // reallocates (contents undefined) if larger

mxl_frame* frame_alloc(mxl_ctx* ctx, uint32_t w, uint32_t h);
void frame_release(mxl_frame* f);
inline mxl_frame* frame_retain(mxl_frame* f) { if (f) f->refs.fetch_add(1); return f; }
void video_slot_set(VideoSlot& s, mxl_frame* f, Rational dur, Rational off);

void frame_layout_yuv420p(uint32_t w, uint32_t h, mxl_frame_layout* out);

}  // namespace mxl

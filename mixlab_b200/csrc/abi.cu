// abi.cu -- C ABI of the module layer: `trait ModuleT` (src/module/mod.rs:7-19) as seen through
// DynModuleHostT (src/engine/module.rs:88-119), plus the stand-alone converters.
#include <math.h>

#include <algorithm>

#include "modules.h"
#include "kernels.h"

using namespace mxl;

namespace {

bool kind_has_params(int kind)
{
    switch (kind) {
    case MXL_MOD_AMPLIFIER: case MXL_MOD_ENVELOPE: case MXL_MOD_EQ_THREE: case MXL_MOD_FM_SINE:
    case MXL_MOD_MIXER: case MXL_MOD_OSCILLATOR: case MXL_MOD_TRIGGER: case MXL_MOD_VIDEO_MIXER:
        return true;
    default:
        return false;   // `type Params = ()`
    }
}

// transient device staging for the stand-alone PCM converters
struct Staging {
    void* p = nullptr;
    ~Staging() { if (p) cudaFree(p); }
};

}  // namespace

extern "C" {

double mxl_db_to_linear(double db) { return pow(10.0, db / 20.0); }   // protocol/src/lib.rs:469-471

mxl_module* mxl_module_create(mxl_ctx* ctx, int kind, const void* params) { return module_create(ctx, kind, params); }

void mxl_module_destroy(mxl_module* m)
{
    if (!m) return;
    if (m->ctx && m->ctx->has_device()) { m->ctx->activate(); cudaStreamSynchronize(m->ctx->stream); }
    delete m;
}

int mxl_module_kind_of(const mxl_module* m) { return m ? m->kind : MXL_ERR_INVALID; }

int mxl_module_update(mxl_module* m, int kind, const void* params)
{
    if (!m) MXL_FAIL(MXL_ERR_INVALID, "NULL module");
    // module.rs:104-110: a ModuleParams variant of another module panics
    if (kind != m->kind) MXL_FAIL(MXL_ERR_PARAMS, "module params mismatch! module = %s, params kind = %d", m->kind_name(), kind);
    if (kind_has_params(kind) && !params) MXL_FAIL(MXL_ERR_PARAMS, "%s: NULL params", m->kind_name());
    if (m->ctx) m->ctx->change_epoch++;
    return m->update(params);
}

int mxl_module_params(const mxl_module* m, void* params_out)
{
    if (!m) MXL_FAIL(MXL_ERR_INVALID, "NULL module");
    if (kind_has_params(m->kind) && !params_out) MXL_FAIL(MXL_ERR_INVALID, "NULL params_out");
    return m->get_params(params_out);
}

int mxl_mixer_params_get(const mxl_module* m, mxl_mixer_channel_params* channels_out, uint32_t cap)
{
    return mixer_params_get(m, channels_out, cap);
}

uint32_t mxl_module_n_inputs(const mxl_module* m) { return m ? (uint32_t)m->inputs.size() : 0; }
uint32_t mxl_module_n_outputs(const mxl_module* m) { return m ? (uint32_t)m->outputs.size() : 0; }

int mxl_module_input_type(const mxl_module* m, uint32_t index)
{
    if (!m || index >= m->inputs.size()) MXL_FAIL(MXL_ERR_INVALID, "no input %u", index);
    return m->inputs[index].type;
}

int mxl_module_output_type(const mxl_module* m, uint32_t index)
{
    if (!m || index >= m->outputs.size()) MXL_FAIL(MXL_ERR_INVALID, "no output %u", index);
    return m->outputs[index].type;
}

const char* mxl_module_input_label(const mxl_module* m, uint32_t index)
{
    if (!m || index >= m->inputs.size() || !m->inputs[index].labeled) return nullptr;
    return m->inputs[index].label.c_str();
}

const char* mxl_module_output_label(const mxl_module* m, uint32_t index)
{
    if (!m || index >= m->outputs.size() || !m->outputs[index].labeled) return nullptr;
    return m->outputs[index].label.c_str();
}

int mxl_module_run_tick(mxl_module* m, uint64_t t, const mxl_line* const* inputs, uint32_t n_inputs,
                        mxl_line* const* outputs, uint32_t n_outputs)
{
    if (!m) MXL_FAIL(MXL_ERR_INVALID, "NULL module");
    if ((n_inputs && !inputs) || (n_outputs && !outputs)) MXL_FAIL(MXL_ERR_INVALID, "NULL line table");
    for (uint32_t i = 0; i < n_inputs; i++)
        if (inputs[i] && inputs[i]->ctx != m->ctx) MXL_FAIL(MXL_ERR_INVALID, "input %u belongs to another context", i);
    for (uint32_t i = 0; i < n_outputs; i++)
        if (outputs[i] && outputs[i]->ctx != m->ctx) MXL_FAIL(MXL_ERR_INVALID, "output %u belongs to another context", i);
    IoSet io{inputs, n_inputs, outputs, n_outputs};
    mxl_module* one = m;
    if (m->ctx->has_device()) MXL_TRY(m->ctx->activate());
    MXL_TRY(m->ctx->compute_begin());
    const int rc = run_batch(m->ctx, m->kind, &one, 1, t, &io, nullptr);
    MXL_TRY(m->ctx->compute_end());
    return rc;
}

// The line cached behind terminal `idx` of a host-slice call, at `frames` frames (audio) or one tick slot (video).
static int host_ref_line(mxl_module* m, std::vector<mxl_line*>& cache, uint32_t idx, int type, uint64_t frames, mxl_line** out)
{
    if (cache.size() <= idx) cache.resize(idx + 1, nullptr);
    mxl_line*& l = cache[idx];
    if (l && l->type != type) { line_free(l); l = nullptr; }      // Mixer::update re-creates its terminals (mixer.rs:40-44)
    if (!l) {
        l = line_alloc(m->ctx, type, frames);
        if (!l) return MXL_ERR_OOM;
    } else if (type == MXL_LINE_VIDEO ? l->slots.size() != frames : l->frames != frames) {
        MXL_TRY(line_resize(l, frames));
    }
    *out = l;
    return MXL_OK;
}

int mxl_module_run_tick_host(mxl_module* m, uint64_t t, const mxl_host_ref* inputs, uint32_t n_inputs,
                             mxl_host_ref* outputs, uint32_t n_outputs)
{
    if (!m) MXL_FAIL(MXL_ERR_INVALID, "NULL module");
    if ((n_inputs && !inputs) || (n_outputs && !outputs)) MXL_FAIL(MXL_ERR_INVALID, "NULL terminal table");
    mxl_ctx* ctx = m->ctx;
    if (!ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "mxl_module_run_tick_host: context has no CUDA device; there is no CPU fallback");
    if (n_inputs != m->inputs.size() || n_outputs != m->outputs.size())
        MXL_FAIL(MXL_ERR_INVALID, "%s: expected %zu inputs / %zu outputs, got %u / %u", m->kind_name(), m->inputs.size(), m->outputs.size(), n_inputs, n_outputs);
    MXL_TRY(ctx->activate());
    // the slices are ordinary host memory and the call is synchronous: everything goes on the compute stream
    // (with copy overlap enabled, first wait for what the side streams still carry)
    if (ctx->overlap) MXL_TRY(mxl_ctx_synchronize(ctx));
    std::vector<const mxl_line*> in(n_inputs, nullptr);
    std::vector<mxl_line*> out(n_outputs, nullptr);
    for (uint32_t i = 0; i < n_inputs; i++) {
        const mxl_host_ref& r = inputs[i];
        if (!r.connected) continue;                                     // InputRef::Disconnected
        const int type = m->inputs[i].type;
        if (r.type != type) MXL_FAIL(MXL_ERR_TYPE_MISMATCH, "%s input %u: expected line type %d, got %d", m->kind_name(), i, type, r.type);
        mxl_line* l = nullptr;
        if (type == MXL_LINE_VIDEO) {
            MXL_TRY(host_ref_line(m, m->host_in, i, type, 1, &l));
            if (r.frame && r.frame->ctx != ctx) MXL_FAIL(MXL_ERR_INVALID, "input %u: frame belongs to another context", i);
            if (r.frame && (r.duration_den == 0 || r.offset_den == 0)) MXL_FAIL(MXL_ERR_INVALID, "input %u: zero denominator", i);
            video_slot_set(l->slots[0], r.frame, r.frame ? Rational::make(r.duration_num, r.duration_den) : Rational(),
                           r.frame ? Rational::make(r.offset_num, r.offset_den) : Rational());
        } else {
            if (type == MXL_LINE_STEREO && (r.len & 1)) MXL_FAIL(MXL_ERR_LENGTH, "%s input %u: stereo slice of odd length %llu", m->kind_name(), i, (unsigned long long)r.len);
            if (r.len && !r.samples) MXL_FAIL(MXL_ERR_INVALID, "input %u: NULL slice", i);
            MXL_TRY(host_ref_line(m, m->host_in, i, type, type == MXL_LINE_STEREO ? r.len / 2 : r.len, &l));
            if (r.len) MXL_CUDA(cudaMemcpyAsync(l->dev, r.samples, r.len * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
            ctx->h2d_bytes += r.len * sizeof(float);
        }
        in[i] = l;
    }
    for (uint32_t i = 0; i < n_outputs; i++) {
        mxl_host_ref& r = outputs[i];
        const int type = m->outputs[i].type;
        if (r.type != type) MXL_FAIL(MXL_ERR_TYPE_MISMATCH, "%s output %u: expected line type %d, got %d", m->kind_name(), i, type, r.type);
        if (type == MXL_LINE_VIDEO) {
            MXL_TRY(host_ref_line(m, m->host_out, i, type, 1, &out[i]));
            video_slot_set(out[i]->slots[0], nullptr, Rational(), Rational());
        } else {
            if (type == MXL_LINE_STEREO && (r.len & 1)) MXL_FAIL(MXL_ERR_LENGTH, "%s output %u: stereo slice of odd length %llu", m->kind_name(), i, (unsigned long long)r.len);
            if (r.len && !r.samples) MXL_FAIL(MXL_ERR_INVALID, "output %u: NULL slice", i);
            MXL_TRY(host_ref_line(m, m->host_out, i, type, type == MXL_LINE_STEREO ? r.len / 2 : r.len, &out[i]));
        }
    }
    IoSet io{in.data(), n_inputs, out.data(), n_outputs};
    mxl_module* one = m;
    MXL_TRY(run_batch(ctx, m->kind, &one, 1, t, &io, nullptr));
    for (uint32_t i = 0; i < n_outputs; i++) {
        mxl_host_ref& r = outputs[i];
        if (m->outputs[i].type == MXL_LINE_VIDEO) {
            const VideoSlot& s = out[i]->slots[0];
            r.frame = frame_retain(s.frame);                            // NULL = None
            r.duration_num = s.duration_hint.num; r.duration_den = s.duration_hint.den;
            r.offset_num = s.tick_offset.num; r.offset_den = s.tick_offset.den;
        } else if (r.len) {
            MXL_CUDA(cudaMemcpyAsync(r.samples, out[i]->dev, r.len * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            ctx->d2h_bytes += r.len * sizeof(float);
        }
    }
    // input frames are referenced by the cached one-slot lines only for the duration of the call
    for (uint32_t i = 0; i < n_inputs; i++)
        if (in[i] && in[i]->type == MXL_LINE_VIDEO) video_slot_set(m->host_in[i]->slots[0], nullptr, Rational(), Rational());
    MXL_CUDA(cudaStreamSynchronize(ctx->stream));
    return MXL_OK;
}

int mxl_eq_three_state(mxl_module* m, double state[11]) { return eq_three_state(m, state); }
int mxl_envelope_state(mxl_module* m, int32_t* state, uint64_t* seq, double* off_amplitude) { return envelope_state(m, state, seq, off_amplitude); }
int mxl_meter_read(mxl_module* m, uint32_t slot, float peak[2], double sumsq[2], int32_t* clip) { return meter_read(m, slot, peak, sumsq, clip); }
int mxl_meter_download(mxl_module* m, mxl_meter_record* records, uint32_t cap) { return meter_download(m, records, cap); }
int mxl_meter_download_async(mxl_module* m, mxl_meter_record* records, uint32_t cap) { return meter_download_async(m, records, cap); }
int mxl_plotter_read(mxl_module* m, float* left, float* right, uint32_t cap_frames) { return plotter_read(m, left, right, cap_frames); }
int mxl_source_set_line(mxl_module* m, mxl_line* line) { return source_set_line(m, line); }
int mxl_pcm_sink_download(mxl_module* m, int16_t* host, uint64_t n_samples) { return pcm_sink_download(m, host, n_samples); }

int mxl_stream_input_write_audio(mxl_module* m, uint64_t source_id, int64_t time_num, int64_t time_den, const int16_t* samples, uint64_t n_samples)
{
    if (time_den == 0) MXL_FAIL(MXL_ERR_INVALID, "mxl_stream_input_write_audio: zero denominator");
    return stream_input_write_audio(m, source_id, Rational::make(time_num, time_den), samples, n_samples);
}

int mxl_stream_input_write_video(mxl_module* m, uint64_t source_id, int64_t time_num, int64_t time_den, mxl_frame* frame,
                                 int64_t duration_num, int64_t duration_den)
{
    if (time_den == 0 || duration_den == 0) MXL_FAIL(MXL_ERR_INVALID, "mxl_stream_input_write_video: zero denominator");
    return stream_input_write_video(m, source_id, Rational::make(time_num, time_den), frame, Rational::make(duration_num, duration_den));
}

int mxl_stream_input_pending(const mxl_module* m, uint32_t* audio_frames, uint32_t* video_frames) { return stream_input_pending(m, audio_frames, video_frames); }
int mxl_monitor_recv_audio(mxl_module* m, mxl_audio_fragment* info, int16_t* pcm, uint32_t cap_samples) { return monitor_recv_audio(m, info, pcm, cap_samples); }
int mxl_monitor_recv_video(mxl_module* m, mxl_video_job* out) { return monitor_recv_video(m, out); }
int mxl_stream_output_set_live(mxl_module* m, int live) { return stream_output_set_live(m, live); }
int64_t mxl_output_device_read(mxl_module* m, float* out, uint64_t cap_samples) { return output_device_read(m, out, cap_samples); }
int mxl_output_device_clip(mxl_module* m, int32_t* clip) { return output_device_clip(m, clip); }

// Device staging for i16 PCM on its way in or out: a ring owned by the context, so the asynchronous
// converters neither allocate nor synchronise per call.  A region is reused only after a wrap, which
// first waits for everything queued on the context.
static int pcm_staging(mxl_ctx* ctx, uint64_t n_samples, int16_t** out)
{
    const size_t need = ((size_t)n_samples * sizeof(int16_t) + 255) & ~(size_t)255;
    if (ctx->pcm_ring_cap < 2 * need || ctx->pcm_ring_used + need > ctx->pcm_ring_cap) {
        MXL_TRY(mxl_ctx_synchronize(ctx));
        if (ctx->pcm_ring_cap < 2 * need) {
            if (ctx->pcm_ring) cudaFree(ctx->pcm_ring);
            ctx->pcm_ring = nullptr;
            ctx->pcm_ring_cap = 0;
            const size_t cap = std::max<size_t>(1 << 20, 4 * need);
            MXL_CUDA(cudaMalloc(&ctx->pcm_ring, cap));
            ctx->pcm_ring_cap = cap;
        }
        ctx->pcm_ring_used = 0;
    }
    *out = (int16_t*)((uint8_t*)ctx->pcm_ring + ctx->pcm_ring_used);
    ctx->pcm_ring_used += need;
    return MXL_OK;
}

// stream_input.rs:110-112,167-173: the i16 PCM crosses the bus (2 B/sample), the divide runs on the device
int mxl_pcm_unpack_i16_async(mxl_ctx* ctx, const int16_t* host_pcm, uint64_t n_samples, mxl_line* dst)
{
    if (!ctx || !dst) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if (!ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "mxl_pcm_unpack_i16: context has no CUDA device; there is no CPU fallback");
    if (dst->type == MXL_LINE_VIDEO) MXL_FAIL(MXL_ERR_LINE_TYPE, "mxl_pcm_unpack_i16: video line");
    if (n_samples > dst->len()) MXL_FAIL(MXL_ERR_LENGTH, "%llu samples, line holds %llu", (unsigned long long)n_samples, (unsigned long long)dst->len());
    if (n_samples == 0) return MXL_OK;
    if (!host_pcm) MXL_FAIL(MXL_ERR_INVALID, "NULL pcm");
    MXL_TRY(ctx->activate());
    int16_t* staging;
    MXL_TRY(pcm_staging(ctx, n_samples, &staging));
    cudaStream_t up;
    MXL_TRY(ctx->upload_stream(&up));
    MXL_CUDA(cudaMemcpyAsync(staging, host_pcm, n_samples * sizeof(int16_t), cudaMemcpyHostToDevice, up));
    ctx->h2d_bytes += n_samples * sizeof(int16_t);
    MXL_TRY(ctx->compute_begin());
    const int st = k::launch_pcm_unpack(ctx, staging, dst->dev, n_samples);
    MXL_TRY(ctx->compute_end());
    return st;
}

int mxl_pcm_unpack_i16(mxl_ctx* ctx, const int16_t* host_pcm, uint64_t n_samples, mxl_line* dst)
{
    MXL_TRY(mxl_pcm_unpack_i16_async(ctx, host_pcm, n_samples, dst));
    return mxl_ctx_synchronize(ctx);
}

// src/video/encode.rs:184-195
int mxl_pcm_pack_i16_async(mxl_ctx* ctx, const mxl_line* src, int16_t* host_pcm, uint64_t n_samples)
{
    if (!ctx || !src) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if (!ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "mxl_pcm_pack_i16: context has no CUDA device; there is no CPU fallback");
    if (src->type == MXL_LINE_VIDEO) MXL_FAIL(MXL_ERR_LINE_TYPE, "mxl_pcm_pack_i16: video line");
    if (n_samples > src->len()) MXL_FAIL(MXL_ERR_LENGTH, "%llu samples, line holds %llu", (unsigned long long)n_samples, (unsigned long long)src->len());
    if (n_samples == 0) return MXL_OK;
    if (!host_pcm) MXL_FAIL(MXL_ERR_INVALID, "NULL pcm");
    MXL_TRY(ctx->activate());
    int16_t* staging;
    MXL_TRY(pcm_staging(ctx, n_samples, &staging));
    MXL_TRY(ctx->compute_begin());
    const int st = k::launch_pcm_pack(ctx, src->dev, staging, n_samples);
    MXL_TRY(ctx->compute_end());
    MXL_TRY(st);
    cudaStream_t down;
    MXL_TRY(ctx->download_stream(&down));
    MXL_CUDA(cudaMemcpyAsync(host_pcm, staging, n_samples * sizeof(int16_t), cudaMemcpyDeviceToHost, down));
    ctx->d2h_bytes += n_samples * sizeof(int16_t);
    return MXL_OK;
}

int mxl_pcm_pack_i16(mxl_ctx* ctx, const mxl_line* src, int16_t* host_pcm, uint64_t n_samples)
{
    MXL_TRY(mxl_pcm_pack_i16_async(ctx, src, host_pcm, n_samples));
    return mxl_ctx_synchronize(ctx);
}

int mxl_frame_to_rgba(const mxl_frame* frame, uint8_t* rgba_host)
{
    if (!frame || !rgba_host) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    mxl_ctx* ctx = frame->ctx;
    mxl_rgba* pic = mxl_rgba_alloc(ctx, frame->layout.width, frame->layout.height, 1);
    if (!pic) return MXL_ERR_OOM;
    mxl_frame* f = const_cast<mxl_frame*>(frame);
    int st = mxl_frames_to_rgba(ctx, &f, 1, pic, 0);
    if (st == MXL_OK) st = mxl_rgba_download(pic, 0, 1, rgba_host);
    mxl_rgba_free(pic);
    return st;
}

mxl_frame* mxl_frame_scale(mxl_frame* src, uint32_t out_w, uint32_t out_h)
{
    if (src && (out_w == 0 || out_h == 0)) { set_error("mxl_frame_scale: empty target %ux%u", out_w, out_h); return nullptr; }
    return frame_scale(src, out_w, out_h);
}

int mxl_frames_scale(mxl_ctx* ctx, mxl_frame* const* src, mxl_frame** dst, uint32_t n, uint32_t out_w, uint32_t out_h)
{
    if (!ctx || (n && (!src || !dst))) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if (!ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "mxl_frames_scale: context has no CUDA device; there is no CPU fallback");
    if (out_w == 0 || out_h == 0) MXL_FAIL(MXL_ERR_INVALID, "mxl_frames_scale: empty target %ux%u", out_w, out_h);
    for (uint32_t i = 0; i < n; i++)
        if (!src[i] || src[i]->ctx != ctx) MXL_FAIL(MXL_ERR_INVALID, "mxl_frames_scale: frame %u is NULL or of another context", i);
    MXL_TRY(ctx->activate());
    MXL_TRY(ctx->compute_begin());
    const int st = frames_scale(ctx, src, dst, n, out_w, out_h);
    MXL_TRY(ctx->compute_end());
    return st;
}

struct mxl_rgba {
    mxl_ctx* ctx = nullptr;
    uint32_t width = 0, height = 0, n = 0;
    uint8_t* dev = nullptr;
    void* jobs = nullptr;          // device staging for ComposeRgbaJob arrays
    size_t jobs_cap = 0;
    size_t picture_bytes() const { return (size_t)width * height * 4; }
};

mxl_rgba* mxl_rgba_alloc(mxl_ctx* ctx, uint32_t width, uint32_t height, uint32_t n_pictures)
{
    if (!ctx || !ctx->has_device() || width == 0 || height == 0 || n_pictures == 0) {
        set_error("mxl_rgba_alloc: needs a device context and a non-empty size");
        return nullptr;
    }
    if (ctx->activate() != MXL_OK) return nullptr;
    mxl_rgba* r = new mxl_rgba();
    r->ctx = ctx; r->width = width; r->height = height; r->n = n_pictures;
    if (cudaMalloc(&r->dev, r->picture_bytes() * n_pictures) != cudaSuccess) {
        set_error("mxl_rgba_alloc: out of device memory (%zu bytes)", r->picture_bytes() * n_pictures);
        delete r;
        return nullptr;
    }
    return r;
}

int mxl_rgba_free(mxl_rgba* pics)
{
    if (!pics) return MXL_OK;
    pics->ctx->activate();
    cudaStreamSynchronize(pics->ctx->stream);
    cudaFree(pics->dev);
    if (pics->jobs) cudaFree(pics->jobs);
    delete pics;
    return MXL_OK;
}

void* mxl_rgba_device_ptr(mxl_rgba* pics, uint32_t index)
{
    return pics && index < pics->n ? pics->dev + pics->picture_bytes() * index : nullptr;
}

int mxl_rgba_download_async(const mxl_rgba* pics, uint32_t first, uint32_t count, uint8_t* host)
{
    if (!pics || !host) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if ((uint64_t)first + count > pics->n) MXL_FAIL(MXL_ERR_LENGTH, "pictures %u..%u of %u", first, first + count, pics->n);
    mxl_ctx* ctx = pics->ctx;
    MXL_TRY(ctx->activate());
    cudaStream_t s;
    MXL_TRY(ctx->download_stream(&s));
    const size_t bytes = pics->picture_bytes() * count;
    MXL_CUDA(cudaMemcpyAsync(host, pics->dev + pics->picture_bytes() * first, bytes, cudaMemcpyDeviceToHost, s));
    ctx->d2h_bytes += bytes;
    return MXL_OK;
}

int mxl_rgba_download(const mxl_rgba* pics, uint32_t first, uint32_t count, uint8_t* host)
{
    MXL_TRY(mxl_rgba_download_async(pics, first, count, host));
    return mxl_ctx_synchronize(pics->ctx);
}

static int compose_rgba(mxl_ctx* ctx, mxl_frame* const* a, mxl_frame* const* b, uint32_t n, uint32_t fade, mxl_rgba* out, uint32_t first)
{
    if (!ctx || !out || (n && !a)) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if (!ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "compose_rgba: context has no CUDA device; there is no CPU fallback");
    if (out->ctx != ctx) MXL_FAIL(MXL_ERR_INVALID, "compose_rgba: pictures belong to another context");
    if ((uint64_t)first + n > out->n) MXL_FAIL(MXL_ERR_LENGTH, "pictures %u..%u of %u", first, first + n, out->n);
    if (n == 0) return MXL_OK;
    mxl_frame_layout lay;
    frame_layout_yuv420p(out->width, out->height, &lay);
    std::vector<k::ComposeRgbaJob> jobs(n);
    for (uint32_t i = 0; i < n; i++) {
        const mxl_frame* fa = a[i];
        const mxl_frame* fb = b ? b[i] : nullptr;
        for (const mxl_frame* f : {fa, fb})
            if (f && (f->ctx != ctx || f->layout.width != out->width || f->layout.height != out->height))
                MXL_FAIL(MXL_ERR_INVALID, "compose_rgba: layer %u is %ux%u, pictures are %ux%u", i, f->layout.width, f->layout.height, out->width, out->height);
        // a layer whose weight is 0 goes to the kernel as missing (never read): (a*255 + b*0) / 255 == a
        jobs[i] = k::ComposeRgbaJob{(fa && fade != 0) ? fa->dev : nullptr, (fb && fade != 255) ? fb->dev : nullptr,
                                    out->dev + out->picture_bytes() * (first + i), fade, 0};
    }
    MXL_TRY(ctx->activate());
    if (out->jobs_cap < n) {
        if (out->jobs) { MXL_CUDA(cudaStreamSynchronize(ctx->stream)); cudaFree(out->jobs); out->jobs = nullptr; }
        MXL_CUDA(cudaMalloc(&out->jobs, (size_t)n * 2 * sizeof(k::ComposeRgbaJob)));
        out->jobs_cap = (size_t)n * 2;
    }
    MXL_TRY(ctx->compute_begin());
    MXL_CUDA(cudaMemcpyAsync(out->jobs, jobs.data(), n * sizeof(k::ComposeRgbaJob), cudaMemcpyHostToDevice, ctx->stream));
    const int st = k::launch_compose_rgba(ctx, lay, (const k::ComposeRgbaJob*)out->jobs, n);
    MXL_TRY(ctx->compute_end());
    return st;
}

int mxl_video_compose_rgba(mxl_ctx* ctx, mxl_frame* const* a, mxl_frame* const* b, uint32_t n, double fader, mxl_rgba* out, uint32_t first_picture)
{
    if (n && !b) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    return compose_rgba(ctx, a, b, n, mxl_fader_to_u8(fader), out, first_picture);      // video_mixer.rs:168
}

int mxl_ctx_fused_profile(mxl_ctx* ctx, uint32_t max_ctas, uint64_t* stamps_out, uint32_t cap_ctas)
{
    if (!ctx || !ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_TRY(ctx->activate());
    MXL_CUDA(cudaStreamSynchronize(ctx->stream));
    int n = 0;
    if (stamps_out && ctx->fused_prof) {
        n = (int)std::min(ctx->fused_prof_ctas, cap_ctas);
        MXL_CUDA(cudaMemcpy(stamps_out, ctx->fused_prof, (size_t)n * k::kFusedProfStamps * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    }
    if (max_ctas != ctx->fused_prof_cap) {
        if (ctx->fused_prof) { cudaFree(ctx->fused_prof); ctx->fused_prof = nullptr; }
        ctx->fused_prof_cap = 0;
        if (max_ctas) {
            MXL_CUDA(cudaMalloc(&ctx->fused_prof, (size_t)max_ctas * k::kFusedProfStamps * sizeof(uint64_t)));
            ctx->fused_prof_cap = max_ctas;
        }
        ctx->fused_prof_ctas = 0;
    }
    return n;
}

int mxl_rgba_upload(mxl_rgba* pics, uint32_t first, uint32_t count, const uint8_t* host)
{
    if (!pics || !host) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if ((uint64_t)first + count > pics->n) MXL_FAIL(MXL_ERR_LENGTH, "pictures %u..%u of %u", first, first + count, pics->n);
    mxl_ctx* ctx = pics->ctx;
    MXL_TRY(ctx->activate());
    const size_t bytes = pics->picture_bytes() * count;
    MXL_CUDA(cudaMemcpyAsync(pics->dev + pics->picture_bytes() * first, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += bytes;
    return mxl_ctx_synchronize(ctx);
}

int mxl_rgba_to_frames(mxl_ctx* ctx, const mxl_rgba* pics, uint32_t first_picture, uint32_t n, mxl_frame* const* frames)
{
    if (!ctx || !pics || (n && !frames)) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if (!ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "mxl_rgba_to_frames: context has no CUDA device; there is no CPU fallback");
    if (pics->ctx != ctx) MXL_FAIL(MXL_ERR_INVALID, "mxl_rgba_to_frames: pictures belong to another context");
    if ((uint64_t)first_picture + n > pics->n) MXL_FAIL(MXL_ERR_LENGTH, "pictures %u..%u of %u", first_picture, first_picture + n, pics->n);
    if (n == 0) return MXL_OK;
    mxl_frame_layout lay;
    frame_layout_yuv420p(pics->width, pics->height, &lay);
    std::vector<k::RgbaToYuvJob> jobs(n);
    for (uint32_t i = 0; i < n; i++) {
        const mxl_frame* f = frames[i];
        if (!f || f->ctx != ctx || f->layout.width != pics->width || f->layout.height != pics->height)
            MXL_FAIL(MXL_ERR_INVALID, "mxl_rgba_to_frames: frame %u is NULL, of another context or not %ux%u", i, pics->width, pics->height);
        jobs[i] = k::RgbaToYuvJob{pics->dev + pics->picture_bytes() * (first_picture + i), f->dev};
    }
    MXL_TRY(ctx->activate());
    mxl_rgba* p = const_cast<mxl_rgba*>(pics);              // the job staging buffer rides on the picture set
    if (p->jobs_cap * sizeof(k::ComposeRgbaJob) < n * sizeof(k::RgbaToYuvJob)) {
        if (p->jobs) { MXL_CUDA(cudaStreamSynchronize(ctx->stream)); cudaFree(p->jobs); p->jobs = nullptr; }
        MXL_CUDA(cudaMalloc(&p->jobs, (size_t)n * 2 * sizeof(k::ComposeRgbaJob)));
        p->jobs_cap = (size_t)n * 2;
    }
    MXL_TRY(ctx->compute_begin());
    MXL_CUDA(cudaMemcpyAsync(p->jobs, jobs.data(), n * sizeof(k::RgbaToYuvJob), cudaMemcpyHostToDevice, ctx->stream));
    const int st = k::launch_rgba_to_yuv(ctx, lay, (const k::RgbaToYuvJob*)p->jobs, n);
    MXL_TRY(ctx->compute_end());
    return st;
}

int mxl_frames_to_rgba(mxl_ctx* ctx, mxl_frame* const* frames, uint32_t n, mxl_rgba* out, uint32_t first_picture)
{
    for (uint32_t i = 0; frames && i < n; i++)
        if (!frames[i]) MXL_FAIL(MXL_ERR_INVALID, "mxl_frames_to_rgba: frame %u is NULL", i);
    // fade = 255: out = (a*255 + blank*0) / 255 = a exactly
    return compose_rgba(ctx, frames, nullptr, n, 255u, out, first_picture);
}

}  // extern "C"

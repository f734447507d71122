// core.cu -- context, line buffers, frames, error state.
#include <string.h>

#include <mutex>

#include <stdlib.h>

#include <ctype.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>

#include "common.h"
#include "kernels.h"

namespace mxl {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

const char* last_error() { return g_error; }

// ---- Rational64 (num_rational semantics: reduced, den > 0) ---------------------------------------
static int64_t gcd64(int64_t a, int64_t b)
{
    if (a < 0) a = -a;
    if (b < 0) b = -b;
    while (b) { int64_t t = a % b; a = b; b = t; }
    return a;
}

Rational Rational::make(int64_t n, int64_t d)
{
    Rational r;
    if (d == 0) { r.num = n; r.den = 1; return r; }     // Rational64::new panics; callers validate
    if (d < 0) { n = -n; d = -d; }
    int64_t g = gcd64(n, d);
    if (g > 1) { n /= g; d /= g; }
    r.num = n; r.den = d;
    return r;
}

Rational Rational::operator+(const Rational& o) const
{
    // lcm-based like num_rational's Add
    int64_t g = gcd64(den, o.den);
    int64_t l = den / g * o.den;
    __int128 n = (__int128)num * (l / den) + (__int128)o.num * (l / o.den);
    return Rational::make((int64_t)n, l);
}

Rational Rational::operator-(const Rational& o) const
{
    int64_t g = gcd64(den, o.den);
    int64_t l = den / g * o.den;
    __int128 n = (__int128)num * (l / den) - (__int128)o.num * (l / o.den);
    return Rational::make((int64_t)n, l);
}

bool Rational::operator>=(const Rational& o) const
{
    return (__int128)num * o.den >= (__int128)o.num * den;
}

bool Rational::operator>(const Rational& o) const
{
    return (__int128)num * o.den > (__int128)o.num * den;
}

// ---- checked line access (io.rs:36-61,100-126) ---------------------------------------------------
static const char* type_name(int t)
{
    return t == MXL_LINE_MONO ? "mono" : (t == MXL_LINE_STEREO ? "stereo" : (t == MXL_LINE_VIDEO ? "video" : "?"));
}

int expect_input(const mxl_line* line, int type, const char* what)
{
    if (!line) return MXL_OK;   // Disconnected: expect_* hands out the static zero buffer / None
    if (line->type != type)
        MXL_FAIL(MXL_ERR_LINE_TYPE, "%s: expected %s input, got %s", what, type_name(type), type_name(line->type));
    return MXL_OK;
}

int expect_output(const mxl_line* line, int type, const char* what)
{
    if (!line) MXL_FAIL(MXL_ERR_INVALID, "%s: output line is NULL", what);
    if (line->type != type)
        MXL_FAIL(MXL_ERR_LINE_TYPE, "%s: expected %s output, got %s", what, type_name(type), type_name(line->type));
    return MXL_OK;
}

// ---- lines ---------------------------------------------------------------------------------------
mxl_line* line_alloc(mxl_ctx* ctx, int type, uint64_t frames)
{
    if (!ctx) { set_error("line_alloc: NULL context"); return nullptr; }
    if (type != MXL_LINE_MONO && type != MXL_LINE_STEREO && type != MXL_LINE_VIDEO) {
        set_error("line_alloc: bad line type %d", type);
        return nullptr;
    }
    mxl_line* l = new mxl_line();
    l->ctx = ctx; l->type = type; l->frames = frames; l->capacity = frames;
    if (type == MXL_LINE_VIDEO) {
        l->slots.resize(frames);
        return l;
    }
    if (!ctx->has_device()) { set_error("line_alloc: context has no device"); delete l; return nullptr; }
    if (ctx->activate() != MXL_OK) { delete l; return nullptr; }
    size_t bytes = (size_t)l->len() * sizeof(float);
    if (bytes) {
        cudaError_t e = cudaMalloc(&l->dev, bytes);
        if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); delete l; return nullptr; }
        ctx->compute_begin();
        e = cudaMemsetAsync(l->dev, 0, bytes, ctx->stream);        // vec![0.0; N]  (io.rs:73-74)
        if (e != cudaSuccess) { set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e)); cudaFree(l->dev); delete l; return nullptr; }
        ctx->compute_end();
    }
    return l;
}

void line_free(mxl_line* l)
{
    if (!l) return;
    if (l->type == MXL_LINE_VIDEO) {
        for (auto& s : l->slots) frame_release(s.frame);
    } else if (l->dev) {
        if (l->ctx) l->ctx->activate();
        cudaFree(l->dev);
    }
    delete l;
}

int line_resize(mxl_line* l, uint64_t frames)
{
    if (l->type == MXL_LINE_VIDEO) {
        for (auto& s : l->slots) { frame_release(s.frame); s.frame = nullptr; }
        l->slots.assign(frames, VideoSlot());
        l->frames = frames;
        return MXL_OK;
    }
    if (frames <= l->capacity) { l->frames = frames; return MXL_OK; }
    MXL_TRY(l->ctx->activate());
    // cudaFree synchronises the device; lines only grow between runs.
    l->ctx->change_epoch++;
    if (l->dev) { MXL_CUDA(cudaFree(l->dev)); l->dev = nullptr; }
    l->frames = frames;
    l->capacity = frames;
    size_t bytes = (size_t)l->len() * sizeof(float);
    if (bytes) {
        MXL_CUDA(cudaMalloc(&l->dev, bytes));
        MXL_TRY(l->ctx->compute_begin());
        MXL_CUDA(cudaMemsetAsync(l->dev, 0, bytes, l->ctx->stream));
        MXL_TRY(l->ctx->compute_end());
    }
    return MXL_OK;
}

// ---- frames --------------------------------------------------------------------------------------
static uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

void frame_layout_yuv420p(uint32_t w, uint32_t h, mxl_frame_layout* out)
{
    // The layout av_frame_get_buffer(frame, 0) gives the reference (codec/src/ffmpeg/frame.rs:84-86):
    // luma linesize = width padded to 32, chroma linesize = ceil(luma/2) padded to 32;
    // chroma rows = ceil(h/2).  All linesizes are multiples of 32, which video_mixer.rs:196-201 asserts.
    uint32_t luma = align_up(w, 32);
    uint32_t chroma = align_up((luma + 1) / 2, 32);
    out->width = w; out->height = h;
    out->stride[0] = luma; out->stride[1] = chroma; out->stride[2] = chroma;
    out->plane_h[0] = h; out->plane_h[1] = (h + 1) / 2; out->plane_h[2] = (h + 1) / 2;
    uint64_t off = 0;
    for (int p = 0; p < 3; p++) { out->offset[p] = off; off += (uint64_t)out->stride[p] * out->plane_h[p]; }
    out->size = off;
}

mxl_frame* frame_alloc(mxl_ctx* ctx, uint32_t w, uint32_t h)
{
    if (!ctx || !ctx->has_device()) { set_error("frame_alloc: context has no device"); return nullptr; }
    if (w == 0 || h == 0) { set_error("frame_alloc: empty picture %ux%u", w, h); return nullptr; }
    if (ctx->activate() != MXL_OK) return nullptr;
    mxl_frame* f = new mxl_frame();
    f->ctx = ctx;
    frame_layout_yuv420p(w, h, &f->layout);
    auto pooled = ctx->frame_pool.find(f->layout.size);
    if (pooled != ctx->frame_pool.end() && !pooled->second.empty()) {
        f->dev = pooled->second.back();      // same stream => reuse is ordered after the last use
        pooled->second.pop_back();
        ctx->frame_pool_bytes -= f->layout.size;
        return f;
    }
    cudaError_t e = cudaMalloc(&f->dev, f->layout.size);
    if (e != cudaSuccess) { set_error("cudaMalloc(%llu) failed: %s", (unsigned long long)f->layout.size, cudaGetErrorString(e)); delete f; return nullptr; }
    return f;
}

void frame_release(mxl_frame* f)
{
    if (!f) return;
    if (f->refs.fetch_sub(1) == 1) {
        if (f->slab) {
            if (f->slab->refs.fetch_sub(1) == 1) {
                f->ctx->activate();
                cudaFree(f->slab->base);      // synchronises with whatever still reads the slab
                delete f->slab;
            }
        } else if (f->dev) {
            mxl_ctx* c = f->ctx;
            if (c->frame_pool_bytes + f->layout.size > c->frame_pool_cap) {      // the pool is full: back to the driver
                c->activate();
                cudaFree(f->dev);             // synchronises with whatever still reads the frame
            } else {
                c->frame_pool[f->layout.size].push_back(f->dev);
                c->frame_pool_bytes += f->layout.size;
            }
        }
        delete f;
    }
}

void video_slot_set(VideoSlot& s, mxl_frame* f, Rational dur, Rational off)
{
    frame_retain(f);
    frame_release(s.frame);
    s.frame = f;
    s.duration_hint = dur;
    s.tick_offset = off;
}

KernelTimer::KernelTimer(mxl_ctx* c, const char* name) : ctx(c)
{
    if (!c || !c->kernel_timing) return;
    if (c->kernel_events.size() >= (1u << 16)) return;      // nobody is reading: stop recording rather than grow without bound
    cudaEvent_t ev[2] = {nullptr, nullptr};
    for (auto& e : ev) {
        if (!c->kernel_event_pool.empty()) { e = c->kernel_event_pool.back(); c->kernel_event_pool.pop_back(); }
        else if (cudaEventCreate(&e) != cudaSuccess) return;
    }
    if (cudaEventRecord(ev[0], c->stream) != cudaSuccess) return;
    c->kernel_events.push_back(mxl_ctx::KernelEvents{name, ev[0], ev[1]});
    slot = (int)c->kernel_events.size() - 1;
}

bool pdl_enabled(const mxl_ctx* ctx)
{
    static const bool off = getenv("MXL_NO_PDL") != nullptr;
    return !off && ctx && !ctx->kernel_timing && !ctx->pdl_hold;
}

KernelTimer::~KernelTimer()
{
    if (slot >= 0) cudaEventRecord(ctx->kernel_events[slot].b, ctx->stream);
}

}  // namespace mxl

int mxl_ctx::activate() const
{
    if (device < 0) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_CUDA(cudaSetDevice(device));
    return MXL_OK;
}

int mxl_ctx::upload_stream(cudaStream_t* s)
{
    if (!overlap) { *s = stream; return MXL_OK; }
    if (in_must_wait) { MXL_CUDA(cudaStreamWaitEvent(stream_in, ev_compute, 0)); in_must_wait = false; }
    in_dirty = true;
    *s = stream_in;
    return MXL_OK;
}

int mxl_ctx::download_stream(cudaStream_t* s)
{
    if (!overlap) { *s = stream; return MXL_OK; }
    if (out_must_wait) { MXL_CUDA(cudaStreamWaitEvent(stream_out, ev_compute, 0)); out_must_wait = false; }
    out_dirty = true;
    *s = stream_out;
    return MXL_OK;
}

int mxl_ctx::compute_begin()
{
    if (!overlap) return MXL_OK;
    if (in_dirty) {
        MXL_CUDA(cudaEventRecord(ev_in, stream_in));
        MXL_CUDA(cudaStreamWaitEvent(stream, ev_in, 0));
        in_dirty = false;
    }
    if (out_dirty) {
        MXL_CUDA(cudaEventRecord(ev_out, stream_out));
        MXL_CUDA(cudaStreamWaitEvent(stream, ev_out, 0));
        out_dirty = false;
    }
    return MXL_OK;
}

int mxl_ctx::compute_end()
{
    if (!overlap) return MXL_OK;
    MXL_CUDA(cudaEventRecord(ev_compute, stream));
    in_must_wait = out_must_wait = true;
    return MXL_OK;
}

using namespace mxl;

// ================================================================================================
// C ABI: context, lines, frames
// ================================================================================================
extern "C" {

static mxl_ctx* ctx_create(int device, uint32_t sample_rate, uint32_t spt, cudaStream_t stream, bool external)
{
    if (sample_rate == 0 || spt == 0) { set_error("mxl_ctx_create: sample_rate and samples_per_tick must be non-zero"); return nullptr; }
    mxl_ctx* c = new mxl_ctx();
    c->sample_rate = sample_rate;
    c->spt = spt;
    c->device = device;
    if (device < 0) { c->device = MXL_DEVICE_NONE; return c; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device >= count) {
        set_error("mxl_ctx_create: CUDA device %d not available (%s, %d devices); there is no CPU fallback", device,
                  e != cudaSuccess ? cudaGetErrorString(e) : "ok", count);
        delete c;
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", device); delete c; return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        c->sm_count = prop.multiProcessorCount;
        if (prop.major < 10) {
            set_error("mxl_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
            delete c;
            return nullptr;
        }
    }
    if (external) {
        c->stream = stream;
        c->own_stream = false;
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); delete c; return nullptr; }
        c->own_stream = true;
    }
    cudaEventCreate(&c->ev_begin);
    cudaEventCreate(&c->ev_end);
    return c;
}

mxl_ctx* mxl_ctx_create(int device, uint32_t sample_rate, uint32_t samples_per_tick)
{
    return ctx_create(device, sample_rate, samples_per_tick, nullptr, false);
}

mxl_ctx* mxl_ctx_create_on_stream(int device, uint32_t sample_rate, uint32_t samples_per_tick, void* cuda_stream)
{
    return ctx_create(device, sample_rate, samples_per_tick, (cudaStream_t)cuda_stream, true);
}

int mxl_ctx_device_memory(mxl_ctx* ctx, uint64_t* free_bytes, uint64_t* total_bytes)
{
    if (!ctx) MXL_FAIL(MXL_ERR_INVALID, "NULL context");
    if (!ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_TRY(mxl_ctx_synchronize(ctx));
    size_t f = 0, t = 0;
    MXL_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return MXL_OK;
}

// Host side of a host-fed session: the engine thread and the pinned buffers it allocates afterwards move next to the GPU.
// On a two-socket box a rank whose staging buffers live on the far socket pulls them through the inter-socket link, and
// N ranks that all start on socket 0 share that socket's memory controllers (bench.py's e2e leg at 2-8 GPUs).
int mxl_ctx_bind_host_to_gpu_node(mxl_ctx* ctx, int32_t* node_out, int32_t* cpus_out)
{
    if (node_out) *node_out = -1;
    if (cpus_out) *cpus_out = 0;
    if (!ctx) MXL_FAIL(MXL_ERR_INVALID, "NULL context");
    if (!ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_TRY(ctx->activate());
    char bdf[32] = {0};
    MXL_CUDA(cudaDeviceGetPCIBusId(bdf, sizeof bdf, ctx->device));
    for (char* p = bdf; *p; p++) *p = (char)tolower(*p);
    auto read_small = [](const std::string& path) -> std::string {
        FILE* f = fopen(path.c_str(), "r");
        if (!f) return "";
        char buf[4096];
        size_t n = fread(buf, 1, sizeof buf - 1, f);
        fclose(f);
        buf[n] = 0;
        while (n && (buf[n - 1] == '\n' || buf[n - 1] == ' ')) buf[--n] = 0;
        return buf;
    };
    // the kernel's view of the topology (nvidia-smi topo is not trustworthy inside a VM)
    const std::string node_s = read_small(std::string("/sys/bus/pci/devices/") + bdf + "/numa_node");
    const int node = node_s.empty() ? -1 : atoi(node_s.c_str());
    if (node < 0) return MXL_OK;                                   // single node, or the platform does not say: nothing to do
    const std::string list = read_small("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
    cpu_set_t set;
    CPU_ZERO(&set);
    int n_cpus = 0;
    for (size_t i = 0; i < list.size();) {                         // "0-15,32-47"
        const int a = atoi(list.c_str() + i);
        int b = a;
        while (i < list.size() && list[i] != '-' && list[i] != ',') i++;
        if (i < list.size() && list[i] == '-') { b = atoi(list.c_str() + i + 1); while (i < list.size() && list[i] != ',') i++; }
        for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, &set); n_cpus++; }
        if (i < list.size()) i++;
    }
    if (n_cpus == 0) return MXL_OK;
    if (sched_setaffinity(0, sizeof set, &set) != 0) return MXL_OK;               // not permitted (cgroup): leave things as they are
    if (node < 64) {
        unsigned long mask = 1ul << node;
        syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, &mask, sizeof(mask) * 8);   // pinned buffers allocated from now on: near node first
    }
    if (node_out) *node_out = node;
    if (cpus_out) *cpus_out = n_cpus;
    return MXL_OK;
}

int mxl_ctx_set_kernel_timing(mxl_ctx* ctx, int enabled)
{
    if (!ctx) MXL_FAIL(MXL_ERR_INVALID, "NULL context");
    ctx->kernel_timing = enabled != 0;
    return MXL_OK;
}

// Synchronises, then folds the event pairs recorded since the last read into one entry per kernel name.
int mxl_ctx_kernel_times(mxl_ctx* ctx, mxl_kernel_time* out, uint32_t cap)
{
    if (!ctx || (cap && !out)) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if (!ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    MXL_TRY(mxl_ctx_synchronize(ctx));
    uint32_t n = 0;
    for (auto& e : ctx->kernel_events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
            uint32_t i = 0;
            for (; i < n; i++) if (!strncmp(out[i].name, e.name, sizeof out[i].name - 1)) break;
            if (i == n && n < cap) {
                memset(&out[n], 0, sizeof out[n]);
                strncpy(out[n].name, e.name, sizeof out[n].name - 1);
                n++;
            }
            if (i < n) { out[i].launches++; out[i].total_ms += ms; }
        }
        ctx->kernel_event_pool.push_back(e.a);
        ctx->kernel_event_pool.push_back(e.b);
    }
    ctx->kernel_events.clear();
    return (int)n;
}

int mxl_ctx_destroy(mxl_ctx* ctx)
{
    if (!ctx) return MXL_OK;
    if (ctx->has_device()) {
        ctx->activate();
        cudaStreamSynchronize(ctx->stream);
        mxl_ctx_comm_destroy(ctx);
        if (ctx->flush_buf) cudaFree(ctx->flush_buf);
        for (auto& e : ctx->kernel_events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
        for (cudaEvent_t e : ctx->kernel_event_pool) cudaEventDestroy(e);
        for (auto& kv : ctx->eq_stream_tables) if (kv.second) cudaFree(kv.second);
        for (auto& kv : ctx->scale_tables) if (kv.second) cudaFree(kv.second);
        if (ctx->scale_jobs) cudaFree(ctx->scale_jobs);
        if (ctx->fused_prof) cudaFree(ctx->fused_prof);
        if (ctx->pcm_ring) cudaFree(ctx->pcm_ring);
        for (auto& kv : ctx->frame_pool)
            for (uint8_t* p : kv.second) cudaFree(p);
        if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
        if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
        if (ctx->stream_aux) { cudaStreamSynchronize(ctx->stream_aux); cudaStreamDestroy(ctx->stream_aux); }
        if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
        if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
        if (ctx->stream_in) { cudaStreamSynchronize(ctx->stream_in); cudaStreamDestroy(ctx->stream_in); }
        if (ctx->stream_out) { cudaStreamSynchronize(ctx->stream_out); cudaStreamDestroy(ctx->stream_out); }
        for (cudaEvent_t e : {ctx->ev_in, ctx->ev_out, ctx->ev_compute, ctx->fences[0], ctx->fences[1], ctx->fences[2], ctx->fences[3]})
            if (e) cudaEventDestroy(e);
        if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    }
    delete ctx;
    return MXL_OK;
}

int mxl_ctx_synchronize(mxl_ctx* ctx)
{
    if (!ctx) MXL_FAIL(MXL_ERR_INVALID, "NULL context");
    MXL_TRY(ctx->activate());
    MXL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->stream_in) MXL_CUDA(cudaStreamSynchronize(ctx->stream_in));
    if (ctx->stream_out) MXL_CUDA(cudaStreamSynchronize(ctx->stream_out));
    return MXL_OK;
}

// Frees pooled frame buffers until at most keep_bytes stay parked (size classes a session no longer uses -- its sources
// changed resolution -- would otherwise stay allocated until the context dies), and sets the pool's cap.
int mxl_ctx_trim_frame_pool(mxl_ctx* ctx, uint64_t keep_bytes, uint64_t new_cap_bytes, uint64_t* freed_out)
{
    if (!ctx) MXL_FAIL(MXL_ERR_INVALID, "NULL context");
    uint64_t freed = 0;
    if (ctx->has_device() && ctx->frame_pool_bytes > keep_bytes) {
        MXL_TRY(mxl_ctx_synchronize(ctx));
        // largest classes first: the ones a changed geometry left behind are usually not the smallest
        std::vector<uint64_t> sizes;
        for (auto& kv : ctx->frame_pool) if (!kv.second.empty()) sizes.push_back(kv.first);
        std::sort(sizes.begin(), sizes.end(), [](uint64_t a, uint64_t b) { return a > b; });
        for (uint64_t sz : sizes) {
            auto& v = ctx->frame_pool[sz];
            while (!v.empty() && ctx->frame_pool_bytes > keep_bytes) {
                cudaFree(v.back());
                v.pop_back();
                ctx->frame_pool_bytes -= sz;
                freed += sz;
            }
        }
    }
    if (new_cap_bytes) ctx->frame_pool_cap = new_cap_bytes;
    if (freed_out) *freed_out = freed;
    return MXL_OK;
}

int mxl_ctx_set_copy_overlap(mxl_ctx* ctx, int enabled)
{
    if (!ctx) MXL_FAIL(MXL_ERR_INVALID, "NULL context");
    if (!ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "mxl_ctx_set_copy_overlap: context has no CUDA device");
    MXL_TRY(mxl_ctx_synchronize(ctx));
    if (enabled && !ctx->stream_in) {
        MXL_CUDA(cudaStreamCreateWithFlags(&ctx->stream_in, cudaStreamNonBlocking));
        MXL_CUDA(cudaStreamCreateWithFlags(&ctx->stream_out, cudaStreamNonBlocking));
        MXL_CUDA(cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming));
        MXL_CUDA(cudaEventCreateWithFlags(&ctx->ev_out, cudaEventDisableTiming));
        MXL_CUDA(cudaEventCreateWithFlags(&ctx->ev_compute, cudaEventDisableTiming));
        for (auto& f : ctx->fences) MXL_CUDA(cudaEventCreateWithFlags(&f, cudaEventDisableTiming));
    }
    ctx->overlap = enabled != 0;
    ctx->in_dirty = ctx->out_dirty = ctx->in_must_wait = ctx->out_must_wait = false;
    return MXL_OK;
}

int mxl_ctx_download_fence(mxl_ctx* ctx, uint32_t slot)
{
    if (!ctx || slot >= 4) MXL_FAIL(MXL_ERR_INVALID, "bad fence slot");
    if (!ctx->overlap) MXL_FAIL(MXL_ERR_INVALID, "fences need copy overlap enabled");
    MXL_TRY(ctx->activate());
    cudaStream_t s;
    MXL_TRY(ctx->download_stream(&s));
    MXL_CUDA(cudaEventRecord(ctx->fences[slot], s));
    return MXL_OK;
}

int mxl_ctx_wait_fence(mxl_ctx* ctx, uint32_t slot)
{
    if (!ctx || slot >= 4) MXL_FAIL(MXL_ERR_INVALID, "bad fence slot");
    if (!ctx->overlap) MXL_FAIL(MXL_ERR_INVALID, "fences need copy overlap enabled");
    MXL_TRY(ctx->activate());
    MXL_CUDA(cudaEventSynchronize(ctx->fences[slot]));
    return MXL_OK;
}

uint64_t mxl_ctx_h2d_bytes(const mxl_ctx* ctx) { return ctx ? ctx->h2d_bytes : 0; }
uint64_t mxl_ctx_d2h_bytes(const mxl_ctx* ctx) { return ctx ? ctx->d2h_bytes : 0; }

void* mxl_ctx_stream(mxl_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint32_t mxl_ctx_sample_rate(const mxl_ctx* ctx) { return ctx ? ctx->sample_rate : 0; }
uint32_t mxl_ctx_samples_per_tick(const mxl_ctx* ctx) { return ctx ? ctx->spt : 0; }
uint64_t mxl_ctx_launch_count(const mxl_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mxl_ctx_timer_begin(mxl_ctx* ctx)
{
    if (!ctx) MXL_FAIL(MXL_ERR_INVALID, "NULL context");
    MXL_TRY(ctx->activate());
    MXL_CUDA(cudaEventRecord(ctx->ev_begin, ctx->stream));
    return MXL_OK;
}

int mxl_ctx_timer_end(mxl_ctx* ctx)
{
    if (!ctx) MXL_FAIL(MXL_ERR_INVALID, "NULL context");
    MXL_TRY(ctx->activate());
    MXL_CUDA(cudaEventRecord(ctx->ev_end, ctx->stream));
    return MXL_OK;
}

int mxl_ctx_timer_elapsed_ms(mxl_ctx* ctx, float* ms)
{
    if (!ctx || !ms) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    MXL_TRY(ctx->activate());
    MXL_CUDA(cudaEventSynchronize(ctx->ev_end));
    MXL_CUDA(cudaEventElapsedTime(ms, ctx->ev_begin, ctx->ev_end));
    return MXL_OK;
}

int mxl_ctx_flush_l2(mxl_ctx* ctx)
{
    if (!ctx) MXL_FAIL(MXL_ERR_INVALID, "NULL context");
    MXL_TRY(ctx->activate());
    if (!ctx->flush_buf) {
        ctx->flush_bytes = (size_t)256 << 20;     // 2x the 126 MB L2
        MXL_CUDA(cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
    }
    // not counted as a path kernel
    uint64_t before = ctx->launches;
    int s = k::launch_fill_bytes(ctx, ctx->flush_buf, ctx->flush_bytes, 0);
    ctx->launches = before;
    return s;
}

const char* mxl_last_error(void) { return last_error(); }
const char* mxl_version(void) { return "mixlab-b200 0.1 (sm_100a)"; }

void* mxl_host_alloc(size_t bytes)
{
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { set_error("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e)); return nullptr; }
    return p;
}

int mxl_host_free(void* p)
{
    if (!p) return MXL_OK;
    MXL_CUDA(cudaFreeHost(p));
    return MXL_OK;
}

// ---- audio lines ---------------------------------------------------------------------------------
mxl_line* mxl_line_alloc(mxl_ctx* ctx, int line_type, uint64_t frames)
{
    if (line_type == MXL_LINE_VIDEO) { set_error("mxl_line_alloc: use mxl_video_line_alloc for video lines"); return nullptr; }
    return line_alloc(ctx, line_type, frames);
}

int mxl_line_free(mxl_line* line) { line_free(line); return MXL_OK; }
int mxl_line_type_of(const mxl_line* line) { return line ? line->type : MXL_ERR_INVALID; }
uint64_t mxl_line_frames(const mxl_line* line) { return line ? line->frames : 0; }
uint64_t mxl_line_len(const mxl_line* line) { return line ? line->len() : 0; }
void* mxl_line_device_ptr(mxl_line* line) { return line ? line->dev : nullptr; }

static int audio_line(const mxl_line* line, uint64_t n, const char* what)
{
    if (!line) MXL_FAIL(MXL_ERR_INVALID, "%s: NULL line", what);
    if (line->type == MXL_LINE_VIDEO) MXL_FAIL(MXL_ERR_LINE_TYPE, "%s: video line has no sample payload", what);
    if (n > line->len()) MXL_FAIL(MXL_ERR_LENGTH, "%s: %llu floats requested, line holds %llu", what,
                                  (unsigned long long)n, (unsigned long long)line->len());
    return line->ctx->activate();
}

int mxl_line_zero(mxl_line* line)
{
    MXL_TRY(audio_line(line, 0, "mxl_line_zero"));
    if (line->len()) {
        MXL_TRY(line->ctx->compute_begin());
        MXL_CUDA(cudaMemsetAsync(line->dev, 0, line->len() * sizeof(float), line->ctx->stream));
        MXL_TRY(line->ctx->compute_end());
    }
    return MXL_OK;
}

int mxl_line_upload_async(mxl_line* line, const float* host, uint64_t n)
{
    MXL_TRY(audio_line(line, n, "mxl_line_upload"));
    if (n && !host) MXL_FAIL(MXL_ERR_INVALID, "mxl_line_upload: NULL host pointer");
    if (n) {
        cudaStream_t st;
        MXL_TRY(line->ctx->upload_stream(&st));
        MXL_CUDA(cudaMemcpyAsync(line->dev, host, n * sizeof(float), cudaMemcpyHostToDevice, st));
        line->ctx->h2d_bytes += n * sizeof(float);
    }
    return MXL_OK;
}

int mxl_line_download_async(const mxl_line* line, float* host, uint64_t n)
{
    MXL_TRY(audio_line(line, n, "mxl_line_download"));
    if (n && !host) MXL_FAIL(MXL_ERR_INVALID, "mxl_line_download: NULL host pointer");
    if (n) {
        cudaStream_t st;
        MXL_TRY(line->ctx->download_stream(&st));
        MXL_CUDA(cudaMemcpyAsync(host, line->dev, n * sizeof(float), cudaMemcpyDeviceToHost, st));
        line->ctx->d2h_bytes += n * sizeof(float);
    }
    return MXL_OK;
}

int mxl_line_upload(mxl_line* line, const float* host, uint64_t n)
{
    MXL_TRY(mxl_line_upload_async(line, host, n));
    return mxl_ctx_synchronize(line->ctx);
}

int mxl_line_download(const mxl_line* line, float* host, uint64_t n)
{
    MXL_TRY(mxl_line_download_async(line, host, n));
    return mxl_ctx_synchronize(line->ctx);
}

// ---- frames --------------------------------------------------------------------------------------
int mxl_frame_layout_yuv420p(uint32_t width, uint32_t height, mxl_frame_layout* out)
{
    if (!out) MXL_FAIL(MXL_ERR_INVALID, "NULL layout");
    frame_layout_yuv420p(width, height, out);
    return MXL_OK;
}

int mxl_unify_picture_settings(uint32_t aw, uint32_t ah, uint32_t bw, uint32_t bh, uint32_t* w, uint32_t* h)
{
    if (!w || !h) MXL_FAIL(MXL_ERR_INVALID, "NULL output");
    // video_mixer.rs:276-297; yuv420p => log2_chroma_w = log2_chroma_h = 1, masks = 1
    uint32_t width = aw > bw ? aw : bw;
    uint32_t height = ah > bh ? ah : bh;
    *w = (width + 1u) & ~1u;
    *h = (height + 1u) & ~1u;
    return MXL_OK;
}

int mxl_scale_geometry_yuv420p(uint32_t in_w, uint32_t in_h, uint32_t out_w, uint32_t out_h, mxl_scale_geometry* out)
{
    if (!out) MXL_FAIL(MXL_ERR_INVALID, "NULL output");
    if (in_w == 0 || in_h == 0) MXL_FAIL(MXL_ERR_INVALID, "empty input picture");   // Ratio::new would panic on /0
    // src/video/encode.rs:355-374: scale_factor = min(W/w, H/h) as an exact ratio
    uint64_t sn, sd;
    if ((uint64_t)out_w * in_h <= (uint64_t)out_h * in_w) { sn = out_w; sd = in_w; } else { sn = out_h; sd = in_h; }
    uint32_t sw = (uint32_t)((sn * in_w) / sd) & ~1u;       // to_integer() then align_horizontal (pixfmt.rs:104-106)
    uint32_t sh = (uint32_t)((sn * in_h) / sd) & ~1u;       // align_vertical (pixfmt.rs:108-110)
    out->scaled_w = sw;
    out->scaled_h = sh;
    out->letterbox_x = ((out_w - sw) / 2) & ~1u;
    out->letterbox_y = ((out_h - sh) / 2) & ~1u;
    return MXL_OK;
}

uint8_t mxl_fader_to_u8(double fader)
{
    // Rust `(fader * 255.0) as u8`: NaN -> 0, saturating, toward zero
    double v = fader * 255.0;
    if (v != v) return 0;
    if (v >= 255.0) return 255;
    if (v <= 0.0) return 0;
    return (uint8_t)v;
}

mxl_frame* mxl_frame_alloc(mxl_ctx* ctx, uint32_t width, uint32_t height) { return frame_alloc(ctx, width, height); }

mxl_frame* mxl_frame_blank(mxl_ctx* ctx, uint32_t width, uint32_t height)
{
    mxl_frame* f = frame_alloc(ctx, width, height);
    if (!f) return nullptr;
    ctx->compute_begin();
    if (k::launch_blank(ctx, f->layout, f->dev) != MXL_OK) { frame_release(f); return nullptr; }
    ctx->compute_end();
    return f;
}

mxl_frame* mxl_frame_retain(mxl_frame* frame) { return frame_retain(frame); }
int mxl_frame_release(mxl_frame* frame) { frame_release(frame); return MXL_OK; }

int mxl_frame_get_layout(const mxl_frame* frame, mxl_frame_layout* out)
{
    if (!frame || !out) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    *out = frame->layout;
    return MXL_OK;
}

void* mxl_frame_device_ptr(mxl_frame* frame) { return frame ? frame->dev : nullptr; }

static int frame_copy_planes(const mxl_frame* f, uint8_t* const host[3], const uint32_t strides[3], bool upload)
{
    if (!f || !host || !strides) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    MXL_TRY(f->ctx->activate());
    cudaStream_t st;
    if (upload) MXL_TRY(f->ctx->upload_stream(&st)); else MXL_TRY(f->ctx->download_stream(&st));
    for (int p = 0; p < 3; p++) {
        uint32_t w = p == 0 ? f->layout.width : (f->layout.width + 1) / 2;
        if (!host[p] || strides[p] < w) MXL_FAIL(MXL_ERR_INVALID, "plane %d: NULL pointer or stride %u < width %u", p, strides[p], w);
        if (upload)
            MXL_CUDA(cudaMemcpy2DAsync(f->dev + f->layout.offset[p], f->layout.stride[p], host[p], strides[p], w,
                                       f->layout.plane_h[p], cudaMemcpyHostToDevice, st));
        else
            MXL_CUDA(cudaMemcpy2DAsync(host[p], strides[p], f->dev + f->layout.offset[p], f->layout.stride[p], w,
                                       f->layout.plane_h[p], cudaMemcpyDeviceToHost, st));
    }
    return mxl_ctx_synchronize(f->ctx);
}

int mxl_frame_upload(mxl_frame* frame, const uint8_t* const planes[3], const uint32_t strides[3])
{
    return frame_copy_planes(frame, const_cast<uint8_t* const*>(planes), strides, true);
}

int mxl_frame_download(const mxl_frame* frame, uint8_t* const planes[3], const uint32_t strides[3])
{
    return frame_copy_planes(frame, planes, strides, false);
}

int mxl_frame_upload_raw_async(mxl_frame* frame, const uint8_t* host, uint64_t size)
{
    if (!frame || !host) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if (size != frame->layout.size) MXL_FAIL(MXL_ERR_LENGTH, "raw size %llu != frame size %llu", (unsigned long long)size, (unsigned long long)frame->layout.size);
    MXL_TRY(frame->ctx->activate());
    cudaStream_t st;
    MXL_TRY(frame->ctx->upload_stream(&st));
    MXL_CUDA(cudaMemcpyAsync(frame->dev, host, size, cudaMemcpyHostToDevice, st));
    frame->ctx->h2d_bytes += size;
    return MXL_OK;
}

int mxl_frame_download_raw_async(const mxl_frame* frame, uint8_t* host, uint64_t size)
{
    if (!frame || !host) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if (size != frame->layout.size) MXL_FAIL(MXL_ERR_LENGTH, "raw size %llu != frame size %llu", (unsigned long long)size, (unsigned long long)frame->layout.size);
    MXL_TRY(frame->ctx->activate());
    cudaStream_t st;
    MXL_TRY(frame->ctx->download_stream(&st));
    MXL_CUDA(cudaMemcpyAsync(host, frame->dev, size, cudaMemcpyDeviceToHost, st));
    frame->ctx->d2h_bytes += size;
    return MXL_OK;
}

int mxl_frames_alloc_batch(mxl_ctx* ctx, uint32_t width, uint32_t height, uint32_t n, mxl_frame** frames_out)
{
    if (!ctx || (n && !frames_out)) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if (!ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "mxl_frames_alloc_batch: context has no CUDA device");
    if (width == 0 || height == 0) MXL_FAIL(MXL_ERR_INVALID, "empty picture %ux%u", width, height);
    if (n == 0) return MXL_OK;
    MXL_TRY(ctx->activate());
    mxl_frame_layout lay;
    frame_layout_yuv420p(width, height, &lay);
    const uint64_t pitch = (lay.size + 255) & ~(uint64_t)255;     // frames stay 256-byte aligned; 1080p: pitch == size
    FrameSlab* slab = new FrameSlab();
    cudaError_t e = cudaMalloc(&slab->base, pitch * n);
    if (e != cudaSuccess) {
        delete slab;
        MXL_FAIL(MXL_ERR_OOM, "cudaMalloc(%llu) failed: %s", (unsigned long long)(pitch * n), cudaGetErrorString(e));
    }
    slab->refs = (int)n;
    for (uint32_t i = 0; i < n; i++) {
        mxl_frame* f = new mxl_frame();
        f->ctx = ctx;
        f->layout = lay;
        f->slab = slab;
        f->dev = slab->base + pitch * i;
        frames_out[i] = f;
    }
    return MXL_OK;
}

// One cudaMemcpyAsync per run of frames that are adjacent on the device (and, by construction, in `host`).
static int frames_copy_raw(mxl_frame* const* frames, uint32_t n, uint8_t* host, uint64_t bytes_each, bool upload)
{
    if (n == 0) return MXL_OK;
    if (!frames || !host) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    mxl_ctx* ctx = frames[0] ? frames[0]->ctx : nullptr;
    for (uint32_t i = 0; i < n; i++) {
        if (!frames[i] || frames[i]->ctx != ctx) MXL_FAIL(MXL_ERR_INVALID, "frame %u is NULL or of another context", i);
        if (frames[i]->layout.size != bytes_each)
            MXL_FAIL(MXL_ERR_LENGTH, "raw size %llu != size %llu of frame %u", (unsigned long long)bytes_each, (unsigned long long)frames[i]->layout.size, i);
    }
    MXL_TRY(ctx->activate());
    cudaStream_t st;
    MXL_TRY(upload ? ctx->upload_stream(&st) : ctx->download_stream(&st));
    uint32_t i = 0;
    while (i < n) {
        uint32_t j = i + 1;
        while (j < n && frames[j]->dev == frames[j - 1]->dev + bytes_each) j++;
        const uint64_t bytes = (uint64_t)(j - i) * bytes_each;
        if (upload) MXL_CUDA(cudaMemcpyAsync(frames[i]->dev, host + (uint64_t)i * bytes_each, bytes, cudaMemcpyHostToDevice, st));
        else MXL_CUDA(cudaMemcpyAsync(host + (uint64_t)i * bytes_each, frames[i]->dev, bytes, cudaMemcpyDeviceToHost, st));
        i = j;
    }
    (upload ? ctx->h2d_bytes : ctx->d2h_bytes) += (uint64_t)n * bytes_each;
    return MXL_OK;
}

int mxl_frames_upload_raw_async(mxl_frame* const* frames, uint32_t n, const uint8_t* host, uint64_t bytes_each)
{
    return frames_copy_raw(frames, n, const_cast<uint8_t*>(host), bytes_each, true);
}

int mxl_frames_download_raw_async(mxl_frame* const* frames, uint32_t n, uint8_t* host, uint64_t bytes_each)
{
    return frames_copy_raw(frames, n, host, bytes_each, false);
}

int mxl_frame_upload_raw(mxl_frame* frame, const uint8_t* host, uint64_t size)
{
    MXL_TRY(mxl_frame_upload_raw_async(frame, host, size));
    return mxl_ctx_synchronize(frame->ctx);
}

int mxl_frame_download_raw(const mxl_frame* frame, uint8_t* host, uint64_t size)
{
    MXL_TRY(mxl_frame_download_raw_async(frame, host, size));
    return mxl_ctx_synchronize(frame->ctx);
}

// ---- video lines ---------------------------------------------------------------------------------
mxl_line* mxl_video_line_alloc(mxl_ctx* ctx, uint32_t ticks) { return line_alloc(ctx, MXL_LINE_VIDEO, ticks); }

int mxl_video_line_set(mxl_line* line, uint32_t slot, mxl_frame* frame, int64_t duration_num, int64_t duration_den,
                       int64_t offset_num, int64_t offset_den)
{
    if (!line) MXL_FAIL(MXL_ERR_INVALID, "NULL line");
    if (line->type != MXL_LINE_VIDEO) MXL_FAIL(MXL_ERR_LINE_TYPE, "mxl_video_line_set: not a video line");
    if (slot >= line->slots.size()) MXL_FAIL(MXL_ERR_LENGTH, "slot %u out of %zu", slot, line->slots.size());
    if (frame && (duration_den == 0 || offset_den == 0)) MXL_FAIL(MXL_ERR_INVALID, "zero denominator");
    video_slot_set(line->slots[slot], frame, frame ? Rational::make(duration_num, duration_den) : Rational(),
                   frame ? Rational::make(offset_num, offset_den) : Rational());
    return MXL_OK;
}

mxl_frame* mxl_video_line_get(const mxl_line* line, uint32_t slot)
{
    if (!line || line->type != MXL_LINE_VIDEO || slot >= line->slots.size()) return nullptr;
    return line->slots[slot].frame;
}

int mxl_video_line_get_timing(const mxl_line* line, uint32_t slot, int64_t duration[2], int64_t offset[2])
{
    if (!line || line->type != MXL_LINE_VIDEO || slot >= line->slots.size()) MXL_FAIL(MXL_ERR_INVALID, "mxl_video_line_get_timing: no such slot");
    const VideoSlot& s = line->slots[slot];
    if (!s.frame) MXL_FAIL(MXL_ERR_INVALID, "mxl_video_line_get_timing: slot %u is empty", slot);
    if (duration) { duration[0] = s.duration_hint.num; duration[1] = s.duration_hint.den; }
    if (offset) { offset[0] = s.tick_offset.num; offset[1] = s.tick_offset.den; }
    return MXL_OK;
}

int mxl_video_line_clear(mxl_line* line)
{
    if (!line) MXL_FAIL(MXL_ERR_INVALID, "NULL line");
    if (line->type != MXL_LINE_VIDEO) MXL_FAIL(MXL_ERR_LINE_TYPE, "mxl_video_line_clear: not a video line");
    for (auto& s : line->slots) { frame_release(s.frame); s = VideoSlot(); }
    return MXL_OK;
}

}  // extern "C"

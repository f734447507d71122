// graph.cu -- Workspace (src/engine/workspace.rs) + Engine::run_tick (src/engine.rs:400-510) with
// device-resident line buffers.
//
// What the reference does every tick -- build the terminal set, DFS-topsort from it, allocate and
// zero one Vec per output, run modules serially -- is done here once per topology change.  The plan
// groups modules by (dependency level, kind) into STAGES; a stage is one batched kernel launch over
// all its modules and over every tick of the call (each audio line holds n_ticks*S frames, legal
// because the reference's run_tick loops are length-agnostic and module state carries across).
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <set>

#include "modules.h"

using namespace mxl;

struct Stage {
    int level = 0;
    int kind = 0;
    std::vector<int> modules;
    int fused = -1;                     // index into mxl_graph::fused for a MXL_STAGE_FUSED_VOICE_MIX stage
    float last_ms = -1.f;
    int last_launches = 0;
    uint64_t last_bytes = 0;
    float last_host_us = 0.f;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

struct mxl_graph {
    mxl_ctx* ctx = nullptr;
    std::vector<mxl_module*> modules;                                  // index = ModuleId; nullptr = removed
    std::map<std::pair<int, uint32_t>, std::pair<int, uint32_t>> connections;   // InputId -> OutputId (workspace.rs:17)
    bool dirty = true;
    bool profiling = false;
    bool timings_pending = false;
    bool split_streams = true;                                         // audio stages on ctx->stream_aux next to video stages
    bool fusion = true;                                                // Osc -> EqThree -> Panner -> Mixer [-> Meter] groups as one launch
    std::set<std::pair<int, uint32_t>> pinned;                         // output lines the host observes: always materialised
    std::set<std::pair<int, uint32_t>> hidden;                         // interior lines of fused groups the current plan does not write
    std::set<int> fusion_veto;                                         // mixers whose group's parameters left the fused kernel's domain
    std::vector<FusedGroup> fused;
    uint64_t runs_since_plan = 0;
    uint64_t checked_epoch = 0;                                        // ctx->change_epoch the per-run validity checks last saw
    std::vector<mxl_line*> resize_list;                                // lines a run sizes to the call: every output the plan writes
    float last_call_host_us = 0.f;                                     // host time of the last run_ticks call
    uint32_t last_call_ticks = 0;

    // plan
    std::vector<int> run_order;
    std::vector<int> position;                                         // module id -> index in run_order, -1 = does not run
    std::vector<Stage> stages;
    std::vector<std::vector<mxl_line*>> out_lines;                     // graph-owned output lines per module
    // resolved inputs: per module, per input: producing (module, output) or (-1, 0) = Disconnected
    std::vector<std::vector<std::pair<int, uint32_t>>> resolved;

    ~mxl_graph()
    {
        for (auto& s : stages) {
            if (s.ev0) cudaEventDestroy(s.ev0);
            if (s.ev1) cudaEventDestroy(s.ev1);
        }
        for (auto& v : out_lines) for (mxl_line* l : v) line_free(l);
        for (mxl_module* m : modules) delete m;
    }
};

namespace {

bool is_source(int kind) { return kind == MXL_MOD_SOURCE_MONO || kind == MXL_MOD_SOURCE_STEREO || kind == MXL_MOD_SOURCE_VIDEO; }

mxl_module* module_at(mxl_graph* g, int id)
{
    return (id >= 0 && (size_t)id < g->modules.size()) ? g->modules[id] : nullptr;
}

// engine.rs:439-457
void traverse(mxl_graph* g, int id, std::vector<char>& seen)
{
    if (seen[id]) return;
    seen[id] = 1;
    mxl_module* m = g->modules[id];
    for (uint32_t i = 0; i < m->inputs.size(); i++) {
        auto it = g->connections.find({id, i});
        if (it != g->connections.end() && module_at(g, it->second.first)) traverse(g, it->second.first, seen);
    }
    g->run_order.push_back(id);
}

// Groups Oscillator -> EqThree -> StereoPanner -> Mixer [-> Meter] (BASELINE configs 2 and 4 are made of them) that can
// run as one launch (fused_voice.cu).  A group is taken only when no module outside it consumes one of its interior
// lines: those consumers would be scheduled by level, before the fused launch has written anything.  Lines the HOST
// observes (mxl_graph_output / mxl_graph_pin_output) stay inside the group and are written by the fused kernel.
void find_fused_groups(mxl_graph* g, const std::vector<int>& level, std::vector<int>& owner)
{
    g->fused.clear();
    g->hidden.clear();
    owner.assign(g->modules.size(), -1);
    static const bool env_off = getenv("MXL_NO_FUSION") != nullptr;
    if (!g->fusion || env_off || !g->ctx->has_device()) return;
    std::map<std::pair<int, uint32_t>, int> consumers;
    for (int id : g->run_order)
        for (const auto& r : g->resolved[id])
            if (r.first >= 0) consumers[r]++;
    auto n_consumers = [&](int m, uint32_t o) { auto it = consumers.find({m, o}); return it == consumers.end() ? 0 : it->second; };
    auto line_if_pinned = [&](int m, uint32_t o) -> mxl_line* { return g->pinned.count({m, o}) ? g->out_lines[m][o] : nullptr; };
    for (int id : g->run_order) {
        mxl_module* mx = g->modules[id];
        if (mx->kind != MXL_MOD_MIXER || g->fusion_veto.count(id) || mx->inputs.empty()) continue;
        FusedGroup grp;
        grp.mixer = mx;
        grp.members.push_back(id);
        std::map<int, int> voice_of_eq, eq_refs;
        bool ok = true;
        for (uint32_t c = 0; ok && c < mx->inputs.size(); c++) {
            FusedChanRef cr{false, -1, -1, nullptr};
            const auto r = g->resolved[id][c];
            if (r.first >= 0) {
                mxl_module* pm = g->modules[r.first];
                if (pm->kind != MXL_MOD_STEREO_PANNER || owner[r.first] >= 0 || n_consumers(r.first, 0) != 1) { ok = false; break; }
                cr.connected = true;
                cr.pan_out = line_if_pinned(r.first, 0);
                grp.members.push_back(r.first);
                for (uint32_t side = 0; ok && side < 2; side++) {
                    const auto rr = g->resolved[r.first][side];
                    if (rr.first < 0) continue;
                    mxl_module* em = g->modules[rr.first];
                    if (em->kind != MXL_MOD_EQ_THREE || owner[rr.first] >= 0) { ok = false; break; }
                    auto it = voice_of_eq.find(rr.first);
                    if (it == voice_of_eq.end()) {
                        FusedVoiceRef v{nullptr, em, g->out_lines[rr.first][0], nullptr, nullptr};
                        const auto ro = g->resolved[rr.first][0];
                        if (ro.first >= 0) {
                            mxl_module* om = g->modules[ro.first];
                            if (om->kind != MXL_MOD_OSCILLATOR || ro.second != 0 || owner[ro.first] >= 0 ||
                                n_consumers(ro.first, 0) != 1 || n_consumers(ro.first, 1) != 0) { ok = false; break; }
                            v.osc = om;
                            v.osc_mono = line_if_pinned(ro.first, 0);
                            v.osc_stereo = line_if_pinned(ro.first, 1);
                            grp.members.push_back(ro.first);
                        }
                        grp.members.push_back(rr.first);
                        it = voice_of_eq.emplace(rr.first, (int)grp.voices.size()).first;
                        grp.voices.push_back(v);
                    }
                    (side == 0 ? cr.left : cr.right) = it->second;
                    eq_refs[rr.first]++;
                }
            }
            grp.chans.push_back(cr);
        }
        for (const auto& e : eq_refs)
            if (ok && n_consumers(e.first, 0) != e.second) ok = false;       // an EqThree line also feeds somebody else
        if (!ok || !fused_group_supported(g->ctx, grp)) continue;
        grp.master = g->out_lines[id][0];
        grp.cue = g->out_lines[id][1];
        if (fused_meter_supported(g->ctx)) {                                 // the first Meter on the master bus rides along
            for (int mid : g->run_order) {
                mxl_module* mm = g->modules[mid];
                if (mm->kind == MXL_MOD_METER && owner[mid] < 0 && g->resolved[mid][0] == std::make_pair(id, 0u)) {
                    grp.meter = mm;
                    grp.members.push_back(mid);
                    break;
                }
            }
        }
        for (int mid : grp.members) owner[mid] = (int)g->fused.size();
        for (int mid : grp.members) {
            const int kind = g->modules[mid]->kind;
            if (kind == MXL_MOD_OSCILLATOR) { for (uint32_t o = 0; o < 2; o++) if (!g->pinned.count({mid, o})) g->hidden.insert({mid, o}); }
            else if (kind == MXL_MOD_STEREO_PANNER) { if (!g->pinned.count({mid, 0u})) g->hidden.insert({mid, 0u}); }
        }
        (void)level;
        g->fused.push_back(std::move(grp));
    }
}

int build_plan(mxl_graph* g)
{
    const size_t n = g->modules.size();
    // terminal modules = modules that feed nobody (engine.rs:408-416)
    std::vector<char> terminal(n, 0), seen(n, 0);
    for (size_t i = 0; i < n; i++) terminal[i] = g->modules[i] != nullptr;
    for (auto& c : g->connections)
        if (module_at(g, c.first.first) && module_at(g, c.second.first)) terminal[c.second.first] = 0;
    // DFS from each terminal (engine.rs:421-430).  The reference walks a HashSet in arbitrary order;
    // ascending ModuleId here, which is one of the orders it can take.
    g->run_order.clear();
    for (size_t i = 0; i < n; i++)
        if (terminal[i]) traverse(g, (int)i, seen);
    g->position.assign(n, -1);
    for (size_t i = 0; i < g->run_order.size(); i++) g->position[g->run_order[i]] = (int)i;

    // an input sees its producer's buffer only if the producer ran EARLIER in this tick
    // (engine.rs:479-482: connections.get(..).and_then(|o| buffers.get(o)) else Disconnected)
    g->resolved.assign(n, {});
    std::vector<int> level(n, 0);
    for (int id : g->run_order) {
        mxl_module* m = g->modules[id];
        g->resolved[id].assign(m->inputs.size(), {-1, 0u});
        int lv = 0;
        for (uint32_t i = 0; i < m->inputs.size(); i++) {
            auto it = g->connections.find({id, i});
            if (it == g->connections.end()) continue;
            const int pm = it->second.first;
            if (!module_at(g, pm) || g->position[pm] < 0 || g->position[pm] >= g->position[id]) continue;
            g->resolved[id][i] = it->second;
            lv = std::max(lv, level[pm] + 1);
        }
        level[id] = lv;
    }

    // graph-owned output lines (sources present their own line)
    if (g->out_lines.size() < n) g->out_lines.resize(n);
    for (size_t id = 0; id < n; id++) {
        mxl_module* m = g->modules[id];
        const size_t want = (m && !is_source(m->kind) && g->position[id] >= 0) ? m->outputs.size() : 0;
        std::vector<mxl_line*>& v = g->out_lines[id];
        bool same = v.size() == want;
        for (size_t o = 0; same && o < want; o++) same = v[o]->type == m->outputs[o].type;
        if (same) continue;
        for (mxl_line* l : v) line_free(l);
        v.clear();
        if (!g->ctx->has_device()) continue;          // planning-only context: no buffers
        for (size_t o = 0; o < want; o++) {
            mxl_line* l = line_alloc(g->ctx, m->outputs[o].type, 0);
            if (!l) return MXL_ERR_OOM;
            v.push_back(l);
        }
    }
    // fused groups first: their modules leave the per-kind stages
    std::vector<int> owner;
    find_fused_groups(g, level, owner);

    // stages by (level, kind)
    for (auto& s : g->stages) {
        if (s.ev0) cudaEventDestroy(s.ev0);
        if (s.ev1) cudaEventDestroy(s.ev1);
    }
    g->stages.clear();
    std::map<std::pair<int, int>, size_t> index;
    for (int id : g->run_order) {
        if (owner[id] >= 0) continue;
        const int kind = g->modules[id]->kind;
        auto key = std::make_pair(level[id], kind);
        auto it = index.find(key);
        if (it == index.end()) {
            index[key] = g->stages.size();
            Stage s;
            s.level = level[id];
            s.kind = kind;
            g->stages.push_back(s);
            it = index.find(key);
        }
        g->stages[it->second].modules.push_back(id);
    }
    for (size_t f = 0; f < g->fused.size(); f++) {                 // one stage per group, where its mixer stood
        Stage s;
        s.level = level[g->fused[f].members[0]];
        s.kind = MXL_STAGE_FUSED_VOICE_MIX;
        s.fused = (int)f;
        s.modules = g->fused[f].members;
        g->stages.push_back(s);
    }
    std::stable_sort(g->stages.begin(), g->stages.end(), [](const Stage& a, const Stage& b) {
        return a.level != b.level ? a.level < b.level : a.kind < b.kind;
    });

    g->resize_list.clear();
    for (int id : g->run_order)
        for (size_t o = 0; o < g->out_lines[id].size(); o++)
            if (!g->hidden.count({id, (uint32_t)o})) g->resize_list.push_back(g->out_lines[id][o]);
    g->ctx->change_epoch++;
    g->dirty = false;
    g->timings_pending = false;
    g->runs_since_plan = 0;
    return MXL_OK;
}

mxl_line* output_line(mxl_graph* g, int id, uint32_t out)
{
    mxl_module* m = module_at(g, id);
    if (!m || out >= m->outputs.size() || g->position[id] < 0) return nullptr;
    if (is_source(m->kind)) return source_line(m);
    return out < g->out_lines[id].size() ? g->out_lines[id][out] : nullptr;
}

void collect_timings(mxl_graph* g)
{
    if (!g->timings_pending) return;
    for (auto& s : g->stages) {
        if (s.ev0 && s.ev1 && cudaEventSynchronize(s.ev1) == cudaSuccess) {
            float ms = -1.f;
            if (cudaEventElapsedTime(&ms, s.ev0, s.ev1) == cudaSuccess) s.last_ms = ms;
        }
    }
    g->timings_pending = false;
}

}  // namespace

extern "C" {

mxl_graph* mxl_graph_create(mxl_ctx* ctx)
{
    if (!ctx) { set_error("mxl_graph_create: NULL context"); return nullptr; }
    mxl_graph* g = new mxl_graph();
    g->ctx = ctx;
    return g;
}

void mxl_graph_destroy(mxl_graph* g)
{
    if (!g) return;
    if (g->ctx->has_device()) { g->ctx->activate(); cudaStreamSynchronize(g->ctx->stream); }
    delete g;
}

int mxl_graph_add_module(mxl_graph* g, mxl_module* m)
{
    if (!g || !m) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if (m->ctx != g->ctx) MXL_FAIL(MXL_ERR_INVALID, "module belongs to another context");
    g->modules.push_back(m);
    g->dirty = true;
    return (int)g->modules.size() - 1;
}

int mxl_graph_remove_module(mxl_graph* g, int module_id)
{
    if (!g) MXL_FAIL(MXL_ERR_INVALID, "NULL graph");
    mxl_module* m = module_at(g, module_id);
    if (!m) MXL_FAIL(MXL_ERR_INVALID, "no module %d", module_id);
    if (g->ctx->has_device()) { MXL_TRY(g->ctx->activate()); MXL_CUDA(cudaStreamSynchronize(g->ctx->stream)); }
    // engine.rs:321-352 DeleteModule drops every connection touching the module
    for (auto it = g->connections.begin(); it != g->connections.end();) {
        if (it->first.first == module_id || it->second.first == module_id) it = g->connections.erase(it);
        else ++it;
    }
    delete m;
    g->modules[module_id] = nullptr;
    for (auto it = g->pinned.begin(); it != g->pinned.end();) it = it->first == module_id ? g->pinned.erase(it) : std::next(it);
    g->fusion_veto.erase(module_id);
    g->fused.clear();                       // the groups hold module pointers: rebuilt by the next plan
    g->dirty = true;
    return MXL_OK;
}

mxl_module* mxl_graph_module(mxl_graph* g, int module_id) { return g ? module_at(g, module_id) : nullptr; }

int mxl_graph_connect(mxl_graph* g, int in_module, uint32_t in_index, int out_module, uint32_t out_index)
{
    if (!g) MXL_FAIL(MXL_ERR_INVALID, "NULL graph");
    // workspace.rs:97-114
    mxl_module* im = module_at(g, in_module);
    if (!im || in_index >= im->inputs.size()) MXL_FAIL(MXL_ERR_NO_INPUT, "connect: no input %d:%u", in_module, in_index);
    mxl_module* om = module_at(g, out_module);
    if (!om || out_index >= om->outputs.size()) MXL_FAIL(MXL_ERR_NO_OUTPUT, "connect: no output %d:%u", out_module, out_index);
    if (im->inputs[in_index].type != om->outputs[out_index].type)
        MXL_FAIL(MXL_ERR_TYPE_MISMATCH, "connect: line type mismatch between input %d:%u and output %d:%u", in_module, in_index, out_module, out_index);
    g->connections[{in_module, in_index}] = {out_module, out_index};
    g->dirty = true;
    return MXL_OK;
}

int mxl_graph_disconnect(mxl_graph* g, int in_module, uint32_t in_index)
{
    if (!g) MXL_FAIL(MXL_ERR_INVALID, "NULL graph");
    g->connections.erase({in_module, in_index});       // workspace.rs:116-118
    g->dirty = true;
    return MXL_OK;
}

int mxl_graph_plan(mxl_graph* g, int* order_out, uint32_t cap)
{
    if (!g) MXL_FAIL(MXL_ERR_INVALID, "NULL graph");
    if (g->dirty) MXL_TRY(build_plan(g));
    for (size_t i = 0; i < g->run_order.size() && i < cap && order_out; i++) order_out[i] = g->run_order[i];
    return (int)g->run_order.size();
}

int mxl_graph_run_ticks(mxl_graph* g, uint64_t tick0, uint32_t n_ticks)
{
    if (!g) MXL_FAIL(MXL_ERR_INVALID, "NULL graph");
    mxl_ctx* ctx = g->ctx;
    if (!ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "mxl_graph_run_ticks: context has no CUDA device; there is no CPU fallback");
    if (n_ticks == 0) return MXL_OK;
    const auto call_t0 = std::chrono::steady_clock::now();
    struct CallTimer {
        mxl_graph* g; std::chrono::steady_clock::time_point t0; uint32_t ticks;
        ~CallTimer() { g->last_call_host_us = std::chrono::duration<float, std::micro>(std::chrono::steady_clock::now() - t0).count(); g->last_call_ticks = ticks; }
    } call_timer{g, call_t0, n_ticks};
    MXL_TRY(ctx->activate());
    if (g->checked_epoch != ctx->change_epoch) {                       // (nothing was updated, resized or re-planned since: skip)
        // a module whose params changed its terminals (Mixer::update re-creates itself, mixer.rs:40-44)
        for (size_t id = 0; id < g->modules.size() && !g->dirty; id++) {
            mxl_module* m = g->modules[id];
            if (m && !is_source(m->kind) && g->position.size() > id && g->position[id] >= 0 && g->out_lines[id].size() != m->outputs.size()) g->dirty = true;
            if (m && g->resolved.size() > id && g->position[id] >= 0 && g->resolved[id].size() != m->inputs.size()) g->dirty = true;
        }
        // a fused group whose parameters left the fused kernel's domain (update() since the plan) goes back to stages
        for (const FusedGroup& fg : g->fused)
            if (!g->dirty && !fused_group_params_ok(fg)) { g->fusion_veto.insert(fg.members[0]); g->dirty = true; }
    }
    if (g->dirty) MXL_TRY(build_plan(g));
    g->checked_epoch = ctx->change_epoch;
    collect_timings(g);
    g->runs_since_plan++;

    const uint64_t frames = (uint64_t)n_ticks * ctx->spt;
    const uint64_t t = tick0 * (uint64_t)ctx->spt;                     // engine.rs:490
    for (mxl_line* l : g->resize_list) MXL_TRY(line_resize(l, l->type == MXL_LINE_VIDEO ? n_ticks : frames));
    // host-fed sources must cover the call
    for (int id : g->run_order) {
        mxl_module* m = g->modules[id];
        if (!is_source(m->kind)) continue;
        mxl_line* l = source_line(m);
        if (l && l->frames < (l->type == MXL_LINE_VIDEO ? (uint64_t)n_ticks : frames))
            MXL_FAIL(MXL_ERR_LENGTH, "source module %d: line holds %llu frames/slots, the call needs %llu", id,
                     (unsigned long long)l->frames, (unsigned long long)(l->type == MXL_LINE_VIDEO ? (uint64_t)n_ticks : frames));
    }

    MXL_TRY(ctx->compute_begin());
    struct EndGuard { mxl_ctx* c; ~EndGuard() { c->compute_end(); } } end_guard{ctx};

    // fork: audio stages -> aux stream when the graph also has video stages to overlap them with
    // (modules with terminals of both kinds -- StreamInput, Monitor -- tie the two sub-graphs together: one stream then)
    bool has_audio = false, has_video = false, has_mixed = false;
    for (const Stage& s : g->stages) {
        if (is_source(s.kind)) continue;
        if (s.kind == MXL_MOD_STREAM_INPUT || s.kind == MXL_MOD_MONITOR || s.kind == MXL_MOD_STREAM_OUTPUT) has_mixed = true;
        else if (s.kind == MXL_MOD_VIDEO_MIXER) has_video = true;
        else has_audio = true;
    }
    // measured on a one-tick live call (the worst case for the fork/join's four API calls): 26.4 us per tick with
    // the split, 29.2 us without -- the overlap of the audio chain with the compositor still pays
    static const uint32_t split_min_ticks = getenv("MXL_SPLIT_MIN_TICKS") ? (uint32_t)atoi(getenv("MXL_SPLIT_MIN_TICKS")) : 1u;
    static const bool env_no_split = getenv("MXL_NO_STREAM_SPLIT") != nullptr;
    const bool split = g->split_streams && has_audio && has_video && !has_mixed && n_ticks >= split_min_ticks && !env_no_split;
    cudaStream_t main_stream = ctx->stream;
    // every exit path re-serialises the two streams: later main-stream work (downloads, the next call's line_resize)
    // must be ordered behind audio kernels already enqueued on the side stream, also when a stage fails half way
    struct StreamGuard {
        mxl_ctx* c; cudaStream_t s; bool forked = false;
        ~StreamGuard()
        {
            c->stream = s; c->pdl_hold = false;
            if (forked && cudaEventRecord(c->ev_join, c->stream_aux) == cudaSuccess) cudaStreamWaitEvent(s, c->ev_join, 0);
        }
    } stream_guard{ctx, main_stream};
    ctx->pdl_hold = split && n_ticks > 8;                            // see common.h: launch_chained
    if (split) {
        if (!ctx->stream_aux) {
            int lo = 0, hi = 0;
            MXL_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            MXL_CUDA(cudaStreamCreateWithPriority(&ctx->stream_aux, cudaStreamNonBlocking, hi));
            MXL_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
            MXL_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        }
        MXL_CUDA(cudaEventRecord(ctx->ev_fork, main_stream));
        MXL_CUDA(cudaStreamWaitEvent(ctx->stream_aux, ctx->ev_fork, 0));
        stream_guard.forked = true;
    }
    std::vector<mxl_module*> mods;
    std::vector<IoSet> ios;
    std::vector<const mxl_line*> in_ptrs;
    std::vector<mxl_line*> out_ptrs;
    for (Stage& s : g->stages) {
        if (is_source(s.kind)) { s.last_launches = 0; s.last_bytes = 0; continue; }
        if (s.fused >= 0) {                                            // one launch for the whole voice group
            ctx->stream = split ? ctx->stream_aux : main_stream;
            if (g->profiling) {
                if (!s.ev0) { MXL_CUDA(cudaEventCreate(&s.ev0)); MXL_CUDA(cudaEventCreate(&s.ev1)); }
                MXL_CUDA(cudaEventRecord(s.ev0, ctx->stream));
            }
            const uint64_t before = ctx->launches;
            uint64_t bytes = 0;
            const auto host_t0 = std::chrono::steady_clock::now();
            MXL_TRY(run_fused_group(ctx, g->fused[s.fused], t, &bytes));
            s.last_host_us = std::chrono::duration<float, std::micro>(std::chrono::steady_clock::now() - host_t0).count();
            s.last_launches = (int)(ctx->launches - before);
            s.last_bytes = bytes;
            if (g->profiling) MXL_CUDA(cudaEventRecord(s.ev1, ctx->stream));
            continue;
        }
        mods.clear(); ios.clear(); in_ptrs.clear(); out_ptrs.clear();
        size_t n_in_total = 0, n_out_total = 0;
        for (int id : s.modules) { n_in_total += g->modules[id]->inputs.size(); n_out_total += g->modules[id]->outputs.size(); }
        in_ptrs.reserve(n_in_total + 1);
        out_ptrs.reserve(n_out_total + 1);
        for (int id : s.modules) {
            mxl_module* m = g->modules[id];
            IoSet io{};
            io.in = in_ptrs.data() + in_ptrs.size();
            io.n_in = (uint32_t)m->inputs.size();
            for (uint32_t i = 0; i < io.n_in; i++) {
                const auto& r = g->resolved[id][i];
                const mxl_line* l = r.first >= 0 ? output_line(g, r.first, r.second) : nullptr;
                // a source line longer than the call is presented at the call's length
                in_ptrs.push_back(l);
            }
            io.out = out_ptrs.data() + out_ptrs.size();
            io.n_out = (uint32_t)m->outputs.size();
            for (mxl_line* l : g->out_lines[id]) out_ptrs.push_back(l);
            mods.push_back(m);
            ios.push_back(io);
        }
        ctx->stream = (split && s.kind != MXL_MOD_VIDEO_MIXER) ? ctx->stream_aux : main_stream;
        if (g->profiling) {
            if (!s.ev0) { MXL_CUDA(cudaEventCreate(&s.ev0)); MXL_CUDA(cudaEventCreate(&s.ev1)); }
            MXL_CUDA(cudaEventRecord(s.ev0, ctx->stream));
        }
        const uint64_t before = ctx->launches;
        uint64_t bytes = 0;
        const auto host_t0 = std::chrono::steady_clock::now();
        MXL_TRY(run_batch(ctx, s.kind, mods.data(), (int)mods.size(), t, ios.data(), &bytes));
        s.last_host_us = std::chrono::duration<float, std::micro>(std::chrono::steady_clock::now() - host_t0).count();
        s.last_launches = (int)(ctx->launches - before);
        s.last_bytes = bytes;
        if (g->profiling) MXL_CUDA(cudaEventRecord(s.ev1, ctx->stream));
    }
    ctx->stream = main_stream;              // (the guard joins the side stream before anything downstream of the run)
    g->timings_pending = g->profiling;
    return MXL_OK;
}

int mxl_graph_set_stream_split(mxl_graph* g, int enabled)
{
    if (!g) MXL_FAIL(MXL_ERR_INVALID, "NULL graph");
    g->split_streams = enabled != 0;
    return MXL_OK;
}

mxl_line* mxl_graph_output(mxl_graph* g, int module_id, uint32_t out_index)
{
    if (!g) { set_error("mxl_graph_output: NULL graph"); return nullptr; }
    if (g->dirty && build_plan(g) != MXL_OK) return nullptr;
    if (g->hidden.count({module_id, out_index})) {
        // a line inside a fused group that nothing observed so far: observed from now on
        const bool stale = g->runs_since_plan > 0;
        g->pinned.insert({module_id, out_index});
        g->dirty = true;
        if (stale) {
            set_error("output %d:%u was interior to a fused voice group in the last run and was not written; it is from the next run on "
                      "(call mxl_graph_pin_output or mxl_graph_output before running, or mxl_graph_set_fusion(g, 0))", module_id, out_index);
            return nullptr;
        }
        if (build_plan(g) != MXL_OK) return nullptr;
    }
    return output_line(g, module_id, out_index);
}

int mxl_graph_pin_output(mxl_graph* g, int module_id, uint32_t out_index)
{
    if (!g) MXL_FAIL(MXL_ERR_INVALID, "NULL graph");
    mxl_module* m = module_at(g, module_id);
    if (!m || out_index >= m->outputs.size()) MXL_FAIL(MXL_ERR_NO_OUTPUT, "pin: no output %d:%u", module_id, out_index);
    if (g->pinned.insert({module_id, out_index}).second) g->dirty = true;
    return MXL_OK;
}

int mxl_graph_set_fusion(mxl_graph* g, int enabled)
{
    if (!g) MXL_FAIL(MXL_ERR_INVALID, "NULL graph");
    if (g->fusion != (enabled != 0)) { g->fusion = enabled != 0; g->dirty = true; }
    return MXL_OK;
}

int mxl_graph_set_profiling(mxl_graph* g, int enabled)
{
    if (!g) MXL_FAIL(MXL_ERR_INVALID, "NULL graph");
    g->profiling = enabled != 0;
    return MXL_OK;
}

int mxl_graph_performance(mxl_graph* g, mxl_perf_account* out, uint32_t cap)
{
    if (!g || (cap && !out)) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if (g->dirty) MXL_TRY(build_plan(g));
    collect_timings(g);
    const float ticks = (float)(g->last_call_ticks ? g->last_call_ticks : 1);
    uint32_t n = 0;
    float staged_host = 0.f;
    for (const Stage& s : g->stages) staged_host += s.last_launches ? s.last_host_us : 0.f;
    if (n < cap) out[n] = mxl_perf_account{-1, -1, -1.f, std::max(0.f, g->last_call_host_us - staged_host) / ticks};   // PerformanceAccount::Engine
    n++;
    for (const Stage& s : g->stages) {
        if (is_source(s.kind) || s.modules.empty()) continue;
        const float share = 1.f / (float)s.modules.size() / ticks;
        for (int id : s.modules) {
            // (a fused voice group's stage lists every module it replaced: each gets an equal share, under its own kind)
            if (n < cap) out[n] = mxl_perf_account{id, g->modules[id] ? g->modules[id]->kind : s.kind, s.last_ms >= 0.f ? s.last_ms * 1000.f * share : -1.f, s.last_host_us * share};
            n++;
        }
    }
    return (int)std::min<uint32_t>(n, cap);
}

int mxl_graph_stage_count(mxl_graph* g)
{
    if (!g) MXL_FAIL(MXL_ERR_INVALID, "NULL graph");
    if (g->dirty) MXL_TRY(build_plan(g));
    return (int)g->stages.size();
}

int mxl_graph_stage_info(mxl_graph* g, uint32_t stage, mxl_stage_info* out)
{
    if (!g || !out) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if (g->dirty) MXL_TRY(build_plan(g));
    if (stage >= g->stages.size()) MXL_FAIL(MXL_ERR_INVALID, "stage %u out of %zu", stage, g->stages.size());
    collect_timings(g);
    const Stage& s = g->stages[stage];
    out->kind = s.kind;
    out->n_modules = (int32_t)s.modules.size();
    out->n_launches = s.last_launches;
    out->last_ms = s.last_ms;
    out->algorithmic_bytes = s.last_bytes;
    out->host_us = s.last_host_us;
    out->_pad = 0.f;
    return MXL_OK;
}

}  // extern "C"

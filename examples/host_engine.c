/* host_engine.c -- the C ABI used from plain C, the way a cgo / Rust-FFI host would use it: BASELINE config 1
 * (4 stereo sources -> Mixer -> Amplifier) run tick by tick twice over,
 *   (a) with the engine's own host slices, one mxl_module_run_tick_host per module per tick (the reference's
 *       dispatch, src/engine.rs:461-494), and
 *   (b) as a graph, all ticks in one mxl_graph_run_ticks,
 * and the two results compared bit for bit.  Prints a checksum; exit status 0 = both paths agree.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/host_engine.c -Lmixlab_b200 -lmixlab_b200 -Wl,-rpath,$PWD/mixlab_b200 -o host_engine
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mixlab_b200.h"

#define SPT 800
#define TICKS 12
#define CHANNELS 4

#define CHECK(expr)                                                                      \
    do {                                                                                 \
        int st_ = (expr);                                                                \
        if (st_ < 0) { fprintf(stderr, "%s -> %d: %s\n", #expr, st_, mxl_last_error()); return 1; } \
    } while (0)

static uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main(int argc, char **argv)
{
    const int plan_only = argc > 1 && !strcmp(argv[1], "--no-device");
    mxl_ctx *ctx = mxl_ctx_create(plan_only ? MXL_DEVICE_NONE : 0, 48000, SPT);
    if (!ctx) { fprintf(stderr, "mxl_ctx_create: %s\n", mxl_last_error()); return 1; }

    const mxl_mixer_channel_params ch[CHANNELS] = {{0.0, 1.0, 0, 0}, {-6.0, 0.8, 1, 0}, {3.0, 0.5, 0, 0}, {-12.0, 0.25, 1, 0}};
    const mxl_mixer_params mp = {ch, CHANNELS};
    const mxl_amplifier_params ap = {0.9, 0.5};
    mxl_module *mixer = mxl_module_create(ctx, MXL_MOD_MIXER, &mp);
    mxl_module *amp = mxl_module_create(ctx, MXL_MOD_AMPLIFIER, &ap);
    if (!mixer || !amp) { fprintf(stderr, "mxl_module_create: %s\n", mxl_last_error()); return 1; }
    printf("mixer: %u inputs, %u outputs (%s, %s)\n", mxl_module_n_inputs(mixer), mxl_module_n_outputs(mixer),
           mxl_module_output_label(mixer, 0), mxl_module_output_label(mixer, 1));
    if (plan_only) {                                   /* no GPU here: compute calls must refuse, not fall back */
        mxl_host_ref none;
        memset(&none, 0, sizeof none);
        int st = mxl_module_run_tick_host(amp, 0, &none, 0, &none, 0);
        printf("compute without a device -> %d (%s)\n", st, mxl_last_error());
        mxl_module_destroy(mixer); mxl_module_destroy(amp); mxl_ctx_destroy(ctx);
        return st == MXL_ERR_NO_DEVICE ? 0 : 1;
    }

    /* synthetic sources */
    const size_t n = (size_t)2 * SPT * TICKS;
    float *src[CHANNELS], *control = malloc(sizeof(float) * SPT * TICKS);
    for (int c = 0; c < CHANNELS; c++) {
        uint64_t s = 1 + (uint64_t)c;
        src[c] = malloc(sizeof(float) * n);
        for (size_t i = 0; i < n; i++) src[c][i] = (float)((double)(splitmix64(&s) >> 40) / (double)(1 << 23) - 1.0);
    }
    { uint64_t s = 5; for (size_t i = 0; i < (size_t)SPT * TICKS; i++) control[i] = (float)((double)(splitmix64(&s) >> 40) / (double)(1 << 24)); }

    /* (a) host slices, module by module, tick by tick */
    float *master = malloc(sizeof(float) * 2 * SPT), *cue = malloc(sizeof(float) * 2 * SPT), *out_a = malloc(sizeof(float) * n);
    for (int k = 0; k < TICKS; k++) {
        mxl_host_ref in[CHANNELS], out[2], ain[2], aout[1];
        memset(in, 0, sizeof in); memset(out, 0, sizeof out); memset(ain, 0, sizeof ain); memset(aout, 0, sizeof aout);
        for (int c = 0; c < CHANNELS; c++) {
            in[c].type = MXL_LINE_STEREO; in[c].connected = 1; in[c].samples = src[c] + (size_t)2 * SPT * k; in[c].len = 2 * SPT;
        }
        out[0].type = out[1].type = MXL_LINE_STEREO;
        out[0].samples = master; out[1].samples = cue; out[0].len = out[1].len = 2 * SPT;
        CHECK(mxl_module_run_tick_host(mixer, (uint64_t)k * SPT, in, CHANNELS, out, 2));
        ain[0].type = MXL_LINE_STEREO; ain[0].connected = 1; ain[0].samples = master; ain[0].len = 2 * SPT;
        ain[1].type = MXL_LINE_MONO; ain[1].connected = 1; ain[1].samples = control + (size_t)SPT * k; ain[1].len = SPT;
        aout[0].type = MXL_LINE_STEREO; aout[0].samples = out_a + (size_t)2 * SPT * k; aout[0].len = 2 * SPT;
        CHECK(mxl_module_run_tick_host(amp, (uint64_t)k * SPT, ain, 2, aout, 1));
    }

    /* (b) the same modules' twins in a graph, all ticks in one call */
    mxl_graph *g = mxl_graph_create(ctx);
    mxl_line *lines[CHANNELS + 1];
    int ids[CHANNELS + 1];
    for (int c = 0; c <= CHANNELS; c++) {
        const int stereo = c < CHANNELS;
        mxl_module *s = mxl_module_create(ctx, stereo ? MXL_MOD_SOURCE_STEREO : MXL_MOD_SOURCE_MONO, NULL);
        lines[c] = mxl_line_alloc(ctx, stereo ? MXL_LINE_STEREO : MXL_LINE_MONO, (uint64_t)SPT * TICKS);
        if (!s || !lines[c]) { fprintf(stderr, "source: %s\n", mxl_last_error()); return 1; }
        CHECK(mxl_line_upload(lines[c], stereo ? src[c] : control, stereo ? n : (size_t)SPT * TICKS));
        CHECK(mxl_source_set_line(s, lines[c]));
        CHECK(ids[c] = mxl_graph_add_module(g, s));
    }
    int gm, ga;
    CHECK(gm = mxl_graph_add_module(g, mxl_module_create(ctx, MXL_MOD_MIXER, &mp)));
    CHECK(ga = mxl_graph_add_module(g, mxl_module_create(ctx, MXL_MOD_AMPLIFIER, &ap)));
    for (int c = 0; c < CHANNELS; c++) CHECK(mxl_graph_connect(g, gm, (uint32_t)c, ids[c], 0));
    CHECK(mxl_graph_connect(g, ga, 0, gm, 0));
    CHECK(mxl_graph_connect(g, ga, 1, ids[CHANNELS], 0));
    if (mxl_graph_connect(g, ga, 0, ids[CHANNELS], 0) != MXL_ERR_TYPE_MISMATCH) {      /* mono into a stereo input: workspace.rs:108-113 */
        fprintf(stderr, "type mismatch was not refused\n");
        return 1;
    }
    CHECK(mxl_graph_run_ticks(g, 0, TICKS));
    float *out_b = malloc(sizeof(float) * n);
    CHECK(mxl_line_download(mxl_graph_output(g, ga, 0), out_b, n));

    uint64_t sum = 0;
    for (size_t i = 0; i < n; i++) { uint32_t u; memcpy(&u, &out_b[i], 4); sum = sum * 1099511628211ull + u; }
    const int same = memcmp(out_a, out_b, sizeof(float) * n) == 0;
    printf("%d ticks, %zu samples, host-slice path %s graph path, checksum %016llx, %llu kernel launches\n", TICKS, n,
           same ? "==" : "!=", (unsigned long long)sum, (unsigned long long)mxl_ctx_launch_count(ctx));

    mxl_graph_destroy(g);
    for (int c = 0; c <= CHANNELS; c++) mxl_line_free(lines[c]);
    mxl_module_destroy(mixer); mxl_module_destroy(amp);
    mxl_ctx_destroy(ctx);
    return same ? 0 : 2;
}

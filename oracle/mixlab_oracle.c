/*
 * mixlab_oracle.c -- CPU ORACLE (test infrastructure, never shipped, never on the product path).
 * See mixlab_oracle.h for the pinning status.  Citations are /root/reference-relative.
 *
 * Rules followed everywhere: f64 exactly where the Rust widens; no FMA
 * contraction (build with -ffp-contract=off); Rust `as i16` / `as u8` float casts saturate,
 * truncate toward zero and map NaN to 0; integer /255 truncates; f32 accumulation in channel
 * order.
 */
#include "mixlab_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* std::f64::consts::PI */
#define ORC_PI 3.14159265358979323846264338327950288

/* ======================================================================================== */
/* scalars                                                                                  */
/* ======================================================================================== */

double orc_db_to_linear(double db)
{
    /* protocol/src/lib.rs:469-471 */
    return pow(10.0, db / 20.0);
}

/* ======================================================================================== */
/* audio                                                                                    */
/* ======================================================================================== */

void orc_mixer(const float *const *inputs, const double *gain_db, const double *fader,
               const uint8_t *cue, int channels, float *master, float *cue_out, size_t len)
{
    /* mixer.rs:54-55  util::zero (util.rs:26-30) */
    for (size_t i = 0; i < len; i++) master[i] = 0.0f;
    for (size_t i = 0; i < len; i++) cue_out[i] = 0.0f;

    /* mixer.rs:57-68 */
    for (int ch = 0; ch < channels; ch++) {
        const float *input = inputs[ch];
        double channel_gain = fader[ch] * orc_db_to_linear(gain_db[ch]);   /* mixer.rs:59 */
        for (size_t i = 0; i < len; i++) {
            float in = input ? input[i] : 0.0f;                             /* io.rs:44-46 */
            master[i] += (float)((double)in * channel_gain);                /* mixer.rs:62 */
            if (cue[ch]) cue_out[i] += in;                                  /* mixer.rs:64-66 */
        }
    }
}

/* amplifier.rs:71-73 */
static double depth(double value, double d) { return 1.0 - d + d * value; }

void orc_amplifier(const float *input, const float *mod, double amplitude, double mod_depth,
                   float *output, size_t len)
{
    /* amplifier.rs:52-57 */
    for (size_t i = 0; i < len; i++) {
        double mod_value = mod ? (double)mod[i / 2] : 1.0;
        output[i] = (float)((double)input[i] * depth(mod_value, mod_depth) * amplitude);
    }
}

#define EQ_FREQ_LO 420.0                    /* eq_three.rs:8  */
#define EQ_FREQ_HI 2700.0                   /* eq_three.rs:9  */
#define EQ_VSA (1.0 / 4294967295.0)         /* eq_three.rs:11 */

void orc_eq_three_create(orc_eq_three *eq, double sample_rate)
{
    memset(eq, 0, sizeof *eq);
    /* eq_three.rs:117-119  set_freq */
    eq->lo_coef = 2.0 * sin(ORC_PI * EQ_FREQ_LO / sample_rate);
    eq->hi_coef = 2.0 * sin(ORC_PI * EQ_FREQ_HI / sample_rate);
}

/* eq_three.rs:121-128 */
static inline double lowpass_pump(double coef, double poles[4], double sample)
{
    poles[0] += coef * (sample - poles[0]) + EQ_VSA;
    poles[1] += coef * (poles[0] - poles[1]);
    poles[2] += coef * (poles[1] - poles[2]);
    poles[3] += coef * (poles[2] - poles[3]);
    return poles[3];
}

void orc_eq_three_run(orc_eq_three *eq, double gain_lo_db, double gain_mid_db, double gain_hi_db,
                      const float *input, float *output, size_t n)
{
    /* eq_three.rs:62-64 */
    double gain_lo = orc_db_to_linear(gain_lo_db);
    double gain_mid = orc_db_to_linear(gain_mid_db);
    double gain_hi = orc_db_to_linear(gain_hi_db);

    /* eq_three.rs:66-86 */
    for (size_t i = 0; i < n; i++) {
        double sample = (double)input[i];
        double lo = lowpass_pump(eq->lo_coef, eq->lo_poles, sample);
        double hi = eq->history[0] - lowpass_pump(eq->hi_coef, eq->hi_poles, sample);
        double mid = eq->history[0] - (hi + lo);
        eq->history[0] = eq->history[1];
        eq->history[1] = eq->history[2];
        eq->history[2] = sample;
        lo = lo * gain_lo;
        mid = mid * gain_mid;
        hi = hi * gain_hi;
        output[i] = (float)(lo + mid + hi);
    }
}

/* oscillator.rs:15-23: decided on the sign BIT, so -0.0 -> -1.0 and the 0.0 arm is dead */
static double osc_sign(double n) { return signbit(n) ? -1.0 : 1.0; }
/* oscillator.rs:25-27 */
static double osc_sine(double n) { return sin(n * 2.0 * ORC_PI); }
/* oscillator.rs:30-32 */
static double osc_saw(double n) { return 2.0 * (n - floor(0.5 + n)); }
/* oscillator.rs:35-37 */
static double osc_triangle(double n) { return 2.0 * fabs(osc_saw(n)) - 1.0; }

void orc_oscillator(uint64_t t, double sample_rate, double freq, int waveform,
                    float *mono, float *stereo, size_t n)
{
    /* oscillator.rs:73-89 */
    for (size_t i = 0; i < n; i++) {
        double t0 = (double)(t + (uint64_t)i) / sample_rate;
        double x = t0 * freq;
        double v;
        switch (waveform) {
        case ORC_WAVE_SINE: v = osc_sine(x); break;
        case ORC_WAVE_SQUARE: v = osc_sign(osc_sine(x)); break;
        case ORC_WAVE_SAW: v = osc_saw(x); break;
        case ORC_WAVE_TRIANGLE: v = osc_triangle(x); break;
        case ORC_WAVE_ON: v = 1.0; break;
        default: v = 0.0; break;
        }
        float sample = (float)v;
        if (mono) mono[i] = sample;
        if (stereo) { stereo[i * 2 + 0] = sample; stereo[i * 2 + 1] = sample; }
    }
}

void orc_envelope_create(orc_envelope *env)
{
    env->state = ORC_ENV_INITIAL;   /* envelope.rs:76 */
    env->seq = 0;
    env->off_amplitude = 0.0;
}

/* envelope.rs:16-18 */
static double seq_duration_ms(uint64_t first, uint64_t last, double sample_rate)
{
    return (double)(last - first) / sample_rate * 1000.0;
}
/* envelope.rs:20-28 */
static double env_clamp(double x) { return x > 1.0 ? 1.0 : (x < 0.0 ? 0.0 : x); }

/* envelope.rs:34-58 */
static double env_amplitude(const orc_envelope *env, uint64_t t, double sr, double attack_ms,
                            double decay_ms, double sustain, double release_ms)
{
    switch (env->state) {
    case ORC_ENV_ON: {
        double ms_since_on = seq_duration_ms(env->seq, t, sr);
        if (ms_since_on < attack_ms) {
            return 1.0 / attack_ms * ms_since_on;
        } else {
            double ms_since_decay_started = ms_since_on - attack_ms;
            double decay_amplitude = 1.0 - env_clamp(1.0 / decay_ms * ms_since_decay_started);
            return sustain + ((1.0 - sustain) * decay_amplitude);
        }
    }
    case ORC_ENV_OFF: {
        double ms_since_off = seq_duration_ms(env->seq, t, sr);
        double release_amplitude = 1.0 - env_clamp(1.0 / release_ms * ms_since_off);
        return env->off_amplitude * release_amplitude;
    }
    default:
        return 0.0;
    }
}

void orc_envelope_run(orc_envelope *env, uint64_t t, double sample_rate, double attack_ms,
                      double decay_ms, double sustain_amplitude, double release_ms,
                      const float *input, float *output, size_t n)
{
    /* envelope.rs:96-117 */
    for (size_t i = 0; i < n; i++) {
        uint64_t sample_seq = t + (uint64_t)i;
        if (env->state == ORC_ENV_ON) {
            if (input[i] == 0.0f) {
                double amp = env_amplitude(env, sample_seq, sample_rate, attack_ms, decay_ms,
                                           sustain_amplitude, release_ms);
                env->state = ORC_ENV_OFF;
                env->seq = sample_seq;
                env->off_amplitude = amp;
            }
        } else {
            if (input[i] == 1.0f) {
                env->state = ORC_ENV_ON;
                env->seq = sample_seq;
            }
        }
        output[i] = (float)env_amplitude(env, sample_seq, sample_rate, attack_ms, decay_ms,
                                         sustain_amplitude, release_ms);
    }
}

void orc_fm_sine(uint64_t t, double sample_rate, double freq_lo, double freq_hi,
                 const float *input, float *output, size_t n)
{
    /* fm_sine.rs:42-43 */
    double freq_amp = (freq_hi - freq_lo) / 2.0;
    double freq_mid = freq_lo + freq_amp;
    /* fm_sine.rs:45-53 */
    for (size_t i = 0; i < n; i++) {
        double ts = (double)(t + (uint64_t)i) / sample_rate;
        double in = input ? (double)input[i] : 0.0;
        double co = (freq_mid + freq_amp * in) * 2.0 * ORC_PI;
        double x = sin(co * ts);
        output[i * 2 + 0] = (float)x;
        output[i * 2 + 1] = (float)x;
    }
}

void orc_stereo_panner(const float *left, const float *right, float *output, size_t n)
{
    /* stereo_panner.rs:35-38 */
    for (size_t i = 0; i < n; i++) {
        output[i * 2 + 0] = left ? left[i] : 0.0f;
        output[i * 2 + 1] = right ? right[i] : 0.0f;
    }
}

void orc_stereo_splitter(const float *input, float *left, float *right, size_t n)
{
    /* stereo_splitter.rs:41-44 */
    for (size_t i = 0; i < n; i++) {
        left[i] = input ? input[i * 2 + 0] : 0.0f;
        right[i] = input ? input[i * 2 + 1] : 0.0f;
    }
}

void orc_trigger(int open, float *output, size_t n)
{
    /* trigger.rs:38-45 */
    float value = open ? 1.0f : 0.0f;
    for (size_t i = 0; i < n; i++) output[i] = value;
}

void orc_pcm_pack_i16(const float *samples, int16_t *pcm, size_t len)
{
    /* src/video/encode.rs:184-195 */
    for (size_t i = 0; i < len; i++) {
        float s = samples[i];
        if (s > 1.0f) s = 1.0f; else if (s < -1.0f) s = -1.0f;
        float scaled = s * 32767.0f;            /* i16::max_value() as f32 */
        /* Rust `as i16`: NaN -> 0, saturate, truncate toward zero */
        int16_t v;
        if (scaled != scaled) v = 0;
        else if (scaled >= 32767.0f) v = 32767;
        else if (scaled <= -32768.0f) v = -32768;
        else v = (int16_t)scaled;
        pcm[i] = v;
    }
}

void orc_pcm_unpack_i16(const int16_t *pcm, float *samples, size_t len)
{
    /* stream_input.rs:167-173: divisor = -(i16::MIN as f32) = 32768.0 */
    for (size_t i = 0; i < len; i++) samples[i] = (float)pcm[i] / 32768.0f;
}

void orc_plotter_tap(const float *stereo, float *left, float *right, size_t n)
{
    /* plotter.rs:43-50 */
    for (size_t i = 0; i < n; i++) {
        left[i] = stereo[i * 2];
        right[i] = stereo[i * 2 + 1];
    }
}

int orc_clip_detect(const float *stereo, size_t len)
{
    /* output_device.rs:188-208 with both channels mapped */
    int clip = 0;
    for (size_t i = 0; i < len; i++)
        if (stereo[i] < -1.0f || stereo[i] > 1.0f) clip = 1;
    return clip;
}

void orc_meter(const float *stereo, size_t n, float peak[2], double sumsq[2], int *clip)
{
    peak[0] = peak[1] = 0.0f;
    sumsq[0] = sumsq[1] = 0.0;
    int c = 0;
    for (size_t i = 0; i < n; i++) {
        for (int ch = 0; ch < 2; ch++) {
            float s = stereo[i * 2 + ch];
            float a = fabsf(s);
            if (a > peak[ch]) peak[ch] = a;           /* NaN never raises the peak */
            sumsq[ch] += (double)s * (double)s;
            if (s < -1.0f || s > 1.0f) c = 1;
        }
    }
    *clip = c;
}

/* ======================================================================================== */
/* video                                                                                    */
/* ======================================================================================== */

static uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

void orc_frame_layout_yuv420p(uint32_t width, uint32_t height, orc_frame_layout *out)
{
    /* av_frame_get_buffer(frame, 0) (codec/src/ffmpeg/frame.rs:84-86; FFmpeg libavutil
     * frame.c get_video_buffer): luma width is padded until linesize[0] is a multiple of the
     * alignment (32), chroma linesize = ceil(padded/2) rounded up to 32. */
    uint32_t luma = align_up(width, 32);
    uint32_t chroma = align_up((luma + 1) / 2, 32);
    out->width = width;
    out->height = height;
    out->stride[0] = luma;
    out->stride[1] = chroma;
    out->stride[2] = chroma;
    out->plane_h[0] = height;
    out->plane_h[1] = (height + 1) / 2;   /* AV_CEIL_RSHIFT(h, 1) */
    out->plane_h[2] = (height + 1) / 2;
    size_t off = 0;
    for (int p = 0; p < 3; p++) {
        out->offset[p] = off;
        off += (size_t)out->stride[p] * out->plane_h[p];
    }
    out->size = off;
}

void orc_frame_blank(const orc_frame_layout *lay, uint8_t *data)
{
    /* frame.rs:94-135 */
    memset(data + lay->offset[0], 0x00, (size_t)lay->stride[0] * lay->plane_h[0]);
    memset(data + lay->offset[1], 0x80, (size_t)lay->stride[1] * lay->plane_h[1]);
    memset(data + lay->offset[2], 0x80, (size_t)lay->stride[2] * lay->plane_h[2]);
}

uint8_t orc_fader_to_u8(double fader)
{
    /* video_mixer.rs:168, Rust float->u8 `as` cast */
    double v = fader * 255.0;
    if (v != v) return 0;
    if (v >= 255.0) return 255;
    if (v <= 0.0) return 0;
    return (uint8_t)v;
}

/* video_mixer.rs:211-235  fade_line: 32 bytes per iteration, `while out < end` */
static void fade_line(uint8_t *out, const uint8_t *a, const uint8_t *b, size_t len, uint8_t fade)
{
    uint16_t a_fade = fade;
    uint16_t b_fade = (uint16_t)(255 - fade);
    uint8_t *end = out + len;
    while (out < end) {
        for (int k = 0; k < 32; k++) {
            uint16_t a_comp = (uint16_t)(a[k] * a_fade);
            uint16_t b_comp = (uint16_t)(b[k] * b_fade);
            out[k] = (uint8_t)((uint16_t)(a_comp + b_comp) / 255);
        }
        a += 32; b += 32; out += 32;
    }
}

void orc_video_crossfade(const orc_frame_layout *lay, const uint8_t *a, const uint8_t *b,
                         uint8_t fade, uint8_t *out)
{
    /* video_mixer.rs:171-237 */
    for (int comp = 0; comp < 3; comp++) {
        uint32_t shift = comp == 0 ? 0 : 1;                /* pixfmt.rs:134-150 */
        uint32_t width = lay->width >> shift;              /* video_mixer.rs:176 */
        uint32_t height = lay->height >> shift;            /* video_mixer.rs:177 */
        const uint8_t *a_ptr = (a ? a : out) + lay->offset[comp];
        const uint8_t *b_ptr = (b ? b : out) + lay->offset[comp];
        uint8_t *out_ptr = out + lay->offset[comp];
        size_t stride = lay->stride[comp];
        for (uint32_t y = 0; y < height; y++)
            fade_line(out_ptr + y * stride, a_ptr + y * stride, b_ptr + y * stride, width, fade);
    }
}

void orc_unify_picture(uint32_t aw, uint32_t ah, uint32_t bw, uint32_t bh, uint32_t *w, uint32_t *h)
{
    /* video_mixer.rs:276-297, yuv420p: log2_chroma_w = log2_chroma_h = 1 */
    uint32_t width = aw > bw ? aw : bw;
    uint32_t height = ah > bh ? ah : bh;
    *w = (width + 1u) & ~1u;
    *h = (height + 1u) & ~1u;
}

static uint64_t gcd_u64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }

void orc_scale_geometry_yuv420p(uint32_t in_w, uint32_t in_h, uint32_t out_w, uint32_t out_h,
                                orc_scale_geometry *g)
{
    /* src/video/encode.rs:355-374.  Ratio<usize> is exact; min by cross-multiplication. */
    uint64_t wn = out_w, wd = in_w, hn = out_h, hd = in_h;
    uint64_t sn, sd;
    if (wn * hd <= hn * wd) { sn = wn; sd = wd; } else { sn = hn; sd = hd; }
    uint64_t k = gcd_u64(sn, sd);
    if (k) { sn /= k; sd /= k; }
    uint32_t scaled_w = (uint32_t)((sn * in_w) / sd) & ~1u;   /* to_integer, align_horizontal */
    uint32_t scaled_h = (uint32_t)((sn * in_h) / sd) & ~1u;
    g->scaled_w = scaled_w;
    g->scaled_h = scaled_h;
    g->letterbox_x = ((out_w - scaled_w) / 2) & ~1u;
    g->letterbox_y = ((out_h - scaled_h) / 2) & ~1u;
}

static uint8_t clip_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

void orc_yuv420p_to_rgba(const orc_frame_layout *lay, const uint8_t *yuv, uint8_t *rgba)
{
    const uint8_t *yp = yuv + lay->offset[0], *up = yuv + lay->offset[1], *vp = yuv + lay->offset[2];
    for (uint32_t y = 0; y < lay->height; y++) {
        for (uint32_t x = 0; x < lay->width; x++) {
            int c = (int)yp[(size_t)y * lay->stride[0] + x] - 16;
            int d = (int)up[(size_t)(y >> 1) * lay->stride[1] + (x >> 1)] - 128;
            int e = (int)vp[(size_t)(y >> 1) * lay->stride[2] + (x >> 1)] - 128;
            uint8_t *px = rgba + ((size_t)y * lay->width + x) * 4;
            px[0] = clip_u8((298 * c + 409 * e + 128) >> 8);
            px[1] = clip_u8((298 * c - 100 * d - 208 * e + 128) >> 8);
            px[2] = clip_u8((298 * c + 516 * d + 128) >> 8);
            px[3] = 255;
        }
    }
}

void orc_rgba_to_yuv420p(const orc_frame_layout *lay, const uint8_t *rgba, uint8_t *yuv)
{
    uint8_t *yp = yuv + lay->offset[0], *up = yuv + lay->offset[1], *vp = yuv + lay->offset[2];
    const uint32_t w = lay->width, h = lay->height;
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            const uint8_t *px = rgba + ((size_t)y * w + x) * 4;
            yp[(size_t)y * lay->stride[0] + x] = (uint8_t)(((66 * px[0] + 129 * px[1] + 25 * px[2] + 128) >> 8) + 16);
        }
    for (uint32_t cy = 0; cy < (h + 1) / 2; cy++)
        for (uint32_t cx = 0; cx < (w + 1) / 2; cx++) {
            int sum[3] = {0, 0, 0};
            for (int dy = 0; dy < 2; dy++)
                for (int dx = 0; dx < 2; dx++) {
                    uint32_t yy = 2 * cy + dy, xx = 2 * cx + dx;
                    if (yy >= h) yy = h - 1;                         /* an odd edge repeats its last row / column */
                    if (xx >= w) xx = w - 1;
                    const uint8_t *px = rgba + ((size_t)yy * w + xx) * 4;
                    sum[0] += px[0]; sum[1] += px[1]; sum[2] += px[2];
                }
            const int r = (sum[0] + 2) >> 2, g = (sum[1] + 2) >> 2, b = (sum[2] + 2) >> 2;
            up[(size_t)cy * lay->stride[1] + cx] = (uint8_t)(((-38 * r - 74 * g + 112 * b + 128) >> 8) + 128);
            vp[(size_t)cy * lay->stride[2] + cx] = (uint8_t)(((112 * r - 94 * g - 18 * b + 128) >> 8) + 128);
        }
}

/* Keys cubic convolution kernel, a = -0.6 */
static double cubic_weight(double x)
{
    const double a = -0.6;
    x = fabs(x);
    if (x <= 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
    if (x < 2.0) return ((a * x - 5.0 * a) * x + 8.0 * a) * x - 4.0 * a;
    return 0.0;
}

/* 4-tap 14-bit coefficient table for one axis: tap k of dst index d reads src[clamp(pos[d]+k)] */
static void bicubic_table(uint32_t src_n, uint32_t dst_n, int32_t *pos, int16_t *coef)
{
    for (uint32_t d = 0; d < dst_n; d++) {
        int64_t num = (int64_t)(2 * (uint64_t)d + 1) * src_n - dst_n;   /* centre-aligned */
        int64_t den = 2 * (int64_t)dst_n;
        int64_t ix = num >= 0 ? num / den : -((-num + den - 1) / den);
        double frac = (double)(num - ix * den) / (double)den;
        int w[4], sum = 0, best = 0;
        for (int k = 0; k < 4; k++) {
            w[k] = (int)lrint(cubic_weight(frac - (double)(k - 1)) * 16384.0);
            sum += w[k];
            if (w[k] > w[best]) best = k;
        }
        w[best] += 16384 - sum;
        pos[d] = (int32_t)ix - 1;
        for (int k = 0; k < 4; k++) coef[d * 4 + k] = (int16_t)w[k];
    }
}

void orc_bicubic_plane(const uint8_t *src, uint32_t sw, uint32_t sh, uint32_t sstride,
                       uint8_t *dst, uint32_t dw, uint32_t dh, uint32_t dstride)
{
    int32_t *xpos = malloc(sizeof(int32_t) * dw), *ypos = malloc(sizeof(int32_t) * dh);
    int16_t *xco = malloc(sizeof(int16_t) * 4 * dw), *yco = malloc(sizeof(int16_t) * 4 * dh);
    uint8_t *tmp = malloc((size_t)dw * sh);
    bicubic_table(sw, dw, xpos, xco);
    bicubic_table(sh, dh, ypos, yco);
    for (uint32_t y = 0; y < sh; y++) {
        for (uint32_t x = 0; x < dw; x++) {
            int acc = 0;
            for (int k = 0; k < 4; k++) {
                int sx = xpos[x] + k;
                sx = sx < 0 ? 0 : (sx >= (int)sw ? (int)sw - 1 : sx);
                acc += xco[x * 4 + k] * (int)src[(size_t)y * sstride + sx];
            }
            tmp[(size_t)y * dw + x] = clip_u8((acc + 8192) >> 14);
        }
    }
    for (uint32_t y = 0; y < dh; y++) {
        for (uint32_t x = 0; x < dw; x++) {
            int acc = 0;
            for (int k = 0; k < 4; k++) {
                int sy = ypos[y] + k;
                sy = sy < 0 ? 0 : (sy >= (int)sh ? (int)sh - 1 : sy);
                acc += yco[y * 4 + k] * (int)tmp[(size_t)sy * dw + x];
            }
            dst[(size_t)y * dstride + x] = clip_u8((acc + 8192) >> 14);
        }
    }
    free(xpos); free(ypos); free(xco); free(yco); free(tmp);
}


/* ---- the scaler as a CPU implementation would arrange it (same arithmetic, bit-identical) ---- */
typedef struct { uint32_t src_n, dst_n; int32_t *pos; int16_t *coef; } tap_cache_entry;
#define TAP_CACHE 32
static __thread tap_cache_entry tap_cache[TAP_CACHE];
static __thread int tap_cache_n = 0, tap_cache_next = 0;

/* per-thread cache of tap tables by (source length, destination length); `keep` is never evicted */
static const tap_cache_entry *taps_for(uint32_t src_n, uint32_t dst_n, const tap_cache_entry *keep)
{
    for (int i = 0; i < tap_cache_n; i++)
        if (tap_cache[i].src_n == src_n && tap_cache[i].dst_n == dst_n) return &tap_cache[i];
    tap_cache_entry *e;
    if (tap_cache_n < TAP_CACHE) {
        e = &tap_cache[tap_cache_n++];
    } else {
        e = &tap_cache[tap_cache_next];
        if (e == keep) e = &tap_cache[(tap_cache_next + 1) % TAP_CACHE];
        tap_cache_next = (int)((e - tap_cache) + 1) % TAP_CACHE;
        free(e->pos); free(e->coef);
    }
    e->src_n = src_n; e->dst_n = dst_n;
    e->pos = malloc(sizeof(int32_t) * dst_n);
    e->coef = malloc(sizeof(int16_t) * 4 * dst_n);
    bicubic_table(src_n, dst_n, e->pos, e->coef);
    return e;
}

void orc_bicubic_plane_fast(const uint8_t *src, uint32_t sw, uint32_t sh, uint32_t sstride,
                            uint8_t *dst, uint32_t dw, uint32_t dh, uint32_t dstride)
{
    const tap_cache_entry *tx = taps_for(sw, dw, NULL), *ty = taps_for(sh, dh, tx);
    uint8_t *tmp = malloc((size_t)dw * sh);
    /* columns whose four taps need no clamp */
    uint32_t x_lo = 0, x_hi = dw;
    while (x_lo < dw && tx->pos[x_lo] < 0) x_lo++;
    while (x_hi > x_lo && tx->pos[x_hi - 1] + 3 >= (int)sw) x_hi--;
    for (uint32_t y = 0; y < sh; y++) {
        const uint8_t *row = src + (size_t)y * sstride;
        uint8_t *out = tmp + (size_t)y * dw;
        for (uint32_t x = 0; x < dw; x++) {
            const int16_t *c = tx->coef + 4 * x;
            int acc;
            if (x >= x_lo && x < x_hi) {
                const uint8_t *p = row + tx->pos[x];
                acc = c[0] * p[0] + c[1] * p[1] + c[2] * p[2] + c[3] * p[3];
            } else {
                acc = 0;
                for (int k = 0; k < 4; k++) {
                    int sx = tx->pos[x] + k;
                    sx = sx < 0 ? 0 : (sx >= (int)sw ? (int)sw - 1 : sx);
                    acc += c[k] * (int)row[sx];
                }
            }
            out[x] = clip_u8((acc + 8192) >> 14);
        }
    }
    for (uint32_t y = 0; y < dh; y++) {
        const uint8_t *r[4];
        for (int k = 0; k < 4; k++) {
            int sy = ty->pos[y] + k;
            sy = sy < 0 ? 0 : (sy >= (int)sh ? (int)sh - 1 : sy);
            r[k] = tmp + (size_t)sy * dw;
        }
        const int c0 = ty->coef[4 * y], c1 = ty->coef[4 * y + 1], c2 = ty->coef[4 * y + 2], c3 = ty->coef[4 * y + 3];
        uint8_t *out = dst + (size_t)y * dstride;
        for (uint32_t x = 0; x < dw; x++) {
            int acc = c0 * r[0][x] + c1 * r[1][x] + c2 * r[2][x] + c3 * r[3][x];
            acc = (acc + 8192) >> 14;
            out[x] = (uint8_t)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
        }
    }
    free(tmp);
}

void orc_letterbox_scale(const orc_frame_layout *li, const uint8_t *in, const orc_frame_layout *lo, uint8_t *out)
{
    if (li->width == lo->width && li->height == lo->height) {          /* encode.rs:342-345 */
        memcpy(out, in, lo->size);
        return;
    }
    orc_scale_geometry g;
    orc_scale_geometry_yuv420p(li->width, li->height, lo->width, lo->height, &g);
    orc_frame_blank(lo, out);                                           /* encode.rs:382 */
    for (int p = 0; p < 3; p++) {
        const int sh = p ? 1 : 0;
        const uint32_t sw = p ? (li->width + 1) / 2 : li->width;
        const uint32_t dw = g.scaled_w >> sh, dh = g.scaled_h >> sh;
        if (!dw || !dh) continue;
        orc_bicubic_plane_fast(in + li->offset[p], sw, li->plane_h[p], li->stride[p],
                               out + lo->offset[p] + (size_t)(g.letterbox_y >> sh) * lo->stride[p] + (g.letterbox_x >> sh),
                               dw, dh, lo->stride[p]);
    }
}

/* ======================================================================================== */
/* engine walker: Engine::run_tick, src/engine.rs:400-510                                   */
/* ======================================================================================== */

typedef struct {
    int kind;
    int n_in, n_out;
    int in_type[ORC_MAX_PORTS];
    int out_type[3];
    int conn_module[ORC_MAX_PORTS];     /* connections: InputId -> OutputId (workspace.rs:17) */
    int conn_output[ORC_MAX_PORTS];
    /* params */
    double p[8];
    int mixer_channels;
    double *mixer_gain_db, *mixer_fader;
    uint8_t *mixer_cue;
    /* state */
    orc_eq_three eq;
    orc_envelope env;
    uint64_t plotter_count;
    const float *source;
    size_t source_frames;
    float meter_peak[2];
    double meter_sumsq[2];
    int meter_clip;
    /* per-tick buffers (engine.rs:461 `buffers` map entries owned until the tick ends) */
    float *out_buf[3];
} orc_module;

struct orc_graph {
    double sample_rate;
    uint32_t spt;
    int n_modules, cap;
    orc_module **modules;
    int *order, n_order;
    /* static zero buffers, src/engine/io.rs:8-9 */
    float *zero_stereo, *zero_mono;
};

orc_graph *orc_graph_create(double sample_rate, uint32_t samples_per_tick)
{
    orc_graph *g = calloc(1, sizeof *g);
    g->sample_rate = sample_rate;
    g->spt = samples_per_tick;
    g->zero_stereo = calloc((size_t)samples_per_tick * 2, sizeof(float));
    g->zero_mono = calloc(samples_per_tick, sizeof(float));
    return g;
}

void orc_graph_destroy(orc_graph *g)
{
    if (!g) return;
    for (int i = 0; i < g->n_modules; i++) {
        orc_module *m = g->modules[i];
        free(m->mixer_gain_db); free(m->mixer_fader); free(m->mixer_cue);
        free(m);
    }
    free(g->modules); free(g->order); free(g->zero_stereo); free(g->zero_mono);
    free(g);
}

int orc_graph_add(orc_graph *g, int kind, const double *params, int n_params)
{
    orc_module *m = calloc(1, sizeof *m);
    m->kind = kind;
    for (int i = 0; i < ORC_MAX_PORTS; i++) m->conn_module[i] = -1;
    for (int i = 0; i < n_params && i < 8 && kind != ORC_MOD_MIXER; i++) m->p[i] = params[i];
    switch (kind) {
    case ORC_MOD_AMPLIFIER:        /* p: amplitude, mod_depth.  amplifier.rs:21-24 */
        m->n_in = 2; m->in_type[0] = ORC_LINE_STEREO; m->in_type[1] = ORC_LINE_MONO;
        m->n_out = 1; m->out_type[0] = ORC_LINE_STEREO; break;
    case ORC_MOD_ENVELOPE:         /* p: attack_ms, decay_ms, sustain, release_ms.  envelope.rs:77-78 */
        m->n_in = 1; m->in_type[0] = ORC_LINE_MONO; m->n_out = 1; m->out_type[0] = ORC_LINE_MONO;
        orc_envelope_create(&m->env); break;
    case ORC_MOD_EQ_THREE:         /* p: gain_lo_db, gain_mid_db, gain_hi_db.  eq_three.rs:42-43 */
        m->n_in = 1; m->in_type[0] = ORC_LINE_MONO; m->n_out = 1; m->out_type[0] = ORC_LINE_MONO;
        orc_eq_three_create(&m->eq, g->sample_rate); break;
    case ORC_MOD_FM_SINE:          /* p: freq_lo, freq_hi.  fm_sine.rs:22-23 */
        m->n_in = 1; m->in_type[0] = ORC_LINE_MONO; m->n_out = 1; m->out_type[0] = ORC_LINE_STEREO; break;
    case ORC_MOD_MIXER: {          /* params: C, then C x (gain_db, fader, cue).  mixer.rs:22-28 */
        int c = (int)params[0];
        m->mixer_channels = c;
        m->mixer_gain_db = malloc(sizeof(double) * (c ? c : 1));
        m->mixer_fader = malloc(sizeof(double) * (c ? c : 1));
        m->mixer_cue = malloc(c ? c : 1);
        for (int i = 0; i < c; i++) {
            m->mixer_gain_db[i] = params[1 + 3 * i];
            m->mixer_fader[i] = params[2 + 3 * i];
            m->mixer_cue[i] = params[3 + 3 * i] != 0.0;
            m->in_type[i] = ORC_LINE_STEREO;
        }
        m->n_in = c; m->n_out = 2; m->out_type[0] = m->out_type[1] = ORC_LINE_STEREO; break;
    }
    case ORC_MOD_OSCILLATOR:       /* p: freq, waveform.  oscillator.rs:47-51 */
        m->n_in = 0; m->n_out = 2; m->out_type[0] = ORC_LINE_MONO; m->out_type[1] = ORC_LINE_STEREO; break;
    case ORC_MOD_PLOTTER:          /* plotter.rs:22-23 */
    case ORC_MOD_METER:
        m->n_in = 1; m->in_type[0] = ORC_LINE_STEREO; m->n_out = 0; break;
    case ORC_MOD_STEREO_PANNER:    /* stereo_panner.rs:17-18 */
        m->n_in = 2; m->in_type[0] = m->in_type[1] = ORC_LINE_MONO;
        m->n_out = 1; m->out_type[0] = ORC_LINE_STEREO; break;
    case ORC_MOD_STEREO_SPLITTER:  /* stereo_splitter.rs:17-21 */
        m->n_in = 1; m->in_type[0] = ORC_LINE_STEREO;
        m->n_out = 2; m->out_type[0] = m->out_type[1] = ORC_LINE_MONO; break;
    case ORC_MOD_TRIGGER:          /* p: open.  trigger.rs:27-28 */
        m->n_in = 0; m->n_out = 1; m->out_type[0] = ORC_LINE_MONO; break;
    case ORC_MOD_SOURCE_STEREO:
        m->n_in = 0; m->n_out = 1; m->out_type[0] = ORC_LINE_STEREO; break;
    case ORC_MOD_SOURCE_MONO:
        m->n_in = 0; m->n_out = 1; m->out_type[0] = ORC_LINE_MONO; break;
    default:
        free(m); return -1;
    }
    if (g->n_modules == g->cap) {
        g->cap = g->cap ? g->cap * 2 : 16;
        g->modules = realloc(g->modules, sizeof(orc_module *) * g->cap);
        g->order = realloc(g->order, sizeof(int) * g->cap);
    }
    g->modules[g->n_modules] = m;
    return g->n_modules++;
}

int orc_graph_connect(orc_graph *g, int in_module, int in_index, int out_module, int out_index)
{
    /* workspace.rs:97-114 */
    if (in_module < 0 || in_module >= g->n_modules || in_index < 0 ||
        in_index >= g->modules[in_module]->n_in) return -1;
    if (out_module < 0 || out_module >= g->n_modules || out_index < 0 ||
        out_index >= g->modules[out_module]->n_out) return -2;
    if (g->modules[in_module]->in_type[in_index] != g->modules[out_module]->out_type[out_index]) return -3;
    g->modules[in_module]->conn_module[in_index] = out_module;
    g->modules[in_module]->conn_output[in_index] = out_index;
    return 0;
}

void orc_graph_set_source(orc_graph *g, int module, const float *data, size_t frames)
{
    g->modules[module]->source = data;
    g->modules[module]->source_frames = frames;
}

/* engine.rs:439-457 */
static void traverse(orc_graph *g, int id, uint8_t *seen)
{
    if (seen[id]) return;
    seen[id] = 1;
    orc_module *m = g->modules[id];
    for (int i = 0; i < m->n_in; i++)
        if (m->conn_module[i] >= 0) traverse(g, m->conn_module[i], seen);
    g->order[g->n_order++] = id;
}

static size_t line_len(const orc_graph *g, int type)
{
    return type == ORC_LINE_STEREO ? (size_t)g->spt * 2 : (type == ORC_LINE_MONO ? g->spt : 0);
}

void orc_graph_run_tick(orc_graph *g, uint64_t tick, int capture_module, int capture_output,
                        float *capture)
{
    int n = g->n_modules;
    /* engine.rs:408-416: terminal modules = modules that feed nobody */
    uint8_t *terminal = malloc(n ? n : 1), *seen = calloc(n ? n : 1, 1);
    memset(terminal, 1, n);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < g->modules[i]->n_in; k++)
            if (g->modules[i]->conn_module[k] >= 0) terminal[g->modules[i]->conn_module[k]] = 0;
    /* engine.rs:421-430.  The reference iterates a HashSet (arbitrary order); ascending id here. */
    g->n_order = 0;
    for (int i = 0; i < n; i++)
        if (terminal[i]) traverse(g, i, seen);

    uint64_t t = tick * (uint64_t)g->spt;                              /* engine.rs:490 */
    size_t S = g->spt;

    /* engine.rs:464-507 */
    for (int oi = 0; oi < g->n_order; oi++) {
        orc_module *m = g->modules[g->order[oi]];
        /* engine.rs:470-472 + io.rs:71-77: fresh zeroed Vec per output, every tick */
        for (int o = 0; o < m->n_out; o++)
            m->out_buf[o] = calloc(line_len(g, m->out_type[o]) + 1, sizeof(float));
        /* engine.rs:475-484: connected AND producer already ran this tick, else Disconnected */
        const float *in[ORC_MAX_PORTS];
        for (int k = 0; k < m->n_in; k++) {
            in[k] = NULL;
            if (m->conn_module[k] >= 0) {
                orc_module *src = g->modules[m->conn_module[k]];
                in[k] = src->out_buf[m->conn_output[k]];               /* NULL if not yet run */
            }
        }
        switch (m->kind) {
        case ORC_MOD_AMPLIFIER:
            /* expect_stereo of Disconnected = zero buffer (io.rs:44-46); control keeps None */
            orc_amplifier(in[0] ? in[0] : g->zero_stereo, in[1], m->p[0], m->p[1], m->out_buf[0], S * 2);
            break;
        case ORC_MOD_ENVELOPE:
            orc_envelope_run(&m->env, t, g->sample_rate, m->p[0], m->p[1], m->p[2], m->p[3],
                             in[0] ? in[0] : g->zero_mono, m->out_buf[0], S);
            break;
        case ORC_MOD_EQ_THREE:
            orc_eq_three_run(&m->eq, m->p[0], m->p[1], m->p[2], in[0] ? in[0] : g->zero_mono,
                             m->out_buf[0], S);
            break;
        case ORC_MOD_FM_SINE:
            orc_fm_sine(t, g->sample_rate, m->p[0], m->p[1], in[0] ? in[0] : g->zero_mono,
                        m->out_buf[0], S);
            break;
        case ORC_MOD_MIXER:
            orc_mixer(in, m->mixer_gain_db, m->mixer_fader, m->mixer_cue, m->mixer_channels,
                      m->out_buf[0], m->out_buf[1], S * 2);
            break;
        case ORC_MOD_OSCILLATOR:
            orc_oscillator(t, g->sample_rate, m->p[0], (int)m->p[1], m->out_buf[0], m->out_buf[1], S);
            break;
        case ORC_MOD_PLOTTER:
            /* plotter.rs:37-56 */
            m->plotter_count += 1;
            if (m->plotter_count % 6 == 0 && in[0]) {
                float *l = malloc(sizeof(float) * S), *r = malloc(sizeof(float) * S);
                orc_plotter_tap(in[0], l, r, S);
                free(l); free(r);
            }
            break;
        case ORC_MOD_METER:
            orc_meter(in[0] ? in[0] : g->zero_stereo, S, m->meter_peak, m->meter_sumsq, &m->meter_clip);
            break;
        case ORC_MOD_STEREO_PANNER:
            orc_stereo_panner(in[0], in[1], m->out_buf[0], S);
            break;
        case ORC_MOD_STEREO_SPLITTER:
            orc_stereo_splitter(in[0], m->out_buf[0], m->out_buf[1], S);
            break;
        case ORC_MOD_TRIGGER:
            orc_trigger(m->p[0] != 0.0, m->out_buf[0], S);
            break;
        case ORC_MOD_SOURCE_STEREO:
        case ORC_MOD_SOURCE_MONO: {
            size_t w = m->kind == ORC_MOD_SOURCE_STEREO ? 2 : 1;
            if (m->source && m->source_frames) {
                size_t base = (size_t)(t % m->source_frames);
                for (size_t i = 0; i < S; i++) {
                    size_t f = (base + i) % m->source_frames;
                    for (size_t c = 0; c < w; c++) m->out_buf[0][i * w + c] = m->source[f * w + c];
                }
            }
            break;
        }
        }
        if (capture && g->order[oi] == capture_module && capture_output < m->n_out)
            memcpy(capture, m->out_buf[capture_output],
                   sizeof(float) * line_len(g, m->out_type[capture_output]));
    }
    /* end of tick: the `buffers` map is dropped (engine.rs:461) */
    for (int i = 0; i < n; i++)
        for (int o = 0; o < 3; o++) { free(g->modules[i]->out_buf[o]); g->modules[i]->out_buf[o] = NULL; }
    free(terminal); free(seen);
}

int orc_graph_last_order(const orc_graph *g, int *order, int cap)
{
    int k = g->n_order < cap ? g->n_order : cap;
    memcpy(order, g->order, sizeof(int) * k);
    return g->n_order;
}

void orc_graph_meter(const orc_graph *g, int module, float peak[2], double sumsq[2], int *clip)
{
    const orc_module *m = g->modules[module];
    peak[0] = m->meter_peak[0]; peak[1] = m->meter_peak[1];
    sumsq[0] = m->meter_sumsq[0]; sumsq[1] = m->meter_sumsq[1];
    *clip = m->meter_clip;
}

/* ======================================================================================== */
/* audio sample-rate converter (definition: see the header)                                  */
/* ======================================================================================== */
struct orc_resampler {
    uint32_t L, M, channels;
    double *coef;                 /* [L][32] */
    float *x; size_t n_in, cap;   /* the whole input stream */
    size_t n_out;
};

static double rs_sinc(double u) { const double pi = 3.14159265358979323846264338327950288; return u == 0.0 ? 1.0 : sin(pi * u) / (pi * u); }
static double rs_bh(double u)
{
    const double pi = 3.14159265358979323846264338327950288;
    return 0.35875 + 0.48829 * cos(pi * u) + 0.14128 * cos(2.0 * pi * u) + 0.01168 * cos(3.0 * pi * u);
}

orc_resampler *orc_resampler_create(uint32_t in_rate, uint32_t out_rate, uint32_t channels)
{
    orc_resampler *r = calloc(1, sizeof *r);
    uint64_t g = gcd_u64(in_rate, out_rate);
    r->L = (uint32_t)(out_rate / g); r->M = (uint32_t)(in_rate / g); r->channels = channels;
    r->coef = malloc(sizeof(double) * 32 * r->L);
    const double ratio = (double)r->L / (double)r->M;
    const double w = 0.92 * (ratio < 1.0 ? ratio : 1.0);
    for (uint32_t ph = 0; ph < r->L; ph++) {
        double row[32], sum = 0.0;
        for (int k = 0; k < 32; k++) {
            const double x = (double)(k - 16 + 1) - (double)ph / (double)r->L;
            row[k] = w * rs_sinc(w * x) * rs_bh(x / 16.0);
            sum += row[k];
        }
        for (int k = 0; k < 32; k++) r->coef[(size_t)ph * 32 + k] = row[k] / sum;
    }
    return r;
}

void orc_resampler_destroy(orc_resampler *r) { if (r) { free(r->coef); free(r->x); free(r); } }
double orc_resampler_coef(const orc_resampler *r, uint32_t phase, uint32_t k) { return r->coef[(size_t)phase * 32 + k]; }
void orc_resampler_ratio(const orc_resampler *r, uint32_t *L, uint32_t *M) { *L = r->L; *M = r->M; }

size_t orc_resampler_push(orc_resampler *r, const float *in, size_t in_frames, float *out, size_t cap_frames)
{
    const size_t C = r->channels;
    if (r->n_in + in_frames > r->cap) {
        r->cap = (r->n_in + in_frames) * 2 + 1024;
        r->x = realloc(r->x, r->cap * C * sizeof(float));
    }
    memcpy(r->x + r->n_in * C, in, in_frames * C * sizeof(float));
    r->n_in += in_frames;
    const size_t total = r->n_in <= 16 ? 0 : (size_t)(((uint64_t)(r->n_in - 16) * r->L + r->M - 1) / r->M);
    size_t made = 0;
    for (size_t m = r->n_out; m < total && made < cap_frames; m++, made++) {
        const uint64_t pos = (uint64_t)m * r->M;
        const int64_t n0 = (int64_t)(pos / r->L);
        const double *c = r->coef + (size_t)(pos % r->L) * 32;
        for (size_t ch = 0; ch < C; ch++) {
            double acc = 0.0;
            for (int k = 0; k < 32; k++) {
                const int64_t n = n0 - 16 + 1 + k;
                const double xv = n < 0 ? 0.0 : (double)r->x[(size_t)n * C + ch];
                acc = fma(c[k], xv, acc);
            }
            out[made * C + ch] = (float)acc;
        }
    }
    r->n_out += made;
    return made;
}

/* ======================================================================================== */
/* one live session, tick after tick (see the header)                                        */
/* ======================================================================================== */
void orc_session_run(orc_session *s, uint64_t tick0, uint32_t n_ticks)
{
    const size_t spt = s->graph->spt, n = 2 * spt;
    for (uint32_t k = 0; k < n_ticks; k++) {
        const uint64_t tick = tick0 + k;
        /* StreamInput x2: the tick's i16 -> f32 (stream_input.rs:110-112) */
        for (int src = 0; src < 2; src++) {
            const size_t off = (size_t)((tick * 2 + (uint64_t)src) * n) % (s->pcm_in_samples - n + 1);
            orc_pcm_unpack_i16(s->pcm_in + off, s->scratch + (size_t)src * n, n);
        }
        /* Engine::run_tick over the audio graph; the master bus goes to the monitor */
        orc_graph_run_tick(s->graph, tick, s->master_module, 0, s->scratch + 2 * n);
        /* VideoMixer: the stored layers change every ticks_per_frame ticks; blank + crossfade every tick */
        const uint64_t fi = tick / s->ticks_per_frame;
        const uint8_t *a = s->layers_a + (size_t)(fi % s->n_layers) * s->lay.size;
        const uint8_t *b = s->layers_b + (size_t)(fi % s->n_layers) * s->lay.size;
        orc_frame_blank(&s->lay, s->composite);                        /* video_mixer.rs:151 */
        orc_video_crossfade(&s->lay, a, b, s->fade, s->composite);
        /* Monitor: scale to the monitor's size, pack the bus */
        orc_letterbox_scale(&s->lay, s->composite, &s->lay_mon, s->monitor_out);
        orc_pcm_pack_i16(s->scratch + 2 * n, s->pcm_out, n);
    }
}

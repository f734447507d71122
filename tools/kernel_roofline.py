#!/usr/bin/env python
"""Per-kernel HBM roofline: every audio/video module kernel alone through the C ABI on lines far larger
than L2, algorithmic bytes / CUDA-event time vs the measured copy peak."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mixlab_b200 as mxl
from mixlab_b200 import workloads as W


def timed(ctx, fn, reps):
    for _ in range(3):
        fn()
    ctx.synchronize()
    l0 = ctx.launch_count
    ctx.timer_begin()
    for _ in range(reps):
        fn()
    ctx.timer_end()
    ms = ctx.timer_elapsed_ms() / reps
    return ms, (ctx.launch_count - l0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1 << 25)      # 32 Mi frames: mono 128 MiB, stereo 256 MiB
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="", help="substring filter on kernel names")
    args = ap.parse_args()
    n = args.frames
    peak = 6549.8
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    ctx = mxl.Context(0, 48000, 800)
    base = W.uniform_pm1(7, 1 << 20)
    mono_h = np.resize(base, n)
    st_h = np.resize(base, 2 * n)
    gate_h = np.resize(np.where(np.arange(1 << 20) % 9000 < 4000, 1.0, 0.0).astype(np.float32), n)
    mono, mono2, stereo, gate = ctx.mono(mono_h), ctx.mono(mono_h), ctx.stereo(st_h), ctx.mono(gate_h)
    o_m, o_m2, o_s, o_s2 = ctx.line(mxl.LINE_MONO, n), ctx.line(mxl.LINE_MONO, n), ctx.line(mxl.LINE_STEREO, n), ctx.line(mxl.LINE_STEREO, n)
    S = n
    rows = []

    def run(name, kind, params, ins, outs, nbytes):
        if args.only and args.only not in name:
            return
        mod = ctx.module(kind, params)
        ms, launches = timed(ctx, lambda: mod.run_tick(0, ins, outs), args.reps)
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append({"kernel": name, "frames": n, "algorithmic_bytes": nbytes, "ms": ms, "gbs": gbs, "frac_of_measured_peak": gbs / peak,
                     "launches": launches})
        print(json.dumps(rows[-1]), flush=True)
        mod.destroy()

    run("Oscillator(sine)", mxl.MOD_OSCILLATOR, (440.0, mxl.WAVE_SINE, 0), [], [o_m, o_s], 12 * S)
    run("Oscillator(saw)", mxl.MOD_OSCILLATOR, (440.0, mxl.WAVE_SAW, 0), [], [o_m, o_s], 12 * S)
    run("EqThree", mxl.MOD_EQ_THREE, (4.0, 0.0, 4.0), [mono], [o_m], 8 * S)
    run("StereoPanner", mxl.MOD_STEREO_PANNER, None, [mono, mono2], [o_s], 16 * S)
    run("StereoSplitter", mxl.MOD_STEREO_SPLITTER, None, [stereo], [o_m, o_m2], 16 * S)
    run("Amplifier(+control)", mxl.MOD_AMPLIFIER, (0.9, 0.5), [stereo, mono], [o_s], 20 * S)
    run("Amplifier", mxl.MOD_AMPLIFIER, (0.9, 0.5), [stereo, None], [o_s], 16 * S)
    run("FmSine", mxl.MOD_FM_SINE, (90.0, 110.0), [mono], [o_s], 12 * S)
    run("Envelope", mxl.MOD_ENVELOPE, (25.0, 500.0, 0.8, 200.0), [gate], [o_m], 8 * S)
    run("Trigger", mxl.MOD_TRIGGER, (mxl.GATE_OPEN,), [], [o_m], 4 * S)
    run("Meter", mxl.MOD_METER, None, [stereo], [], 8 * S)
    run("PcmSink(pack i16)", mxl.MOD_PCM_SINK, None, [stereo], [], 12 * S)
    run("Mixer(2)", mxl.MOD_MIXER, [(0.0, 1.0, True), (-6.0, 0.5, False)], [stereo, o_s2], [o_s, ctx.line(mxl.LINE_STEREO, n)], 8 * S * 4)
    if not args.only or args.only in "Resampler":
        # audio sample-rate converter (self-specified): 44.1 kHz -> 48 kHz, stereo f32 line in, f32 line out
        rs = ctx.resampler(44100, 48000, 2)
        n_in = min(n, 1 << 24)
        src_line = ctx.line(mxl.LINE_STEREO, n_in)
        L = mxl.lib()
        def push():
            mxl.check(L.mxl_resampler_push_line(rs.h, src_line.h, rs.out.h))
        ms, launches = timed(ctx, push, args.reps)
        n_out = n_in * 160 // 147
        nb = 8 * n_in + 8 * n_out
        rows.append({"kernel": "Resampler 44.1k->48k stereo", "frames": n_in, "algorithmic_bytes": nb, "ms": ms, "gbs": nb / ms / 1e6,
                     "frac_of_measured_peak": nb / ms / 1e6 / peak, "launches": launches,
                     "note": "FP64-bound by design: 64 fma per output frame against 15.3 bytes of line traffic"})
        print(json.dumps(rows[-1]), flush=True)
        rs.close(); src_line.free()
    if not args.only or args.only in "FusedVoiceMix":
        # the fused voice group (BASELINE config 2's graph) on a call far larger than L2: 2^21 frames x 10 voices
        d = W.config2_graph()
        g, ids = W.build_graph(ctx, d)
        T2 = (1 << 21) // 800
        tick = [0]
        def step():
            g.run_ticks(tick[0], T2); tick[0] += T2
        ms, launches = timed(ctx, step, args.reps)
        nb = W.algorithmic_bytes_per_tick(d, 800) * T2
        rows.append({"kernel": "FusedVoiceMix (config 2 graph, %d ticks per call)" % T2, "frames": T2 * 800, "algorithmic_bytes": nb, "ms": ms,
                     "gbs": nb / ms / 1e6, "frac_of_measured_peak": nb / ms / 1e6 / peak, "launches": launches,
                     "compulsory_bytes": T2 * (16 * 800 + 32), "note": "API-level bytes of the 32 modules it replaces (464*S per tick); FP64-bound"})
        print(json.dumps(rows[-1]), flush=True)
        g.destroy()
    if args.only and args.only not in "VideoMixer NOAUDIO":
        ctx.close()
        return
    # video: 64 frames of 1080p per launch through the VideoMixer module
    T = 64
    fa = [ctx.frame(1920, 1080, blank=True) for _ in range(T)]
    fb = [ctx.frame(1920, 1080, blank=True) for _ in range(T)]
    la, lb = ctx.video_line(T), ctx.video_line(T)
    for k in range(T):
        la.set(k, fa[k], duration=(1, 60)); lb.set(k, fb[k], duration=(1, 60))
    outs = [ctx.video_line(T) for _ in range(3)]
    vm = ctx.module(mxl.MOD_VIDEO_MIXER, (0, 1, 0.5))
    ms, launches = timed(ctx, lambda: vm.run_tick(0, [la, lb, None, None], outs), args.reps)
    nb = T * W.CROSSFADE_BYTES_PER_FRAME
    print(json.dumps({"kernel": "VideoMixer crossfade 1080p x64", "algorithmic_bytes": nb, "ms": ms, "gbs": nb / ms / 1e6,
                      "frac_of_measured_peak": nb / ms / 1e6 / peak, "launches": launches}), flush=True)
    vm.update((0, -1, 0.5))
    ms, launches = timed(ctx, lambda: vm.run_tick(0, [la, lb, None, None], outs), args.reps)
    nb = T * 2 * W.FRAME_BYTES
    print(json.dumps({"kernel": "VideoMixer crossfade 1080p x64, layer B missing", "algorithmic_bytes": nb, "ms": ms, "gbs": nb / ms / 1e6,
                      "frac_of_measured_peak": nb / ms / 1e6 / peak, "launches": launches}), flush=True)
    # one-pass compositor: crossfade + yuv420p -> RGBA8 (BASELINE config 3 output ii), 64 frames per launch
    pics = ctx.rgba(1920, 1080, T)
    ms, launches = timed(ctx, lambda: ctx.compose_rgba(fa, fb, 0.5, pics), args.reps)
    nb = T * (2 * W.FRAME_BYTES + 1920 * 1080 * 4)
    print(json.dumps({"kernel": "compose_rgba 1080p x64 (2 layers -> RGBA8)", "algorithmic_bytes": nb, "ms": ms, "gbs": nb / ms / 1e6,
                      "frac_of_measured_peak": nb / ms / 1e6 / peak, "launches": launches}), flush=True)
    ms, launches = timed(ctx, lambda: ctx.frames_to_rgba(fa, pics), args.reps)
    nb = T * (W.FRAME_BYTES + 1920 * 1080 * 4)
    print(json.dumps({"kernel": "frames_to_rgba 1080p x64", "algorithmic_bytes": nb, "ms": ms, "gbs": nb / ms / 1e6,
                      "frac_of_measured_peak": nb / ms / 1e6 / peak, "launches": launches}), flush=True)
    # the other direction: RGBA8 -> yuv420p (self-specified), 64 pictures per launch
    ms, launches = timed(ctx, lambda: pics.to_frames(fa), args.reps)
    nb = T * (W.FRAME_BYTES + 1920 * 1080 * 4)
    print(json.dumps({"kernel": "rgba_to_yuv 1080p x64", "algorithmic_bytes": nb, "ms": ms, "gbs": nb / ms / 1e6,
                      "frac_of_measured_peak": nb / ms / 1e6 / peak, "launches": launches}), flush=True)
    pics.free()
    # tiled letterbox scaler: 32 frames per launch, 720p -> 1080p and 2160p -> 1080p (bytes: source read + scaled frame written)
    for (sw, sh) in ((1280, 720), (3840, 2160)):
        src = [ctx.frame(sw, sh, blank=True) for _ in range(32)]

        def scale_once():
            for f in ctx.frames_scale(src, 1920, 1080):
                f.release()
        ms, launches = timed(ctx, scale_once, args.reps)
        nb = 32 * (sw * sh * 3 // 2 + W.FRAME_BYTES)
        print(json.dumps({"kernel": "scale_tiled %dx%d -> 1080p x32" % (sw, sh), "algorithmic_bytes": nb, "ms": ms, "gbs": nb / ms / 1e6,
                          "frac_of_measured_peak": nb / ms / 1e6 / peak, "launches": launches}), flush=True)
        for f in src:
            f.release()
    ctx.close()


if __name__ == "__main__":
    main()

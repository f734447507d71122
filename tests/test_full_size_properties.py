"""GPU: properties that hold at ANY size, checked at BASELINE.json's full sizes (128 ticks of 1080p frames,
128-tick audio lines, the widest mixer bus of the config-5 sweep) where running the scalar oracle would take
minutes: identities, symmetries, batching invariance, round trips.  Everything through the C ABI."""
import numpy as np
import pytest

from mixlab_b200 import workloads as W

pytestmark = pytest.mark.gpu

SPT, TICKS = 800, 128


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def _mixer_lines(mxl, ctx, frames_list, ticks):
    line = ctx.video_line(ticks)
    for k in range(ticks):
        line.set(k, frames_list[k % len(frames_list)])
    return line


def _run_video(mxl, ctx, a_frames, b_frames, fader, ticks, swap=False):
    la = _mixer_lines(mxl, ctx, a_frames, ticks)
    lb = _mixer_lines(mxl, ctx, b_frames, ticks)
    outs = [ctx.video_line(ticks) for _ in range(3)]
    mod = ctx.module(mxl.MOD_VIDEO_MIXER, (0, 1, fader))
    mod.run_tick(0, [lb, la, None, None] if swap else [la, lb, None, None], outs)
    return [outs[0].get(k).download_raw() for k in range(ticks)]


def test_crossfade_identities_and_symmetry_128_frames_1080p(mxl, ctx48):
    """video_mixer.rs:211-235: f = 255 returns layer A, f = 0 layer B, and swapping the layers with the
    complementary fader gives the same picture -- 128 composited 1080p frames per launch."""
    data_a = [W.random_bytes(10 + i, W.FRAME_BYTES) for i in range(3)]
    data_b = [W.random_bytes(20 + i, W.FRAME_BYTES) for i in range(3)]
    fa = [ctx48.frame(W.FRAME_W, W.FRAME_H, d) for d in data_a]
    fb = [ctx48.frame(W.FRAME_W, W.FRAME_H, d) for d in data_b]
    for k, out in enumerate(_run_video(mxl, ctx48, fa, fb, 1.0, TICKS)):
        assert np.array_equal(out, data_a[k % 3]), k
    for k, out in enumerate(_run_video(mxl, ctx48, fa, fb, 0.0, TICKS)):
        assert np.array_equal(out, data_b[k % 3]), k
    # f = 63 with (A, B)  ==  f = 192 with (B, A):  a*63 + b*192 is the same sum
    assert mxl.fader_to_u8(0.25) == 63 and mxl.fader_to_u8(192.5 / 255.0) == 192
    x = _run_video(mxl, ctx48, fa, fb, 0.25, TICKS)
    y = _run_video(mxl, ctx48, fa, fb, 192.5 / 255.0, TICKS, swap=True)
    for k in range(TICKS):
        assert np.array_equal(x[k], y[k]), k
    # every byte of the blend lies between the two layers
    lo, hi = np.minimum(data_a[1], data_b[1]), np.maximum(data_a[1], data_b[1])
    assert np.all((x[1] >= lo) & (x[1] <= hi))


def test_compose_rgba_alpha_and_identity_128_frames_1080p(mxl, ctx48):
    """The one-pass compositor at full size: alpha is opaque everywhere, f = 255 equals the plain conversion of
    layer A, and grey chroma (U = V = 128) gives R = G = B."""
    n = TICKS
    grey = np.full(W.FRAME_BYTES, 128, np.uint8)
    grey[:W.FRAME_W * W.FRAME_H] = W.random_bytes(5, W.FRAME_W * W.FRAME_H)
    fa = [ctx48.frame(W.FRAME_W, W.FRAME_H, W.random_bytes(40 + i, W.FRAME_BYTES)) for i in range(2)]
    fg = ctx48.frame(W.FRAME_W, W.FRAME_H, grey)
    pics, ref = ctx48.rgba(W.FRAME_W, W.FRAME_H, n), ctx48.rgba(W.FRAME_W, W.FRAME_H, 2)
    ctx48.compose_rgba([fa[k % 2] for k in range(n)], [fg] * n, 1.0, pics)
    ctx48.frames_to_rgba(fa, ref)
    got, want = pics.download(), ref.download()
    for k in range(n):
        assert np.array_equal(got[k], want[k % 2]), k
    assert np.all(got[:, 3::4] == 255)
    ctx48.compose_rgba([fg], [fa[0]], 1.0, pics)
    g = pics.download(0, 1)[0].reshape(-1, 4)
    assert np.array_equal(g[:, 0], g[:, 1]) and np.array_equal(g[:, 1], g[:, 2])
    pics.free(); ref.free()


def test_mixer_identity_on_the_widest_bus(mxl, ctx48):
    """mixer.rs:54-68 at C = 256, S = 65 536: with one live channel at 0 dB / fader 1.0 the master and cue buses
    ARE that channel (x * 1.0 is exact, the other 255 channels add +0.0), whichever channel it is."""
    C, frames = 256, 65536 * 4
    x = W.uniform_pm1(77, 2 * frames)
    lx = ctx48.stereo(x)
    master, cue = ctx48.line(mxl.LINE_STEREO, frames), ctx48.line(mxl.LINE_STEREO, frames)
    mod = ctx48.module(mxl.MOD_MIXER, [(0.0, 1.0, True)] * C)
    for live in (0, 159, 160, 255):                     # both sides of the 160-channel launch split
        ins = [None] * C
        ins[live] = lx
        mod.run_tick(0, ins, [master, cue])
        assert np.array_equal(_bits(master.download()), _bits(x)), live
        assert np.array_equal(_bits(cue.download()), _bits(x)), live
    # two live channels at -6.0206 dB ~ 0.5: master = f32(f64(x) * g) + f32(f64(x) * g), cue = x + x
    g = 10.0 ** (-6.0 / 20.0)
    mod2 = ctx48.module(mxl.MOD_MIXER, [(-6.0, 1.0, True)] * C)
    ins = [None] * C
    ins[3] = lx
    ins[200] = lx
    mod2.run_tick(0, ins, [master, cue])
    h = (x.astype(np.float64) * g).astype(np.float32)
    assert np.array_equal(_bits(master.download()), _bits(h + h))
    assert np.array_equal(_bits(cue.download()), _bits(x + x))


def test_audio_graph_batching_invariance_128_ticks(mxl, ctx48):
    """Engine::run_tick for 128 ticks in one call == 128 calls of one tick (engine.rs:490: t = tick * S), for the
    modules whose kernels are exact: the config-2 graph's master bus up to the EqThree chunk-carry noise, the
    Oscillator / Envelope lines bit for bit."""
    d = W.config2_graph()
    g1, ids1 = W.build_graph(ctx48, d)
    g2, ids2 = W.build_graph(ctx48, d)
    g1.run_ticks(0, TICKS)
    m = d.taps["master"]
    whole = g1.output(ids1[m[0]], m[1]).download()
    parts = []
    for k in range(TICKS):
        g2.run_ticks(k, 1)
        parts.append(g2.output(ids2[m[0]], m[1]).download())
    parts = np.concatenate(parts)
    assert np.count_nonzero(_bits(whole) != _bits(parts)) <= 4      # ~1e-9 flips per EqThree sample
    assert np.allclose(whole, parts, rtol=1e-6, atol=1e-7)
    g1.destroy(); g2.destroy()
    # exact modules: one call vs per-tick calls, bit for bit
    osc = ctx48.module(mxl.MOD_OSCILLATOR, (440.0, mxl.WAVE_SINE, 0))
    env = ctx48.module(mxl.MOD_ENVELOPE, (25.0, 500.0, 0.8, 200.0))
    env2 = ctx48.module(mxl.MOD_ENVELOPE, (25.0, 500.0, 0.8, 200.0))
    n = SPT * TICKS
    gate = np.where((np.arange(n) // 9000) % 2 == 0, 1.0, 0.0).astype(np.float32)
    mono, stereo, eo = ctx48.line(mxl.LINE_MONO, n), ctx48.line(mxl.LINE_STEREO, n), ctx48.line(mxl.LINE_MONO, n)
    t0 = 48000 * 3600 * 5
    osc.run_tick(t0, [], [mono, stereo])
    env.run_tick(t0, [ctx48.mono(gate)], [eo])
    whole_osc, whole_env = mono.download(), eo.download()
    m1, s1, e1 = ctx48.line(mxl.LINE_MONO, SPT), ctx48.line(mxl.LINE_STEREO, SPT), ctx48.line(mxl.LINE_MONO, SPT)
    for k in range(TICKS):
        osc.run_tick(t0 + k * SPT, [], [m1, s1])
        assert np.array_equal(_bits(m1.download()), _bits(whole_osc[k * SPT:(k + 1) * SPT])), k
        gl = ctx48.mono(gate[k * SPT:(k + 1) * SPT])
        env2.run_tick(t0 + k * SPT, [gl], [e1])
        assert np.array_equal(_bits(e1.download()), _bits(whole_env[k * SPT:(k + 1) * SPT])), k
        gl.free()


def test_pcm_round_trip_and_meter_at_full_length(mxl, ctx48):
    """encode.rs:184-195 / stream_input.rs:167-173 on 2^24 samples: unpack(pack(x)) * 32768 == trunc(clamp(x) * 32767);
    the meter's per-tick peaks and sums of squares against numpy in f64."""
    n = 1 << 24
    x = (W.uniform_pm1(9, n) * np.float32(1.2)).astype(np.float32)
    line = ctx48.stereo(x)
    pcm = np.empty(n, np.int16)
    mxl.check(mxl.lib().mxl_pcm_pack_i16(ctx48.h, line.h, pcm.ctypes.data, n))
    want = np.trunc(np.clip(x, -1.0, 1.0) * np.float32(32767.0)).astype(np.int16)
    assert np.array_equal(pcm, want)
    back = ctx48.line(mxl.LINE_STEREO, n // 2)
    mxl.check(mxl.lib().mxl_pcm_unpack_i16(ctx48.h, pcm.ctypes.data, n, back.h))
    assert np.array_equal(back.download() * np.float32(32768.0), want.astype(np.float32))
    meter = ctx48.module(mxl.MOD_METER)
    meter.run_tick(0, [line], [])
    ticks = n // 2 // SPT
    rec = meter.meter_download(ticks)
    xs = x[:ticks * SPT * 2].reshape(ticks, SPT, 2).astype(np.float64)
    assert np.array_equal(rec["peak"], np.abs(xs).max(axis=1).astype(np.float32))
    assert np.allclose(rec["sumsq"], (xs * xs).sum(axis=1), rtol=1e-12, atol=0)
    assert np.array_equal(rec["clip"] != 0, (np.abs(xs) > 1.0).any(axis=(1, 2)))

"""GPU parity: every audio module through the C ABI (mxl_module_run_tick) against the CPU oracle on
the same seeded inputs.  Exact (bit-for-bit f32) wherever the module has no transcendental;
<= 1e-6 relative for the sin-based ones (north_star tolerance)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, assert_close_audio, load_f32, mismatch_count
from mixlab_b200 import workloads as W

pytestmark = pytest.mark.gpu

LENGTHS = [0, 1, 3, 4, 5, 735, 800, 1023, 4096, 65536 + 7]


def bits_equal(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


# ---- Mixer (mixer.rs:46-71) -------------------------------------------------------------------
@pytest.mark.parametrize("channels", [1, 2, 4, 8, 10, 33, 161, 256])
@pytest.mark.parametrize("frames", [1, 7, 800, 4099])
def test_mixer(mxl, oracle, ctx48, channels, frames):
    rng = np.random.default_rng(channels * 1000 + frames)
    ins = [W.uniform_pm1(channels * (1 << 20) + frames + c, 2 * frames) for c in range(channels)]
    if channels >= 4:
        ins[2] = None                                   # a disconnected channel (static zero buffer)
    gains = rng.uniform(-24.0, 6.0, channels)
    faders = rng.uniform(0.0, 1.0, channels)
    cues = rng.integers(0, 2, channels)
    want_m, want_c = oracle.mixer(ins, gains, faders, cues, frames)
    mod = ctx48.module(mxl.MOD_MIXER, list(zip(gains, faders, cues)))
    lines = [None if x is None else ctx48.stereo(x) for x in ins]
    master, cue = ctx48.line(mxl.LINE_STEREO, frames), ctx48.line(mxl.LINE_STEREO, frames)
    # outputs start dirty: the module must zero them itself (util::zero, mixer.rs:54-55)
    master.upload(np.full(2 * frames, 7.0, np.float32))
    cue.upload(np.full(2 * frames, -3.0, np.float32))
    mod.run_tick(0, lines, [master, cue])
    assert bits_equal(master.download(), want_m)
    assert bits_equal(cue.download(), want_c)
    assert [lab for lab, _ in mod.inputs()] == [str(i + 1) for i in range(channels)]
    assert mod.outputs() == [("Master", mxl.LINE_STEREO), ("Cue", mxl.LINE_STEREO)]
    for ln in lines + [master, cue]:
        if ln is not None:
            ln.free()
    mod.destroy()


def test_mixer_zero_length_and_zero_channels(mxl, oracle, ctx48):
    mod = ctx48.module(mxl.MOD_MIXER, [])
    master, cue = ctx48.stereo(np.ones(16, np.float32)), ctx48.stereo(np.ones(16, np.float32))
    mod.run_tick(0, [], [master, cue])
    assert np.all(master.download() == 0.0) and np.all(cue.download() == 0.0)
    e1, e2 = ctx48.line(mxl.LINE_STEREO, 0), ctx48.line(mxl.LINE_STEREO, 0)
    mod.run_tick(0, [], [e1, e2])
    ctx48.synchronize()


# ---- Amplifier (amplifier.rs:38-73) -----------------------------------------------------------
@pytest.mark.parametrize("frames", LENGTHS)
@pytest.mark.parametrize("with_mod", [True, False])
def test_amplifier(mxl, oracle, ctx48, frames, with_mod):
    x = W.uniform_pm1(100 + frames, 2 * frames)
    m = W.uniform_01(200 + frames, frames) if with_mod else None
    want = oracle.amplifier(x, m, 0.9, 0.5)
    mod = ctx48.module(mxl.MOD_AMPLIFIER, (0.9, 0.5))
    out = ctx48.line(mxl.LINE_STEREO, frames)
    mod.run_tick(0, [ctx48.stereo(x), ctx48.mono(m) if with_mod else None], [out])
    assert bits_equal(out.download(), want)
    assert mod.inputs() == [("Input", mxl.LINE_STEREO), ("Control", mxl.LINE_MONO)]


def test_amplifier_update_and_extreme_params(mxl, oracle, ctx48):
    x = W.uniform_pm1(5, 2 * 800)
    m = W.uniform_01(6, 800)
    mod = ctx48.module(mxl.MOD_AMPLIFIER, (1.0, 0.0))
    out = ctx48.line(mxl.LINE_STEREO, 800)
    lx, lm = ctx48.stereo(x), ctx48.mono(m)
    for amp, depth in [(1.0, 0.0), (0.0, 1.0), (0.333, 0.777), (1e-300, 1.0), (3.5, -0.5)]:
        mod.update((amp, depth))
        p = mod.params()
        assert (p.amplitude, p.mod_depth) == (amp, depth)
        mod.run_tick(0, [lx, lm], [out])
        assert bits_equal(out.download(), oracle.amplifier(x, m, amp, depth))


# ---- EqThree (eq_three.rs:58-89,106-125) ------------------------------------------------------
@pytest.mark.parametrize("chunk", ["auto", "s16", "s32", "s64", "64", "256", "1024", "4096"])
def test_eq_three_golden(mxl, ctx44, chunk, monkeypatch):
    """The reference's own golden vector (eq_three.rs:150-167), one run_tick over 355 285 samples.
    "auto"/"sN": the single-launch kernel eq_stream_kernel<N>; plain numbers: the two-launch kernel."""
    monkeypatch.delenv("MXL_EQ_CHUNK", raising=False)
    monkeypatch.delenv("MXL_EQ_STREAM_CHUNK", raising=False)
    if chunk.startswith("s"):
        monkeypatch.setenv("MXL_EQ_STREAM_CHUNK", chunk[1:])
    elif chunk != "auto":
        monkeypatch.setenv("MXL_EQ_CHUNK", chunk)
    x = load_f32(os.path.join(GOLDEN, "eq_three", "chronos.f32.raw"))
    want = load_f32(os.path.join(GOLDEN, "eq_three", "chronos-eq.f32.raw"))
    mod = ctx44.module(mxl.MOD_EQ_THREE, (4.0, 0.0, 4.0))
    out = ctx44.line(mxl.LINE_MONO, x.size)
    mod.run_tick(0, [ctx44.mono(x)], [out])
    got = out.download()
    assert mismatch_count(got, want) == 0


@pytest.mark.parametrize("split", [735, 800, 4096, 100000])
def test_eq_three_state_carries_across_calls(mxl, oracle, ctx44, split):
    x = load_f32(os.path.join(GOLDEN, "eq_three", "chronos.f32.raw"))[:300000]
    want = load_f32(os.path.join(GOLDEN, "eq_three", "chronos-eq.f32.raw"))[:300000]
    mod = ctx44.module(mxl.MOD_EQ_THREE, (4.0, 0.0, 4.0))
    got = []
    for i in range(0, x.size, split):
        seg = x[i:i + split]
        out = ctx44.line(mxl.LINE_MONO, seg.size)
        inp = ctx44.mono(seg)
        mod.run_tick(i, [inp], [out])
        got.append(out.download())
        out.free(); inp.free()
    got = np.concatenate(got)
    assert mismatch_count(got, want) == 0
    # final state equals the oracle's to f64 rounding noise of the chunk carry
    eq = oracle.EqThree(44100.0)
    eq.run((4.0, 0.0, 4.0), x)
    st = mod.eq_three_state()
    ref = np.array(list(eq.state.lo_poles) + list(eq.state.hi_poles) + list(eq.state.history))
    assert np.allclose(st, ref, rtol=1e-12, atol=1e-15)
    assert np.array_equal(st[8:], ref[8:])


@pytest.mark.parametrize("frames", [1, 2, 3, 5, 31, 33, 255, 257, 800, 7071, 7072, 7073, 65536 + 3, 1 << 20])
@pytest.mark.parametrize("gains", [(-6.0, 0.0, 4.0), (0.0, 0.0, 0.0), (6.0, -24.0, 3.0)])
def test_eq_three_random(mxl, oracle, ctx48, frames, gains):
    x = W.uniform_pm1(300 + frames, frames)
    want = oracle.EqThree(48000.0).run(gains, x)
    mod = ctx48.module(mxl.MOD_EQ_THREE, gains)
    out = ctx48.line(mxl.LINE_MONO, frames)
    mod.run_tick(0, [ctx48.mono(x)], [out])
    assert mismatch_count(out.download(), want) == 0


@pytest.mark.parametrize("seed,scale", [(11, 1.0), (12, 0.02), (13, 30.0)])
def test_eq_three_long_differential_run(mxl, oracle, ctx48, seed, scale):
    """The time-parallel kernel starts every chunk but the call's first from a carried state that equals the reference's
    serially rounded poles only to f64 rounding noise (~1e-16 relative); the final `as f32` absorbs that except where a sum
    sits within that noise of a rounding boundary -- about 1e-8 per sample.  So "bit-exact" is a statement about
    probability, not a theorem: this run counts.  2^23 samples (three minutes of audio) in ONE call, i.e. 130 000 carried
    chunk starts: every differing sample must differ by exactly one f32 ulp, and there may be at most 4 of them."""
    n = 1 << 23
    x = (W.uniform_pm1(seed, n) * np.float32(scale)).astype(np.float32)
    gains = (3.0, -2.0, 1.5)
    want = oracle.EqThree(48000.0).run(gains, x)
    mod = ctx48.module(mxl.MOD_EQ_THREE, gains)
    out = ctx48.line(mxl.LINE_MONO, n)
    mod.run_tick(0, [ctx48.mono(x)], [out])
    got = out.download()
    diff = np.flatnonzero(got.view(np.uint32) != want.view(np.uint32))
    print("EqThree differential run: %d of %d samples differ (scale %g)" % (diff.size, n, scale))
    assert diff.size <= 4
    if diff.size:
        ulps = np.abs(got.view(np.int32)[diff].astype(np.int64) - want.view(np.int32)[diff].astype(np.int64))
        assert ulps.max() == 1


@pytest.mark.parametrize("sr_spt", [(22050, 368), (96000, 1600), (192000, 3200)])
def test_eq_three_other_sample_rates(mxl, oracle, sr_spt):
    """The chunk/halo plan follows the pole decay at the context's sample rate."""
    sr, spt = sr_spt
    x = W.uniform_pm1(sr, 200000)
    want = oracle.EqThree(float(sr)).run((4.0, -3.0, 2.0), x)
    with mxl.Context(0, sr, spt) as ctx:
        mod = ctx.module(mxl.MOD_EQ_THREE, (4.0, -3.0, 2.0))
        out = ctx.line(mxl.LINE_MONO, x.size)
        mod.run_tick(0, [ctx.mono(x)], [out])
        assert mismatch_count(out.download(), want) == 0


def test_eq_three_many_instances_one_launch(mxl, oracle, ctx48):
    """Ten EqThree modules of a graph level share one launch (blockIdx.y = instance)."""
    d = W.GraphDesc("eqs")
    n, spt, ticks = 10, 800, 20
    srcs, eqs = [], []
    for k in range(n):
        s = d.add("SourceMono")
        e = d.add("EqThree", (k - 5.0, 0.5 * k, -1.0 * k))
        d.connect(e, 0, s, 0)
        srcs.append(s); eqs.append(e)
    g, ids = W.build_graph(ctx48, d)
    datas = [W.uniform_pm1(900 + k, spt * ticks) for k in range(n)]
    lines = [ctx48.mono(x) for x in datas]
    for k in range(n):
        g.module(ids[srcs[k]]).set_source_line(lines[k])
    before = ctx48.launch_count
    g.run_ticks(0, ticks)
    assert ctx48.launch_count - before == 1
    for k in range(n):
        want = oracle.EqThree(48000.0).run((k - 5.0, 0.5 * k, -1.0 * k), datas[k])
        assert mismatch_count(g.output(ids[eqs[k]], 0).download(), want) == 0
    g.destroy()


@pytest.mark.parametrize("path", ["stream", "two_launch"])
@pytest.mark.parametrize("bad", [np.nan, np.inf, -np.inf])
@pytest.mark.parametrize("pos", [0, 1000, 70001, 299999])
def test_eq_three_non_finite_sample_is_sticky(mxl, oracle, ctx48, bad, pos, path, monkeypatch):
    """eq_three.rs:121-128: once a NaN or an infinity is in the poles it never leaves -- every later output is
    NaN, in this call and the following ones.  The time-parallel carry forgets by construction, so the kernel
    has to re-impose that; samples before the bad one stay bit-exact."""
    monkeypatch.delenv("MXL_EQ_CHUNK", raising=False)
    if path == "two_launch":
        monkeypatch.setenv("MXL_EQ_CHUNK", "256")
    n = 300000
    x = W.uniform_pm1(4321, n)
    x[pos] = bad
    ref = oracle.EqThree(48000.0)
    want = ref.run((3.0, -2.0, 1.0), x)
    mod = ctx48.module(mxl.MOD_EQ_THREE, (3.0, -2.0, 1.0))
    out = ctx48.line(mxl.LINE_MONO, n)
    mod.run_tick(0, [ctx48.mono(x)], [out])
    got = out.download()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.all(np.isnan(got[pos + 1:])) and not np.any(np.isnan(got[:pos]))
    assert mismatch_count(got[:pos], want[:pos]) == 0
    # the next call starts from poisoned poles: all NaN, as in the reference
    y = W.uniform_pm1(55, 5000)
    want2 = ref.run((3.0, -2.0, 1.0), y)
    out2 = ctx48.line(mxl.LINE_MONO, y.size)
    mod.run_tick(n, [ctx48.mono(y)], [out2])
    assert np.all(np.isnan(want2)) and np.all(np.isnan(out2.download()))
    # and a clean module is not disturbed by another one's poison (the flags are per instance)
    clean = ctx48.module(mxl.MOD_EQ_THREE, (3.0, -2.0, 1.0))
    mod.run_tick(n + 5000, [ctx48.mono(y)], [out2])
    clean.run_tick(0, [ctx48.mono(y)], [out2])
    assert mismatch_count(out2.download(), oracle.EqThree(48000.0).run((3.0, -2.0, 1.0), y)) == 0


def test_eq_three_disconnected_input(mxl, oracle, ctx48):
    # Disconnected = static zero buffer (io.rs:36-41): only the VSA terms drive the poles
    want = oracle.EqThree(48000.0).run((4.0, 0.0, 4.0), np.zeros(5000, np.float32))
    mod = ctx48.module(mxl.MOD_EQ_THREE, (4.0, 0.0, 4.0))
    out = ctx48.line(mxl.LINE_MONO, 5000)
    mod.run_tick(0, [None], [out])
    assert mismatch_count(out.download(), want) == 0


# ---- Oscillator (oscillator.rs:15-37,65-92) ---------------------------------------------------
@pytest.mark.parametrize("waveform", ["WAVE_SAW", "WAVE_TRIANGLE", "WAVE_ON", "WAVE_OFF"])
@pytest.mark.parametrize("t0", [0, 800 * 12345, (1 << 40) + 17])
def test_oscillator_exact_waveforms(mxl, oracle, ctx48, waveform, t0):
    frames = 4099
    wf = getattr(mxl, waveform)
    for freq in (100.0, 55.0 * 2 ** (4 / 3.0), 19999.5):
        want_m, want_s = oracle.oscillator(t0, 48000.0, freq, getattr(oracle, waveform), frames)
        mod = ctx48.module(mxl.MOD_OSCILLATOR, (freq, wf, 0))
        mono, stereo = ctx48.line(mxl.LINE_MONO, frames), ctx48.line(mxl.LINE_STEREO, frames)
        mod.run_tick(t0, [], [mono, stereo])
        assert bits_equal(mono.download(), want_m), (waveform, freq)
        assert bits_equal(stereo.download(), want_s)
        mono.free(); stereo.free(); mod.destroy()


@pytest.mark.parametrize("t0", [0, 48000 * 3600, 48000 * 3600 * 24 * 30])
@pytest.mark.parametrize("sr_spt", [(48000, 800), (44100, 735)])
def test_oscillator_sine(mxl, oracle, t0, sr_spt):
    sr, spt = sr_spt
    frames = 20000
    with mxl.Context(0, sr, spt) as ctx:
        worst = 0
        for freq in (100.0, 440.0, 55.0 * 2 ** (7 / 3.0), 12000.0):
            want_m, want_s = oracle.oscillator(t0, float(sr), freq, oracle.WAVE_SINE, frames)
            mod = ctx.module(mxl.MOD_OSCILLATOR, (freq, mxl.WAVE_SINE, 0))
            mono, stereo = ctx.line(mxl.LINE_MONO, frames), ctx.line(mxl.LINE_STEREO, frames)
            mod.run_tick(t0, [], [mono, stereo])
            got = mono.download()
            assert_close_audio(got, want_m, what="sine %g" % freq)
            st = stereo.download()
            assert bits_equal(st[0::2], got) and bits_equal(st[1::2], got)
            worst = max(worst, mismatch_count(got, want_m))
        # 1-ulp(f64) differences between the device sine and glibc's survive `as f32` only when they
        # straddle an f32 rounding boundary: a handful of samples at most
        assert worst <= 4


def test_oscillator_square(mxl, oracle, ctx48):
    frames = 48000
    for freq in (100.0, 55.0 * 2 ** (3 / 3.0), 1234.5):
        want_m, _ = oracle.oscillator(800, 48000.0, freq, oracle.WAVE_SQUARE, frames)
        mod = ctx48.module(mxl.MOD_OSCILLATOR, (freq, mxl.WAVE_SQUARE, 0))
        mono, stereo = ctx48.line(mxl.LINE_MONO, frames), ctx48.line(mxl.LINE_STEREO, frames)
        mod.run_tick(800, [], [mono, stereo])
        assert bits_equal(mono.download(), want_m), freq


# ---- FmSine (fm_sine.rs:37-56) ----------------------------------------------------------------
@pytest.mark.parametrize("frames", [1, 5, 800, 4099])
@pytest.mark.parametrize("t0", [0, 48000 * 600])
def test_fm_sine(mxl, oracle, ctx48, frames, t0):
    x = W.uniform_pm1(400 + frames, frames)
    want = oracle.fm_sine(t0, 48000.0, 90.0, 110.0, x)
    mod = ctx48.module(mxl.MOD_FM_SINE, (90.0, 110.0))
    out = ctx48.line(mxl.LINE_STEREO, frames)
    mod.run_tick(t0, [ctx48.mono(x)], [out])
    got = out.download()
    assert_close_audio(got, want, what="fm_sine")
    assert mismatch_count(got, want) <= 4
    out2 = ctx48.line(mxl.LINE_STEREO, frames)
    mod.run_tick(t0, [None], [out2])
    assert_close_audio(out2.download(), oracle.fm_sine(t0, 48000.0, 90.0, 110.0, None, frames), what="fm_sine disconnected")


# ---- Panner / Splitter / Trigger --------------------------------------------------------------
@pytest.mark.parametrize("frames", LENGTHS)
def test_panner_splitter_trigger(mxl, oracle, ctx48, frames):
    l, r = W.uniform_pm1(1 + frames, frames), W.uniform_pm1(2 + frames, frames)
    pan = ctx48.module(mxl.MOD_STEREO_PANNER)
    st = ctx48.line(mxl.LINE_STEREO, frames)
    pan.run_tick(0, [ctx48.mono(l), ctx48.mono(r)], [st])
    assert bits_equal(st.download(), oracle.stereo_panner(l, r, frames))
    st2 = ctx48.line(mxl.LINE_STEREO, frames)
    pan.run_tick(0, [ctx48.mono(l), None], [st2])
    assert bits_equal(st2.download(), oracle.stereo_panner(l, None, frames))
    sp = ctx48.module(mxl.MOD_STEREO_SPLITTER)
    lo, ro = ctx48.line(mxl.LINE_MONO, frames), ctx48.line(mxl.LINE_MONO, frames)
    sp.run_tick(0, [st], [lo, ro])
    assert bits_equal(lo.download(), l) and bits_equal(ro.download(), r)
    assert pan.inputs() == [("L", mxl.LINE_MONO), ("R", mxl.LINE_MONO)] and sp.outputs() == [("L", mxl.LINE_MONO), ("R", mxl.LINE_MONO)]
    for gate, val in ((mxl.GATE_OPEN, 1.0), (mxl.GATE_CLOSED, 0.0)):
        trg = ctx48.module(mxl.MOD_TRIGGER, (gate,))
        o = ctx48.line(mxl.LINE_MONO, frames)
        o.upload(np.full(frames, 9.0, np.float32))
        trg.run_tick(0, [], [o])
        assert np.all(o.download() == val)


# ---- Envelope (envelope.rs:16-58,91-120) ------------------------------------------------------
def gate_pattern(seed, frames, density):
    rng = np.random.default_rng(seed)
    g = rng.uniform(0.01, 0.99, frames).astype(np.float32)       # inert samples
    n_ev = max(1, int(frames * density))
    pos = rng.integers(0, frames, n_ev)
    g[pos] = rng.integers(0, 2, n_ev).astype(np.float32)         # exact 0.0 / 1.0 events
    neg = rng.integers(0, frames, max(1, n_ev // 8))
    g[neg] = -0.0                                                # -0.0 == 0.0 is an off event
    return g


@pytest.mark.parametrize("frames", [1, 3, 800, 1024, 1025, 2047, 2048, 2049, 5000, 70000])
@pytest.mark.parametrize("density", [0.0005, 0.01, 0.5])
def test_envelope_random_gates(mxl, oracle, ctx48, frames, density):
    g = gate_pattern(frames * 7 + int(density * 1e4), frames, density)
    env = oracle.Envelope()
    want = env.run(4800, 48000.0, 25.0, 500.0, 0.8, 200.0, g)
    mod = ctx48.module(mxl.MOD_ENVELOPE, (25.0, 500.0, 0.8, 200.0))
    out = ctx48.line(mxl.LINE_MONO, frames)
    mod.run_tick(4800, [ctx48.mono(g)], [out])
    assert bits_equal(out.download(), want)
    st, seq, off = mod.envelope_state()
    assert (st, seq) == (env.state.state, env.state.seq) or (st == 0 and env.state.state == 0)
    if st == 2:
        assert off == env.state.off_amplitude


def test_envelope_state_across_calls(mxl, oracle, ctx48):
    frames, spt = 800 * 40, 800
    g = np.full(frames, 0.5, np.float32)
    g[1000] = 1.0
    g[9000] = 0.0
    g[9100] = 1.0
    g[20000:20100] = 0.0
    g[25000:] = 1.0
    env = oracle.Envelope()
    want = env.run(0, 48000.0, 10.0, 100.0, 0.5, 50.0, g)
    mod = ctx48.module(mxl.MOD_ENVELOPE, (10.0, 100.0, 0.5, 50.0))
    got = []
    for i in range(0, frames, spt * 5):
        out = ctx48.line(mxl.LINE_MONO, spt * 5)
        mod.run_tick(i, [ctx48.mono(g[i:i + spt * 5])], [out])
        got.append(out.download())
    assert bits_equal(np.concatenate(got), want)
    out = ctx48.line(mxl.LINE_MONO, 800)
    mod.run_tick(frames, [None], [out])          # disconnected gate = zeros: releases
    want2 = env.run(frames, 48000.0, 10.0, 100.0, 0.5, 50.0, np.zeros(800, np.float32))
    assert bits_equal(out.download(), want2)


@pytest.mark.parametrize("kind", ["sparse", "dense", "all_on", "all_off", "inert", "tile_edges"])
def test_envelope_long_call_many_tiles(mxl, oracle, ctx48, kind):
    """A call of 1.2 M samples = 586 tiles of the streaming kernel: the look-back crosses several
    32-tile windows; events sit on tile boundaries; whole tiles are inert or all events."""
    frames = 1_200_003
    if kind == "sparse":
        g = gate_pattern(5, frames, 4e-6)
    elif kind == "dense":
        g = gate_pattern(6, frames, 0.3)
    elif kind == "all_on":
        g = np.ones(frames, np.float32)
    elif kind == "all_off":
        g = np.zeros(frames, np.float32)
        g[0] = 1.0
    elif kind == "inert":
        g = np.full(frames, 0.25, np.float32)
        g[7] = 1.0
    else:
        g = np.full(frames, 0.5, np.float32)
        for k in range(1, frames // 2048, 3):
            g[k * 2048 - 1] = 1.0
            g[k * 2048] = 0.0
            g[k * 2048 + 2047] = 1.0 if k % 2 else 0.0
    env = oracle.Envelope()
    want = env.run(123456789, 48000.0, 25.0, 500.0, 0.8, 200.0, g)
    mod = ctx48.module(mxl.MOD_ENVELOPE, (25.0, 500.0, 0.8, 200.0))
    out = ctx48.line(mxl.LINE_MONO, frames)
    gl = ctx48.mono(g)
    mod.run_tick(123456789, [gl], [out])
    assert bits_equal(out.download(), want)
    st, seq, off = mod.envelope_state()
    assert st == env.state.state and (st == 0 or seq == env.state.seq) and (st != 2 or off == env.state.off_amplitude)
    # the same module again: state carries, tile descriptors of the first launch are stale (epoch)
    want2 = env.run(123456789 + frames, 48000.0, 25.0, 500.0, 0.8, 200.0, g[::-1].copy())
    mod.run_tick(123456789 + frames, [ctx48.mono(g[::-1].copy())], [out])
    assert bits_equal(out.download(), want2)


@pytest.mark.parametrize("params", [
    (0.0, 0.0, 0.3, 0.0),            # 1/0 = inf everywhere: inf * 0 = NaN at the sample a phase begins
    (0.001, 5.0, 0.0, 1.0),          # sustain and "released" reached within a tile: the constant arms
    (2.0, 30.0, 0.6, 10.0),
    (50.0, -10.0, 1.2, -5.0),        # negative times: the clamp's lower bound, amplitudes above 1
    (25.0, 500.0, -0.0, 200.0),      # sustain -0.0: the sign of the sustained zero
    (float("inf"), 100.0, 0.5, 100.0),
    (10.0, float("nan"), 0.5, float("nan")),
])
def test_envelope_parameter_corners(mxl, oracle, ctx48, params):
    """The output arms are specialised per thread (attack only / decay without clamp / sustain / release without clamp /
    released) on the strength of monotonicity arguments that need finite, non-negative slopes: corner parameters must fall
    back to the general arm and still give the reference's bits.  Trigger-like gate: long runs of exact 1.0 / 0.0 with
    inert stretches in between, so that every arm occurs, over 40 tiles."""
    frames = 2048 * 40 + 77
    g = np.full(frames, 0.5, np.float32)
    rng = np.random.default_rng(99)
    pos = 0
    on = True
    while pos < frames:
        run = int(rng.integers(5, 9000))
        kind = rng.integers(0, 3)
        if kind < 2:
            g[pos:pos + run] = 1.0 if on else 0.0
            on = not on
        pos += run
    env = oracle.Envelope()
    want = env.run(777, 48000.0, *params, g)
    mod = ctx48.module(mxl.MOD_ENVELOPE, params)
    out = ctx48.line(mxl.LINE_MONO, frames)
    mod.run_tick(777, [ctx48.mono(g)], [out])
    got = out.download()
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), (int((~same).sum()), int(np.flatnonzero(~same)[0]))


def test_pcm_async_ring_many_calls(mxl, oracle, ctx48):
    """N2/N3 hand-off without per-call allocation: 40 blocks of i16 PCM in, Amplifier, i16 PCM out, all queued
    before one synchronise; 5 MB pass through the 1 MB staging ring, which wraps several times
    (encode.rs:184-195, stream_input.rs:167-173)."""
    spt, ticks = 16384, 40
    n = 2 * spt
    pin_in = mxl.PinnedBuffer(ticks * n * 2, np.int16)
    pin_out = mxl.PinnedBuffer(ticks * n * 2, np.int16)
    pcm = (W.splitmix64(11, ticks * n) & np.uint64(0xFFFF)).astype(np.uint16).view(np.int16)
    pin_in.array[:] = pcm
    amp = ctx48.module(mxl.MOD_AMPLIFIER, (0.5, 0.0))
    lines = [(ctx48.line(mxl.LINE_STEREO, spt), ctx48.line(mxl.LINE_STEREO, spt)) for _ in range(ticks)]
    L = mxl.lib()
    for k, (src, dst) in enumerate(lines):
        mxl.check(L.mxl_pcm_unpack_i16_async(ctx48.h, pin_in.ptr + k * n * 2, n, src.h))
        amp.run_tick(k * spt, [src, None], [dst])
        mxl.check(L.mxl_pcm_pack_i16_async(ctx48.h, dst.h, pin_out.ptr + k * n * 2, n))
    ctx48.synchronize()
    want = oracle.pcm_pack_i16(oracle.amplifier(oracle.pcm_unpack_i16(pcm), None, 0.5, 0.0))
    assert np.array_equal(pin_out.array, want)
    pin_in.free(); pin_out.free()


# ---- Meter / Plotter / PCM --------------------------------------------------------------------
def test_meter(mxl, oracle, ctx48):
    spt, ticks = 800, 7
    x = W.uniform_pm1(77, 2 * spt * ticks) * np.float32(1.01)
    x[2 * spt * 3 + 11] = np.float32(-1.5)
    mod = ctx48.module(mxl.MOD_METER)
    mod.run_tick(0, [ctx48.stereo(x)], [])
    clips = []
    for k in range(ticks):
        peak, sumsq, clip = mod.meter_read(k)
        wp, ws, wc = oracle.meter(x[2 * spt * k:2 * spt * (k + 1)])
        assert peak == wp and clip == wc
        assert np.allclose(sumsq, ws, rtol=1e-13, atol=0)
        clips.append(clip)
    assert clips[3] is True


def test_meter_many_ticks_odd_tick_length(mxl, oracle, ctx44):
    """Enough slots for the warp-per-slot shape of the meter, 735-sample ticks (odd slots start 8-byte
    aligned only: the float2 path), a ragged last tick, a NaN (dropped by the peak, poisons the sum) and
    one sample beyond +-1."""
    spt, ticks, tail = 735, 2600, 201
    frames = spt * ticks + tail
    x = W.uniform_pm1(78, 2 * frames)
    x[2 * spt * 1001 + 5] = np.float32(1.25)
    x[2 * spt * 1002 + 8] = np.float32(np.nan)
    mod = ctx44.module(mxl.MOD_METER)
    mod.run_tick(0, [ctx44.stereo(x)], [])
    rec = mod.meter_download(ticks + 1)
    for k in list(range(0, ticks + 1, 97)) + [1001, 1002, ticks]:
        wp, ws, wc = oracle.meter(x[2 * spt * k:2 * spt * (k + 1)])
        assert tuple(rec["peak"][k]) == tuple(wp) and bool(rec["clip"][k]) == wc, k
        assert np.allclose(rec["sumsq"][k], ws, rtol=1e-13, atol=0, equal_nan=True), k
    assert bool(rec["clip"][1001]) and np.isnan(rec["sumsq"][1002]).any()


def test_plotter_tap(mxl, oracle, ctx48):
    # plotter.rs:37-56: every 6th tick the de-interleaved tick is reported
    spt = 800
    mod = ctx48.module(mxl.MOD_PLOTTER)
    x = W.uniform_pm1(9, 2 * spt * 13)
    line = ctx48.stereo(x)
    mod.run_tick(0, [line], [])                  # ticks 1..13 -> taps at count 6 and 12; last is tick index 11
    l, r = mod.plotter_read(spt)
    wl, wr = oracle.plotter_tap(x[2 * spt * 11:2 * spt * 12])
    assert bits_equal(l, wl) and bits_equal(r, wr)
    one = ctx48.stereo(x[:2 * spt])
    for k in range(4):                           # counts 14..17: no tap
        mod.run_tick(0, [one], [])
        assert mod.plotter_read(spt)[0].size == 0
    mod.run_tick(0, [one], [])                   # count 18
    l, r = mod.plotter_read(spt)
    assert bits_equal(l, x[0:2 * spt:2]) and bits_equal(r, x[1:2 * spt:2])


def test_pcm_pack_unpack(mxl, oracle, ctx48):
    n = 2 * 4099
    x = W.uniform_pm1(31, n) * np.float32(1.2)
    x[:12] = [0.0, 1.0, -1.0, 2.0, -2.0, 0.5, -0.5, 0.99999, np.nan, 1e-9, -1e-9, -0.0]
    line = ctx48.stereo(x)
    got = np.empty(n, np.int16)
    mxl.check(mxl.lib().mxl_pcm_pack_i16(ctx48.h, line.h, got.ctypes.data, n))
    assert np.array_equal(got, oracle.pcm_pack_i16(x))
    sink = ctx48.module(mxl.MOD_PCM_SINK)
    sink.run_tick(0, [line], [])
    assert np.array_equal(sink.pcm_download(n), oracle.pcm_pack_i16(x))
    pcm = (W.splitmix64(5, n) & np.uint64(0xFFFF)).astype(np.uint16).view(np.int16)
    pcm[:4] = [-32768, -1, 0, 32767]
    dst = ctx48.line(mxl.LINE_STEREO, n // 2)
    mxl.check(mxl.lib().mxl_pcm_unpack_i16(ctx48.h, pcm.ctypes.data, n, dst.h))
    assert bits_equal(dst.download(), oracle.pcm_unpack_i16(pcm))


# ---- error behaviour (io.rs:40-41,49-50 panics -> status codes) -------------------------------
def test_line_type_mismatch_is_an_error(mxl, ctx48):
    mod = ctx48.module(mxl.MOD_EQ_THREE, (0.0, 0.0, 0.0))
    st = ctx48.line(mxl.LINE_STEREO, 16)
    mo = ctx48.line(mxl.LINE_MONO, 16)
    with pytest.raises(mxl.MxlError) as e:
        mod.run_tick(0, [st], [mo])
    assert e.value.status == mxl.ERR_LINE_TYPE and "expected mono input, got stereo" in str(e.value)
    with pytest.raises(mxl.MxlError) as e:
        mod.run_tick(0, [mo], [st])
    assert e.value.status == mxl.ERR_LINE_TYPE
    with pytest.raises(mxl.MxlError) as e:
        mod.update((1.0, 2.0), kind=mxl.MOD_AMPLIFIER)
    assert e.value.status == mxl.ERR_PARAMS
    amp = ctx48.module(mxl.MOD_AMPLIFIER, (1.0, 0.0))
    with pytest.raises(mxl.MxlError) as e:
        amp.run_tick(0, [ctx48.line(mxl.LINE_STEREO, 32), None], [ctx48.line(mxl.LINE_STEREO, 16)])
    assert e.value.status == mxl.ERR_LENGTH

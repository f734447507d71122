"""GPU parity of the module just after the path -- SURVEY.md §8(f) N3: Monitor::run_tick (src/module/monitor.rs:112-140),
the codec thread's loop body (235-247) and EncodeStream (src/video/encode.rs:34-107,184-221,279-287) up to the two
encoder calls.  What aac::Encoder::encode would be given (2 x 1024 packed i16 + timestamps) must be bit-exact; what
AvcEncoder::send_frame would be given (pts, duration, blank gap fillers, the picture letterbox-scaled to 560 x 350):
timing exact, pixels equal to the oracle's scaler (identity / self-specified bicubic, see DESIGN.md "unpinned").
Oracle: oracle/pyoracle.py::MonitorFeed."""
from fractions import Fraction

import numpy as np
import pytest

from mixlab_b200 import workloads as W

pytestmark = pytest.mark.gpu

SR, SPT = 48000, 800


def drain(mod):
    a, v = [], []
    while True:
        x = mod.monitor_recv_audio()
        if x is None:
            break
        a.append(x)
    while True:
        x = mod.monitor_recv_video()
        if x is None:
            break
        v.append(x)
    return a, v


def compare(mod, orc, n_audio_before=0, n_video_before=0):
    a, v = drain(mod)
    wa, wv = orc.audio_out[n_audio_before:], orc.video_out[n_video_before:]
    assert len(a) == len(wa), (len(a), len(wa))
    for (dec, dur, frag), (wdec, wdur, wfrag) in zip(a, wa):
        assert Fraction(*dec) == wdec and Fraction(*dur) == wdur
        assert np.array_equal(frag, wfrag)
    assert len(v) == len(wv), (len(v), len(wv))
    for i, ((pts, dur, tb, blank, fr), (wpts, wdur, wblank, wpix)) in enumerate(zip(v, wv)):
        assert (pts, dur, tb, blank) == (wpts, wdur, orc.time_base, wblank), i
        assert np.array_equal(fr.download_raw(), wpix), i
    return len(orc.audio_out), len(orc.video_out)


def test_terminals_and_defaults(mxl, ctx48):
    mod = ctx48.module(mxl.MOD_MONITOR)
    assert mod.inputs() == [("Video", mxl.LINE_VIDEO), ("Audio", mxl.LINE_STEREO)] and mod.outputs() == []   # monitor.rs:97-100,144
    p = mod.params()
    assert (p.width, p.height) == (560, 350)                                                                  # monitor.rs:21-22


def test_audio_fragments_tick_by_tick_and_batched(mxl, oracle, ctx48):
    """800-frame ticks (1600 samples) against 2048-sample fragments: a fragment leaves only when the buffer holds MORE
    than 2048 samples, one per tick at most (encode.rs:199)."""
    x = (W.uniform_pm1(3, 2 * SPT * 40) * np.float32(1.3)).astype(np.float32)       # some samples clip
    x[5] = np.float32(np.nan)
    orc = oracle.MonitorFeed(SR)
    for k in range(40):
        orc.run_tick((100 + k) * SPT, x[2 * SPT * k:2 * SPT * (k + 1)], None)
    mod = ctx48.module(mxl.MOD_MONITOR)
    for k in range(7):                                                               # tick by tick ...
        mod.run_tick((100 + k) * SPT, [None, ctx48.stereo(x[2 * SPT * k:2 * SPT * (k + 1)])], [])
    mod.run_tick(107 * SPT, [None, ctx48.stereo(x[2 * SPT * 7:])], [])               # ... and 33 ticks in one call
    na, nv = compare(mod, orc)
    assert na == (2 * SPT * 40 - 1) // 2048 and nv > 0                               # video side: blank gap fillers only
    assert all(j[2] for j in orc.video_out)


def test_disconnected_audio_is_silence(mxl, oracle, ctx48):
    orc = oracle.MonitorFeed(SR)
    mod = ctx48.module(mxl.MOD_MONITOR)
    for k in range(5):
        orc.run_tick(k * SPT, np.zeros(2 * SPT, np.float32), None)
        mod.run_tick(k * SPT, [None, None], [])
    compare(mod, orc)


@pytest.mark.parametrize("size", [(560, 350), (1280, 720), (64, 36)])
def test_video_jobs_drop_rule_gap_filling_and_scaling(mxl, oracle, ctx48, size):
    w, h = size
    lay = oracle.frame_layout(w, h)
    tick = Fraction(SPT, SR)
    # (tick, duration_hint, tick_offset): steady 30 fps, a frame that ends before the video clock (dropped), a gap
    # (blank filler from the barrier), a late long frame, back-to-back frames inside one batched call
    plan = {0: (Fraction(1, 30), Fraction(0)), 2: (Fraction(1, 30), Fraction(0)), 3: (Fraction(1, 120), Fraction(0)),
            9: (Fraction(1, 30), tick / 3), 10: (Fraction(1, 10), Fraction(0)), 11: (Fraction(1, 60), Fraction(0)),
            20: (Fraction(1, 30), Fraction(0)), 21: (Fraction(1, 30), Fraction(0)), 22: (Fraction(1, 30), tick)}
    frames = {k: (ctx48.frame(w, h, W.random_bytes(500 + k, lay.size)), W.random_bytes(500 + k, lay.size)) for k in plan}
    n = 26
    x = W.uniform_pm1(9, 2 * SPT * n)
    orc = oracle.MonitorFeed(SR)
    for k in range(n):
        v = None
        if k in plan:
            v = (frames[k][1], lay, plan[k][0], plan[k][1])
        orc.run_tick(k * SPT, x[2 * SPT * k:2 * SPT * (k + 1)], v)
    mod = ctx48.module(mxl.MOD_MONITOR)

    def call(k0, k1):
        vl = ctx48.video_line(k1 - k0)
        for k in range(k0, k1):
            if k in plan:
                d, o = plan[k]
                vl.set(k - k0, frames[k][0], duration=(d.numerator, d.denominator), offset=(o.numerator, o.denominator))
        mod.run_tick(k0 * SPT, [vl, ctx48.stereo(x[2 * SPT * k0:2 * SPT * k1])], [])

    for k in range(6):
        call(k, k + 1)
    call(6, 19)
    call(19, n)
    compare(mod, orc)
    assert any(not j[2] for j in orc.video_out) and any(j[2] for j in orc.video_out)
    assert len([j for j in orc.video_out if not j[2]]) < len(plan)                  # at least one frame was dropped


def test_in_a_graph_behind_the_mixers(mxl, oracle, ctx48):
    """Oscillator -> Mixer -> Monitor.Audio and VideoMixer -> Monitor.Video in one graph call."""
    g = ctx48.graph()
    osc = g.add(mxl.MOD_OSCILLATOR, (330.0, mxl.WAVE_SAW, 0))
    mix = g.add(mxl.MOD_MIXER, [(0.0, 1.0, False)])
    sv = g.add(mxl.MOD_SOURCE_VIDEO)
    vm = g.add(mxl.MOD_VIDEO_MIXER, (0, -1, 1.0))
    mon = g.add(mxl.MOD_MONITOR)
    g.connect(mix, 0, osc, 1)
    g.connect(vm, 0, sv, 0)
    g.connect(mon, 0, vm, 0)
    g.connect(mon, 1, mix, 0)
    n = 8
    lay = oracle.frame_layout(560, 350)
    pix = [W.random_bytes(900 + k, lay.size) for k in range(n)]
    vl = ctx48.video_line(n)
    for k in range(n):
        vl.set(k, ctx48.frame(560, 350, pix[k]), duration=(1, 60))
    g.module(sv).set_source_line(vl)
    g.run_ticks(0, n)
    orc = oracle.MonitorFeed(SR)
    for k in range(n):
        _, st = oracle.oscillator(k * SPT, float(SR), 330.0, oracle.WAVE_SAW, SPT)
        master, _ = oracle.mixer([st], [0.0], [1.0], [0], SPT)
        # VideoMixer output: crossfade at f = 255 of the stored frame, duration one tick, offset zero (video_mixer.rs:241-247)
        orc.run_tick(k * SPT, master, (oracle.video_crossfade(lay, pix[k], None, 255), lay, Fraction(SPT, SR), Fraction(0)))
    compare(g.module(mon), orc)
    g.destroy()


def test_stream_output_feeds_only_while_live(mxl, oracle, ctx48):
    """StreamOutput (src/module/stream_output.rs): Offline at creation (42), nothing is sent until the connection is
    Live (112-151); the tick the connection completes is the epoch (126); 1120 x 700 pictures (13-14); a re-connect
    starts a new EncodeStream (328-366)."""
    mod = ctx48.module(mxl.MOD_STREAM_OUTPUT)
    assert mod.inputs() == [("Video", mxl.LINE_VIDEO), ("Audio", mxl.LINE_STEREO)] and mod.outputs() == []
    p = mod.params()
    assert (p.width, p.height) == (1120, 700)
    lay = oracle.frame_layout(560, 350)
    x = W.uniform_pm1(31, 2 * SPT * 12)
    pix = {k: W.random_bytes(40 + k, lay.size) for k in (2, 5, 9)}

    def tick(k):
        vl = ctx48.video_line(1)
        if k in pix:
            vl.set(0, ctx48.frame(560, 350, pix[k]), duration=(1, 30))
        mod.run_tick(k * SPT, [vl, ctx48.stereo(x[2 * SPT * k:2 * SPT * (k + 1)])], [])

    for k in range(3):
        tick(k)                                                       # Offline: nothing
    assert drain(mod) == ([], [])
    for start, stop in ((3, 8), (8, 12)):                            # two connections
        mod.stream_output_set_live(True)
        orc = oracle.MonitorFeed(SR, 1120, 700)
        for k in range(start, stop):
            tick(k)
            orc.run_tick(k * SPT, x[2 * SPT * k:2 * SPT * (k + 1)],
                         (pix[k], lay, Fraction(1, 30), Fraction(0)) if k in pix else None)
        compare(mod, orc)
        mod.stream_output_set_live(False)
        tick(stop)
        assert drain(mod) == ([], [])


@pytest.mark.parametrize("seed", range(10))
def test_randomized_ticks_against_the_oracle(mxl, oracle, ctx48, seed):
    """random picture sizes, durations, offsets and gaps, calls of 1..12 ticks, audio beyond +-1: every fragment, every
    job (timing, blank flag, pixels) as the oracle's."""
    rng = np.random.default_rng(77 + seed)
    sizes = [(560, 350), (640, 360), (320, 240)]
    lays = {s: oracle.frame_layout(*s) for s in sizes}
    orc = oracle.MonitorFeed(SR)
    mod = ctx48.module(mxl.MOD_MONITOR)
    tick = int(rng.integers(0, 5000))
    na = nv = 0
    for _ in range(10):
        n = int(rng.integers(1, 13))
        x = (W.uniform_pm1(int(rng.integers(0, 1 << 30)), 2 * SPT * n) * np.float32(1.2)).astype(np.float32)
        vl = ctx48.video_line(n)
        vids = {}
        for k in range(n):
            if rng.random() < 0.45:
                s = sizes[int(rng.integers(0, len(sizes)))]
                pix = W.random_bytes(int(rng.integers(0, 1 << 30)), lays[s].size)
                dur = Fraction(1, int(rng.integers(10, 130)))
                off = Fraction(int(rng.integers(0, SPT + 1)), SR)
                vl.set(k, ctx48.frame(s[0], s[1], pix), duration=(dur.numerator, dur.denominator), offset=(off.numerator, off.denominator))
                vids[k] = (pix, lays[s], dur, off)
        mod.run_tick(tick * SPT, [vl, ctx48.stereo(x)], [])
        for k in range(n):
            orc.run_tick((tick + k) * SPT, x[2 * SPT * k:2 * SPT * (k + 1)], vids.get(k))
        tick += n
        if rng.random() < 0.2:
            tick += int(rng.integers(1, 4))                          # the engine skipped ticks: a gap for the barrier
        na, nv = compare(mod, orc, na, nv)

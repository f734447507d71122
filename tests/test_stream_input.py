"""GPU parity of StreamInput::run_tick (src/module/stream_input.rs:72-147) -- SURVEY.md §8(f) N2, the module just
before the path: audio frames of arbitrary size assembled into ticks (converted i16 -> f32 on the device), zero
fill on underrun, the source clock re-based when the source id changes, video frames held back until due.
Oracle: oracle/pyoracle.py::StreamInput (the reference's control flow with fractions.Fraction as Rational64)."""
from fractions import Fraction

import numpy as np
import pytest

from mixlab_b200 import workloads as W

pytestmark = pytest.mark.gpu

SR, SPT = 48000, 800


def bits_equal(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


def pcm(seed, n):
    return np.ascontiguousarray(W.random_bytes(seed, 2 * n)).view(np.int16).copy()


class Pair:
    """The same pushes into the device module and the oracle."""

    def __init__(self, mxl, oracle, ctx):
        self.mxl, self.ctx = mxl, ctx
        self.mod = ctx.module(mxl.MOD_STREAM_INPUT)
        self.orc = oracle.StreamInput(SR)
        self.frames = {}

    def audio(self, source_id, t, data):
        self.mod.stream_write_audio(source_id, (t.numerator, t.denominator), data)
        assert self.orc.write_audio(source_id, t, data)

    def video(self, source_id, t, tag, duration=Fraction(1, 30)):
        fr = self.ctx.frame(64, 36, W.random_bytes(1000 + tag, self.mxl.frame_layout(64, 36).size))
        self.frames[tag] = fr
        self.mod.stream_write_video(source_id, (t.numerator, t.denominator), fr, (duration.numerator, duration.denominator))
        assert self.orc.write_video(source_id, t, tag, duration)

    def run(self, tick0, n_ticks, spt=SPT):
        """one device call over n_ticks; the oracle tick by tick.  Returns nothing: asserts."""
        vline = self.ctx.video_line(n_ticks)
        aline = self.ctx.line(self.mxl.LINE_STEREO, n_ticks * spt)
        self.mod.run_tick(tick0 * spt, [], [vline, aline])
        got = aline.download()
        for k in range(n_ticks):
            vo, ao = self.orc.run_tick((tick0 + k) * spt, 2 * spt)
            assert bits_equal(got[2 * spt * k:2 * spt * (k + 1)], ao), ("audio", tick0 + k)
            fr = vline.get(k)
            if vo is None:
                assert fr is None, ("video", tick0 + k)
            else:
                tag, dur, off = vo
                assert fr is not None and fr.h == self.frames[tag].h, ("video", tick0 + k)
                d, o = vline.timing(k)
                assert Fraction(*d) == dur and Fraction(*o) == off, ("timing", tick0 + k, d, o, dur, off)
        vline.free(); aline.free()


def test_module_terminals(mxl, ctx48):
    mod = ctx48.module(mxl.MOD_STREAM_INPUT)
    assert mod.inputs() == []
    assert mod.outputs() == [("Video", mxl.LINE_VIDEO), ("Audio", mxl.LINE_STEREO)]      # stream_input.rs:44-47


def test_underrun_is_silence_and_no_video(mxl, oracle, ctx48):
    p = Pair(mxl, oracle, ctx48)
    p.run(0, 3)
    assert p.mod.stream_pending() == (0, 0)


@pytest.mark.parametrize("frame_samples", [2048, 1600, 333, 4801])
def test_audio_frames_assembled_into_ticks(mxl, oracle, ctx48, frame_samples):
    """AAC-sized (1024 stereo frames = 2048 i16), exactly one tick, tiny and larger-than-a-tick frames (odd lengths
    included: the reference counts interleaved samples, not stereo frames)."""
    p = Pair(mxl, oracle, ctx48)
    t = Fraction(0)
    for i in range(40):
        p.audio(1, t, pcm(10 + i, frame_samples))
        t += Fraction(frame_samples // 2, SR)
    p.run(0, 1)          # tick by tick ...
    p.run(1, 1)
    p.run(2, 5)          # ... and batched
    p.run(7, 40)         # the short-frame cases run dry part-way: the rest of that tick and all later ones are zeros
    assert p.mod.stream_pending()[0] == len(p.orc.audio_rx) + (p.orc.audio_frame is not None)
    if frame_samples == 333:
        assert p.mod.stream_pending() == (0, 0)


def test_i16_extremes_convert_exactly(mxl, oracle, ctx48):
    p = Pair(mxl, oracle, ctx48)
    data = np.array([-32768, -32767, -1, 0, 1, 32766, 32767] * 300, np.int16)
    p.audio(1, Fraction(0), data)
    p.run(0, 2)


def test_source_change_rebases_the_clock_and_video_waits_until_due(mxl, oracle, ctx48):
    p = Pair(mxl, oracle, ctx48)
    tick = Fraction(SPT, SR)
    # source 7 starts its own clock at 100 s; audio for 6 ticks
    for i in range(6):
        p.audio(7, Fraction(100) + i * tick, pcm(50 + i, 2 * SPT))
    # video at source times 100 s (due at once), +2.5 ticks (held back twice), +2.6 ticks (same tick as the former is
    # taken one tick later: one frame per tick), +5 ticks
    p.video(7, Fraction(100), 0)
    p.video(7, Fraction(100) + tick * Fraction(5, 2), 1)
    p.video(7, Fraction(100) + tick * Fraction(13, 5), 2)
    p.video(7, Fraction(100) + 5 * tick, 3)
    p.run(10, 4)
    assert p.mod.stream_pending()[1] >= 1
    # a new source id mid-stream: the epoch is recomputed from the first frame of the new source (stream_input.rs:100-106);
    # two frames of the new source inside one tick both re-base (existing_source_id is read once per tick, line 88)
    p.audio(8, Fraction(5), pcm(70, SPT))
    p.audio(8, Fraction(5) + tick / 2, pcm(71, SPT))
    p.audio(8, Fraction(5) + tick, pcm(72, 2 * SPT))
    p.video(8, Fraction(5) + tick, 4)
    p.run(14, 6)
    p.run(20, 2)


def test_video_before_any_audio_has_zero_offset(mxl, oracle, ctx48):
    """no SourceTiming yet -> tick_offset = zero (stream_input.rs:127-132: map / filter / unwrap_or)."""
    p = Pair(mxl, oracle, ctx48)
    p.video(3, Fraction(12345, 1000), 0)
    p.run(0, 1)
    # a frame whose re-based time lies in the past: negative offsets are filtered to zero
    p.audio(3, Fraction(50), pcm(5, 2 * SPT))
    p.video(3, Fraction(49), 1)
    p.run(1, 2)


def test_in_a_graph_feeding_a_mixer_and_a_video_mixer(mxl, oracle, ctx48):
    """StreamInput as a graph node: Audio -> Amplifier, Video -> VideoMixer channel 1, several ticks per call."""
    g = ctx48.graph()
    si = g.add(mxl.MOD_STREAM_INPUT)
    amp = g.add(mxl.MOD_AMPLIFIER, (0.5, 0.0))
    vm = g.add(mxl.MOD_VIDEO_MIXER, (0, -1, 1.0))
    g.connect(amp, 0, si, 1)
    g.connect(vm, 0, si, 0)
    orc = oracle.StreamInput(SR)
    data = [pcm(200 + i, 2 * SPT) for i in range(3)]
    lay = oracle.frame_layout(64, 36)
    pix = W.random_bytes(77, lay.size)
    fr = ctx48.frame(64, 36, pix)
    for i, d in enumerate(data):
        g.module(si).stream_write_audio(1, (i * SPT, SR), d)
        orc.write_audio(1, Fraction(i * SPT, SR), d)
    g.module(si).stream_write_video(1, (SPT, SR), fr, (1, 30))
    orc.write_video(1, Fraction(SPT, SR), "f", Fraction(1, 30))
    n = 4
    g.run_ticks(0, n)
    got = g.output(amp, 0).download()
    vout = g.output(vm, 0)
    arrivals = []
    for k in range(n):
        vo, ao = orc.run_tick(k * SPT, 2 * SPT)
        arrivals.append(vo)
        assert bits_equal(got[2 * SPT * k:2 * SPT * (k + 1)], oracle.amplifier(ao, None, 0.5, 0.0)), k
    # the frame is one tick ahead of tick 0: tick_offset == tick_duration is still "due" (stream_input.rs:134, `>`), so
    # it goes out at tick 0 with offset 1/60 and the VideoMixer keeps it until 0 + 1/60 + 1/30 = tick 3 (video_mixer.rs:139-143)
    assert arrivals[0] == ("f", Fraction(1, 30), Fraction(1, 60)) and arrivals[1:] == [None, None, None]
    want = oracle.video_crossfade(lay, pix, None, 255)
    for k in range(3):
        assert np.array_equal(vout.get(k).download_raw(), want), k
    assert vout.get(3) is None
    g.destroy()


def test_queue_capacity_is_the_ring_buffers(mxl, ctx48):
    mod = ctx48.module(mxl.MOD_STREAM_INPUT)
    one = np.zeros(2, np.int16)
    for _ in range(65536):
        mod.stream_write_audio(1, (0, 1), one)
    with pytest.raises(mxl.MxlError) as e:
        mod.stream_write_audio(1, (0, 1), one)
    assert e.value.status == mxl.ERR_LENGTH


@pytest.mark.parametrize("seed", range(12))
def test_randomized_pushes_against_the_oracle(mxl, oracle, ctx48, seed):
    """random frame sizes (1..6000 samples), source ids, clock jumps, video frames early / late / in bursts, calls of
    1..9 ticks: the device module and the oracle must agree on every sample, every gating decision and every offset."""
    rng = np.random.default_rng(1000 + seed)
    p = Pair(mxl, oracle, ctx48)
    t_src = Fraction(int(rng.integers(0, 1000)), 7)
    source_id, tick, tag = 1, int(rng.integers(0, 10_000)), 0
    for _ in range(25):
        for _ in range(int(rng.integers(0, 5))):
            n = int(rng.integers(1, 6000))
            p.audio(source_id, t_src, pcm(int(rng.integers(0, 1 << 30)), n))
            t_src += Fraction(n // 2, SR)
            if rng.random() < 0.08:
                source_id += 1                                      # a new publisher connects
                t_src = Fraction(int(rng.integers(0, 100_000)), 1000)
        for _ in range(int(rng.integers(0, 3))):
            jitter = Fraction(int(rng.integers(-3 * SPT, 6 * SPT)), SR)
            p.video(source_id, t_src + jitter, tag, Fraction(1, int(rng.integers(24, 61))))
            tag += 1
        n_ticks = int(rng.integers(1, 10))
        p.run(tick, n_ticks)
        tick += n_ticks

"""The whole neighbourhood of the path in one graph, as a Mixlab workspace would wire it:
StreamInput --Audio--> StereoSplitter -> EqThree (left) -> StereoPanner -> Mixer -> Monitor.Audio
StreamInput --Video--> VideoMixer ------------------------------------------> Monitor.Video
fed with decoder-shaped input (i16 audio frames of 1024 stereo frames, 30 fps pictures) and read back as the encoders'
input (AAC fragments, video jobs), against the same chain built from the oracle's pieces."""
from fractions import Fraction

import numpy as np
import pytest

from mixlab_b200 import workloads as W

pytestmark = pytest.mark.gpu

SR, SPT = 48000, 800


def test_stream_input_to_monitor_chain(mxl, oracle, ctx48):
    g = ctx48.graph()
    si = g.add(mxl.MOD_STREAM_INPUT)
    sp = g.add(mxl.MOD_STEREO_SPLITTER)
    eq = g.add(mxl.MOD_EQ_THREE, (4.0, 0.0, -6.0))
    pan = g.add(mxl.MOD_STEREO_PANNER)
    mix = g.add(mxl.MOD_MIXER, [(-3.0, 0.9, False)])
    vm = g.add(mxl.MOD_VIDEO_MIXER, (0, -1, 1.0))
    mon = g.add(mxl.MOD_MONITOR)
    g.connect(sp, 0, si, 1)
    g.connect(eq, 0, sp, 0)
    g.connect(pan, 0, eq, 0)
    g.connect(pan, 1, sp, 1)
    g.connect(mix, 0, pan, 0)
    g.connect(vm, 0, si, 0)
    g.connect(mon, 0, vm, 0)
    g.connect(mon, 1, mix, 0)

    n_ticks = 30
    lay = oracle.frame_layout(1280, 720)
    o_si, o_eq, o_mon = oracle.StreamInput(SR), oracle.EqThree(float(SR)), oracle.MonitorFeed(SR)
    # the receiver's pushes: 1024-frame AAC-sized audio frames, a 30 fps picture every other tick
    t, k = Fraction(7), 0
    while t < Fraction(7) + Fraction(n_ticks * SPT, SR):
        data = np.ascontiguousarray(W.random_bytes(300 + k, 4096)).view(np.int16).copy()
        g.module(si).stream_write_audio(5, (t.numerator, t.denominator), data)
        o_si.write_audio(5, t, data)
        t += Fraction(1024, SR)
        k += 1
    pix = {}
    for j in range(n_ticks // 2):
        ts = Fraction(7) + Fraction(j, 30)
        pix[j] = W.random_bytes(700 + j, lay.size)
        g.module(si).stream_write_video(5, (ts.numerator, ts.denominator), ctx48.frame(1280, 720, pix[j]), (1, 30))
        o_si.write_video(5, ts, j, Fraction(1, 30))

    # device: three calls of 1, 9 and 20 ticks
    tick = 0
    for n in (1, 9, 20):
        g.run_ticks(tick, n)
        tick += n

    # oracle: tick by tick through the same modules; VideoMixer with one layer at fader 1.0 keeps the received picture
    # while its stored frame lives (video_mixer.rs:92-101,139-143) and emits one-tick frames (241-247)
    stored, active_until = None, None
    for kk in range(n_ticks):
        vo, audio = o_si.run_tick(kk * SPT, 2 * SPT)
        left, right = audio[0::2].copy(), audio[1::2].copy()
        left = o_eq.run((4.0, 0.0, -6.0), left)
        stereo = oracle.stereo_panner(left, right, SPT)
        master, _ = oracle.mixer([stereo], [-3.0], [0.9], [0], SPT)
        now = Fraction(kk * SPT, SR)
        if stored is not None and now >= active_until:
            stored = None
        if vo is not None:
            tag, dur, off = vo
            stored, active_until = pix[tag], now + off + dur
        vin = None
        if stored is not None:
            vin = (oracle.video_crossfade(lay, stored, None, 255), lay, Fraction(SPT, SR), Fraction(0))
        o_mon.run_tick(kk * SPT, master, vin)

    m = g.module(mon)
    audio_got, video_got = [], []
    while True:
        x = m.monitor_recv_audio()
        if x is None:
            break
        audio_got.append(x)
    while True:
        x = m.monitor_recv_video()
        if x is None:
            break
        video_got.append(x)
    assert len(audio_got) == len(o_mon.audio_out) > 10
    for (dec, dur, frag), (wdec, wdur, wfrag) in zip(audio_got, o_mon.audio_out):
        assert Fraction(*dec) == wdec and Fraction(*dur) == wdur and np.array_equal(frag, wfrag)
    assert len(video_got) == len(o_mon.video_out) > 10
    for i, ((pts, dur, tb, blank, fr), (wpts, wdur, wblank, wpix)) in enumerate(zip(video_got, o_mon.video_out)):
        assert (pts, dur, blank) == (wpts, wdur, wblank), i
        assert np.array_equal(fr.download_raw(), wpix), i
    g.destroy()

"""GPU parity of the literal drop-in: ModuleT::run_tick with HOST slices (mxl_module_run_tick_host), the call the
unmodified engine loop makes (src/engine.rs:461-494), against the CPU oracle.  The first test is the reference's
own test (src/module/eq_three.rs:150-167) line for line."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, assert_close_audio, load_f32
from mixlab_b200 import workloads as W

pytestmark = pytest.mark.gpu


def bits_equal(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


def test_eq_three_basic_smoke_test(mxl, ctx44):
    """eq_three.rs:150-167 `basic_smoke_test`: create(+4, 0, +4 dB), one run_tick over the whole fixture with
    InputRef::Mono(&input) / OutputRef::Mono(&mut output), `assert!(output == expected_output)`."""
    inp = load_f32(os.path.join(GOLDEN, "eq_three", "chronos.f32.raw"))
    eq = ctx44.module(mxl.MOD_EQ_THREE, (4.0, 0.0, 4.0))
    output = np.zeros(inp.size, np.float32)
    eq.run_tick_host(0, [inp], [output])
    expected_output = load_f32(os.path.join(GOLDEN, "eq_three", "chronos-eq.f32.raw"))
    assert bits_equal(output, expected_output)


def test_eq_three_tick_by_tick_host_slices(mxl, oracle, ctx44):
    """the engine's real cadence: one 735-sample tick per call, state carried in the module (eq_three.rs:33-46)."""
    inp = load_f32(os.path.join(GOLDEN, "eq_three", "chronos.f32.raw"))[:735 * 40]
    want = load_f32(os.path.join(GOLDEN, "eq_three", "chronos-eq.f32.raw"))[:735 * 40]
    eq = ctx44.module(mxl.MOD_EQ_THREE, (4.0, 0.0, 4.0))
    got = np.zeros_like(inp)
    for k in range(40):
        eq.run_tick_host(735 * k, [inp[735 * k:735 * (k + 1)]], [got[735 * k:735 * (k + 1)]])
    assert bits_equal(got, want)


@pytest.mark.parametrize("frames", [0, 1, 735, 800, 4099])
def test_mixer_amplifier_chain_host_slices(mxl, oracle, ctx48, frames):
    """BASELINE config 1 wiring (4-channel Mixer -> Amplifier with a mono control), every hop through host slices."""
    ins = [W.uniform_pm1(1 + c, 2 * frames) for c in range(4)]
    ins[2] = None                                                    # InputRef::Disconnected
    gains, faders, cues = [0.0, -6.0, 3.0, -12.0], [1.0, 0.8, 0.5, 0.25], [0, 1, 0, 1]
    control = (W.uniform_pm1(5, frames) * np.float32(0.5) + np.float32(0.5)).astype(np.float32)
    want_m, want_c = oracle.mixer(ins, gains, faders, cues, frames)
    want = oracle.amplifier(want_m, control, 0.9, 0.5)
    mixer = ctx48.module(mxl.MOD_MIXER, list(zip(gains, faders, cues)))
    amp = ctx48.module(mxl.MOD_AMPLIFIER, (0.9, 0.5))
    master = np.full(2 * frames, 7.0, np.float32)                    # dirty: the module zeroes (mixer.rs:54-55)
    cue = np.full(2 * frames, -3.0, np.float32)
    mixer.run_tick_host(0, ins, [master, cue])
    assert bits_equal(master, want_m) and bits_equal(cue, want_c)
    out = np.empty(2 * frames, np.float32)
    amp.run_tick_host(0, [master, control], [out])
    assert bits_equal(out, want)
    amp.run_tick_host(0, [master, None], [out])                      # control disconnected -> 1.0 (amplifier.rs:54)
    assert bits_equal(out, oracle.amplifier(want_m, None, 0.9, 0.5))


def test_generators_and_routing_host_slices(mxl, oracle, ctx48):
    n, t0 = 800, 48000 * 7
    osc = ctx48.module(mxl.MOD_OSCILLATOR, (220.0, mxl.WAVE_SAW, 0))
    mono, stereo = np.empty(n, np.float32), np.empty(2 * n, np.float32)
    osc.run_tick_host(t0, [], [mono, stereo])
    want_m, want_s = oracle.oscillator(t0, 48000.0, 220.0, oracle.WAVE_SAW, n)
    assert bits_equal(mono, want_m) and bits_equal(stereo, want_s)
    osc.update((220.0, mxl.WAVE_SINE, 0))
    osc.run_tick_host(t0, [], [mono, stereo])
    assert_close_audio(mono, oracle.oscillator(t0, 48000.0, 220.0, oracle.WAVE_SINE, n)[0], what="sine")

    fm = ctx48.module(mxl.MOD_FM_SINE, (90.0, 110.0))
    out = np.empty(2 * n, np.float32)
    fm.run_tick_host(t0, [want_m], [out])
    assert_close_audio(out, oracle.fm_sine(t0, 48000.0, 90.0, 110.0, want_m), what="fm")

    pan = ctx48.module(mxl.MOD_STEREO_PANNER)
    pan.run_tick_host(0, [want_m, None], [out])
    assert bits_equal(out, oracle.stereo_panner(want_m, None, n))
    split = ctx48.module(mxl.MOD_STEREO_SPLITTER)
    left, right = np.empty(n, np.float32), np.empty(n, np.float32)
    split.run_tick_host(0, [want_s], [left, right])
    assert bits_equal(left, want_s[0::2]) and bits_equal(right, want_s[1::2])
    trig = ctx48.module(mxl.MOD_TRIGGER, (0,))
    trig.run_tick_host(0, [], [left])
    assert bits_equal(left, oracle.trigger(True, n))


def test_envelope_state_carries_across_host_ticks(mxl, oracle, ctx48):
    spt, ticks = 800, 30
    g = np.full(spt * ticks, 0.5, np.float32)
    g[1000] = 1.0
    g[9000] = 0.0
    g[9100] = 1.0
    g[20000:20100] = 0.0
    env = oracle.Envelope()
    want = env.run(0, 48000.0, 10.0, 100.0, 0.5, 50.0, g)
    mod = ctx48.module(mxl.MOD_ENVELOPE, (10.0, 100.0, 0.5, 50.0))
    got = np.empty_like(g)
    for k in range(ticks):
        mod.run_tick_host(spt * k, [g[spt * k:spt * (k + 1)]], [got[spt * k:spt * (k + 1)]])
    assert bits_equal(got, want)


def test_line_type_mismatch_is_an_error_not_a_panic(mxl, ctx48):
    """io.rs:40-41,49-50: expect_mono on a stereo ref panics; across the C ABI that is MXL_ERR_TYPE_MISMATCH."""
    eq = ctx48.module(mxl.MOD_EQ_THREE, (0.0, 0.0, 0.0))
    x = np.zeros(16, np.float32)
    with pytest.raises(mxl.MxlError) as e:
        eq.run_tick_host(0, [(mxl.LINE_STEREO, x)], [x.copy()])
    assert e.value.status == mxl.ERR_TYPE_MISMATCH
    with pytest.raises(mxl.MxlError) as e:
        eq.run_tick_host(0, [x], [(mxl.LINE_STEREO, x.copy())])
    assert e.value.status == mxl.ERR_TYPE_MISMATCH
    with pytest.raises(mxl.MxlError):
        eq.run_tick_host(0, [x, x], [x.copy()])                       # wrong terminal count
    pan = ctx48.module(mxl.MOD_STEREO_PANNER)
    with pytest.raises(mxl.MxlError) as e:
        pan.run_tick_host(0, [x, x], [np.zeros(15, np.float32)])      # stereo slice of odd length
    assert e.value.status == mxl.ERR_LENGTH


def test_video_mixer_host_refs(mxl, oracle, ctx48):
    """VideoMixer through Option<VideoFrame> refs: frame handles in, an owned frame out; the stored frame of
    channel A outlives its input (video_mixer.rs:92-101,139-143)."""
    w, h = 560, 350
    lay = oracle.frame_layout(w, h)
    da, db = W.random_bytes(11, lay.size), W.random_bytes(12, lay.size)
    fa, fb = ctx48.frame(w, h, da), ctx48.frame(w, h, db)
    mod = ctx48.module(mxl.MOD_VIDEO_MIXER, (0, 1, 0.25))
    f = oracle.fader_to_u8(0.25)
    spt = 800
    outs = mod.run_tick_host(0, [("video", fa, (1, 30), (0, 1)), ("video", fb, (1, 60), (0, 1)), None, None],
                             ["video", "video", "video"])
    assert np.array_equal(outs[0].download_raw(), oracle.video_crossfade(lay, da, db, f))
    assert np.array_equal(outs[1].download_raw(), da) and np.array_equal(outs[2].download_raw(), db)
    # tick 1: A sends nothing (its 1/30 s frame is still stored), B sends None
    outs = mod.run_tick_host(spt, [("video", None, (0, 1), (0, 1)), ("video", None, (0, 1), (0, 1)), None, None],
                             ["video", "video", "video"])
    assert np.array_equal(outs[0].download_raw(), oracle.video_crossfade(lay, da, None, f))
    assert outs[1] is None and outs[2] is None
    # tick 2: everything expired -> Output stays None (video_mixer.rs:113-119)
    outs = mod.run_tick_host(2 * spt, [None, None, None, None], ["video", "video", "video"])
    assert outs == [None, None, None]

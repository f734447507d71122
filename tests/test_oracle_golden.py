"""CPU: pins the oracle on the reference's only golden vector and checks its other functions against
hand-derived values of the reference source text (file:line cited per test)."""
import hashlib
import math
import os

import numpy as np
import pytest

from helpers import GOLDEN, load_f32, oracle_run
from mixlab_b200 import workloads as W

IN = os.path.join(GOLDEN, "eq_three", "chronos.f32.raw")
OUT = os.path.join(GOLDEN, "eq_three", "chronos-eq.f32.raw")


def test_fixture_checksums():
    # SHA-256 of the reference's fixtures/module/eq_three files
    assert hashlib.sha256(open(IN, "rb").read()).hexdigest().startswith("55d75ec7")
    assert hashlib.sha256(open(OUT, "rb").read()).hexdigest().startswith("e35db948")


def test_eq_three_golden_single_call(oracle):
    # src/module/eq_three.rs:150-167: gains (+4, 0, +4) dB, one run_tick over the whole file, SR 44100
    x, want = load_f32(IN), load_f32(OUT)
    assert x.size == want.size == 355285
    got = oracle.EqThree(44100.0).run((4.0, 0.0, 4.0), x)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_eq_three_golden_tickwise(oracle):
    # state carries across run_tick calls (eq_three.rs:17-27): 735-sample ticks give the same stream
    x, want = load_f32(IN), load_f32(OUT)
    eq = oracle.EqThree(44100.0)
    got = np.concatenate([eq.run((4.0, 0.0, 4.0), x[i:i + 735]) for i in range(0, x.size, 735)])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_db_to_linear(oracle):
    # protocol/src/lib.rs:469-471
    assert oracle.db_to_linear(0.0) == 1.0
    assert oracle.db_to_linear(20.0) == 10.0
    assert oracle.db_to_linear(4.0) == math.pow(10.0, 4.0 / 20.0)


def test_mixer_semantics(oracle):
    # mixer.rs:54-68: f64 multiply rounded to f32, f32 accumulation in channel order; cue adds raw input
    a = np.array([0.5, -0.25, 1.0, 0.0], np.float32)
    b = np.array([0.1, 0.2, -0.3, 0.4], np.float32)
    master, cue = oracle.mixer([a, None, b], [0.0, 0.0, -6.0], [1.0, 1.0, 0.5], [0, 1, 1], 2)
    g = 0.5 * math.pow(10.0, -6.0 / 20.0)
    want = (np.float32(0) + a.astype(np.float64).astype(np.float32))
    want = want + np.float32(0)                              # disconnected channel adds 0.0
    want = want + (b.astype(np.float64) * g).astype(np.float32)
    assert np.array_equal(master, want.astype(np.float32))
    assert np.array_equal(cue, b)


def test_amplifier_semantics(oracle):
    # amplifier.rs:52-57,71-73
    x = np.array([1.0, -1.0, 0.5, 0.25], np.float32)
    mod = np.array([0.0, 1.0], np.float32)
    got = oracle.amplifier(x, mod, 0.9, 0.5)
    want = np.array([(1.0 * (1 - 0.5 + 0.5 * 0.0)) * 0.9, (-1.0 * 0.5) * 0.9, 0.5 * 1.0 * 0.9, 0.25 * 1.0 * 0.9], np.float32)
    assert np.array_equal(got, want)
    assert np.array_equal(oracle.amplifier(x, None, 0.9, 0.5), (x.astype(np.float64) * 1.0 * 0.9).astype(np.float32))


def test_oscillator_waveforms(oracle):
    # oscillator.rs:15-37,73-89
    sr, f, n = 48000.0, 1000.0, 96
    mono, stereo = oracle.oscillator(0, sr, f, oracle.WAVE_SINE, n)
    ph = (np.arange(n, dtype=np.float64) / sr) * f
    assert np.array_equal(mono, np.sin(ph * 2.0 * math.pi).astype(np.float32))
    assert np.array_equal(stereo[0::2], mono) and np.array_equal(stereo[1::2], mono)
    saw, _ = oracle.oscillator(7, sr, f, oracle.WAVE_SAW, n)
    ph7 = ((np.arange(n, dtype=np.float64) + 7) / sr) * f
    assert np.array_equal(saw, (2.0 * (ph7 - np.floor(0.5 + ph7))).astype(np.float32))
    tri, _ = oracle.oscillator(7, sr, f, oracle.WAVE_TRIANGLE, n)
    assert np.array_equal(tri, (2.0 * np.abs(2.0 * (ph7 - np.floor(0.5 + ph7))) - 1.0).astype(np.float32))
    sq, _ = oracle.oscillator(0, sr, f, oracle.WAVE_SQUARE, n)
    assert sq[0] == 1.0 and set(np.unique(sq)) <= {-1.0, 1.0}          # sign bit of +0.0 is positive
    on, _ = oracle.oscillator(0, sr, f, oracle.WAVE_ON, 4)
    off, _ = oracle.oscillator(0, sr, f, oracle.WAVE_OFF, 4)
    assert np.all(on == 1.0) and np.all(off == 0.0)


def test_envelope_state_machine(oracle):
    # envelope.rs:34-58,96-117; defaults 25/500/0.8/200 ms (protocol lib.rs:318-327)
    sr = 48000.0
    gate = np.zeros(4800, np.float32)
    gate[100:2500] = 1.0
    gate[2500:] = 0.0
    gate[50] = 0.5                       # neither == 1.0 nor == 0.0: inert
    env = oracle.Envelope()
    out = env.run(0, sr, 25.0, 500.0, 0.8, 200.0, gate)
    assert np.all(out[:100] == 0.0)                                        # Initial
    ms = (np.arange(100, 2500) - 100) / sr * 1000.0
    att = ms < 25.0
    assert np.array_equal(out[100:2500][att], (1.0 / 25.0 * ms[att]).astype(np.float32))
    dec = 1.0 - np.clip(1.0 / 500.0 * (ms[~att] - 25.0), 0.0, 1.0)
    assert np.array_equal(out[100:2500][~att], (0.8 + (1.0 - 0.8) * dec).astype(np.float32))
    off_amp = 0.8 + 0.2 * (1.0 - (1.0 / 500.0 * ((2500 - 100) / sr * 1000.0 - 25.0)))
    ms_off = (np.arange(2500, 4800) - 2500) / sr * 1000.0
    want_off = off_amp * (1.0 - np.clip(1.0 / 200.0 * ms_off, 0.0, 1.0))
    assert np.array_equal(out[2500:], want_off.astype(np.float32))
    assert env.state.state == 2 and env.state.seq == 2500


def test_fm_sine_panner_splitter_trigger(oracle):
    sr, n = 48000.0, 64
    x = W.uniform_pm1(11, n)
    out = oracle.fm_sine(5, sr, 90.0, 110.0, x)                             # fm_sine.rs:42-53
    amp = (110.0 - 90.0) / 2.0
    mid = 90.0 + amp
    ts = (np.arange(n, dtype=np.float64) + 5) / sr
    want = np.sin(((mid + amp * x.astype(np.float64)) * 2.0 * math.pi) * ts).astype(np.float32)
    assert np.array_equal(out[0::2], want) and np.array_equal(out[1::2], want)
    l, r = W.uniform_pm1(1, n), W.uniform_pm1(2, n)
    st = oracle.stereo_panner(l, r, n)                                      # stereo_panner.rs:35-38
    assert np.array_equal(st[0::2], l) and np.array_equal(st[1::2], r)
    l2, r2 = oracle.stereo_splitter(st, n)                                  # stereo_splitter.rs:41-44
    assert np.array_equal(l2, l) and np.array_equal(r2, r)
    assert np.all(oracle.trigger(True, 8) == 1.0) and np.all(oracle.trigger(False, 8) == 0.0)


def test_pcm_pack_unpack(oracle):
    # src/video/encode.rs:184-195: clamp, * 32767, `as i16` (toward zero, saturating, NaN -> 0)
    x = np.array([0.0, 1.0, -1.0, 2.0, -2.0, 0.5, -0.5, 0.99999, np.nan, 1e-9, -1e-9], np.float32)
    got = oracle.pcm_pack_i16(x)
    want = [0, 32767, -32767, 32767, -32767, 16383, -16383, int(np.float32(0.99999) * np.float32(32767.0)), 0, 0, 0]
    assert got.tolist() == want
    pcm = np.array([-32768, -1, 0, 1, 32767], np.int16)
    assert np.array_equal(oracle.pcm_unpack_i16(pcm), pcm.astype(np.float32) / np.float32(32768.0))   # stream_input.rs:167-173


def test_meter_and_clip(oracle):
    x = np.array([0.5, -0.25, -1.5, 0.75], np.float32)
    peak, sumsq, clip = oracle.meter(x)
    assert peak == (1.5, 0.75) and clip is True
    assert sumsq == (0.25 + 2.25, 0.0625 + 0.5625)
    assert oracle.clip_detect(np.array([1.0, -1.0], np.float32)) is False   # output_device.rs:192: strict
    assert oracle.clip_detect(np.array([1.0000001, 0.0], np.float32)) is True


def test_frame_layout_and_blank(oracle):
    lay = oracle.frame_layout(1920, 1080)
    assert (list(lay.stride), list(lay.plane_h), lay.size) == ([1920, 960, 960], [1080, 540, 540], 3110400)
    lay = oracle.frame_layout(560, 350)            # monitor.rs:21-22
    assert list(lay.stride) == [576, 288, 288] and list(lay.plane_h) == [350, 175, 175]
    data = oracle.frame_blank(lay)                 # frame.rs:128-134
    assert np.all(data[:lay.offset[1]] == 0) and np.all(data[lay.offset[1]:] == 0x80)


def test_crossfade_values(oracle):
    # video_mixer.rs:211-235: (a*f + b*(255-f)) / 255 truncating
    lay = oracle.frame_layout(64, 4)
    a = W.random_bytes(1, lay.size)
    b = W.random_bytes(2, lay.size)
    for f in (0, 1, 63, 127, 254, 255):
        out = oracle.video_crossfade(lay, a, b, f)
        want = ((a.astype(np.uint32) * f + b.astype(np.uint32) * (255 - f)) // 255).astype(np.uint8)
        assert np.array_equal(out, want), f
    # missing layer aliases the blank output (video_mixer.rs:180-188)
    out = oracle.video_crossfade(lay, a, None, 127)
    blank = oracle.frame_blank(lay)
    want = ((a.astype(np.uint32) * 127 + blank.astype(np.uint32) * 128) // 255).astype(np.uint8)
    assert np.array_equal(out, want)


def test_crossfade_rounds_width_up_to_32(oracle):
    # `while out < end` with 32-byte steps touches stride padding up to ceil(w/32)*32 and no further
    lay = oracle.frame_layout(40, 2)               # luma stride 64, processed 64; chroma w=20 -> processed 32 of stride 32
    a = np.full(lay.size, 200, np.uint8)
    b = np.full(lay.size, 100, np.uint8)
    out = oracle.video_crossfade(lay, a, b, 255)
    assert np.all(out[:lay.offset[1]] == 200)
    lay = oracle.frame_layout(70, 2)               # luma stride 96, processed 96; chroma w=35 -> 64 of stride 64
    a = np.full(lay.size, 200, np.uint8)
    out = oracle.video_crossfade(lay, a, None, 255)
    assert np.all(out == 200)


def test_picture_geometry(oracle):
    assert oracle.unify_picture(1920, 1080, 1280, 720) == (1920, 1080)      # video_mixer.rs:276-297
    assert oracle.unify_picture(1279, 719, 640, 1001) == (1280, 1002)
    assert oracle.scale_geometry(1920, 1080, 1920, 1080) == (1920, 1080, 0, 0)
    assert oracle.scale_geometry(1280, 720, 1920, 1080) == (1920, 1080, 0, 0)
    assert oracle.scale_geometry(640, 480, 1920, 1080) == (1440, 1080, 240, 0)   # encode.rs:355-374
    assert oracle.scale_geometry(1080, 1920, 1920, 1080) == (606, 1080, 656, 0)  # 607.5 -> 607 -> even 606; (1920-606)/2=657 -> 656
    assert oracle.fader_to_u8(0.25) == 63 and oracle.fader_to_u8(0.999) == 254 and oracle.fader_to_u8(1.0) == 255
    assert oracle.fader_to_u8(-1.0) == 0 and oracle.fader_to_u8(7.0) == 255 and oracle.fader_to_u8(float("nan")) == 0


def test_engine_walker_matches_direct_calls(oracle):
    # Engine::run_tick (engine.rs:400-510) over config 1 equals calling the module functions by hand
    spt, sr = 800, 48000
    d = W.config1_graph()
    n_ticks = 3
    srcs, datas = {}, {}
    for mid, (kind, seed) in d.sources.items():
        if kind == "stereo":
            data = W.uniform_pm1(seed, 2 * spt * n_ticks)
            srcs[mid] = (data, 2)
        else:
            data = W.uniform_01(seed, spt * n_ticks)
            srcs[mid] = (data, 1)
        datas[mid] = data
    got, g, ids = oracle_run(oracle, d, sr, spt, 0, n_ticks, d.taps["out"], 2, srcs)
    ins = [datas[m] for m in range(4)]
    mixer_params = d.modules[5][1]
    master, cue = oracle.mixer(ins, [p[0] for p in mixer_params], [p[1] for p in mixer_params],
                               [p[2] for p in mixer_params], spt * n_ticks)
    want = oracle.amplifier(master, datas[4], 0.9, 0.5)
    assert np.array_equal(got, want)
    assert g.last_order() == [0, 1, 2, 3, 5, 4, 6]      # DFS from the terminal (Amplifier) through its inputs


def test_engine_walker_cycle_reads_disconnected(oracle):
    # engine.rs:440-442,479-482: a back edge's producer has no buffer yet -> input reads as zero
    g = oracle.Graph(48000.0, 16)
    a = g.add(oracle.MOD_AMPLIFIER, (1.0, 0.0))
    s = g.add(oracle.MOD_STEREO_SPLITTER)
    p = g.add(oracle.MOD_STEREO_PANNER)
    m = g.add(oracle.MOD_METER)
    assert g.connect(s, 0, a, 0) == 0
    assert g.connect(p, 0, s, 0) == 0
    assert g.connect(a, 0, p, 0) == 0          # cycle a -> s -> p -> a
    assert g.connect(m, 0, p, 0) == 0
    g.run_tick(0)
    order = g.last_order()
    assert order[-1] == m and sorted(order) == [a, s, p, m]
    assert g.connect(s, 0, p, 0) == 0 and g.connect(m, 0, s, 0) == -3 and g.connect(m, 9, a, 0) == -1 and g.connect(m, 0, a, 5) == -2


# ---- the neighbours of the path (Python restatements): internal consistency, no GPU -----------------------------
def test_monitor_feed_oracle_invariants(oracle):
    """EncodeStream's bookkeeping (src/video/encode.rs:46-100,184-221): fragments are consecutive 2048-sample windows of
    the packed stream with a clock advancing 1024 / SR each; video jobs tile the time line without gaps or overlaps
    (each starts where the previous one ended, in time-base units) and blank fillers appear only where no frame covers."""
    from fractions import Fraction
    sr, spt = 48000, 800
    rng = np.random.default_rng(5)
    orc = oracle.MonitorFeed(sr)
    lay = oracle.frame_layout(560, 350)
    stream = []
    for k in range(60):
        x = rng.uniform(-1.3, 1.3, 2 * spt).astype(np.float32)
        stream.append(oracle.pcm_pack_i16(x))
        v = None
        if rng.random() < 0.4:
            v = (np.full(lay.size, k, np.uint8), lay, Fraction(1, int(rng.integers(20, 70))), Fraction(int(rng.integers(0, spt)), sr))
        orc.run_tick((1000 + k) * spt, x, v)
    packed = np.concatenate(stream)
    for i, (dec, dur, frag) in enumerate(orc.audio_out):
        assert dec == Fraction(1024 * i, sr) and dur == Fraction(1024, sr)
        assert np.array_equal(frag, packed[2048 * i:2048 * (i + 1)])
    assert len(orc.audio_out) == (packed.size - 1) // 2048                     # a fragment leaves only when MORE than 2048 are buffered
    end = 0
    for pts, dur, blank, pix in orc.video_out:
        assert pts == end and dur >= 0
        end = pts + dur
        assert pix.size == lay.size and (not blank or np.array_equal(pix, orc.blank))
    assert end == orc._round_to_base(orc.video_timestamp, orc.time_base)


def test_stream_input_oracle_conserves_samples(oracle):
    """StreamInput (stream_input.rs:92-124): every pushed sample comes out exactly once, in order, converted by
    sample / 32768; ticks after the queue ran dry are silence."""
    from fractions import Fraction
    sr, spt = 48000, 800
    rng = np.random.default_rng(6)
    orc = oracle.StreamInput(sr)
    pushed = []
    t = Fraction(0)
    for _ in range(30):
        n = int(rng.integers(1, 5000))
        d = rng.integers(-32768, 32768, n).astype(np.int16)
        pushed.append(d)
        orc.write_audio(1, t, d)
        t += Fraction(n // 2, sr)
    total = sum(d.size for d in pushed)
    out = np.concatenate([orc.run_tick(k * spt, 2 * spt)[1] for k in range(total // (2 * spt) + 3)])
    want = np.concatenate(pushed).astype(np.float32) / np.float32(32768.0)
    assert np.array_equal(out[:total], want) and not out[total:].any()


def test_output_device_oracle_routes_and_clips(oracle):
    """OutputDevice (output_device.rs:177-207): routed channels carry the side's samples, the others silence; clip iff
    a routed sample is outside [-1, 1]."""
    x = np.array([0.5, -2.0] * 8, np.float32)                                 # left 0.5, right -2.0
    dev = oracle.OutputDevice(2, None, 4)                                     # only the left side is routed
    assert dev.run_tick(x) is False
    out = dev.pop(1 << 10).reshape(8, 4)
    assert np.all(out[:, 2] == 0.5) and not out[:, [0, 1, 3]].any()
    dev.update(2, 3)
    assert dev.run_tick(x) is True                                            # now the right side (-2.0) is routed: clip
    out = dev.pop(1 << 10).reshape(8, 4)
    assert np.all(out[:, 3] == -2.0) and np.all(out[:, 2] == 0.5)


@pytest.mark.parametrize("src,dst", [((1920, 1080), (560, 350)), ((1920, 1080), (1120, 700)), ((640, 480), (1280, 720)),
                                     ((70, 50), (560, 350)), ((34, 18), (36, 20)), ((2, 2), (64, 64)), ((3840, 2160), (320, 180)),
                                     ((560, 350), (560, 350))])
def test_cpu_arranged_scaler_equals_the_definition(oracle, src, dst):
    """orc_letterbox_scale (tap tables cached, clamps hoisted, row-contiguous vertical pass: what bench.py's CPU arm
    is timed with) against the plain two-pass definition, byte for byte."""
    from mixlab_b200 import workloads as W
    lay_in = oracle.frame_layout(*src)
    data = W.random_bytes(src[0] * 7 + dst[1], lay_in.size)
    assert np.array_equal(oracle.letterbox_scale_fast(data, lay_in, dst[0], dst[1]), oracle.letterbox_scale(data, lay_in, dst[0], dst[1]))


def _float_bicubic(src, dw, dh, a=-0.6):
    """Independent statement of the scaler's definition in float: separable Keys cubic convolution (a = -0.6), four taps,
    pixel centres aligned (u = (d + 0.5) * src / dst - 0.5), edge pixels repeated, no intermediate rounding."""
    def weights(src_n, dst_n):
        u = (np.arange(dst_n) + 0.5) * src_n / dst_n - 0.5
        i0 = np.floor(u).astype(np.int64)
        f = u - i0
        taps = np.stack([i0 - 1, i0, i0 + 1, i0 + 2], axis=1)
        x = np.abs(f[:, None] - np.array([-1.0, 0.0, 1.0, 2.0])[None, :])
        w = np.where(x <= 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0, np.where(x < 2.0, ((a * x - 5.0 * a) * x + 8.0 * a) * x - 4.0 * a, 0.0))
        w = w / w.sum(axis=1, keepdims=True)
        return np.clip(taps, 0, src_n - 1), w
    sh, sw = src.shape
    tx, wx = weights(sw, dw)
    ty, wy = weights(sh, dh)
    s = src.astype(np.float64)
    mid = (s[:, tx] * wx[None, :, :]).sum(axis=2)                  # [sh][dw]
    out = (mid[ty, :] * wy[:, :, None]).sum(axis=1)                # [dh][dw]
    return out


@pytest.mark.parametrize("src,dst", [((640, 360), (960, 540)), ((960, 540), (560, 315)), ((320, 240), (321, 241)), ((64, 48), (640, 480))])
def test_scaler_definition_against_an_independent_float_bicubic(oracle, src, dst):
    """The scaler is self-specified (it stands in for un-vendored swscale: parity UNPINNED).  What can be checked without
    swscale is that the integer two-pass definition is the filter it claims to be: against a float separable Keys bicubic
    written independently here (same taps and alignment, no 14-bit weights, no intermediate u8 rounding) a smooth picture
    must agree to rounding -- PSNR >= 50 dB, no pixel more than 2 levels off.  A half-pixel shift, a transposed weight or a
    wrong clamp shows up as tens of levels."""
    sw, sh = src
    dw, dh = dst
    yy, xx = np.mgrid[0:sh, 0:sw]
    pic = 128 + 60 * np.sin(xx / 17.0) * np.cos(yy / 11.0) + 40 * np.sin((xx + 2 * yy) / 29.0) + 0.05 * xx
    pic = np.clip(np.rint(pic), 0, 255).astype(np.uint8)
    got = oracle.bicubic_plane(pic.reshape(-1), sw, sh, sw, dw, dh, dw).reshape(dh, dw).astype(np.float64)
    want = np.clip(_float_bicubic(pic, dw, dh), 0, 255)
    err = got - want
    psnr = 10 * np.log10(255.0 ** 2 / np.mean(err ** 2))
    assert psnr >= 50.0, psnr
    assert np.abs(err).max() <= 2.0, np.abs(err).max()


def test_session_driver_equals_its_parts(oracle):
    """orc_session_run (the timed CPU arm of the session variant) does per tick exactly what the restated functions
    do when called one by one: unpack, Engine::run_tick, blank + crossfade of the stored layers, monitor scaler, pack."""
    from mixlab_b200 import workloads as W
    sr, spt, w, h = 48000, 800, 128, 72
    desc = W.config4_audio_graph()
    lay = oracle.frame_layout(w, h)
    la = W.random_bytes(1, 3 * lay.size)
    lb = W.random_bytes(2, 3 * lay.size)
    pcm = (W.splitmix64(3, 16 * 2 * spt) & np.uint64(0xFFFF)).astype(np.uint16).view(np.int16)
    g, ids = oracle.build_graph(desc, sr, spt)
    sess = oracle.Session(g, ids[desc.taps["master"][0]], w, h, 56, 34, la, lb, 2, pcm, 0.3)
    g2, ids2 = oracle.build_graph(desc, sr, spt)
    for tick in range(7):
        sess.run(tick, 1)
        master = g2.run_tick(tick, (ids2[desc.taps["master"][0]], 0), 2 * spt)
        fi = (tick // 2) % 3
        comp = oracle.video_crossfade(lay, la[fi * lay.size:(fi + 1) * lay.size], lb[fi * lay.size:(fi + 1) * lay.size], oracle.fader_to_u8(0.3))
        assert np.array_equal(sess.composite, comp), tick
        assert np.array_equal(sess.monitor_out, oracle.letterbox_scale(comp, lay, 56, 34)), tick
        assert np.array_equal(sess.pcm_out, oracle.pcm_pack_i16(master)), tick
    # several ticks in one call leave the last tick's outputs
    sess2 = oracle.Session(oracle.build_graph(desc, sr, spt)[0], ids[desc.taps["master"][0]], w, h, 56, 34, la, lb, 2, pcm, 0.3)
    sess2.run(0, 7)
    assert np.array_equal(sess2.monitor_out, sess.monitor_out) and np.array_equal(sess2.pcm_out, sess.pcm_out)


def test_rgba_to_yuv420p_known_colours_and_round_trip(oracle):
    """UNPINNED (the reference never converts colour): the integer BT.601 definition on the primaries, on a cut 2x2
    block, and against its own inverse (yuv420p -> RGBA -> yuv420p) where nothing clipped."""
    lay = oracle.frame_layout(4, 2)
    colours = {(255, 255, 255): (235, 128, 128), (0, 0, 0): (16, 128, 128), (255, 0, 0): (82, 90, 240),
               (0, 255, 0): (144, 54, 34), (0, 0, 255): (41, 240, 110)}
    for rgb, yuv in colours.items():
        rgba = np.tile(np.array(rgb + (7,), np.uint8), 8)               # alpha is ignored
        out = oracle.rgba_to_yuv420p(lay, rgba)
        assert set(out[lay.offset[0]:lay.offset[0] + 4]) == {yuv[0]} and out[lay.offset[1]] == yuv[1] and out[lay.offset[2]] == yuv[2], rgb
    # odd picture: the last chroma block repeats its last column / row
    lay3 = oracle.frame_layout(3, 3)
    rgba = W.random_bytes(5, 3 * 3 * 4)
    out = oracle.rgba_to_yuv420p(lay3, rgba)
    px = rgba.reshape(3, 3, 4).astype(np.int64)
    corner = px[2, 2, :3]                                               # the bottom-right block is one pixel, four times
    assert out[lay3.offset[1] + lay3.stride[1] + 1] == ((-38 * corner[0] - 74 * corner[1] + 112 * corner[2] + 128) >> 8) + 128
    # round trip from a frame whose conversion to RGB does not clip: luma back within 2 levels, chroma within 3
    lay = oracle.frame_layout(64, 36)
    yuv = oracle.frame_blank(lay)
    rng = np.random.default_rng(3)
    yuv[lay.offset[0]:lay.offset[0] + lay.stride[0] * 36] = rng.integers(60, 180, lay.stride[0] * 36)
    for p in (1, 2):
        yuv[lay.offset[p]:lay.offset[p] + lay.stride[p] * 18] = rng.integers(108, 148, lay.stride[p] * 18)
    rgba = oracle.yuv420p_to_rgba(lay, yuv)
    assert rgba.reshape(-1, 4)[:, :3].min() > 0 and rgba.reshape(-1, 4)[:, :3].max() < 255
    back = oracle.rgba_to_yuv420p(lay, rgba)
    y0 = yuv[lay.offset[0]:lay.offset[0] + lay.stride[0] * 36].reshape(36, -1)[:, :64].astype(int)
    y1 = back[lay.offset[0]:lay.offset[0] + lay.stride[0] * 36].reshape(36, -1)[:, :64].astype(int)
    assert np.abs(y0 - y1).max() <= 2
    for p in (1, 2):
        c0 = yuv[lay.offset[p]:lay.offset[p] + lay.stride[p] * 18].reshape(18, -1)[:, :32].astype(int)
        c1 = back[lay.offset[p]:lay.offset[p] + lay.stride[p] * 18].reshape(18, -1)[:, :32].astype(int)
        assert np.abs(c0 - c1).max() <= 3

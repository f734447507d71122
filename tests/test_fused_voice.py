"""GPU: the fused voice group (Oscillator -> EqThree -> StereoPanner -> Mixer [-> Meter] as ONE launch,
csrc/fused_voice.cu) against the staged path (one launch per stage) and against the oracle's tick-by-tick
engine walker (src/engine.rs:400-510).  Both paths run the same arithmetic in the same order, so every
observable line must be bit-identical between them."""
import numpy as np
import pytest

from helpers import assert_close_audio, mismatch_count, oracle_run, sine_mismatch_budget
from mixlab_b200 import workloads as W

pytestmark = pytest.mark.gpu


def stage_kinds(g):
    return [s["kind"] for s in g.stages() if s["n_launches"] > 0]


def run_pair(mxl, ctx, desc, calls, taps, meter=None, eqs=(), pins=(), tick0=0):
    """Runs `desc` twice -- fused and staged -- over the same sequence of calls; returns per path the concatenated
    tap lines, the meter records and the EqThree states after the last call."""
    out = {}
    for fused in (True, False):
        g, ids = W.build_graph(ctx, desc)
        g.set_fusion(fused)
        for (m, o) in pins:
            g.pin_output(ids[m], o)
        lines = {t: [] for t in taps}
        records = []
        tick = tick0
        for n in calls:
            g.run_ticks(tick, n)
            tick += n
            for t in taps:
                lines[t].append(g.output(ids[t[0]], t[1]).download())
            if meter is not None:
                records.append(np.array(g.module(ids[meter]).meter_download(n)))
        kinds = stage_kinds(g)
        states = [g.module(ids[e]).eq_three_state() for e in eqs]
        out[fused] = ({t: np.concatenate(v) for t, v in lines.items()}, np.concatenate(records) if records else None, states, kinds)
        g.destroy()
    return out


def eq_ids(desc):
    return [i for i, (k, _) in enumerate(desc.modules) if k == "EqThree"]


def assert_same_state(a_states, b_states):
    """EqThree poles after the call.  Chunks other than a call's first start from a scanned state that carries f64
    rounding noise (~1e-16 relative; the `as f32` of the outputs absorbs it -- the lines above are bit-identical), and
    the two paths cut the call into chunks differently: the stored f64 poles agree to that noise, not to the bit.
    The history (the last three inputs) is exact."""
    for a, b in zip(a_states, b_states):
        assert np.allclose(a[:8], b[:8], rtol=1e-12, atol=1e-300) and np.array_equal(a[8:], b[8:])


def assert_same_meter(a, b):
    assert np.array_equal(a["peak"], b["peak"]) and np.array_equal(a["clip"], b["clip"])
    assert np.array_equal(a["sumsq"], b["sumsq"])          # same summation order in both paths


@pytest.mark.parametrize("sr_spt", [(48000, 800), (44100, 735)])
@pytest.mark.parametrize("calls", [[1], [3, 1, 7], [128], [15, 25]])
def test_config2_fused_equals_staged(mxl, sr_spt, calls):
    sr, spt = sr_spt
    d = W.config2_graph()
    taps = [d.taps["master"], d.taps["cue"]] + [(e, 0) for e in eq_ids(d)]
    with mxl.Context(0, sr, spt) as ctx:
        r = run_pair(mxl, ctx, d, calls, taps, meter=d.taps["meter"][0], eqs=eq_ids(d))
    (fl, fm, fs, fk), (sl, sm, ss, sk) = r[True], r[False]
    assert W.STAGE_KIND["FusedVoiceMix"] in fk and W.STAGE_KIND["FusedVoiceMix"] not in sk
    assert W.KIND["EqThree"] in sk and W.KIND["EqThree"] not in fk and W.KIND["Mixer"] not in fk
    # ticks of 800 frames tile the EqThree chunks, so the meter rides in the fused launch; ticks of 735 do not
    assert (W.KIND["Meter"] in fk) == (spt % 2 == 1)
    for t in taps:
        assert mismatch_count(fl[t], sl[t]) == 0, t
    assert_same_meter(fm, sm)
    assert_same_state(fs, ss)


def test_config2_is_two_launches_per_call(mxl, ctx48):
    d = W.config2_graph()
    g, ids = W.build_graph(ctx48, d)
    g.run_ticks(0, 4)
    before = ctx48.launch_count
    g.run_ticks(4, 4)
    assert ctx48.launch_count - before == 2             # the voices, then the ordered sum + meter
    st = [s for s in g.stages() if s["n_launches"]]
    assert len(st) == 1 and st[0]["kind"] == W.STAGE_KIND["FusedVoiceMix"] and st[0]["n_modules"] == 32
    assert st[0]["algorithmic_bytes"] == 464 * 800 * 4      # reported against the unfused API bytes
    g.destroy()


@pytest.mark.parametrize("n_ticks", [8, 33])
def test_config4_audio_fused_vs_oracle(mxl, oracle, ctx48, n_ticks):
    """BASELINE config 4's audio half (5 voices, no meter) against the oracle walker."""
    spt = 800
    d = W.config4_audio_graph()
    oscs = [i for i, (k, _) in enumerate(d.modules) if k == "Oscillator"]
    g, ids = W.build_graph(ctx48, d)
    for o in oscs:
        g.pin_output(ids[o], 0)                 # observed oscillator lines: still one fused launch
    g.run_ticks(7, n_ticks)
    assert stage_kinds(g) == [W.STAGE_KIND["FusedVoiceMix"]]
    got = g.output(ids[d.taps["master"][0]], 0).download()
    cue = g.output(ids[d.taps["cue"][0]], 1).download()
    osc_lines = {o: g.output(ids[o], 0).download() for o in oscs}
    want, og, oids = oracle_run(oracle, d, 48000, spt, 7, n_ticks, d.taps["master"], 2)
    want_cue, _, _ = oracle_run(oracle, d, 48000, spt, 7, n_ticks, d.taps["cue"], 2)
    assert_close_audio(got, want, what="config4 master")
    assert_close_audio(cue, want_cue, what="config4 cue")
    budget, n_bad = sine_mismatch_budget(oracle, d, 48000, spt, 7, n_ticks, osc_lines)
    assert mismatch_count(got, want) <= budget, (mismatch_count(got, want), n_bad)
    assert mismatch_count(cue, want_cue) <= budget
    g.destroy()


def test_pinned_interior_lines_equal_staged(mxl, ctx48):
    d = W.config2_graph()
    oscs = [i for i, (k, _) in enumerate(d.modules) if k == "Oscillator"]
    pans = [i for i, (k, _) in enumerate(d.modules) if k == "StereoPanner"]
    pins = [(oscs[0], 0), (oscs[1], 1), (oscs[3], 0), (oscs[3], 1), (pans[2], 0), (pans[9], 0)]
    taps = pins + [d.taps["master"]]
    r = run_pair(mxl, ctx48, d, [5, 2], taps, pins=pins, tick0=3)
    assert W.STAGE_KIND["FusedVoiceMix"] in r[True][3] and W.KIND["Oscillator"] not in r[True][3]
    for t in taps:
        assert mismatch_count(r[True][0][t], r[False][0][t]) == 0, t


def test_hidden_line_is_refused_after_a_run_then_materialised(mxl, ctx48):
    d = W.config2_graph()
    g, ids = W.build_graph(ctx48, d)
    osc = [i for i, (k, _) in enumerate(d.modules) if k == "Oscillator"][2]
    g.run_ticks(0, 2)
    with pytest.raises(mxl.MxlError):
        g.output(ids[osc], 0)                   # was not written by the run above: loud, not stale
    g.run_ticks(2, 2)                           # observed from now on
    got = g.output(ids[osc], 0).download()
    ref = mxl.Context.module(ctx48, mxl.MOD_OSCILLATOR, d.modules[osc][1])
    mono, stereo = ctx48.line(mxl.LINE_MONO, 1600), ctx48.line(mxl.LINE_STEREO, 1600)
    ref.run_tick(2 * 800, [], [mono, stereo])
    assert mismatch_count(got, mono.download()) == 0
    # asked BEFORE the first run of a fresh graph: no error
    g2, ids2 = W.build_graph(ctx48, d)
    line = g2.output(ids2[osc], 0)
    g2.run_ticks(2, 2)
    assert mismatch_count(line.download(), got) == 0
    for x in (g, g2):
        x.destroy()
    ref.destroy()


def odd_group():
    """A mixer whose channels exercise every shape the fused kernel takes: L != R, one side disconnected, a voice
    shared by two panners, an EqThree without input, a mixer input without panner."""
    d = W.GraphDesc("odd_group")
    o = [d.add("Oscillator", (110.0 * (k + 1), w, 0)) for k, w in enumerate([2, 5, 4, 3])]   # sine, saw, triangle, square
    e = [d.add("EqThree", (3.0 - k, -2.0 + k, 1.5 * k)) for k in range(5)]
    for k in range(4):
        d.connect(e[k], 0, o[k], 0)             # e[4] has no input
    p = [d.add("StereoPanner") for _ in range(5)]
    d.connect(p[0], 0, e[0], 0); d.connect(p[0], 1, e[1], 0)      # L != R
    d.connect(p[1], 0, e[2], 0)                                    # R disconnected
    d.connect(p[2], 1, e[2], 0)                                    # the same voice again, L disconnected
    d.connect(p[3], 0, e[3], 0); d.connect(p[3], 1, e[3], 0)
    d.connect(p[4], 0, e[4], 0); d.connect(p[4], 1, e[0], 0)      # silent EqThree (VSA only) | shared voice
    mix = d.add("Mixer", [(0.0, 1.0, True), (-3.0, 0.7, False), (2.0, 0.9, True), (-12.0, 1.0, False), (6.0, 0.3, True), (0.0, 1.0, True)])
    for k in range(5):
        d.connect(mix, k, p[k], 0)              # input 5 stays disconnected
    met = d.add("Meter")
    d.connect(met, 0, mix, 0)
    d.taps.update(master=(mix, 0), cue=(mix, 1), meter=(met, None))
    return d, e


@pytest.mark.parametrize("sr_spt", [(48000, 800), (44100, 735)])
def test_odd_shapes_fused_equals_staged_and_oracle(mxl, oracle, sr_spt):
    sr, spt = sr_spt
    d, e = odd_group()
    taps = [d.taps["master"], d.taps["cue"]] + [(x, 0) for x in e]
    calls = [3, 1, 5]                           # 735 * 3 frames: an odd number of frames in a call
    with mxl.Context(0, sr, spt) as ctx:
        r = run_pair(mxl, ctx, d, calls, taps, meter=d.taps["meter"][0], eqs=e)
    (fl, fm, fs, fk), (sl, sm, ss, sk) = r[True], r[False]
    assert W.STAGE_KIND["FusedVoiceMix"] in fk
    for t in taps:
        assert mismatch_count(fl[t], sl[t]) == 0, t
    assert_same_meter(fm, sm)
    assert_same_state(fs, ss)
    want, _, _ = oracle_run(oracle, d, sr, spt, 0, sum(calls), d.taps["master"], 2)
    assert_close_audio(fl[d.taps["master"]], want, what="odd group master")


def test_group_with_outside_consumer_stays_staged(mxl, ctx48):
    """A Plotter on a panner's line would be scheduled before the fused launch: such a group is not fused."""
    d = W.config4_audio_graph()
    pan = [i for i, (k, _) in enumerate(d.modules) if k == "StereoPanner"][1]
    plo = d.add("Plotter")
    d.connect(plo, 0, pan, 0)
    g, ids = W.build_graph(ctx48, d)
    g.run_ticks(0, 6)
    kinds = stage_kinds(g)
    assert W.STAGE_KIND["FusedVoiceMix"] not in kinds and W.KIND["Mixer"] in kinds
    g.destroy()


def test_parameter_leaving_the_fused_domain_falls_back(mxl, oracle, ctx48):
    """An oscillator frequency of +inf makes NaN samples, which the reference's EqThree keeps for ever: the staged
    EqThree kernel models that, the fused one refuses -- the group returns to stages before the run."""
    d = W.config4_audio_graph()
    oscs = [i for i, (k, _) in enumerate(d.modules) if k == "Oscillator"]
    res = {}
    for fused in (True, False):
        g, ids = W.build_graph(ctx48, d)
        g.set_fusion(fused)
        g.run_ticks(0, 3)
        if fused:
            assert W.STAGE_KIND["FusedVoiceMix"] in stage_kinds(g)
        g.module(ids[oscs[1]]).update((float("inf"), mxl.WAVE_SINE, 0))
        g.run_ticks(3, 3)
        assert W.STAGE_KIND["FusedVoiceMix"] not in stage_kinds(g)
        g.module(ids[oscs[1]]).update((330.0, mxl.WAVE_SINE, 0))
        g.run_ticks(6, 2)                       # poles are NaN for good (eq_three.rs: nothing resets them)
        res[fused] = g.output(ids[d.taps["master"][0]], 0).download()
        g.destroy()
    assert np.all(np.isnan(res[True])) and np.all(np.isnan(res[False]))


def test_switching_fusion_between_calls_continues_state(mxl, ctx48):
    d = W.config2_graph()
    m = d.taps["master"]
    g, ids = W.build_graph(ctx48, d)
    parts = []
    for i, n in enumerate([4, 9, 2, 6]):
        g.set_fusion(i % 2 == 0)
        g.run_ticks(sum([4, 9, 2, 6][:i]), n)
        parts.append(g.output(ids[m[0]], m[1]).download())
    g.destroy()
    g2, ids2 = W.build_graph(ctx48, d)
    g2.set_fusion(False)
    g2.run_ticks(0, 21)
    whole = g2.output(ids2[m[0]], m[1]).download()
    g2.destroy()
    assert mismatch_count(np.concatenate(parts), whole) == 0


@pytest.mark.parametrize("chunk", ["16", "32", "64"])
def test_every_chunk_length_of_the_voice_kernel(mxl, ctx48, chunk, monkeypatch):
    """The voice kernel picks 16-, 32- or 64-sample chunks by call length; each of them forced on a 300-tick call and on
    short ones: still bit-identical to stages."""
    monkeypatch.setenv("MXL_FUSED_CHUNK", chunk)
    d = W.config2_graph()
    taps = [d.taps["master"], d.taps["cue"]]
    r = run_pair(mxl, ctx48, d, [300, 1, 7], taps, meter=d.taps["meter"][0], eqs=eq_ids(d), tick0=12345)
    for t in taps:
        assert mismatch_count(r[True][0][t], r[False][0][t]) == 0, t
    assert_same_meter(r[True][1], r[False][1])
    assert_same_state(r[True][2], r[False][2])

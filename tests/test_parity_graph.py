"""GPU parity: whole graphs through mxl_graph_run_ticks (Engine::run_tick, src/engine.rs:400-510, for
many ticks per call) against the oracle's tick-by-tick engine walker."""
import numpy as np
import pytest

from helpers import assert_close_audio, build_oracle_graph, mismatch_count, oracle_run, sine_mismatch_budget
from mixlab_b200 import workloads as W

pytestmark = pytest.mark.gpu


def upload_sources(ctx, mxl, g, ids, desc, spt, n_ticks):
    srcs, keep = {}, []
    for mid, (kind, seed) in desc.sources.items():
        if kind == "stereo":
            data = W.uniform_pm1(seed, 2 * spt * n_ticks)
            line = ctx.stereo(data)
            srcs[mid] = (data, 2)
        else:
            data = W.uniform_01(seed, spt * n_ticks)
            line = ctx.mono(data)
            srcs[mid] = (data, 1)
        g.module(ids[mid]).set_source_line(line)
        keep.append(line)
    return srcs, keep


@pytest.mark.parametrize("n_ticks", [1, 3, 64])
def test_config1_mixer_amplifier_exact(mxl, oracle, ctx48, n_ticks):
    spt = 800
    d = W.config1_graph()
    g, ids = W.build_graph(ctx48, d)
    srcs, keep = upload_sources(ctx48, mxl, g, ids, d, spt, n_ticks)
    g.run_ticks(0, n_ticks)
    got = g.output(ids[d.taps["out"][0]], 0).download()
    want, og, oids = oracle_run(oracle, d, 48000, spt, 0, n_ticks, d.taps["out"], 2, srcs)
    assert mismatch_count(got, want) == 0
    got_cue = g.output(ids[d.taps["cue"][0]], 1).download()
    want_cue, _, _ = oracle_run(oracle, d, 48000, spt, 0, n_ticks, d.taps["cue"], 2, srcs)
    assert mismatch_count(got_cue, want_cue) == 0
    assert g.plan() == [ids[i] for i in og.last_order()]
    assert W.algorithmic_bytes_per_tick(d, spt) == 68 * spt
    g.destroy()


@pytest.mark.parametrize("sr_spt", [(48000, 800), (44100, 735)])
def test_config2_32_module_graph(mxl, oracle, sr_spt):
    sr, spt = sr_spt
    n_ticks = 40
    d = W.config2_graph()
    assert W.algorithmic_bytes_per_tick(d, spt) == 464 * spt
    with mxl.Context(0, sr, spt) as ctx:
        g, ids = W.build_graph(ctx, d)
        assert len(g.plan()) == 32
        oscs = [i for i, (k, _) in enumerate(d.modules) if k == "Oscillator"]
        for o in oscs:
            g.pin_output(ids[o], 0)             # observe the oscillator lines (they stay inside the fused launch)
        # two calls of unequal length: module state (EqThree poles) and `t` carry across calls
        g.run_ticks(0, 15)
        first = g.output(ids[d.taps["master"][0]], 0).download()
        osc_first = {o: g.output(ids[o], 0).download() for o in oscs}
        g.run_ticks(15, n_ticks - 15)
        second = g.output(ids[d.taps["master"][0]], 0).download()
        cue = g.output(ids[d.taps["cue"][0]], 1).download()
        osc_lines = {o: np.concatenate([osc_first[o], g.output(ids[o], 0).download()]) for o in oscs}
        got = np.concatenate([first, second])
        want, og, oids = oracle_run(oracle, d, sr, spt, 0, n_ticks, d.taps["master"], 2)
        assert_close_audio(got, want, what="config2 master")
        # Everything behind the oscillators is exact arithmetic in the reference's order, so the bus may differ from
        # the oracle only where a device sine differs from glibc's after `as f32` (isolated samples, normally none):
        # the budget is derived from the oscillator lines themselves, 0 mismatches there -> 0 allowed here
        budget, n_bad = sine_mismatch_budget(oracle, d, sr, spt, 0, n_ticks, osc_lines)
        assert mismatch_count(got, want) <= budget, (mismatch_count(got, want), n_bad)
        want_cue, _, _ = oracle_run(oracle, d, sr, spt, 0, n_ticks, d.taps["cue"], 2)
        assert_close_audio(cue, want_cue[15 * 2 * spt:], what="config2 cue")
        # meter of the last tick
        meter = g.module(ids[d.taps["meter"][0]])
        peak, sumsq, clip = meter.meter_read(n_ticks - 15 - 1)
        opeak, osumsq, oclip = og.meter(oids[d.taps["meter"][0]])
        assert abs(peak[0] - opeak[0]) <= 1e-6 * opeak[0] and abs(peak[1] - opeak[1]) <= 1e-6 * opeak[1]
        assert np.allclose(sumsq, osumsq, rtol=1e-5) and clip == oclip
        assert g.plan() == [ids[i] for i in og.last_order()]
        g.destroy()


def test_exact_modules_graph_bit_exact(mxl, oracle, ctx48):
    """Same topology with Saw/Triangle/On oscillators only: no transcendental anywhere, so the whole
    32-module graph must be bit-exact."""
    spt, n_ticks = 800, 25
    d = W.config2_graph()
    waves = [mxl.WAVE_SAW, mxl.WAVE_TRIANGLE, mxl.WAVE_ON]
    k = 0
    for i, (kind, params) in enumerate(d.modules):
        if kind == "Oscillator":
            d.modules[i] = (kind, (params[0], waves[k % 3], 0))
            k += 1
    g, ids = W.build_graph(ctx48, d)
    g.run_ticks(1000, n_ticks)
    got = g.output(ids[d.taps["master"][0]], 0).download()
    want, _, _ = oracle_run(oracle, d, 48000, spt, 1000, n_ticks, d.taps["master"], 2)
    assert mismatch_count(got, want) == 0
    g.destroy()


def test_all_audio_modules_graph(mxl, oracle, ctx48):
    """Trigger -> Envelope -> FmSine -> Splitter -> Panner -> Amplifier(control = Envelope) -> Mixer."""
    spt, n_ticks = 800, 12
    d = W.GraphDesc("all_audio")
    trg = d.add("Trigger", (mxl.GATE_OPEN,))
    env = d.add("Envelope", (25.0, 500.0, 0.8, 200.0))
    fm = d.add("FmSine", (90.0, 110.0))
    spl = d.add("StereoSplitter")
    pan = d.add("StereoPanner")
    amp = d.add("Amplifier", (0.7, 0.9))
    osc = d.add("Oscillator", (220.0, mxl.WAVE_TRIANGLE, 0))
    mix = d.add("Mixer", [(0.0, 1.0, True), (-6.0, 0.5, False), (3.0, 0.25, True)])
    d.connect(env, 0, trg, 0)
    d.connect(fm, 0, env, 0)
    d.connect(spl, 0, fm, 0)
    d.connect(pan, 0, spl, 1)
    d.connect(pan, 1, spl, 0)
    d.connect(amp, 0, pan, 0)
    d.connect(amp, 1, env, 0)
    d.connect(mix, 0, amp, 0)
    d.connect(mix, 1, osc, 1)        # channel 2 stays disconnected
    g, ids = W.build_graph(ctx48, d)
    g.run_ticks(0, n_ticks)
    got = g.output(ids[mix], 0).download()
    want, og, _ = oracle_run(oracle, d, 48000, spt, 0, n_ticks, (mix, 0), 2)
    assert_close_audio(got, want, what="all_audio master")
    env_got = g.output(ids[env], 0).download()
    env_want, _, _ = oracle_run(oracle, d, 48000, spt, 0, n_ticks, (env, 0), 1)
    assert mismatch_count(env_got, env_want) == 0
    assert g.plan() == [ids[i] for i in og.last_order()]
    g.destroy()


def test_batched_envelopes_one_launch(mxl, oracle, ctx48):
    """Five Envelopes on one dependency level = one batched launch (blockIdx.y = instance): square-wave
    gates of different rates (+1.0 opens, -1.0 is inert), an always-open and an always-closed gate."""
    spt, n_ticks = 800, 16
    d = W.GraphDesc("envs")
    envs = []
    for k, (freq, wave) in enumerate([(3.0, mxl.WAVE_SQUARE), (41.0, mxl.WAVE_SQUARE), (0.7, mxl.WAVE_SQUARE),
                                       (1.0, mxl.WAVE_ON), (1.0, mxl.WAVE_OFF)]):
        osc = d.add("Oscillator", (freq, wave, 0))
        env = d.add("Envelope", (5.0 + k, 80.0 + 10 * k, 0.6, 30.0 + k))
        d.connect(env, 0, osc, 0)
        envs.append(env)
    g, ids = W.build_graph(ctx48, d)
    before = ctx48.launch_count
    g.run_ticks(0, n_ticks)
    assert ctx48.launch_count - before == 2            # one Oscillator launch, one Envelope launch
    g.run_ticks(n_ticks, n_ticks)                      # state carries into a second call
    for env in envs:
        got = g.output(ids[env], 0).download()
        want, _, _ = oracle_run(oracle, d, 48000, spt, 0, 2 * n_ticks, (env, 0), 1)
        assert mismatch_count(got, want[spt * n_ticks:]) == 0
    g.destroy()


def test_cycle_reads_disconnected(mxl, oracle, ctx48):
    # engine.rs:440-442,479-482
    d = W.GraphDesc("cycle")
    src = d.add("Oscillator", (330.0, mxl.WAVE_SAW, 0))
    a = d.add("Amplifier", (0.5, 0.0))
    s = d.add("StereoSplitter")
    p = d.add("StereoPanner")
    mix = d.add("Mixer", [(0.0, 1.0, False), (0.0, 1.0, True)])
    d.connect(s, 0, a, 0)
    d.connect(p, 0, s, 0)
    d.connect(p, 1, s, 1)
    d.connect(a, 0, p, 0)            # back edge: a reads p before p has run -> zeros
    d.connect(mix, 0, p, 0)
    d.connect(mix, 1, src, 1)
    g, ids = W.build_graph(ctx48, d)
    g.run_ticks(0, 4)
    got = g.output(ids[mix], 0).download()
    want, og, _ = oracle_run(oracle, d, 48000, 800, 0, 4, (mix, 0), 2)
    assert mismatch_count(got, want) == 0
    assert g.plan() == [ids[i] for i in og.last_order()]
    g.destroy()


def test_params_update_and_topology_edit_between_runs(mxl, oracle, ctx48):
    spt = 800
    d = W.config1_graph()
    g, ids = W.build_graph(ctx48, d)
    srcs, keep = upload_sources(ctx48, mxl, g, ids, d, spt, 4)
    g.run_ticks(0, 2)
    # UpdateModuleParams between ticks (engine.rs:307-319) and a disconnect (workspace.rs:116-118)
    g.module(ids[6]).update((0.3, 1.0))
    g.disconnect(ids[5], 1)
    d2 = W.config1_graph()
    d2.modules[6] = ("Amplifier", (0.3, 1.0))
    d2.connections = [c for c in d2.connections if not (c[0] == 5 and c[1] == 1)]
    # sources are rings read at t % frames: run the oracle over ticks 2..3 of the same data
    want, _, _ = oracle_run(oracle, d2, 48000, spt, 2, 2, d2.taps["out"], 2, srcs)
    # the device source line is presented from its start each call, so feed it the tail
    for mid, (data, w) in srcs.items():
        keep_line = ctx48.stereo(data[2 * spt * 2:]) if w == 2 else ctx48.mono(data[spt * 2:])
        g.module(ids[mid]).set_source_line(keep_line)
    g.run_ticks(2, 2)
    got = g.output(ids[6], 0).download()
    assert mismatch_count(got, want) == 0
    g.destroy()


@pytest.mark.parametrize("fusion", [True, False])
def test_kernel_timing_brackets_every_launch(mxl, ctx48, fusion):
    """mxl_ctx_set_kernel_timing: one event pair per kernel launch, folded per kernel name."""
    g, ids = W.build_graph(ctx48, W.config2_graph())
    g.set_fusion(fusion)
    g.run_ticks(0, 4)
    ctx48.kernel_times()
    ctx48.set_kernel_timing(True)
    before = ctx48.launch_count
    g.run_ticks(4, 8)
    g.run_ticks(12, 8)
    launched = ctx48.launch_count - before
    times = ctx48.kernel_times()
    ctx48.set_kernel_timing(False)
    assert sum(n for n, _ in times.values()) == launched
    staged = {"oscillator_kernel", "eq_stream_kernel", "panner_kernel", "mixer_kernel", "meter_kernel"}
    assert set(times) == ({"fused_voice_kernel", "fused_mix_kernel"} if fusion else staged)
    assert all(n == 2 and 0.0 < ms < 50.0 for n, ms in times.values())
    g.run_ticks(20, 8)
    assert ctx48.kernel_times() == {}                  # disabled: nothing recorded
    g.destroy()


def test_graph_stage_info_and_launch_count(mxl, ctx48):
    d = W.config2_graph()
    g, ids = W.build_graph(ctx48, d)
    g.set_profiling(True)
    before = ctx48.launch_count
    g.run_ticks(0, 16)
    ctx48.synchronize()
    launched = ctx48.launch_count - before
    stages = g.stages()
    assert sum(s["n_launches"] for s in stages) == launched
    assert sum(s["algorithmic_bytes"] for s in stages) == 464 * 800 * 16
    assert sum(s["n_modules"] for s in stages) == 32
    assert all(s["last_ms"] >= 0 for s in stages if s["n_launches"])
    # one launch serves all ten modules of a kind -- or, fused, two launches the whole graph
    assert launched == 2
    g.set_fusion(False)
    before = ctx48.launch_count
    g.run_ticks(16, 16)
    assert 5 <= ctx48.launch_count - before <= 6
    stages = g.stages()
    assert sum(s["algorithmic_bytes"] for s in stages) == 464 * 800 * 16 and sum(s["n_modules"] for s in stages) == 32
    g.destroy()


@pytest.mark.parametrize("fusion", [True, False])
def test_performance_accounts_shaped_like_engine_stat(mxl, ctx48, fusion):
    """EngineStat::report (src/engine/timing.rs:45-60): an Engine account and one account per module that ran."""
    d = W.config2_graph()
    g, ids = W.build_graph(ctx48, d)
    g.set_fusion(fusion)
    g.set_profiling(True)
    g.run_ticks(0, 16)
    g.run_ticks(16, 16)
    acc = g.performance()
    assert acc[0][0] == -1 and acc[0][3] >= 0.0                      # PerformanceAccount::Engine: host time outside the stages
    mods = acc[1:]
    assert sorted(a[0] for a in mods) == sorted(g.plan())            # every module of the run order, once
    assert all(a[2] > 0.0 and a[3] > 0.0 for a in mods)              # device and host microseconds per tick
    # the ten oscillators share one launch: equal shares
    osc = [a[2] for a in mods if a[1] == mxl.MOD_OSCILLATOR]
    assert len(osc) == 10 and max(osc) == min(osc)
    g.destroy()

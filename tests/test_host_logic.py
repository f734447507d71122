"""CPU (no GPU): the C-ABI library loads, exports every declared symbol, and its host-side logic
(terminals, Workspace::connect, run-order planning, picture geometry) agrees with the reference text
and with the oracle.  No compute entry point is called: they must report MXL_ERR_NO_DEVICE."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mixlab_b200 import workloads as W


@pytest.fixture()
def host_ctx(mxl):
    c = mxl.Context(device=mxl.DEVICE_NONE, sample_rate=48000, samples_per_tick=800)
    yield c
    c.close()


def test_library_exports_every_declared_symbol(mxl):
    declared = mxl.declared_symbols()
    assert len(declared) >= 85
    out = subprocess.check_output(["nm", "-D", "--defined-only", mxl.api.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    # nothing but the ABI leaks out of the library
    assert all(s.startswith("mxl_") for s in exported), sorted(s for s in exported if not s.startswith("mxl_"))[:5]


def test_library_does_not_link_the_oracle(mxl):
    out = subprocess.check_output(["nm", "-D", mxl.api.LIB_PATH], text=True)
    assert "orc_" not in out
    ldd = subprocess.check_output(["ldd", mxl.api.LIB_PATH], text=True)
    assert "liboracle" not in ldd


def test_version_and_error_string(mxl):
    assert b"sm_100a" in mxl.lib().mxl_version()
    assert mxl.lib().mxl_ctx_create(mxl.DEVICE_NONE, 0, 800) is None
    assert "non-zero" in mxl.last_error()


def test_no_device_no_compute(mxl, host_ctx):
    # there is no CPU fallback: every compute entry point refuses a device-less context
    mod = host_ctx.module(mxl.MOD_OSCILLATOR, (440.0, mxl.WAVE_SINE, 0))
    with pytest.raises(mxl.MxlError) as e:
        mod.run_tick(0, [], [])
    assert e.value.status == mxl.ERR_NO_DEVICE
    with pytest.raises(mxl.MxlError):
        host_ctx.line(mxl.LINE_MONO, 16)
    with pytest.raises(mxl.MxlError):
        host_ctx.frame(64, 36)
    g = host_ctx.graph()
    g.add(mxl.MOD_TRIGGER, (mxl.GATE_OPEN,))
    with pytest.raises(mxl.MxlError) as e:
        g.run_ticks(0, 1)
    assert e.value.status == mxl.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    with pytest.raises(mxl.MxlError) as e:
        mxl.Context(device=0)          # this container has no GPU
    assert e.value.status == mxl.ERR_NO_DEVICE


def test_terminals_match_the_reference(mxl, host_ctx):
    M, S, V = mxl.LINE_MONO, mxl.LINE_STEREO, mxl.LINE_VIDEO
    expect = {
        mxl.MOD_AMPLIFIER: ([("Input", S), ("Control", M)], [(None, S)]),                 # amplifier.rs:21-25
        mxl.MOD_ENVELOPE: ([(None, M)], [(None, M)]),                                    # envelope.rs:77-78
        mxl.MOD_EQ_THREE: ([(None, M)], [(None, M)]),                                    # eq_three.rs:42-43
        mxl.MOD_FM_SINE: ([(None, M)], [(None, S)]),                                     # fm_sine.rs:22-23
        mxl.MOD_OSCILLATOR: ([], [("Mono", M), ("Stereo", S)]),                          # oscillator.rs:47-51
        mxl.MOD_PLOTTER: ([(None, S)], []),                                              # plotter.rs:22-23
        mxl.MOD_STEREO_PANNER: ([("L", M), ("R", M)], [(None, S)]),                      # stereo_panner.rs:17-21
        mxl.MOD_STEREO_SPLITTER: ([(None, S)], [("L", M), ("R", M)]),                    # stereo_splitter.rs:17-21
        mxl.MOD_TRIGGER: ([], [(None, M)]),                                              # trigger.rs:27-28
        mxl.MOD_VIDEO_MIXER: ([(str(i + 1), V) for i in range(4)], [("Output", V), ("A", V), ("B", V)]),   # video_mixer.rs:29-36
        mxl.MOD_METER: ([(None, S)], []),
    }
    for kind, (ins, outs) in expect.items():
        mod = host_ctx.module(kind, None)
        assert mod.inputs() == ins and mod.outputs() == outs, kind
        mod.destroy()
    mix = host_ctx.module(mxl.MOD_MIXER, [(0.0, 1.0, False)] * 8)
    assert mix.inputs() == [(str(i + 1), S) for i in range(8)]                           # mixer.rs:22-25
    assert mix.outputs() == [("Master", S), ("Cue", S)]                                  # mixer.rs:26-29
    mix.update([(-3.0, 0.5, True)] * 2)                                                  # update re-creates (mixer.rs:40-44)
    assert len(mix.inputs()) == 2 and mix.params() == [(-3.0, 0.5, True)] * 2


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/mixlab_b200.h must compile as C99 on its own (cgo / bindgen / cffi read it so)."""
    src = tmp_path / "use_header.c"
    src.write_text('#include "mixlab_b200.h"\n'
                   'int probe(void) { mxl_host_ref r; mxl_stage_info s; mxl_video_job j; mxl_audio_fragment f;\n'
                   '  r.len = 0; s.host_us = 0; j.pts = 0; f.n_samples = 0;\n'
                   '  return (int)sizeof(mxl_frame_layout) + (int)r.len + (int)s.host_us + (int)j.pts + (int)f.n_samples + MXL_MOD_MONITOR; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", inc, "-c", str(src), "-o", str(tmp_path / "use_header.o")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout


def _build_c_example(mxl, tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "host_engine")
    libdir = os.path.dirname(mxl.api.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                        os.path.join(root, "examples", "host_engine.c"), "-L", libdir, "-lmixlab_b200",
                        "-Wl,-rpath," + libdir, "-o", exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


def test_c_host_links_and_refuses_to_compute_without_a_device(mxl, tmp_path):
    """examples/host_engine.c: the ABI from plain C (what a cgo / Rust FFI host does).  Without a GPU the compute entry
    points must refuse with MXL_ERR_NO_DEVICE -- there is no CPU fallback."""
    exe = _build_c_example(mxl, tmp_path)
    r = subprocess.run([exe, "--no-device"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_c_host_engine_host_slices_equal_graph(mxl, oracle, tmp_path):
    """the same program on a GPU: config 1 tick by tick through host slices == one graph call, bit for bit; and the
    checksum equals the oracle's."""
    exe = _build_c_example(mxl, tmp_path)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert "host-slice path == graph path" in r.stdout
    # the program's sources are W.uniform_pm1(1..4) / W.uniform_01(5): recompute its checksum through the oracle
    spt, ticks = 800, 12
    ins = [W.uniform_pm1(1 + c, 2 * spt * ticks) for c in range(4)]
    control = W.uniform_01(5, spt * ticks)
    master, _ = oracle.mixer(ins, [0.0, -6.0, 3.0, -12.0], [1.0, 0.8, 0.5, 0.25], [0, 1, 0, 1], spt * ticks)
    want = oracle.amplifier(master, control, 0.9, 0.5)
    s = 0
    for u in want.view(np.uint32).tolist():
        s = (s * 1099511628211 + u) & 0xFFFFFFFFFFFFFFFF
    assert ("checksum %016x" % s) in r.stdout, r.stdout


def test_io_edge_kinds_are_not_provided(mxl, host_ctx):
    for kind in (mxl.MOD_MEDIA_SOURCE, 99):
        with pytest.raises(mxl.MxlError):
            host_ctx.module(kind, None)


def test_stream_input_queues_without_a_device(mxl, host_ctx):
    """StreamInput's run_tick is provided; its queues are host state and work without a GPU."""
    si = host_ctx.module(mxl.MOD_STREAM_INPUT, None)
    assert si.inputs() == [] and si.outputs() == [("Video", mxl.LINE_VIDEO), ("Audio", mxl.LINE_STEREO)]   # stream_input.rs:44-47
    si.stream_write_audio(1, (0, 1), np.zeros(2048, np.int16))
    si.stream_write_audio(1, (1024, 48000), np.zeros(2048, np.int16))
    assert si.stream_pending() == (2, 0)
    with pytest.raises(mxl.MxlError):
        mxl.check(mxl.lib().mxl_stream_input_write_audio(si.h, 1, 0, 0, None, 0))       # zero denominator


def test_params_roundtrip_and_defaults(mxl, host_ctx):
    osc = host_ctx.module(mxl.MOD_OSCILLATOR, None)
    p = osc.params()
    assert (p.freq, p.waveform) == (100.0, mxl.WAVE_SINE)           # frontend default, workspace.rs:469
    env = host_ctx.module(mxl.MOD_ENVELOPE, None)
    p = env.params()
    assert (p.attack_ms, p.decay_ms, p.sustain_amplitude, p.release_ms) == (25.0, 500.0, 0.8, 200.0)   # lib.rs:318-327
    eq = host_ctx.module(mxl.MOD_EQ_THREE, (4.0, 0.0, -6.0))
    p = eq.params()
    assert (p.gain_lo_db, p.gain_mid_db, p.gain_hi_db) == (4.0, 0.0, -6.0)
    eq.update((1.0, 2.0, 3.0))
    assert eq.params().gain_hi_db == 3.0
    with pytest.raises(mxl.MxlError) as e:                           # module.rs:104-110 "module params mismatch!"
        eq.update((1.0, 2.0), kind=mxl.MOD_AMPLIFIER)
    assert e.value.status == mxl.ERR_PARAMS and "mismatch" in str(e.value)
    vm = host_ctx.module(mxl.MOD_VIDEO_MIXER, None)
    p = vm.params()
    assert (p.a, p.b, p.fader) == (-1, -1, 1.0)                      # lib.rs:412-420


def test_connect_errors(mxl, host_ctx):
    g = host_ctx.graph()
    osc = g.add(mxl.MOD_OSCILLATOR, None)
    eq = g.add(mxl.MOD_EQ_THREE, None)
    mix = g.add(mxl.MOD_MIXER, [(0.0, 1.0, False)] * 2)
    g.connect(eq, 0, osc, 0)
    for args, status in [((eq, 1, osc, 0), mxl.ERR_NO_INPUT), ((77, 0, osc, 0), mxl.ERR_NO_INPUT),
                         ((eq, 0, osc, 2), mxl.ERR_NO_OUTPUT), ((eq, 0, 77, 0), mxl.ERR_NO_OUTPUT),
                         ((eq, 0, osc, 1), mxl.ERR_TYPE_MISMATCH), ((mix, 0, eq, 0), mxl.ERR_TYPE_MISMATCH)]:
        with pytest.raises(mxl.MxlError) as e:                      # workspace.rs:97-114
            g.connect(*args)
        assert e.value.status == status, args
    g.connect(mix, 0, osc, 1)
    assert g.plan() == [osc, eq, mix]
    g.remove(osc)                                                   # DeleteModule drops its connections (engine.rs:321-352)
    assert g.plan() == [eq, mix]
    g.destroy()


@pytest.mark.parametrize("desc_fn", [W.config1_graph, W.config2_graph, W.config4_audio_graph])
def test_run_order_matches_the_oracle_engine_walker(mxl, oracle, host_ctx, desc_fn):
    # terminal set + DFS topsort, engine.rs:408-457 (ascending ModuleId as the HashSet order)
    d = desc_fn()
    g, ids = W.build_graph(host_ctx, d)
    og, oids = oracle.build_graph(d, 48000, 800)
    if hasattr(d, "sources"):
        for mid, (kind, seed) in d.sources.items():
            og.set_source(oids[mid], np.zeros(1600, np.float32), 2 if kind == "stereo" else 1)
    og.run_tick(0)
    assert g.plan() == og.last_order()
    g.destroy()


def test_run_order_random_graphs_match_oracle(mxl, oracle, host_ctx):
    rng = np.random.default_rng(5)
    kinds = ["Oscillator", "EqThree", "StereoPanner", "StereoSplitter", "Amplifier", "Envelope", "FmSine", "Trigger", "Meter"]
    for trial in range(40):
        d = W.GraphDesc("rand%d" % trial)
        for _ in range(int(rng.integers(3, 14))):
            k = kinds[int(rng.integers(0, len(kinds)))]
            params = {"Oscillator": (100.0, 2, 0), "EqThree": (0.0, 0.0, 0.0), "Amplifier": (1.0, 0.0),
                      "Envelope": (25.0, 500.0, 0.8, 200.0), "FmSine": (90.0, 110.0), "Trigger": (0,)}.get(k)
            d.add(k, params)
        g, ids = W.build_graph(host_ctx, d)
        og, oids = oracle.build_graph(d, 48000, 64)
        # random type-correct connections, cycles allowed (engine.rs:440-442)
        for _ in range(30):
            im, om = int(rng.integers(0, d.n_modules())), int(rng.integers(0, d.n_modules()))
            mi, mo = g.module(ids[im]), g.module(ids[om])
            if not mi.inputs() or not mo.outputs():
                continue
            ii, oi = int(rng.integers(0, len(mi.inputs()))), int(rng.integers(0, len(mo.outputs())))
            rc = og.connect(oids[im], ii, oids[om], oi)
            if mi.inputs()[ii][1] == mo.outputs()[oi][1]:
                assert rc == 0
                g.connect(ids[im], ii, ids[om], oi)
            else:
                assert rc == -3
                with pytest.raises(mxl.MxlError):
                    g.connect(ids[im], ii, ids[om], oi)
        og.run_tick(0)
        assert g.plan() == og.last_order(), trial
        g.destroy()


def test_picture_geometry_matches_oracle(mxl, oracle):
    rng = np.random.default_rng(9)
    for _ in range(500):
        aw, ah, bw, bh = (int(x) for x in rng.integers(2, 4097, 4))
        assert mxl.unify_picture_settings(aw, ah, bw, bh) == oracle.unify_picture(aw, ah, bw, bh)
        ow, oh = (aw + 1) & ~1, (ah + 1) & ~1
        assert mxl.scale_geometry(bw, bh, ow, oh) == oracle.scale_geometry(bw, bh, ow, oh)
        ml, ol = mxl.frame_layout(aw, ah), oracle.frame_layout(aw, ah)
        assert (list(ml.stride), list(ml.plane_h), list(ml.offset), ml.size) == (list(ol.stride), list(ol.plane_h), list(ol.offset), ol.size)
        assert all(s % 32 == 0 for s in ml.stride)                  # video_mixer.rs:196-201 asserts
    for f in list(np.linspace(-0.5, 1.5, 401)) + [float("nan"), float("inf"), -float("inf")]:
        assert mxl.fader_to_u8(f) == oracle.fader_to_u8(f)          # video_mixer.rs:168
    for db in np.linspace(-60, 12, 145):
        assert mxl.db_to_linear(db) == oracle.db_to_linear(db)      # protocol lib.rs:469-471


def test_div255_identity_used_by_the_crossfade_kernel():
    # video_kernels.cu fade4: x / 255 == (x + 1 + (x >> 8)) >> 8 for every x = a*f + b*(255-f) <= 65025,
    # and the intermediate fits 16 bits (no carry across SIMD-in-register lanes)
    x = np.arange(0, 65026, dtype=np.uint32)
    y = x + 1 + (x >> 8)
    assert np.array_equal(y >> 8, x // 255) and int(y.max()) < 65536


def test_workload_byte_accounting():
    # algorithmic bytes per tick
    assert W.algorithmic_bytes_per_tick(W.config2_graph(), 800) == 464 * 800 == 371200
    assert W.algorithmic_bytes_per_tick(W.config4_audio_graph(), 800) == 236 * 800
    assert W.algorithmic_bytes_per_tick(W.config1_graph(), 800) == 68 * 800 == 54400
    assert W.CROSSFADE_BYTES_PER_FRAME == 9331200
    a = W.uniform_pm1(1, 1000)
    assert a.dtype == np.float32 and a.min() >= -1.0 and a.max() < 1.0 and np.array_equal(a, W.uniform_pm1(1, 1000))
    assert int(W.splitmix64(0, 1)[0]) == 0xE220A8397B1DCDAF        # published splitmix64 first output for seed 0


def test_rust_sys_declares_every_exported_symbol(mxl):
    """rust/src/sys.rs (generated from the header by tools/gen_rust_sys.py; uncompiled: no rustc in this image) declares
    every entry point the header does, once, and is up to date with the header."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "rust", "src", "sys.rs")
    before = open(path).read()
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "gen_rust_sys.py")], stdout=subprocess.DEVNULL)
    assert open(path).read() == before, "rust/src/sys.rs is stale: run tools/gen_rust_sys.py"
    declared = re.findall(r"pub fn (mxl_[a-z0-9_]+)\(", before)
    assert sorted(declared) == sorted(mxl.declared_symbols()) and len(set(declared)) == len(declared)


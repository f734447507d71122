"""Shared helpers of the parity tests: the same GraphDesc instantiated on the CPU oracle."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def oracle_kind(po, name):
    return {
        "Amplifier": po.MOD_AMPLIFIER, "Envelope": po.MOD_ENVELOPE, "EqThree": po.MOD_EQ_THREE,
        "FmSine": po.MOD_FM_SINE, "Mixer": po.MOD_MIXER, "Oscillator": po.MOD_OSCILLATOR,
        "Plotter": po.MOD_PLOTTER, "StereoPanner": po.MOD_STEREO_PANNER,
        "StereoSplitter": po.MOD_STEREO_SPLITTER, "Trigger": po.MOD_TRIGGER, "Meter": po.MOD_METER,
        "SourceStereo": po.MOD_SOURCE_STEREO, "SourceMono": po.MOD_SOURCE_MONO,
    }[name]


def oracle_params(name, params):
    if params is None:
        return ()
    if name == "Mixer":
        flat = [float(len(params))]
        for g, f, c in params:
            flat += [float(g), float(f), 1.0 if c else 0.0]
        return flat
    if name == "Oscillator":
        return [float(params[0]), float(params[1])]
    if name == "Trigger":
        return [1.0 if params[0] == 0 else 0.0]      # GATE_OPEN = 0
    return [float(p) for p in params]


def build_oracle_graph(po, desc, sample_rate, spt):
    g = po.Graph(float(sample_rate), spt)
    ids = [g.add(oracle_kind(po, kind), oracle_params(kind, params)) for kind, params in desc.modules]
    for im, ii, om, oi in desc.connections:
        rc = g.connect(ids[im], ii, ids[om], oi)
        assert rc == 0, (rc, im, ii, om, oi)
    return g, ids


def oracle_run(po, desc, sample_rate, spt, tick0, n_ticks, tap, width, sources=None):
    """Runs the oracle engine tick by tick; returns the concatenated tap line."""
    g, ids = build_oracle_graph(po, desc, sample_rate, spt)
    for mid, (data, w) in (sources or {}).items():
        g.set_source(ids[mid], data, w)
    out = np.empty(n_ticks * spt * width, np.float32)
    for k in range(n_ticks):
        buf = g.run_tick(tick0 + k, (ids[tap[0]], tap[1]), spt * width)
        out[k * spt * width:(k + 1) * spt * width] = buf
    return out, g, ids


def load_f32(path):
    return np.fromfile(path, dtype="<f4")


def assert_close_audio(got, ref, rtol=1e-6, what=""):
    """north_star tolerance: f32 audio within 1e-6 relative; the absolute floor is 1e-6 of the
    line's peak so that samples near a zero crossing are not held to 1e-6 of ~0."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    peak = float(np.max(np.abs(ref))) if ref.size else 0.0
    err = np.abs(got - ref)
    bound = rtol * np.abs(ref) + rtol * max(peak, 1e-30)
    bad = np.nonzero(err > bound)[0]
    assert bad.size == 0, "%s: %d / %d samples out of tolerance, first at %d: got %r ref %r" % (
        what, bad.size, ref.size, bad[0], got[bad[0]], ref[bad[0]])


def mismatch_count(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return int(np.count_nonzero(a.view(np.uint32) != b.view(np.uint32))) if a.dtype == np.float32 else int(np.count_nonzero(a != b))

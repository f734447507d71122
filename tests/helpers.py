"""Shared helpers of the parity tests: the same GraphDesc instantiated on the CPU oracle."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def build_oracle_graph(po, desc, sample_rate, spt):
    return po.build_graph(desc, sample_rate, spt)


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


ABS_FLOOR = 1e-7      # f32 audio within 1e-6 relative, with an absolute floor of 1e-7 near zero


def assert_close_audio(got, ref, rtol=1e-6, what="", atol=ABS_FLOOR):
    """north_star tolerance: |got - ref| <= 1e-6 * |ref| + 1e-7 per sample."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = np.abs(got - ref)
    bound = rtol * np.abs(ref) + atol
    bad = np.nonzero(~(err <= bound))[0]
    assert bad.size == 0, "%s: %d / %d samples out of tolerance, first at %d: got %r ref %r" % (
        what, bad.size, ref.size, bad[0], got[bad[0]], ref[bad[0]])


def sine_mismatch_budget(po, desc, sample_rate, spt, tick0, n_ticks, gpu_lines, span=4096):
    """How many f32 of a bus downstream of the oscillators may differ from the oracle: the device sine (dsp_math.cuh)
    and glibc's agree after `as f32` on all but isolated samples (tests/test_dsp_math.py).  `gpu_lines` maps the
    GraphDesc index of every Oscillator to its Mono line as the device produced it; each oscillator sample that
    differs from the oracle's may disturb the filters behind it for `span` samples per channel (the EqThree cascades
    forget a 1-ulp kick well inside that), and nothing else may differ at all: 0 oscillator mismatches -> 0 allowed."""
    n_bad = 0
    for mid, got in gpu_lines.items():
        kind, params = desc.modules[mid]
        assert kind == "Oscillator"
        want, _ = po.oscillator(tick0 * spt, sample_rate, params[0], params[1], n_ticks * spt)
        n_bad += mismatch_count(np.asarray(got, np.float32), want)
    return n_bad * span * 2, n_bad


def mismatch_count(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return int(np.count_nonzero(a.view(np.uint32) != b.view(np.uint32))) if a.dtype == np.float32 else int(np.count_nonzero(a != b))

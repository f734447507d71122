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

"""CPU: mixlab_b200/csrc/dsp_math.cuh compiled for the host (same source as the device code) against
glibc -- what Rust's f64::sin / `/` resolve to on the reference's Linux target."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_math", "dsp_math_host.cpp")
LIB = os.path.join(HERE, "host_math", "libdsp_math_host.so")


@pytest.fixture(scope="module")
def hm():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(
            os.path.join(HERE, "..", "mixlab_b200", "csrc", "dsp_math.cuh"))):
        subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", LIB, SRC])
    return C.CDLL(LIB)


def _ulps(a, b):
    return np.abs(a.view(np.int64) - b.view(np.int64))


def test_sine_within_one_ulp_of_libm_on_oscillator_phases(hm):
    worst = 0
    n = 200000
    for t0, sr, freq in [(0, 48000.0, 100.0), (48000 * 3600, 48000.0, 440.0), (48000 * 3600 * 24 * 365, 48000.0, 19999.0),
                         (1 << 40, 44100.0, 55.0 * 2 ** (7 / 3.0)), (12345678901, 44100.0, 12000.0)]:
        x = np.empty(n)
        hm.mxl_host_osc_phase(C.c_ulonglong(t0), C.c_double(sr), C.c_double(freq), x.ctypes.data_as(C.c_void_p), C.c_size_t(n))
        got, want = np.empty(n), np.empty(n)
        hm.mxl_host_sin(x.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p), C.c_size_t(n))
        hm.mxl_host_libm_sin(x.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p), C.c_size_t(n))
        big = np.abs(want) > 1e-300
        worst = max(worst, int(_ulps(got[big], want[big]).max()))
        # after `as f32` (oscillator.rs:86) virtually every sample is identical
        assert np.count_nonzero(got.astype(np.float32) != want.astype(np.float32)) <= 2
        assert np.array_equal(np.signbit(got), np.signbit(want))           # Square = sign(sin), oscillator.rs:15-23
    assert worst <= 2


def test_sine_random_arguments(hm):
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-1e6, 1e6, 200000), rng.uniform(-10, 10, 200000), rng.uniform(-1e12, 1e12, 100000),
                        np.array([0.0, -0.0, np.pi, -np.pi, np.pi / 2, 1e-310, 3e13])])
    got, want = np.empty_like(x), np.empty_like(x)
    hm.mxl_host_sin(x.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p), C.c_size_t(x.size))
    hm.mxl_host_libm_sin(x.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p), C.c_size_t(x.size))
    assert np.max(np.abs(got - want)) < 4e-16
    assert np.array_equal(np.signbit(got), np.signbit(want))


def test_four_wide_sine_equals_scalar_sine(hm):
    """sin_f64x4 (what the oscillator kernels call) is the scalar routine step for step, incl. the groups that
    fall back because one member is +-0, huge, inf or NaN."""
    rng = np.random.default_rng(8)
    x = np.concatenate([rng.uniform(-1e7, 1e7, 40000), np.array([0.0, 1.0, 2.0, 3.0, -0.0, 5.0, 1e300, 7.0, np.inf, 1.0, np.nan, 2.0])])
    a, b = np.empty_like(x), np.empty_like(x)
    hm.mxl_host_sin(x.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p), C.c_size_t(x.size))
    hm.mxl_host_sin4(x.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_size_t(x.size))
    assert np.array_equal(a.view(np.uint64)[~np.isnan(a)], b.view(np.uint64)[~np.isnan(b)])
    assert np.array_equal(np.isnan(a), np.isnan(b))


def test_division_by_sample_rate_is_correctly_rounded(hm):
    rng = np.random.default_rng(4)
    for sr in (48000.0, 44100.0, 96000.0, 22050.0):
        a = np.concatenate([np.arange(0, 300000, dtype=np.float64), rng.integers(0, 1 << 52, 300000).astype(np.float64)])
        q = np.empty_like(a)
        hm.mxl_host_div(a.ctypes.data_as(C.c_void_p), C.c_double(sr), q.ctypes.data_as(C.c_void_p), C.c_size_t(a.size))
        assert np.array_equal(q, a / sr)

"""CPU: the EqThree sample arithmetic shared with the device kernel (mixlab_b200/csrc/eq_core.cuh,
eq_plan.h) compiled for the host.  The skewed chunk runner must equal the sequential form bit for
bit; the CPU model of the whole time-parallel scheme must reproduce the reference's golden vector."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_math", "eq_core_host.cpp")
LIB = os.path.join(HERE, "host_math", "libeq_core_host.so")
CSRC = os.path.join(HERE, "..", "mixlab_b200", "csrc")
GOLD = os.path.join(HERE, "golden", "eq_three")


@pytest.fixture(scope="module")
def hm():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("eq_core.cuh", "eq_plan.h", "dsp_math.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", LIB, SRC])
    lib = C.CDLL(LIB)
    lib.mxl_host_eq_plan_size.restype = C.c_size_t
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _gains(db):
    return np.array([10.0 ** (d / 20.0) for d in db], dtype=np.float64)


def _run(hm, fn, x, sr, db, lc=None, state=None):
    out = np.empty_like(x)
    st = np.zeros(11) if state is None else state.copy()
    g = _gains(db)
    args = [_ptr(x), _ptr(out), C.c_size_t(len(x)), _ptr(st), C.c_uint(sr), _ptr(g)]
    if lc is not None:
        args.append(C.c_int(lc))
    getattr(hm, fn)(*args)
    return out, st


def _signal(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    x[n // 3:n // 3 + 4000] = 0.0                      # silence: poles decay to the VSA floor
    x[n // 2:n // 2 + 100] *= 1e-30                    # denormal-range products
    return x


@pytest.mark.parametrize("lc", [8, 16, 32, 64])
def test_skewed_chunks_equal_sequential_bit_for_bit(hm, lc):
    x = _signal(50003, lc)
    st0 = np.random.default_rng(7).uniform(-0.5, 0.5, 11)
    want, sw = _run(hm, "mxl_host_eq_seq", x, 48000, (-6, 0, 4), state=st0)
    got, sg = _run(hm, "mxl_host_eq_skewed", x, 48000, (-6, 0, 4), lc=lc, state=st0)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
    assert np.array_equal(sw.view(np.uint64), sg.view(np.uint64))


@pytest.mark.parametrize("sr,lc", [(48000, 16), (48000, 32), (48000, 64), (44100, 32), (96000, 64)])
def test_plan_halo_and_levels(hm, sr, lc):
    buf = np.zeros(hm.mxl_host_eq_plan_size(), np.uint8)
    assert hm.mxl_host_eq_plan(C.c_uint(sr), C.c_uint(lc), C.c_uint(128), _ptr(buf)) == 1
    halo, lev_lo, lev_hi = (hm.mxl_host_eq_plan_field(_ptr(buf), i) for i in range(3))
    assert 1 <= halo <= 128 and (1 << lev_lo) >= halo and lev_hi <= lev_lo
    # the low cascade forgets like (1-c)^n * n^3: the halo covers about 75 bits of decay
    c_lo = 2.0 * np.sin(np.pi * 420.0 / sr)
    assert halo * lc * -np.log2(1.0 - c_lo) > 75


@pytest.mark.parametrize("lc", [16, 32, 64])
def test_parallel_model_reproduces_golden_vector(hm, lc):
    x = np.fromfile(os.path.join(GOLD, "chronos.f32.raw"), dtype="<f4")
    want = np.fromfile(os.path.join(GOLD, "chronos-eq.f32.raw"), dtype="<f4")
    got, _ = _run(hm, "mxl_host_eq_parallel", x, 44100, (4, 0, 4), lc=lc)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


@pytest.mark.parametrize("lc", [16, 32])
def test_parallel_model_matches_sequential_and_continues_across_calls(hm, lc):
    x = _signal(200000, 99)
    want, sw = _run(hm, "mxl_host_eq_seq", x, 48000, (-6, 3, 4))
    # two calls, the second from the first one's state, ragged split
    cut = 77777
    a, st = _run(hm, "mxl_host_eq_parallel", x[:cut].copy(), 48000, (-6, 3, 4), lc=lc)
    b, st = _run(hm, "mxl_host_eq_parallel", x[cut:].copy(), 48000, (-6, 3, 4), lc=lc, state=st)
    got = np.concatenate([a, b])
    mism = np.count_nonzero(want.view(np.uint32) != got.view(np.uint32))
    assert mism <= 2, mism                             # carry noise is ~1e-16 relative: flips are ~1e-9 per sample
    assert np.allclose(st, sw, rtol=1e-12, atol=1e-24)

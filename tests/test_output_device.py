"""GPU parity of OutputDevice::run_tick (src/module/output_device.rs:173-207): routing of the stereo line into the
device's interleaved buffer, the clip predicate, the 65536-sample ring.  Oracle: oracle/pyoracle.py::OutputDevice."""
import numpy as np
import pytest

from mixlab_b200 import workloads as W

pytestmark = pytest.mark.gpu

SPT = 800


def bits_equal(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


def opt(v):
    return -1 if v is None else v


@pytest.mark.parametrize("left,right,channels", [(0, 1, 2), (1, 0, 2), (0, None, 2), (None, 3, 6), (2, 2, 4), (0, 7, 2), (None, None, 2)])
def test_routing_and_clip(mxl, oracle, ctx48, left, right, channels):
    x = (W.uniform_pm1(5, 2 * SPT * 6) * np.float32(0.99)).astype(np.float32)
    x[2 * SPT * 2 + 11] = np.float32(1.5)          # tick 2, right side
    x[2 * SPT * 4 + 10] = np.float32(-1.25)        # tick 4, left side
    x[2 * SPT * 5 + 3] = np.float32(np.nan)        # NaN fails both comparisons: no clip
    mod = ctx48.module(mxl.MOD_OUTPUT_DEVICE, (opt(left), opt(right), channels, 0))
    orc = oracle.OutputDevice(left, right, channels)
    for k in range(6):
        tick = x[2 * SPT * k:2 * SPT * (k + 1)]
        mod.run_tick(k * SPT, [ctx48.stereo(tick)], [])
        assert mod.output_device_clip() == orc.run_tick(tick), k
        got = mod.output_device_read(SPT * channels + 5)
        assert bits_equal(got, orc.pop(SPT * channels + 5)), k
    # several ticks in one call route the same way
    mod.run_tick(0, [ctx48.stereo(x)], [])
    clip = orc.run_tick(x)
    assert mod.output_device_clip() == clip
    assert bits_equal(mod.output_device_read(1 << 20), orc.pop(1 << 20))


def test_reassignment_clears_left_over_data(mxl, oracle, ctx48):
    x = W.uniform_pm1(9, 2 * SPT)
    mod = ctx48.module(mxl.MOD_OUTPUT_DEVICE, (0, 1, 4, 0))
    orc = oracle.OutputDevice(0, 1, 4)
    for left, right in ((0, 1), (2, 3), (3, None), (3, None), (9, 1)):
        mod.update((opt(left), opt(right), 4, 0))
        orc.update(left, right)
        mod.run_tick(0, [ctx48.stereo(x)], [])
        orc.run_tick(x)
        assert bits_equal(mod.output_device_read(4 * SPT), orc.pop(4 * SPT)), (left, right)
    p = mod.params()
    assert (p.left, p.right) == (-1, 1)             # 9 is beyond the device's channels (output_device.rs:164-168)


def test_ring_takes_only_what_fits_and_disconnected_input(mxl, oracle, ctx48):
    mod = ctx48.module(mxl.MOD_OUTPUT_DEVICE, (0, 1, 2, 0))
    orc = oracle.OutputDevice(0, 1, 2)
    x = W.uniform_pm1(3, 2 * 30000)
    for _ in range(3):                              # 3 x 60000 samples into a 65536-sample ring
        mod.run_tick(0, [ctx48.stereo(x)], [])
        orc.run_tick(x)
    got = mod.output_device_read(1 << 20)
    want = orc.pop(1 << 20)
    assert got.size == want.size == 65536 and bits_equal(got, want)
    mod.run_tick(0, [None], [])                     # the engine's static zero buffer: one tick of silence (io.rs:8-9)
    orc.run_tick(np.zeros(2 * SPT, np.float32))
    assert bits_equal(mod.output_device_read(1 << 20), orc.pop(1 << 20))
    assert mod.output_device_clip() is False


def test_no_stream_queues_nothing(mxl, ctx48):
    mod = ctx48.module(mxl.MOD_OUTPUT_DEVICE, (0, 1, 0, 0))
    mod.run_tick(0, [ctx48.stereo(np.full(2 * SPT, 2.0, np.float32))], [])
    assert mod.output_device_read(10).size == 0 and mod.output_device_clip() is False   # clip is only looked for with a stream (178)

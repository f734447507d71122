"""One tick of a wide mixer bus: host time per call and device time per kernel launch (us), with the wide-bus kernel and,
under MXL_MIXER_NO_WIDE=1, without it."""
import sys, os, json, time
sys.path.insert(0, os.getcwd())
import numpy as np
import mixlab_b200 as mxl
from mixlab_b200 import workloads as W
ctx = mxl.Context(0, 48000, 800)
for C, S in ((256, 64), (256, 800), (64, 800), (32, 4096)):
    ins = [ctx.stereo(W.uniform_pm1(c + 1, 2 * S)) for c in range(C)]
    mod = ctx.module(mxl.MOD_MIXER, [(-3.0, 0.5, c % 2 == 0) for c in range(C)])
    m, q = ctx.line(mxl.LINE_STEREO, S), ctx.line(mxl.LINE_STEREO, S)
    for _ in range(50):
        mod.run_tick(0, ins, [m, q])
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(500):
        mod.run_tick(0, ins, [m, q])
    host = (time.perf_counter() - t0) / 500 * 1e6
    ctx.synchronize()
    ctx.set_kernel_timing(True)
    for _ in range(200):
        mod.run_tick(0, ins, [m, q])
    times = ctx.kernel_times()
    ctx.set_kernel_timing(False)
    print(json.dumps({"C": C, "S": S, "host_us_per_call": round(host, 1), "kernel": {k: round(v[1] / v[0] * 1e3, 2) for k, v in times.items()}, "launches_per_call": sum(v[0] for v in times.values()) / 200}))

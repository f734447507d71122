"""Times the fused voice kernel (and the staged path beside it) over voices x ticks-per-call; run under gpurun.
MXL_DEBUG=1 makes the library print the resident-cluster count it found for each (chunk, voices)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mixlab_b200 as mxl
from mixlab_b200 import workloads as W


def run(ctx, nv, T, fusion, K=30):
    d = W.osc_eq_pan_mixer(nv, True)
    g, ids = W.build_graph(ctx, d)
    g.set_fusion(fusion)
    tick = 0
    for _ in range(3):
        g.run_ticks(tick, T); tick += T
    ctx.synchronize()
    import time
    t0 = time.perf_counter()
    ctx.timer_begin()
    for _ in range(K):
        g.run_ticks(tick, T); tick += T
    ctx.timer_end()
    host_us = (time.perf_counter() - t0) / K * 1e6
    ms = ctx.timer_elapsed_ms() / K
    host_stage = sum(s["host_us"] for s in g.stages())
    ctx.kernel_times()
    ctx.set_kernel_timing(True)
    for _ in range(K):
        g.run_ticks(tick, T); tick += T
    kt = ctx.kernel_times()
    ctx.set_kernel_timing(False)
    g.destroy()
    return ms, {k: round(v[1] / v[0] * 1e3, 2) for k, v in kt.items()}, round(host_us, 2), round(host_stage, 2)


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else None
    with mxl.Context(0, 48000, 800) as ctx:
        for nv in (1, 2, 5, 10, 16):
            for T in (1, 8, 128, 1024):
                if only and only != "%d:%d" % (nv, T):
                    continue
                row = {"voices": nv, "ticks": T}
                for fusion in (True, False):
                    ms, kt, host_us, host_stage = run(ctx, nv, T, fusion)
                    row["fused" if fusion else "staged"] = {"us_per_call": round(ms * 1e3, 2), "kernel_us": kt, "host_enqueue_us_per_call": host_us, "host_us_in_stages": host_stage}
                print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()

"""Where a live (one tick per call) engine thread spends its host time: median host enqueue time per stage of
mxl_graph_run_ticks (mxl_stage_info.host_us) and per call, for the bench's A/V session.

    python tools/live_profile.py [--ticks 3000] [--ticks-per-call 1]
"""
import argparse
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import mixlab_b200 as mxl  # noqa: E402
from mixlab_b200 import workloads as W  # noqa: E402
from mixlab_b200.session import AVSession  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ticks", type=int, default=3000)
    ap.add_argument("--ticks-per-call", type=int, default=1)
    ap.add_argument("--no-video", action="store_true")
    args = ap.parse_args()
    T = args.ticks_per_call
    ctx = mxl.Context(device=0, sample_rate=48000, samples_per_tick=800)
    sess = AVSession(ctx, W.config2_graph(), T, video=not args.no_video, unique_frames=2)
    sess.upload_inputs()
    kind_names = {v: k for k, v in W.KIND.items()}
    kind_names[mxl.STAGE_FUSED_VOICE_MIX] = "FusedVoiceMix"
    tick = 0
    for _ in range(100):
        sess.run_step(tick)
        tick += T
    ctx.synchronize()
    per_stage, per_call = {}, []
    t_all = time.perf_counter()
    for _ in range(args.ticks):
        t0 = time.perf_counter()
        sess.run_step(tick)
        per_call.append((time.perf_counter() - t0) * 1e6)
        tick += T
        for s in sess.graph.stages():
            if s["n_launches"]:
                per_stage.setdefault(kind_names.get(s["kind"], "kind%d" % s["kind"]), []).append(s["host_us"])
    ctx.synchronize()
    wall = (time.perf_counter() - t_all) / args.ticks * 1e6
    out = {"ticks_per_call": T, "calls": args.ticks, "wall_us_per_call_incl_stage_queries": wall,
           "host_us_per_call_median": statistics.median(per_call),
           "stage_host_us_median": {k: round(statistics.median(v), 2) for k, v in per_stage.items()}}
    out["stage_sum_us"] = round(sum(out["stage_host_us_median"].values()), 2)
    print(json.dumps(out))
    sess.close()
    ctx.close()


if __name__ == "__main__":
    main()

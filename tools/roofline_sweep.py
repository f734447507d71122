#!/usr/bin/env python
"""BASELINE config 5: HBM roofline sweep of the mixer bus, C channels x S samples per tick, through the
C ABI (mxl_module_run_tick).  Two variants per point: one tick per launch (launch-latency bound) and
256 ticks per launch (bandwidth bound).  Algorithmic bytes = 8*S*(C+2) per tick.
Inputs rotate over enough buffer sets to exceed the 126 MB L2.  Prints one JSON object per point."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mixlab_b200 as mxl
from mixlab_b200 import workloads as W


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", default="2,4,8,16,32,64,128,256")
    ap.add_argument("--samples", default="64,256,1024,4096,16384,65536")
    ap.add_argument("--batch-ticks", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--max-bytes", type=float, default=6e9)
    args = ap.parse_args()
    peak = 6549.8
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    ctx = mxl.Context(0, 48000, 800)
    L = mxl.lib()
    for C in [int(x) for x in args.channels.split(",")]:
        rng = np.random.default_rng(C)
        gains = rng.uniform(-24.0, 6.0, C)
        faders = rng.uniform(0.1, 1.0, C)
        cues = (np.arange(C) % 2 == 0)
        mod = ctx.module(mxl.MOD_MIXER, list(zip(gains, faders, cues)))
        for S in [int(x) for x in args.samples.split(",")]:
            for ticks in (1, args.batch_ticks):
                frames = S * ticks
                bytes_per_launch = 8 * frames * (C + 2)
                if bytes_per_launch > args.max_bytes:
                    continue
                n_sets = int(max(1, min(64, np.ceil(300e6 / bytes_per_launch))))
                while n_sets > 1 and n_sets * bytes_per_launch > args.max_bytes:
                    n_sets -= 1
                base = W.uniform_pm1(C * (1 << 20) + S, 2 * min(frames, 1 << 16))
                host = np.resize(base, 2 * frames)
                sets = []
                for _ in range(n_sets):
                    ins = [ctx.stereo(host) for _ in range(C)]
                    sets.append((ins, ctx.line(mxl.LINE_STEREO, frames), ctx.line(mxl.LINE_STEREO, frames)))
                for i in range(3):
                    ins, m, c = sets[i % n_sets]
                    mod.run_tick(0, ins, [m, c])
                ctx.synchronize()
                launches0 = ctx.launch_count
                ctx.timer_begin()
                for i in range(args.reps):
                    ins, m, c = sets[i % n_sets]
                    mod.run_tick(0, ins, [m, c])
                ctx.timer_end()
                ms = ctx.timer_elapsed_ms() / args.reps
                gbs = bytes_per_launch / (ms * 1e-3) / 1e9
                print(json.dumps({"channels": C, "samples_per_tick": S, "ticks_per_launch": ticks, "bytes_per_launch": bytes_per_launch,
                                  "ms_per_launch": ms, "gbs": gbs, "frac_of_measured_peak": gbs / peak, "buffer_sets": n_sets,
                                  "launches_per_call": (ctx.launch_count - launches0) / args.reps,
                                  "stereo_frames_per_s": frames / (ms * 1e-3)}), flush=True)
                for ins, m, c in sets:
                    for ln in ins + [m, c]:
                        ln.free()
        mod.destroy()
    ctx.close()


if __name__ == "__main__":
    main()

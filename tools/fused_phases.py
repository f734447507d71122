"""Per-phase SM-clock breakdown of fused_voice_kernel (mxl_ctx_fused_profile: one stamp per phase by the first owner
thread of every CTA); run under gpurun.  Cycles of the SM clock (1965 MHz: 1000 cycles = 0.51 us)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mixlab_b200 as mxl
from mixlab_b200 import workloads as W

NAMES = ["tables", "generate+zero", "scan", "osc-out/hist", "exact", "store", "cluster-wait", "mix"]


def main():
    with mxl.Context(0, 48000, 800) as ctx:
        for nv, T in ((1, 1), (10, 1), (10, 8), (10, 128), (10, 1024)):
            d = W.osc_eq_pan_mixer(nv, True)
            g, ids = W.build_graph(ctx, d)
            tick = 0
            for _ in range(3):
                g.run_ticks(tick, T); tick += T
            ctx.fused_profile(4096, read=False)
            g.run_ticks(tick, T); tick += T
            st = ctx.fused_profile(0).astype(np.int64)
            g.destroy()
            if not len(st):
                continue
            dur = np.diff(st[:, :6], axis=1)                      # stamps: start, tables issued, generated, scanned, filtered, stored
            total = st[:, 5] - st[:, 0]
            row = {"voices": nv, "ticks": T, "ctas": int(len(st)), "cta_total_mean": float(total.mean()), "cta_total_max": int(total.max())}
            for i, name in enumerate(["tables issued", "generate+zero (incl. table wait)", "scan", "exact", "store + products"]):
                row[name] = [float(dur[:, i].mean()), int(dur[:, i].max())]
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()

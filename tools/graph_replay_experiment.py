"""EXPERIMENT (not product code): what would a CUDA graph buy a live one-tick call?  Captures one
mxl_graph_run_ticks(tick, 1) of the bench's A/V session into a CUDA graph by stream capture and replays that same graph
(same parameters every time: the results are not a valid tick sequence, only the timing means something), against the
normal eager call.  Upper bound of what per-tick graphs could give before any parameter-update cost.

    python tools/graph_replay_experiment.py [--no-video]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import mixlab_b200 as mxl  # noqa: E402
from mixlab_b200 import workloads as W  # noqa: E402
from mixlab_b200.session import AVSession  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-video", action="store_true")
    ap.add_argument("--calls", type=int, default=3000)
    args = ap.parse_args()
    rt = C.CDLL("libcudart.so.12")
    ctx = mxl.Context(device=0, sample_rate=48000, samples_per_tick=800)
    sess = AVSession(ctx, W.config2_graph(), 1, video=not args.no_video, unique_frames=2)
    sess.upload_inputs()
    for k in range(200):
        sess.run_step(k)
    ctx.synchronize()
    stream = C.c_void_p(mxl.lib().mxl_ctx_stream(ctx.h))

    def timed(fn, n):
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        host = time.perf_counter() - t0
        ctx.synchronize()
        return host / n * 1e6, (time.perf_counter() - t0) / n * 1e6

    tick = [1000]

    def eager():
        sess.run_step(tick[0])
        tick[0] += 1

    eager_host, eager_total = timed(eager, args.calls)

    graph, gexec = C.c_void_p(), C.c_void_p()
    st = rt.cudaStreamBeginCapture(stream, 2)                       # cudaStreamCaptureModeRelaxed
    assert st == 0, ("cudaStreamBeginCapture", st)
    sess.run_step(tick[0])
    st = rt.cudaStreamEndCapture(stream, C.byref(graph))
    assert st == 0 and graph.value, ("cudaStreamEndCapture", st)
    n_nodes = C.c_size_t()
    rt.cudaGraphGetNodes(graph, None, C.byref(n_nodes))
    st = rt.cudaGraphInstantiate(C.byref(gexec), graph, C.c_ulonglong(0))
    assert st == 0, ("cudaGraphInstantiate", st)

    def replay():
        rt.cudaGraphLaunch(gexec, stream)

    replay_host, replay_total = timed(replay, args.calls)
    print(json.dumps({"workload": "audio" if args.no_video else "av", "graph_nodes": n_nodes.value,
                      "eager_us_per_tick": round(eager_total, 2), "eager_host_us": round(eager_host, 2),
                      "graph_replay_us_per_tick": round(replay_total, 2), "graph_replay_host_us": round(replay_host, 2),
                      "note": "replay re-runs ONE captured tick with frozen parameters: timing only"}))
    sess.close()
    ctx.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- graph ticks/s of the Mixlab tick hot path on B200 (BASELINE.json metric).

One step = `--ticks-per-step` engine ticks of one live A/V session: BASELINE config 2's 32-module audio
graph (10 x [Oscillator -> EqThree -> StereoPanner] -> Mixer(10) -> Meter, S = 800 samples @ 48 kHz)
plus config 3's 1080p yuv420p 2-layer VideoMixer crossfade, one composited frame per tick.

  value  device-resident: inputs (the video layers) already in HBM, K steps timed with CUDA events on
         the launching stream, max over ranks.
  e2e    the same ticks through the C ABI with HOST buffers: every step uploads its 2*T layers from
         pinned memory and downloads T composited frames + master bus + meter records.
  roofline  the dominant kernel (VideoMixer crossfade), algorithmic bytes / CUDA-event time vs the
         measured HBM peak.
  cpu_baseline / --impl reference  the C restatement of the reference's CPU engine (oracle/; the Rust
         reference cannot be built here: no cargo) on the box's host cores.

N > 1: one process per GPU (torchrun), one independent session per rank, no data-path collective
("scaling": "weak").
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "graph ticks/sec (48 kHz stereo x32 modules + 1080p composite)"
UNIT = "ticks/s"
SAMPLE_RATE, SPT = 48000, 800


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ticks-per-step", type=int, default=128)
    ap.add_argument("--workload", default="av", choices=["av", "audio", "video"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="bound of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps")
    ap.add_argument("--e2e-mode", default="pipelined", choices=["pipelined", "serial"],
                    help="pipelined: uploads of step k+1 overlap downloads of step k (copy/compute overlap)")
    return ap.parse_args()


def workload_name(args):
    parts = []
    if args.workload in ("av", "audio"):
        parts.append("cfg2 32-module audio graph (10x[Osc->EqThree->Panner]+Mixer(10)+Meter), S=800 @48kHz")
    if args.workload in ("av", "video"):
        parts.append("cfg3 VideoMixer 1080p yuv420p 2-layer crossfade, 1 frame/tick")
    return " + ".join(parts)


# ---------------------------------------------------------------------------------------------
# distributed plumbing (torch.distributed only for barrier / max-reduce; never on the data path)
# ---------------------------------------------------------------------------------------------
class Dist:
    def __init__(self, n_gpus):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.torch, self.dist = torch, dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0]))
                    mx.append(float(f[1]))
                except ValueError:
                    continue
                for name, v in zip(names, f[3:7]):
                    if v == "Active":
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------
# CPU baseline: the oracle (C restatement of the reference CPU engine) on host cores
# ---------------------------------------------------------------------------------------------
class CpuEngines:
    """`threads` independent reference engines (src/engine.rs:78-93: one engine thread per session), each
    on its own host thread (ctypes releases the GIL).  Engines, graphs and input frames are built once;
    run(n) advances every engine by n ticks: per tick Engine::run_tick over the audio graph (topsort +
    alloc/zero + serial dispatch) and VideoMixer's blank + crossfade (video_mixer.rs:151-237)."""

    def __init__(self, po, desc, args, threads):
        from mixlab_b200 import workloads as W
        self.po, self.threads = po, threads
        self.do_audio = args.workload in ("av", "audio")
        self.do_video = args.workload in ("av", "video")
        self.n_ticks = 0
        self.stop = False
        self.go = threading.Barrier(threads + 1)
        self.done = threading.Barrier(threads + 1)
        self.engines = []
        for i in range(threads):
            e = {"tick": 0}
            if self.do_audio:
                e["graph"], _ = po.build_graph(desc, SAMPLE_RATE, SPT)
            if self.do_video:
                e["lay"] = po.frame_layout(W.FRAME_W, W.FRAME_H)
                e["a"] = W.random_bytes(0xC0DE + 7 * i, e["lay"].size)
                e["b"] = W.random_bytes(0xC0DE + 7 * i + 1, e["lay"].size)
                e["f"] = po.fader_to_u8(0.5)
            self.engines.append(e)
        self.workers = [threading.Thread(target=self._work, args=(e,), daemon=True) for e in self.engines]
        for w in self.workers:
            w.start()

    def _work(self, e):
        po = self.po
        while True:
            self.go.wait()
            if self.stop:
                return
            for _ in range(self.n_ticks):
                if self.do_audio:
                    e["graph"].run_tick(e["tick"])
                if self.do_video:
                    po.video_crossfade(e["lay"], e["a"], e["b"], e["f"])
                e["tick"] += 1
            self.done.wait()

    def run(self, n_ticks):
        """Every engine runs n_ticks ticks; returns the wall time of the slowest (seconds)."""
        self.n_ticks = n_ticks
        self.go.wait()
        t0 = time.perf_counter()
        self.done.wait()
        return time.perf_counter() - t0

    def close(self):
        self.stop = True
        self.go.wait()
        for w in self.workers:
            w.join()


def cpu_baseline_leg(args):
    from oracle import pyoracle as po
    from mixlab_b200 import workloads as W
    po.build()
    desc = W.config2_graph()
    cores = os.cpu_count() or 1
    one_engine = CpuEngines(po, desc, args, 1)
    one_engine.run(8)
    one = 24.0 / one_engine.run(24)
    # ~1/4 of the budget on the single engine thread (faithful to the reference), the rest on all cores
    t1 = max(16, int(one * args.cpu_seconds * 0.25))
    dt1 = one_engine.run(t1)
    one_engine.close()
    single = t1 / dt1
    all_engines = CpuEngines(po, desc, args, cores)
    all_engines.run(8)
    tn = max(8, int(single * args.cpu_seconds * 0.75 * 0.6))
    dtn = all_engines.run(tn)
    all_engines.close()
    multi = cores * tn / dtn
    return {
        "value": multi, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": "%d independent engines x %d ticks of the same workload (%.1f s); single engine thread: %d ticks (%.1f s)"
                  % (cores, tn, dtn, t1, dt1),
        "single_thread_value": single,
        "note": "C restatement of the reference CPU engine (oracle/), not the Rust binary: no cargo/rustc in this image",
    }


def reference_arm(args):
    """--impl reference: the oracle port of the reference CPU engine on all host threads -- one
    independent engine (session) per core, built once; a step = every engine advances `tpt` ticks."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import pyoracle as po
    from mixlab_b200 import workloads as W
    po.build()
    desc = W.config2_graph()
    cores = os.cpu_count() or 1
    engines = CpuEngines(po, desc, args, cores)
    engines.run(4)
    per_engine = 16.0 / engines.run(16)                 # ticks/s of one engine with all of them busy
    # sized so that K+W steps stay within ~2 minutes
    budget = 90.0
    tpt = max(1, min(args.ticks_per_step, int(per_engine * budget / max(1, args.steps + args.warmup))))
    for _ in range(args.warmup):
        engines.run(tpt)
    dt = 0.0
    for _ in range(args.steps):
        dt += engines.run(tpt)
    engines.close()
    value = args.steps * cores * tpt / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32 audio + u8 video",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "ticks_per_step": cores * tpt, "samples_per_tick": SPT,
                   "sample_rate": SAMPLE_RATE, "frame": "1920x1080 yuv420p", "sessions": cores,
                   "parallelism": "%d independent engines, one host thread each (the reference runs one engine thread per session)" % cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "each step: %d independent engines x %d ticks" % (cores, tpt),
                         "note": "C restatement of the reference CPU engine (oracle/); the Rust reference cannot be built here"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------
def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def pcie_ceiling():
    """H2D GB/s of the 2-up : 1-down pinned-copy mix measured with tools/pcie_peak.cu on this pool
    (profiles/r1_pcie_peak.jsonl): the ceiling of the e2e leg, which moves two layers up per frame down."""
    try:
        for line in open(os.path.join(ROOT, "profiles", "r1_pcie_peak.jsonl")):
            d = json.loads(line)
            if d.get("mix"):
                return float(d["h2d_gbs"])
    except Exception:
        pass
    return None


def ncu_traffic(kernel, ticks_per_step):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        for e in json.load(open(path)):
            if e["kernel"] == kernel and e["ticks_per_step"] == ticks_per_step:
                return e["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def b200_arm(args):
    import mixlab_b200 as mxl
    from mixlab_b200 import workloads as W
    from mixlab_b200.session import AVSession, session_seed

    dist = Dist(args.gpus)
    if dist.world != args.gpus and dist.world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, dist.world))
    T, K, Wm = args.ticks_per_step, args.steps, max(args.warmup, 3)
    ctx = mxl.Context(device=dist.local_rank, sample_rate=SAMPLE_RATE, samples_per_tick=SPT)
    desc = W.config2_graph() if args.workload in ("av", "audio") else None
    sess = AVSession(ctx, desc, T, video=args.workload in ("av", "video"), seed=session_seed(0xA11CE, dist.rank))
    sess.upload_inputs()
    ctx.synchronize()

    sampler = ClockSampler(dist.local_rank)

    # ---- value: device-resident ----
    tick = 0
    for _ in range(Wm):
        sess.run_step(tick)
        tick += T
    ctx.synchronize()
    sampler.start()
    dist.barrier()
    launches0 = ctx.launch_count
    host_t0 = time.perf_counter()
    ctx.timer_begin()
    for _ in range(K):
        sess.run_step(tick)
        tick += T
    ctx.timer_end()
    host_enqueue_s = time.perf_counter() - host_t0
    ms = ctx.timer_elapsed_ms()
    ctx.synchronize()
    dist.barrier()
    launches = ctx.launch_count - launches0
    ms_max = dist.max(ms)
    total_ticks = K * T * dist.world
    value = total_ticks / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel: per-stage CUDA events over the same K steps.  Two passes:
    # the kernel's own launch duration with the stages serialised on one stream ("ms"), and its
    # duration in the configuration of the timed region, where the audio stages run beside the
    # compositor on a second stream ("ms_overlapped") ----
    kind_names = {v: k for k, v in W.KIND.items()}

    def stage_pass(split):
        nonlocal tick
        sess.graph.set_stream_split(split)
        sess.graph.set_profiling(True)
        acc, nbytes, nlaunch, host = {}, {}, {}, {}
        for _ in range(K):
            sess.run_step(tick)
            tick += T
            for s in sess.graph.stages():
                if s["n_launches"] == 0:
                    continue
                acc.setdefault(s["kind"], []).append(s["last_ms"])
                nbytes[s["kind"]] = s["algorithmic_bytes"]
                nlaunch[s["kind"]] = s["n_launches"]
                host.setdefault(s["kind"], []).append(s["host_us"])
        sess.graph.set_profiling(False)
        sess.graph.set_stream_split(True)
        return {kind_names[k]: {"ms": statistics.mean(v), "launches": nlaunch[k], "algorithmic_bytes": nbytes[k],
                                "host_enqueue_us": statistics.median(host[k])}
                for k, v in acc.items()}

    stages = stage_pass(False)
    for name, st in stage_pass(True).items():
        stages[name]["ms_overlapped"] = st["ms"]
    peak, peak_src = measured_peak()
    for st in stages.values():
        st["gbs"] = st["algorithmic_bytes"] / (st["ms"] * 1e-3) / 1e9
        st["frac_of_hbm_peak"] = st["gbs"] / peak

    # ---- per-kernel device time: CUDA event pairs directly around each launch (mxl_ctx_set_kernel_timing),
    # over K more steps with the stages serialised on one stream (each kernel alone on the device, as in the
    # ncu launch list under profiles/).  Unlike the stage times above these exclude the host-side preparation
    # of a stage and its table copies ----
    ctx.kernel_times()
    sess.graph.set_stream_split(False)
    ctx.set_kernel_timing(True)
    for _ in range(K):
        sess.run_step(tick)
        tick += T
    kt = ctx.kernel_times()
    ctx.set_kernel_timing(False)
    sess.graph.set_stream_split(True)
    kernels = {name: {"launches": n, "avg_launch_ms": ms / n} for name, (n, ms) in kt.items() if n}
    stage_kernel = {"VideoMixer": "crossfade_flat_kernel", "EqThree": "eq_stream_kernel", "Oscillator": "oscillator_kernel",
                    "StereoPanner": "panner_kernel", "Mixer": "mixer_kernel", "Meter": "meter_kernel"}
    for sname, kname in stage_kernel.items():
        if sname in stages and kname in kernels:
            kernels[kname]["algorithmic_bytes_per_launch"] = stages[sname]["algorithmic_bytes"]
            kernels[kname]["gbs"] = stages[sname]["algorithmic_bytes"] / (kernels[kname]["avg_launch_ms"] * 1e-3) / 1e9
            kernels[kname]["frac_of_hbm_peak"] = kernels[kname]["gbs"] / peak
    if args.workload in ("av", "video"):
        dom, dom_kernel = "VideoMixer", "crossfade_flat_kernel"
    else:
        dom = max(stages, key=lambda n: stages[n]["ms"])
        dom_kernel = stage_kernel.get(dom, dom)
    d = stages[dom]
    dk = kernels.get(dom_kernel, {"avg_launch_ms": d["ms"]})
    achieved = d["algorithmic_bytes"] / (dk["avg_launch_ms"] * 1e-3) / 1e9
    total_kernel_ms = sum(v["avg_launch_ms"] * v["launches"] for v in kernels.values()) / K
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(dom_kernel, T), "kernel": dom_kernel, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": d["algorithmic_bytes"], "avg_launch_ms": dk["avg_launch_ms"],
                "stage_ms_serialised": d["ms"], "stage_ms_overlapped": d["ms_overlapped"],
                "step_share": dk["avg_launch_ms"] * dk.get("launches", K) / K / total_kernel_ms if total_kernel_ms else None,
                "timing": "CUDA event pair recorded directly around each launch of the kernel on its launching stream "
                          "(mxl_ctx_set_kernel_timing), averaged over K steps, stages serialised on one stream; "
                          "step_share = this kernel's device time / all kernels' device time per step; "
                          "stage_ms_overlapped = the stage as it runs in the timed region, beside the audio stream"}
    whole = sess.algorithmic_bytes_per_step / (ms_max / K * 1e-3) / 1e9

    # ---- e2e: host buffers through the C ABI ----
    e2e = None
    if not args.no_e2e:
        Ke = args.e2e_steps or K
        if args.e2e_mode == "pipelined":
            sess.enable_pipelining()
        else:
            ctx.set_copy_overlap(False)

        def e2e_steps(n, tick):
            if args.e2e_mode == "pipelined":
                # step i's results are consumed (fence wait) while step i+1 is in flight
                for i in range(n):
                    sess.enqueue_step_host(tick, i & 1)
                    tick += T
                    if i > 0:
                        sess.wait_step((i - 1) & 1)
                sess.wait_step((n - 1) & 1)
            else:
                for _ in range(n):
                    sess.run_step_host(tick)
                    tick += T
            ctx.synchronize()
            return tick

        tick = e2e_steps(3, tick)
        dist.barrier()
        h2d0, d2h0 = ctx.h2d_bytes, ctx.d2h_bytes
        t0 = time.perf_counter()
        tick = e2e_steps(Ke, tick)
        dt = time.perf_counter() - t0
        h2d_step, d2h_step = (ctx.h2d_bytes - h2d0) // Ke, (ctx.d2h_bytes - d2h0) // Ke
        dist.barrier()
        dt_max = dist.max(dt)
        assert h2d_step == sess.h2d_bytes_per_step and d2h_step == sess.d2h_bytes_per_step, (h2d_step, d2h_step)
        e2e = {"value": Ke * T * dist.world / dt_max, "unit": UNIT, "h2d_bytes_per_step": h2d_step,
               "d2h_bytes_per_step": d2h_step, "steps": Ke, "ms_per_step": dt_max / Ke * 1e3, "mode": args.e2e_mode,
               "h2d_gbs": h2d_step * Ke / dt_max / 1e9, "d2h_gbs": d2h_step * Ke / dt_max / 1e9,
               "bound": "pcie" if sess.video else "launch", "pcie_h2d_ceiling_gbs": pcie_ceiling(),
               "timing": "host wall clock around K steps incl. pinned-host copies, synchronised both sides, max over ranks"}
    clocks = sampler.stop()

    cpu = None
    if dist.rank == 0 and dist.world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(args)

    if dist.rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": dist.world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64/f32 audio + u8 video", "data": "synthetic",
            "config": {"workload": workload_name(args), "ticks_per_step": T, "samples_per_tick": SPT,
                       "sample_rate": SAMPLE_RATE, "frame": "1920x1080 yuv420p", "sessions": dist.world,
                       "parallelism": "1 independent session per GPU, no collective",
                       "l2": "inputs larger than L2: %.0f MB read + %.0f MB written per step vs 126 MB L2"
                             % (sess.h2d_bytes_per_step / 1e6 if sess.video else 0, (sess.T * sess.frame_bytes if sess.video else 0) / 1e6)},
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                       "samples": clocks["samples"]},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "stereo_frames_per_s": value * SPT if args.workload != "video" else None,
            "video_fps": value if args.workload != "audio" else None,
            "whole_step_gbs": whole, "whole_step_frac": whole / peak,
            "host_enqueue_ms_per_step": host_enqueue_s / K * 1e3, "stages": stages, "kernels": kernels,
        }
        print(json.dumps(line))
    sess.close()
    ctx.close()
    dist.close()
    return 0


def main():
    args = parse_args()
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.impl == "reference":
        return reference_arm(args)
    return b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())

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
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-benchmarks of the other BASELINE configs (N = 1 only)")
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


class CpuSessions:
    """`threads` independent sessions of the SESSION variant on the CPU (oracle/: orc_session_run, all per-tick work in
    C, one ctypes call per block of ticks so the threads never meet on the GIL): per tick two StreamInput i16 unpacks,
    Engine::run_tick over the audio graph, VideoMixer's blank + crossfade of its stored 1080p layers (new pictures
    every second tick, as a 30 fps source delivers them), the Monitor's scaler to 560 x 350 and its PCM pack."""

    def __init__(self, po, desc, threads):
        from mixlab_b200 import workloads as W
        self.threads = threads
        self.sessions = []
        lay = po.frame_layout(W.FRAME_W, W.FRAME_H)
        for i in range(threads):
            g, ids = po.build_graph(desc, SAMPLE_RATE, SPT)
            la = W.random_bytes(0xC0DE + 7 * i, 2 * lay.size)
            lb = W.random_bytes(0xC0DE + 7 * i + 1, 2 * lay.size)
            pcm = (W.splitmix64(0x51 + i, 64 * 2 * SPT) & np.uint64(0xFFFF)).astype(np.uint16).view(np.int16)
            self.sessions.append(po.Session(g, ids[desc.taps["master"][0]], W.FRAME_W, W.FRAME_H, 560, 350, la, lb, 2, pcm, 0.5))
        self.tick = 0

    def run(self, n_ticks):
        """Every session runs n_ticks ticks on its own thread; returns the wall time of the slowest (seconds)."""
        t0 = time.perf_counter()
        workers = [threading.Thread(target=s.run, args=(self.tick, n_ticks)) for s in self.sessions]
        for w in workers:
            w.start()
        for w in workers:
            w.join()
        self.tick += n_ticks
        return time.perf_counter() - t0


def cpu_session_rate(po, desc, cores, seconds):
    """ticks/s of `cores` CPU sessions (session variant) over about `seconds` of wall time."""
    eng = CpuSessions(po, desc, cores)
    eng.run(2)
    per = 4.0 / eng.run(4)
    n = max(4, int(per * seconds))
    dt = eng.run(n)
    return cores * n / dt, n, dt


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
    sess_rate, sess_n, sess_dt = cpu_session_rate(po, desc, cores, 4.0) if args.workload == "av" else (None, 0, 0.0)
    return {
        "value": multi, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": "%d independent engines x %d ticks of the same workload (%.1f s); single engine thread: %d ticks (%.1f s)"
                  % (cores, tn, dtn, t1, dt1),
        "single_thread_value": single,
        "session_value": sess_rate,
        "session_sample": "%d independent sessions x %d ticks of the session variant (%.1f s): StreamInput unpack x2, audio graph, "
                          "blank + crossfade, scaler to 560x350 (this repo's two-pass 4-tap definition standing in for swscale), PCM pack"
                          % (cores, sess_n, sess_dt),
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
    sess_rate, sess_n, sess_dt = cpu_session_rate(po, desc, cores, 10.0)
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
        "e2e_session": {"value": sess_rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "sample": "%d independent sessions x %d ticks (%.1f s)" % (cores, sess_n, sess_dt),
                        "work_per_tick": "StreamInput i16 unpack x2, Engine::run_tick over the 32-module audio graph, VideoMixer blank + crossfade "
                                         "of its stored 1080p layers (a new picture every second tick), Monitor scaler to 560x350 + PCM pack",
                        "note": "the scaler is this repo's two-pass 4-tap bicubic definition in C (stands in for swscale, which the "
                                "reference calls; not available here)"},
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


MIN_TIMED_S = 0.5          # a K-step block is repeated until this much device time has been timed
MAX_BLOCKS = 400


def timed_blocks(ctx, dist, run_step, K, min_s=MIN_TIMED_S):
    """Times blocks of EXACTLY K steps with CUDA events on the context's stream (barrier + synchronize on both sides
    of every block) until >= min_s of device time has been seen, and returns (median block ms over ranks' max,
    number of blocks, all block ms of this rank).  One block is what the contract asks for; the repeats only make
    the figure robust against a scheduling hiccup in a 4 ms region."""
    def one_block():
        ctx.synchronize()
        dist.barrier()
        ctx.timer_begin()
        for _ in range(K):
            run_step()
        ctx.timer_end()
        ms = ctx.timer_elapsed_ms()
        ctx.synchronize()
        dist.barrier()
        return ms
    first = dist.max(one_block())
    n = int(min(MAX_BLOCKS, max(1, np.ceil(min_s * 1e3 / max(first, 1e-3)))))
    blocks = [one_block() for _ in range(n)]
    return dist.max(statistics.median(blocks)), n, blocks


class NoDist:
    world, rank = 1, 0

    def barrier(self):
        pass

    def max(self, x):
        return float(x)


def sub_benchmarks(mxl, ctx, args, peak):
    """Other BASELINE configurations measured in the same run with the same event timing (N = 1 only): the audio graph
    alone (config 2) at 128 and 1024 ticks per call, fused and staged; the compositor alone (config 3 i); the RGBA
    compositor (config 3 ii); config 4's session (16 audio modules + composite); a live one-tick call; config 1's CPU
    plumbing on the oracle walker."""
    from mixlab_b200 import workloads as W
    from mixlab_b200.session import AVSession
    nd = NoDist()
    K = max(5, min(args.steps, 50))
    sub = {}

    def run_session(desc, T, video, fusion=True, min_s=0.25):
        sess = AVSession(ctx, desc, T, video=video)
        sess.graph.set_fusion(fusion)
        sess.upload_inputs()
        tick = [0]

        def step():
            sess.run_step(tick[0])
            tick[0] += T
        for _ in range(3):
            step()
        l0 = ctx.launch_count
        step()
        launches = ctx.launch_count - l0
        ms, nb, _ = timed_blocks(ctx, nd, step, K, min_s)
        out = {"ticks_per_s": K * T / (ms * 1e-3), "ms_per_step": ms / K, "ticks_per_step": T, "blocks": nb,
               "launches_per_step": int(launches), "algorithmic_bytes_per_step": sess.algorithmic_bytes_per_step}
        out["gbs"] = sess.algorithmic_bytes_per_step / (ms / K * 1e-3) / 1e9
        out["frac_of_hbm_peak"] = out["gbs"] / peak
        sess.close()
        return out

    for T in (128, 1024):
        for fusion in (True, False):
            r = run_session(W.config2_graph(), T, False, fusion)
            r["stereo_frames_per_s"] = r["ticks_per_s"] * SPT
            # what MUST cross HBM for this graph: master + cue written (16*S B/tick) + one 32-byte meter record per tick
            r["compulsory_bytes_per_step"] = T * (16 * SPT + 32)
            r["compulsory_gbs"] = r["compulsory_bytes_per_step"] / (r["ms_per_step"] * 1e-3) / 1e9
            sub["audio_T%d%s" % (T, "" if fusion else "_staged")] = r
    sub["video_T128"] = run_session(None, 128, True)
    sub["video_T128"]["fps"] = sub["video_T128"]["ticks_per_s"]
    r = run_session(W.config4_audio_graph(), 128, True)
    r["note"] = "BASELINE config 4, one of its 8 sessions: 16 audio modules (236*S B/tick) + 1080p 2-layer composite"
    sub["config4_session_T128"] = r
    live = run_session(W.config2_graph(), 1, True, min_s=0.1)
    sub["live_one_tick_av"] = {"us_per_tick": live["ms_per_step"] * 1e3, "ticks_per_s": live["ticks_per_s"],
                               "launches_per_tick": live["launches_per_step"]}
    live = run_session(W.config2_graph(), 1, False, min_s=0.1)
    sub["live_one_tick_audio"] = {"us_per_tick": live["ms_per_step"] * 1e3, "ticks_per_s": live["ticks_per_s"],
                                  "launches_per_tick": live["launches_per_step"]}

    # config 3 (ii): crossfade + yuv420p -> RGBA8 in one pass, 64 frames per launch
    n = 64
    fa = ctx.frames_batch(W.FRAME_W, W.FRAME_H, n)
    fb = ctx.frames_batch(W.FRAME_W, W.FRAME_H, n)
    pics = ctx.rgba(W.FRAME_W, W.FRAME_H, n)
    for k in range(n):
        fa[k].upload_raw(W.random_bytes(0xA11CE + k % 3, W.FRAME_BYTES))
        fb[k].upload_raw(W.random_bytes(0xB0B + k % 3, W.FRAME_BYTES))

    def compose():
        ctx.compose_rgba(fa, fb, 0.5, pics)
    for _ in range(3):
        compose()
    ms, nb, _ = timed_blocks(ctx, nd, compose, K, 0.25)
    per_frame = 2 * W.FRAME_BYTES + W.FRAME_W * W.FRAME_H * 4
    sub["compose_rgba_64"] = {"fps": K * n / (ms * 1e-3), "ms_per_launch": ms / K, "frames_per_launch": n,
                              "algorithmic_bytes_per_frame": per_frame, "gbs": per_frame * n / (ms / K * 1e-3) / 1e9,
                              "frac_of_hbm_peak": per_frame * n / (ms / K * 1e-3) / 1e9 / peak, "blocks": nb,
                              "note": "self-specified BT.601 conversion on top of the pinned blend (the reference never converts colour)"}
    pics.free()
    for f in fa + fb:
        f.release()

    # config 5: the mixer bus alone, C channels x 800-sample ticks, 128 ticks per launch (bandwidth regime) and one tick
    # per launch (latency regime); algorithmic bytes 8*S*(C+2) per tick.  The full C x S sweep is tools/roofline_sweep.py.
    points = []
    for C in (2, 16, 64, 256):
        rng = np.random.default_rng(C)
        mod = ctx.module(mxl.MOD_MIXER, list(zip(rng.uniform(-24.0, 6.0, C), rng.uniform(0.1, 1.0, C), (np.arange(C) % 2 == 0))))
        for T in (128, 1):
            frames = SPT * T
            nbytes = 8 * frames * (C + 2)
            n_sets = int(max(1, min(16, np.ceil(300e6 / nbytes))))           # rotate over more than L2 holds
            host = np.resize(W.uniform_pm1(C * 131 + T, 2 * min(frames, 1 << 16)), 2 * frames)
            sets = [([ctx.stereo(host) for _ in range(C)], ctx.line(mxl.LINE_STEREO, frames), ctx.line(mxl.LINE_STEREO, frames))
                    for _ in range(n_sets)]
            it = [0]

            def one():
                ins, m, c = sets[it[0] % n_sets]
                it[0] += 1
                mod.run_tick(0, ins, [m, c])
            for _ in range(3):
                one()
            ms, nb, _ = timed_blocks(ctx, nd, one, K, 0.1)
            points.append({"channels": C, "ticks_per_launch": T, "ms_per_launch": ms / K, "gbs": nbytes / (ms / K * 1e-3) / 1e9,
                           "frac_of_hbm_peak": nbytes / (ms / K * 1e-3) / 1e9 / peak})
            for ins, m, c in sets:
                for ln in ins:
                    ln.free()
                m.free(); c.free()
        mod.destroy()
    sub["config5_mixer_bus"] = {"samples_per_tick": SPT, "points": points,
                                "note": "BASELINE config 5 at S = 800; the whole C x S sweep: tools/roofline_sweep.py -> profiles/r1_mixer_sweep.jsonl"}
    return sub


def config1_cpu_plumbing(seconds=1.5):
    """BASELINE config 1: 4-channel Mixer + Amplifier on the reference CPU engine's tick (per-tick topsort + alloc/zero +
    serial dispatch), oracle walker, ONE core (the reference's engine thread, src/engine.rs:78)."""
    from oracle import pyoracle as po
    from mixlab_b200 import workloads as W
    po.build()
    d = W.config1_graph()
    g, ids = po.build_graph(d, SAMPLE_RATE, SPT)
    n_src = 64
    for mid, (kind, seed) in d.sources.items():
        if kind == "stereo":
            g.set_source(ids[mid], W.uniform_pm1(seed, 2 * SPT * n_src), 2)
        else:
            g.set_source(ids[mid], W.uniform_01(seed, SPT * n_src), 1)
    for k in range(16):
        g.run_tick(k % n_src)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for k in range(256):
            g.run_tick(k % n_src)
        n += 256
    dt = time.perf_counter() - t0
    return {"ticks_per_s": n / dt, "cores": 1, "kind": "port", "ticks": n, "bytes_per_tick": 68 * SPT,
            "note": "oracle walker (C restatement of Engine::run_tick), no GPU involved"}


def shared_source_check(mxl, ctx, dist):
    """N > 1, after the timed region: the optional shared-source mode (north_star; extends the reference's one receiver
    per mountpoint, src/source.rs:93-95).  Rank 0 broadcasts one stereo line of a tick and one 1080p frame to every
    session's GPU with NCCL over NVLink (mxl_line_broadcast / mxl_frame_broadcast); every rank checks bit-equality."""
    from mixlab_b200 import workloads as W
    uid = [mxl.comm_unique_id() if dist.rank == 0 else None]
    dist.dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], dist.rank, dist.world)
    want_line = W.uniform_pm1(4242, 2 * SPT)
    want_frame = W.random_bytes(99, W.FRAME_BYTES)
    line = ctx.stereo(want_line if dist.rank == 0 else np.full(2 * SPT, 7.0, np.float32))
    frame = ctx.frame(W.FRAME_W, W.FRAME_H, data=want_frame if dist.rank == 0 else np.zeros(W.FRAME_BYTES, np.uint8))
    line.broadcast(0)
    frame.broadcast(0)
    ok = bool(np.array_equal(line.download().view(np.uint32), want_line.view(np.uint32))) and \
        bool(np.array_equal(frame.download_raw(), want_frame))
    for _ in range(5):                                   # absorbs the skew between the processes
        frame.broadcast(0)
        line.broadcast(0)
    ctx.synchronize()
    dist.barrier()
    ctx.timer_begin()
    for _ in range(50):
        frame.broadcast(0)
    ctx.timer_end()
    frame_ms = dist.max(ctx.timer_elapsed_ms()) / 50
    ctx.synchronize()
    dist.barrier()
    ctx.timer_begin()
    for _ in range(200):
        line.broadcast(0)
    ctx.timer_end()
    line_us = dist.max(ctx.timer_elapsed_ms()) / 200 * 1e3
    all_ok = dist.sum(1.0 if ok else 0.0) == dist.world
    ctx.synchronize()
    dist.barrier()
    ctx.comm_destroy()
    line.free()
    frame.release()
    return {"ok": bool(all_ok), "ranks": dist.world, "frame_bytes": W.FRAME_BYTES, "frame_ms": frame_ms,
            "frame_gbs": W.FRAME_BYTES / (frame_ms * 1e-3) / 1e9, "line_bytes": 8 * SPT, "line_us": line_us,
            "note": "ncclBroadcast root 0 on each context's stream, device-timed, max over ranks; 50 frames / 200 lines back to back"}


def e2e_session_leg(mxl, ctx, dist, args, T, Ke):
    """The same ticks as a LIVE SESSION moves them (next to `e2e`, which is the worst case: a new full-size picture per
    layer per tick up, every full-size composite down).  The reference's sources deliver pictures at their own frame rate
    (30 fps into 60 Hz ticks; VideoMixer re-uses its stored frames, video_mixer.rs:92-143), audio arrives as i16
    (stream_input.rs:110-112), and its sinks take the composite at the monitor's size (560 x 350, monitor.rs:21-22) plus
    1024-sample PCM fragments: StreamInput x2 -> VideoMixer -> Monitor beside the 32-module audio graph, through the C
    ABI with pinned host buffers, pipelined (uploads of step k+1 overlap the downloads of step k)."""
    from mixlab_b200 import workloads as W
    from mixlab_b200.session import StreamSession, session_seed
    ss = StreamSession(ctx, W.config2_graph(), T, seed=session_seed(0x5E55, dist.rank))

    def steps(n, tick):
        for i in range(n):
            ss.enqueue_step(tick, i & 1)
            tick += T
            if i > 0:
                ss.finish_step((i - 1) & 1)
        pics, frags = ss.finish_step((n - 1) & 1)
        ctx.synchronize()
        return tick, pics, frags

    tick, _, _ = steps(3, 0)
    dist.barrier()
    h2d0, d2h0, l0 = ctx.h2d_bytes, ctx.d2h_bytes, ctx.launch_count
    t0 = time.perf_counter()
    tick, pics, frags = steps(Ke, tick)
    dt = time.perf_counter() - t0
    h2d_step, d2h_step = (ctx.h2d_bytes - h2d0) // Ke, (ctx.d2h_bytes - d2h0) // Ke
    launches = (ctx.launch_count - l0) // Ke
    dist.barrier()
    dt_max = dist.max(dt)
    out = {"value": Ke * T * dist.world / dt_max, "unit": UNIT, "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
           "d2h_bytes_per_tick": d2h_step / T, "h2d_bytes_per_tick": h2d_step / T, "steps": Ke, "ms_per_step": dt_max / Ke * 1e3,
           "h2d_gbs": h2d_step * Ke / dt_max / 1e9, "d2h_gbs": d2h_step * Ke / dt_max / 1e9, "launches_per_step": int(launches),
           "pictures_per_step": pics, "pcm_fragments_per_step": frags, "source_fps": ss.fps, "monitor": "560x350",
           "graph": "2 x StreamInput (1080p pictures at 30 fps + i16 audio) -> VideoMixer -> Monitor, + cfg2 32-module audio graph -> Monitor.Audio",
           "timing": "host wall clock around K pipelined steps incl. pinned-host copies, synchronised both sides, max over ranks"}
    ss.close()
    return out


def b200_arm(args):
    import mixlab_b200 as mxl
    from mixlab_b200 import workloads as W
    from mixlab_b200.session import AVSession, session_seed

    dist = Dist(args.gpus)
    if dist.world != args.gpus and dist.world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, dist.world))
    T, K, Wm = args.ticks_per_step, args.steps, max(args.warmup, 3)
    ctx = mxl.Context(device=dist.local_rank, sample_rate=SAMPLE_RATE, samples_per_tick=SPT)
    # host side of the host-fed legs: this rank's thread and the pinned staging buffers it allocates from here on move
    # next to its GPU (a no-op on a single-node host)
    numa = (-1, 0) if os.environ.get("MXL_NO_NUMA_BIND") else ctx.bind_host_to_gpu_node()
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

    def value_step():
        nonlocal tick
        sess.run_step(tick)
        tick += T
    launches0 = ctx.launch_count
    host_t0 = time.perf_counter()
    for _ in range(K):
        value_step()
    host_enqueue_s = time.perf_counter() - host_t0
    launches = ctx.launch_count - launches0            # kernels of K steps
    ms_max, n_blocks, block_ms = timed_blocks(ctx, dist, value_step, K)
    total_ticks = K * T * dist.world
    value = total_ticks / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel: per-stage CUDA events over the same K steps.  Two passes:
    # the kernel's own launch duration with the stages serialised on one stream ("ms"), and its
    # duration in the configuration of the timed region, where the audio stages run beside the
    # compositor on a second stream ("ms_overlapped") ----
    kind_names = {v: k for k, v in W.STAGE_KIND.items()}

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
                    "StereoPanner": "panner_kernel", "Mixer": "mixer_kernel", "Meter": "meter_kernel",
                    "FusedVoiceMix": "fused_voice_mix_kernel"}
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
    l2_note = "inputs larger than L2: %.0f MB read + %.0f MB written per step vs 126 MB L2" % (
        sess.h2d_bytes_per_step / 1e6 if sess.video else 0, (sess.T * sess.frame_bytes if sess.video else 0) / 1e6)

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
    e2e_session = None
    if not args.no_e2e and args.workload == "av":
        e2e_session = e2e_session_leg(mxl, ctx, dist, args, T, args.e2e_steps or K)
    clocks = sampler.stop()

    shared = shared_source_check(mxl, ctx, dist) if dist.world > 1 else None
    sub = None
    if dist.world == 1 and not args.no_sub:
        sess.close()
        sess = None
        sub = sub_benchmarks(mxl, ctx, args, peak)
        if not args.no_cpu_baseline:
            sub["config1_cpu_plumbing"] = config1_cpu_plumbing()

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
                       "host_numa": {"node_of_rank0_gpu": numa[0], "cpus_bound": numa[1]},
                       "l2": l2_note},
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                       "samples": clocks["samples"]},
            "e2e": e2e, "e2e_session": e2e_session, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "stereo_frames_per_s": value * SPT if args.workload != "video" else None,
            "video_fps": value if args.workload != "audio" else None,
            "whole_step_gbs": whole, "whole_step_frac": whole / peak,
            "host_enqueue_ms_per_step": host_enqueue_s / K * 1e3, "stages": stages, "kernels": kernels,
            "timed_blocks": {"blocks": n_blocks, "steps_per_block": K, "median_block_ms": ms_max,
                             "min_block_ms": min(block_ms), "max_block_ms": max(block_ms),
                             "note": "value = K*T / median block; a block = exactly K steps between CUDA events, repeated until >= %.1f s" % MIN_TIMED_S},
            "shared_source": shared, "sub": sub,
        }
        print(json.dumps(line))
    if sess is not None:
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

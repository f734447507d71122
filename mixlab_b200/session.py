"""One live A/V session = one reference `Engine` (src/engine.rs:229-234): a module graph that is run
tick after tick.  This is the host loop a Mixlab engine thread would run against the C ABI:

  * device-resident mode  -- inputs already in HBM, `run_step` = mxl_graph_run_ticks over a batch of
    ticks (what bench.py reports as `value`);
  * host-fed mode         -- `run_step_host` uploads that step's video layers from pinned host memory
    (the frames StreamInput / MediaSource hand over, src/module/stream_input.rs:126-147), runs the
    ticks and downloads what the reference's sinks consume: the composited frames (Monitor /
    StreamOutput, monitor.rs:124-128), the master bus and the meter records (bench.py's `e2e`).

Sessions share nothing: N GPUs run N sessions, one per rank, with no data-path
collective.
"""
import ctypes as C

import numpy as np

from . import api
from . import workloads as W


def shard_sessions(n_sessions, world_size):
    """session i -> rank i mod world_size.  Returns per-rank lists of session ids."""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    return [[s for s in range(n_sessions) if s % world_size == r] for r in range(world_size)]


def session_seed(base_seed, session_id):
    """distinct synthetic content per session (base + rank)."""
    return (base_seed + 0x1000 * session_id) & W.MASK64


class AVSession:
    """Audio graph (GraphDesc) + optional VideoMixer compositing two host-/device-fed 1080p layers."""

    def __init__(self, ctx, audio_desc, ticks_per_step, video=True, width=W.FRAME_W, height=W.FRAME_H,
                 fader=0.5, seed=0xA11CE, unique_frames=4):
        self.ctx = ctx
        self.T = int(ticks_per_step)
        self.video = video
        self.audio_desc = audio_desc
        self.graph, self.ids = W.build_graph(ctx, audio_desc) if audio_desc is not None else (ctx.graph(), [])
        self.master = audio_desc.taps.get("master") if audio_desc is not None else None
        self.meter = audio_desc.taps.get("meter") if audio_desc is not None else None
        self.width, self.height = width, height
        self.fader = fader
        self.layout = api.frame_layout(width, height)
        self.frame_bytes = int(self.layout.size)
        self.audio_bytes_per_tick = W.algorithmic_bytes_per_tick(audio_desc, ctx.spt) if audio_desc is not None else 0
        self.video_bytes_per_tick = 3 * self.frame_bytes if video else 0
        self._pinned = []
        if video:
            g = self.graph
            self.src_a = g.add(api.MOD_SOURCE_VIDEO)
            self.src_b = g.add(api.MOD_SOURCE_VIDEO)
            self.vmix = g.add(api.MOD_VIDEO_MIXER, (0, 1, fader))
            g.connect(self.vmix, 0, self.src_a, 0)
            g.connect(self.vmix, 1, self.src_b, 0)
            # T device frames per layer, presented as the sources' video lines (one slot per tick)
            # the T frames of a layer are adjacent on the device: a step's layer goes up as one copy
            self.frames_a = ctx.frames_batch(width, height, self.T)
            self.frames_b = ctx.frames_batch(width, height, self.T)
            self._fa = (C.c_void_p * self.T)(*[f.h for f in self.frames_a])
            self._fb = (C.c_void_p * self.T)(*[f.h for f in self.frames_b])
            self.line_a, self.line_b = ctx.video_line(self.T), ctx.video_line(self.T)
            rate = (ctx.spt, ctx.sample_rate)
            for k in range(self.T):
                self.line_a.set(k, self.frames_a[k], duration=rate)
                self.line_b.set(k, self.frames_b[k], duration=rate)
            g.module(self.src_a).set_source_line(self.line_a)
            g.module(self.src_b).set_source_line(self.line_b)
            # pinned host staging: T slots per layer, filled from a few unique synthetic frames
            self.host_a = self._pin(self.T * self.frame_bytes)
            self.host_b = self._pin(self.T * self.frame_bytes)
            self.host_out = self._pin(self.T * self.frame_bytes)
            self.host_out2 = None          # second result buffer, allocated by enable_pipelining()
            ua = [W.random_bytes(seed + 17 * i, self.frame_bytes) for i in range(unique_frames)]
            ub = [W.random_bytes(seed + 0xB0B + 31 * i, self.frame_bytes) for i in range(unique_frames)]
            va = self.host_a.array.reshape(self.T, self.frame_bytes)
            vb = self.host_b.array.reshape(self.T, self.frame_bytes)
            for k in range(self.T):
                va[k] = ua[k % unique_frames]
                vb[k] = ub[k % unique_frames]
        if self.master is not None:
            self.host_master = self._pin(self.T * ctx.spt * 2 * 4, np.float32)
        self.meter_records = np.zeros(self.T, api.METER_RECORD)
        self._L = api.lib()
        self._pipelined = False

    def _pin(self, nbytes, dtype=np.uint8):
        p = api.PinnedBuffer(nbytes, dtype)
        self._pinned.append(p)
        return p

    # ---- bytes -------------------------------------------------------------------------------
    @property
    def algorithmic_bytes_per_step(self):
        return self.T * (self.audio_bytes_per_tick + self.video_bytes_per_tick)

    @property
    def h2d_bytes_per_step(self):
        return 2 * self.T * self.frame_bytes if self.video else 0

    @property
    def d2h_bytes_per_step(self):
        n = self.T * self.frame_bytes if self.video else 0
        if self.master is not None:
            n += self.T * self.ctx.spt * 8
        if self.meter is not None:
            n += self.T * api.METER_RECORD.itemsize
        return n

    # ---- steps -------------------------------------------------------------------------------
    def upload_inputs(self):
        """All layers of one step from pinned host memory to the device frames (asynchronous)."""
        if not self.video:
            return
        L, fb = self._L, self.frame_bytes
        api.check(L.mxl_frames_upload_raw_async(self._fa, self.T, self.host_a.ptr, fb))
        api.check(L.mxl_frames_upload_raw_async(self._fb, self.T, self.host_b.ptr, fb))

    def run_step(self, tick0):
        """Device-resident step: T ticks of the whole graph, asynchronous on the context stream."""
        self.graph.run_ticks(tick0, self.T)

    def download_outputs(self):
        """What the reference's sinks read each tick, to pinned host memory; synchronises."""
        L, fb = self._L, self.frame_bytes
        if self.video:
            api.check(L.mxl_frames_download_raw_async(self._out_frames(), self.T, self.host_out.ptr, fb))
        if self.master is not None:
            line = L.mxl_graph_output(self.graph.h, self.ids[self.master[0]], self.master[1])
            api.check(L.mxl_line_download_async(line, self.host_master.ptr, self.T * self.ctx.spt * 2))
        if self.meter is not None:
            self.graph.module(self.ids[self.meter[0]]).meter_download(self.T, self.meter_records)
        self.ctx.synchronize()

    def run_step_host(self, tick0):
        """Host-fed step: upload -> T ticks -> download, through the C ABI with host buffers."""
        self.upload_inputs()
        self.run_step(tick0)
        self.download_outputs()

    # ---- pipelined host-fed steps (copy/compute overlap, double-buffered results) -------------
    def enable_pipelining(self):
        """Uploads, ticks and downloads of consecutive steps overlap (mxl_ctx_set_copy_overlap); results
        alternate between two pinned buffers so the consumer of step k reads while step k+1 runs."""
        if self._pipelined:
            return
        self.ctx.set_copy_overlap(True)
        self._out = [self.host_out if self.video else None, self._pin(self.T * self.frame_bytes) if self.video else None]
        if self.master is not None:
            self._master = [self.host_master, self._pin(self.T * self.ctx.spt * 2 * 4, np.float32)]
        self._meter = [self._pin(self.T * api.METER_RECORD.itemsize), self._pin(self.T * api.METER_RECORD.itemsize)]
        self._pipelined = True

    def enqueue_step_host(self, tick0, slot):
        """Host-fed step without host synchronisation: upload -> T ticks -> download into result buffer
        `slot` (0/1), then a download fence on `slot`.  wait_step(slot) blocks until its results landed."""
        L, fb = self._L, self.frame_bytes
        self.upload_inputs()
        self.run_step(tick0)
        if self.video:
            api.check(L.mxl_frames_download_raw_async(self._out_frames(), self.T, self._out[slot].ptr, fb))
        if self.master is not None:
            line = L.mxl_graph_output(self.graph.h, self.ids[self.master[0]], self.master[1])
            api.check(L.mxl_line_download_async(line, self._master[slot].ptr, self.T * self.ctx.spt * 2))
        if self.meter is not None:
            api.check(L.mxl_meter_download_async(self.graph.module(self.ids[self.meter[0]]).h, self._meter[slot].ptr, self.T))
        self.ctx.download_fence(slot)

    def _out_frames(self):
        """Handles of this step's composited frames (the VideoMixer output line's slots)."""
        L = self._L
        out = L.mxl_graph_output(self.graph.h, self.vmix, 0)
        arr = (C.c_void_p * self.T)()
        for k in range(self.T):
            arr[k] = L.mxl_video_line_get(out, k)
        return arr

    def wait_step(self, slot):
        self.ctx.wait_fence(slot)

    def result_frames(self, slot):
        return self._out[slot].array.reshape(self.T, self.frame_bytes)

    def result_master(self, slot):
        return self._master[slot].array

    def result_meter(self, slot):
        return self._meter[slot].array.view(api.METER_RECORD)

    def output_frame(self, k):
        out = self.graph.output(self.vmix, 0)
        return out.get(k)

    def close(self):
        self.ctx.synchronize()
        self.graph.destroy()
        if self.video:
            for f in self.frames_a + self.frames_b:
                f.release()
            self.line_a.free()
            self.line_b.free()
        for p in self._pinned:
            p.free()
        self._pinned = []


class StreamSession:
    """One live session fed and drained the way the reference's is (bench.py's `e2e_session`):

        StreamInput A --video--> VideoMixer(a=A, b=B, fader) --> Monitor.Video      (stream_input.rs, video_mixer.rs)
        StreamInput B --video-->                                 Monitor.Audio <-- audio graph's master bus
        (both StreamInputs also deliver their i16 audio, converted on the device)

    The receivers' side -- what RTMP / Icecast ingest pushes with SourceSend::write_audio / write_video
    (src/source.rs:156-190) -- is this class: decoded pictures arrive at the SOURCE frame rate (30 fps into 60 Hz
    ticks: VideoMixer re-uses its stored frames in between, video_mixer.rs:92-143) from pinned host memory, audio as
    interleaved i16 in AAC-sized frames of 1024 per channel.  The consumers' side is what the Monitor's codec thread
    takes (monitor.rs:235-247): monitor-sized pictures (560 x 350) and 1024-sample PCM fragments, downloaded every step.
    """

    AUDIO_FRAME = 1024                    # samples per channel per pushed frame (rtmp/mod.rs:213-245: AAC frames)

    def __init__(self, ctx, audio_desc, ticks_per_step, width=W.FRAME_W, height=W.FRAME_H, source_fps=30, fader=0.5,
                 seed=0xA11CE, unique_frames=4, monitor=(560, 350), ring=3):
        self.ctx, self.T = ctx, int(ticks_per_step)
        L = self._L = api.lib()
        tick_rate = ctx.sample_rate // ctx.spt
        assert tick_rate % source_fps == 0 and self.T % (tick_rate // source_fps) == 0
        self.tpf = tick_rate // source_fps                       # ticks per source frame
        self.fps = source_fps
        self.n_frames = self.T // self.tpf                       # pictures per source per step
        self.audio_desc = audio_desc
        self.graph, self.ids = W.build_graph(ctx, audio_desc)
        g = self.graph
        self.master = audio_desc.taps["master"]
        self.meter = audio_desc.taps.get("meter")
        self.in_a, self.in_b = g.add(api.MOD_STREAM_INPUT), g.add(api.MOD_STREAM_INPUT)
        self.vmix = g.add(api.MOD_VIDEO_MIXER, (0, 1, fader))
        self.mon = g.add(api.MOD_MONITOR, (monitor[0], monitor[1], 0))
        g.connect(self.vmix, 0, self.in_a, 0)
        g.connect(self.vmix, 1, self.in_b, 0)
        g.connect(self.mon, 0, self.vmix, 0)
        g.connect(self.mon, 1, self.ids[self.master[0]], self.master[1])
        self.fader = fader
        self.layout = api.frame_layout(width, height)
        self.frame_bytes = int(self.layout.size)
        self.mon_layout = api.frame_layout(*monitor)
        self.mon_bytes = int(self.mon_layout.size)
        self._pinned = []
        # pictures: `ring` slabs of n_frames adjacent device frames per source, filled in turn (a slab is rewritten
        # `ring` steps after the graph last looked at it)
        self.slabs = [[ctx.frames_batch(width, height, self.n_frames) for _ in range(ring)] for _ in range(2)]
        self._slab_arr = [[(C.c_void_p * self.n_frames)(*[f.h for f in slab]) for slab in per] for per in self.slabs]
        self.host_pix = [self._pin(self.n_frames * self.frame_bytes), self._pin(self.n_frames * self.frame_bytes)]
        for s, base in ((0, seed), (1, seed + 0xB0B)):
            uniq = [W.random_bytes(base + 17 * i, self.frame_bytes) for i in range(unique_frames)]
            view = self.host_pix[s].array.reshape(self.n_frames, self.frame_bytes)
            for k in range(self.n_frames):
                view[k] = uniq[k % unique_frames]
        # audio: one step of interleaved i16 per source
        n_pcm = self.T * ctx.spt * 2
        self.host_pcm = []
        for s in range(2):
            r = (W.splitmix64(seed + 0x51 * (s + 1), n_pcm) & np.uint64(0xFFFF)).astype(np.uint16).view(np.int16)
            self.host_pcm.append(np.ascontiguousarray(r))
        self.audio_pushed = [0, 0]                               # samples per channel pushed so far, per source
        self.frames_pushed = 0
        self.steps = 0
        # results, double-buffered: monitor pictures, PCM fragments, meter records
        self.max_jobs = self.T + 8
        self.max_frags = self.T * ctx.spt * 2 // 2048 + 2
        self.out_pix = [self._pin(self.max_jobs * self.mon_bytes), self._pin(self.max_jobs * self.mon_bytes)]
        self.out_pcm = [np.empty((self.max_frags, 2048), np.int16), np.empty((self.max_frags, 2048), np.int16)]
        self.out_meter = [self._pin(self.T * api.METER_RECORD.itemsize), self._pin(self.T * api.METER_RECORD.itemsize)]
        self.jobs = [[], []]                                     # per slot: [(pts, duration, blank, frame handle)]
        self.frags = [[], []]                                    # per slot: [(decode timestamp, duration)]
        self.frag_due = [0, 0]                                   # per slot: fragments complete once that step has run
        self.frag_popped = 0
        ctx.set_copy_overlap(True)

    def _pin(self, nbytes, dtype=np.uint8):
        p = api.PinnedBuffer(nbytes, dtype)
        self._pinned.append(p)
        return p

    @property
    def h2d_bytes_per_step(self):
        return 2 * (self.n_frames * self.frame_bytes + self.T * self.ctx.spt * 2 * 2)

    def feed(self):
        """The receivers' pushes for the next step: n_frames pictures per source (one copy per source from pinned
        memory) and the step's audio in frames of 1024 samples per channel."""
        L, ctx = self._L, self.ctx
        slab = self.steps % len(self.slabs[0])
        for s, mod in ((0, self.in_a), (1, self.in_b)):
            api.check(L.mxl_frames_upload_raw_async(self._slab_arr[s][slab], self.n_frames, self.host_pix[s].ptr, self.frame_bytes))
            h = self.graph.module(mod).h
            for k in range(self.n_frames):
                api.check(L.mxl_stream_input_write_video(h, 1, self.frames_pushed + k, self.fps, self.slabs[s][slab][k].h, 1, self.fps))
            pcm, n = self.host_pcm[s], self.T * ctx.spt
            base = pcm.ctypes.data
            for off in range(0, n, self.AUDIO_FRAME):
                cnt = min(self.AUDIO_FRAME, n - off)
                api.check(L.mxl_stream_input_write_audio(h, 1, self.audio_pushed[s] + off, ctx.sample_rate, base + 4 * off, 2 * cnt))
            self.audio_pushed[s] += n
        self.frames_pushed += self.n_frames

    def enqueue_step(self, tick0, slot):
        """feed -> T ticks -> start the downloads of what the Monitor's codec thread would take, into result slot 0/1."""
        L = self._L
        self.feed()
        self.graph.run_ticks(tick0, self.T)
        self.steps += 1
        # AudioCtx::send_audio hands a fragment on when its buffer holds MORE than 2 * 1024 samples (encode.rs:197-221)
        self.frag_due[slot] = (self.steps * self.T * self.ctx.spt * 2 - 1) // 2048
        mon = self.graph.module(self.mon)
        jobs = []
        job = api.VideoJob()
        while len(jobs) < self.max_jobs and api.check(L.mxl_monitor_recv_video(mon.h, C.byref(job))) == 1:
            jobs.append((job.pts, job.duration, bool(job.blank), job.frame))
        self.jobs[slot] = jobs
        if jobs:
            arr = (C.c_void_p * len(jobs))(*[j[3] for j in jobs])
            api.check(L.mxl_frames_download_raw_async(arr, len(jobs), self.out_pix[slot].ptr, self.mon_bytes))
        if self.meter is not None:
            api.check(L.mxl_meter_download_async(self.graph.module(self.ids[self.meter[0]]).h, self.out_meter[slot].ptr, self.T))
        self.ctx.download_fence(slot)

    def finish_step(self, slot):
        """Blocks until result slot `slot` is complete; pops its PCM fragments; releases the pictures' device frames.
        Returns (number of pictures, number of fragments)."""
        L = self._L
        self.ctx.wait_fence(slot)
        for j in self.jobs[slot]:
            L.mxl_frame_release(j[3])
        mon = self.graph.module(self.mon)
        info = api.AudioFragment()
        frags, buf = [], self.out_pcm[slot]
        # only this step's fragments: the next step may already be in flight, and its PCM is still on its way
        while self.frag_popped < self.frag_due[slot] and len(frags) < self.max_frags and \
                api.check(L.mxl_monitor_recv_audio(mon.h, C.byref(info), buf[len(frags)].ctypes.data, 2048)) == 1:
            frags.append(((info.decode_num, info.decode_den), (info.duration_num, info.duration_den)))
            self.frag_popped += 1
        self.frags[slot] = frags
        return len(self.jobs[slot]), len(frags)

    def pictures(self, slot):
        return self.out_pix[slot].array[:len(self.jobs[slot]) * self.mon_bytes].reshape(len(self.jobs[slot]), self.mon_bytes)

    def close(self):
        self.ctx.synchronize()
        self.graph.destroy()
        for per in self.slabs:
            for slab in per:
                for f in slab:
                    f.release()
        for p in self._pinned:
            p.free()
        self._pinned = []
        self.ctx.set_copy_overlap(False)

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

"""Session driver: sharding logic (CPU, incl. a world_size-2 gloo run) and, on the GPU, the host-fed
step in serial and pipelined (copy/compute overlap) mode against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from mixlab_b200 import workloads as W
from mixlab_b200.session import session_seed, shard_sessions

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_sessions_round_robin():
    assert shard_sessions(8, 8) == [[i] for i in range(8)]
    assert shard_sessions(8, 2) == [[0, 2, 4, 6], [1, 3, 5, 7]]
    assert shard_sessions(3, 4) == [[0], [1], [2], []]
    assert shard_sessions(0, 2) == [[], []]
    with pytest.raises(ValueError):
        shard_sessions(4, 0)
    seeds = {session_seed(0xA11CE, s) for s in range(64)}
    assert len(seeds) == 64


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch
import torch.distributed as dist
from mixlab_b200.session import shard_sessions, session_seed
from mixlab_b200 import workloads as W
from oracle import pyoracle as po

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = shard_sessions(6, world)[rank]
# every rank runs its own sessions on the CPU oracle (stand-in for its GPU): no data-path exchange
desc = W.config4_audio_graph()
sums = []
for s in mine:
    g, ids = po.build_graph(desc, 48000, 800)
    out = g.run_tick(s, (ids[desc.taps["master"][0]], 0), 1600)
    sums.append(float(np.abs(out).sum()) + session_seed(1, s) % 7)
# the only collectives: a barrier and the reductions bench.py uses for its report
dist.barrier()
t = torch.tensor([float(len(mine)), 1.0 + rank], dtype=torch.float64)
tot = t.clone(); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
gathered = [None] * world
dist.all_gather_object(gathered, mine)
if rank == 0:
    flat = sorted(x for part in gathered for x in part)
    assert flat == list(range(6)), flat
    assert tot[0].item() == 6.0 and mx[1].item() == float(world)
    print("GLOO_OK", gathered)
dist.destroy_process_group()
"""


def test_two_rank_gloo_sharding(tmp_path):
    """world_size 2 on CPU: sessions shard round-robin, ranks exchange nothing but the report reductions."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), str(script), ROOT]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "GLOO_OK [[0, 2, 4], [1, 3, 5]]" in r.stdout


def _check_step(oracle, desc, sess, frames, master, meter, tick0, T, spt=800, sr=48000):
    lay = oracle.frame_layout(sess.width, sess.height)
    fb = sess.frame_bytes
    f = oracle.fader_to_u8(sess.fader)
    for k in range(T):
        a = sess.host_a.array[k * fb:(k + 1) * fb]
        b = sess.host_b.array[k * fb:(k + 1) * fb]
        assert np.array_equal(frames[k], oracle.video_crossfade(lay, a, b, f)), k
    og, oids = oracle.build_graph(desc, sr, spt)
    # oscillators are pure functions of t and the EQs are re-run from tick 0 to carry their state
    want = None
    for k in range(tick0 + T):
        buf = og.run_tick(k, (oids[desc.taps["master"][0]], 0), 2 * spt)
        if k >= tick0:
            want = buf if want is None else np.concatenate([want, buf])
    err = np.abs(master.astype(np.float64) - want.astype(np.float64))
    assert np.all(err <= 1e-6 * np.abs(want) + 1e-7)                     # SURVEY 8(d) gate
    if "meter" in desc.taps:
        pk, sq, clip = og.meter(oids[desc.taps["meter"][0]])
        assert abs(meter[T - 1]["peak"][0] - pk[0]) <= 1e-6 * pk[0] + 1e-7


@pytest.mark.gpu
def test_host_fed_steps_serial_and_pipelined(mxl, oracle):
    from mixlab_b200.session import AVSession
    T = 3
    desc = W.config2_graph()
    with mxl.Context(0, 48000, 800) as ctx:
        sess = AVSession(ctx, desc, T, video=True, width=560, height=350, unique_frames=3)
        sess.run_step_host(0)
        _check_step(oracle, desc, sess, sess.host_out.array.reshape(T, sess.frame_bytes), np.array(sess.host_master.array),
                    sess.meter_records, 0, T)
        h2d0, d2h0 = ctx.h2d_bytes, ctx.d2h_bytes
        sess.enable_pipelining()
        # 4 pipelined steps; inputs change between steps while earlier steps are still in flight
        for i in range(4):
            if i == 2:
                sess.wait_step(0)                 # results of step 0 must be complete and stable
                keep = np.array(sess.result_frames(0))
            sess.enqueue_step_host(T * (1 + i), i & 1)
            if i > 0:
                sess.wait_step((i - 1) & 1)
        sess.wait_step(1)
        ctx.synchronize()
        assert ctx.h2d_bytes - h2d0 == 4 * sess.h2d_bytes_per_step
        assert ctx.d2h_bytes - d2h0 == 4 * sess.d2h_bytes_per_step
        _check_step(oracle, desc, sess, sess.result_frames(1), np.array(sess.result_master(1)), sess.result_meter(1), T * 4, T)
        _check_step(oracle, desc, sess, sess.result_frames(0), np.array(sess.result_master(0)), sess.result_meter(0), T * 3, T)
        assert np.array_equal(keep, sess.result_frames(0))     # same inputs every step -> same composite
        sess.close()


@pytest.mark.gpu
@pytest.mark.parametrize("session_id", [0, 3, 7])
def test_config4_av_session_vs_oracle(mxl, oracle, session_id):
    """BASELINE config 4: one of the 8 independent A/V sessions (16 audio modules + 1080p 2-layer composite), with
    that session's own synthetic content (base seed + session id), host-fed through the C ABI for 9 ticks
    and then 8 more: audio against the oracle's engine walker, every composited frame bit-exact."""
    from mixlab_b200.session import AVSession, session_seed
    T = 9
    desc = W.config4_audio_graph()
    assert W.algorithmic_bytes_per_tick(desc, 800) == 236 * 800
    with mxl.Context(0, 48000, 800) as ctx:
        sess = AVSession(ctx, desc, T, video=True, unique_frames=3, seed=session_seed(0xA11CE, session_id))
        assert sess.frame_bytes == W.FRAME_BYTES
        if session_id:                                       # distinct sessions carry distinct pictures
            other = AVSession(ctx, None, 1, video=True, unique_frames=1, seed=session_seed(0xA11CE, 0))
            assert not np.array_equal(other.host_a.array[:4096], sess.host_a.array[:4096])
            other.close()
        sess.run_step_host(0)
        _check_step(oracle, desc, sess, sess.host_out.array.reshape(T, sess.frame_bytes), np.array(sess.host_master.array),
                    sess.meter_records, 0, T)
        sess.run_step_host(T)
        _check_step(oracle, desc, sess, sess.host_out.array.reshape(T, sess.frame_bytes), np.array(sess.host_master.array),
                    sess.meter_records, T, T)
        sess.close()


@pytest.mark.gpu
def test_overlap_orders_upload_after_compute(mxl, oracle, ctx48):
    """WAR hazard: an upload enqueued right after a run must not overwrite inputs the run still reads."""
    ctx48.set_copy_overlap(True)
    n = 1 << 22
    a = W.uniform_pm1(1, n)
    b = W.uniform_pm1(2, n)
    pa, pb = mxl.PinnedBuffer(n * 4, np.float32), mxl.PinnedBuffer(n * 4, np.float32)
    pa.array[:] = a
    pb.array[:] = b
    src = ctx48.line(mxl.LINE_STEREO, n // 2)
    out = ctx48.line(mxl.LINE_STEREO, n // 2)
    amp = ctx48.module(mxl.MOD_AMPLIFIER, (0.5, 0.0))
    res = mxl.PinnedBuffer(n * 4, np.float32)
    L = mxl.lib()
    for _ in range(5):
        mxl.check(L.mxl_line_upload_async(src.h, pa.ptr, n))
        amp.run_tick(0, [src, None], [out])
        mxl.check(L.mxl_line_upload_async(src.h, pb.ptr, n))       # must wait for the run above
        mxl.check(L.mxl_line_download_async(out.h, res.ptr, n))
        ctx48.synchronize()
        assert np.array_equal(res.array, oracle.amplifier(a, None, 0.5, 0.0))
        amp.run_tick(0, [src, None], [out])
        mxl.check(L.mxl_line_download_async(out.h, res.ptr, n))
        ctx48.synchronize()
        assert np.array_equal(res.array, oracle.amplifier(b, None, 0.5, 0.0))
    ctx48.set_copy_overlap(False)
    for p in (pa, pb, res):
        p.free()


@pytest.mark.gpu
def test_steady_state_holds_device_memory_constant(mxl, oracle):
    """A live session must not grow: frames recycle through the context's pool, job tables and staging rings are
    reused, tap tables are cached.  300 host-fed steps of an A/V session whose layer B is smaller than layer A
    (so the scaler, the blank fill and the crossfade all run every tick), then the free device memory is what it
    was after warm-up."""
    from mixlab_b200.session import AVSession
    with mxl.Context(0, 48000, 800) as ctx:
        sess = AVSession(ctx, W.config4_audio_graph(), 4, video=True, unique_frames=2)
        small = [ctx.frame(1280, 720, W.random_bytes(70 + k, oracle.frame_layout(1280, 720).size)) for k in range(4)]
        for k in range(4):
            sess.line_b.set(k, small[k], duration=(ctx.spt, ctx.sample_rate))
        for step in range(20):
            sess.run_step_host(step * 4)
        free0, _ = ctx.device_memory()
        for step in range(20, 320):
            sess.run_step_host(step * 4)
        free1, _ = ctx.device_memory()
        assert free1 >= free0 - (1 << 20), (free0, free1)
        sess.close()


@pytest.mark.gpu
def test_frame_pool_cap_and_trim(mxl):
    """Released frames park in a per-size pool for reuse; the pool is capped and can be trimmed (a session whose sources
    changed resolution must not keep the old size class until the context dies)."""
    with mxl.Context(0, 48000, 800) as ctx:
        ctx.synchronize()
        frames = [ctx.frame(1920, 1080, blank=True) for _ in range(8)]
        ctx.synchronize()
        free_held, _ = ctx.device_memory()
        for f in frames:
            f.release()
        free_parked, _ = ctx.device_memory()
        assert free_parked <= free_held + (1 << 20)            # parked, not freed
        again = ctx.frame(1920, 1080, blank=True)              # comes out of the pool: no new allocation
        assert ctx.device_memory()[0] <= free_parked + (1 << 20)
        again.release()
        freed = ctx.trim_frame_pool(0)
        assert freed == 8 * 3110400
        assert ctx.device_memory()[0] >= free_parked + 6 * 3110400
        # with a small cap a released frame goes straight back to the driver
        ctx.trim_frame_pool(0, new_cap_bytes=2 * 3110400)
        frames = [ctx.frame(1920, 1080, blank=True) for _ in range(6)]
        for f in frames:
            f.release()
        assert ctx.trim_frame_pool(0) == 2 * 3110400


class VideoMixerModel:
    """VideoMixer::run_tick (video_mixer.rs:70-250) for two channels of one picture size (no scaler involved): stored
    frames kept until `active_until`, one composite per tick from whatever is stored."""

    def __init__(self, oracle, lay, fader, sr):
        from fractions import Fraction
        self.o, self.lay, self.f8, self.sr, self.F = oracle, lay, oracle.fader_to_u8(fader), sr, Fraction
        self.stored = [None, None]                       # (pixels, active_until)

    def run_tick(self, t, inputs):
        now = self.F(int(t), self.sr)
        for c in range(2):
            if self.stored[c] is not None and now >= self.stored[c][1]:        # 94-101
                self.stored[c] = None
        if all(i is None for i in inputs) and all(s is None for s in self.stored):
            return None                                                         # 113-119
        for c, inp in enumerate(inputs):
            if inp is not None:
                pix, dur, off = inp
                self.stored[c] = (pix, now + off + dur)                         # 139-143
        a = self.stored[0][0] if self.stored[0] else None
        b = self.stored[1][0] if self.stored[1] else None
        return self.o.video_crossfade(self.lay, a, b, self.f8)


@pytest.mark.gpu
@pytest.mark.parametrize("T", [6, 16])
def test_stream_session_against_the_oracle_chain(mxl, oracle, T):
    """bench.py's session variant end to end: two StreamInputs (30 fps pictures from pinned memory, i16 audio) ->
    VideoMixer -> Monitor beside the audio graph, three pipelined steps.  Every monitor picture (timing, pixels) and
    every PCM fragment against the oracle's StreamInput, engine walker, VideoMixer model and MonitorFeed."""
    from fractions import Fraction
    from mixlab_b200.session import StreamSession
    sr, spt, w, h, mon = 48000, 800, 128, 72, (56, 34)
    desc = W.config4_audio_graph()
    steps = 3
    with mxl.Context(0, sr, spt) as ctx:
        sess = StreamSession(ctx, desc, T, width=w, height=h, monitor=mon, unique_frames=3, seed=0x5E55)
        h2d0, d2h0 = ctx.h2d_bytes, ctx.d2h_bytes
        got_jobs, got_pix, got_frags, got_pcm = [], [], [], []
        for i in range(steps):
            sess.enqueue_step(i * T, i & 1)
            if i > 0:
                n_pic, n_frag = sess.finish_step((i - 1) & 1)
                got_jobs += [j[:3] for j in sess.jobs[(i - 1) & 1]]
                got_pix += [np.array(p) for p in sess.pictures((i - 1) & 1)]
                got_frags += sess.frags[(i - 1) & 1]
                got_pcm += [np.array(sess.out_pcm[(i - 1) & 1][k]) for k in range(n_frag)]
        slot = (steps - 1) & 1
        n_pic, n_frag = sess.finish_step(slot)
        got_jobs += [j[:3] for j in sess.jobs[slot]]
        got_pix += [np.array(p) for p in sess.pictures(slot)]
        got_frags += sess.frags[slot]
        got_pcm += [np.array(sess.out_pcm[slot][k]) for k in range(n_frag)]
        ctx.synchronize()
        assert ctx.h2d_bytes - h2d0 == steps * sess.h2d_bytes_per_step
        host_pix = [np.array(sess.host_pix[s].array).reshape(sess.n_frames, sess.frame_bytes) for s in range(2)]
        host_pcm = [np.array(p) for p in sess.host_pcm]
        n_frames, tpf, fps = sess.n_frames, sess.tpf, sess.fps
        sess.close()

    # ---- the oracle chain, fed the same pushes at the same ticks ----
    lay = oracle.frame_layout(w, h)
    ins = [oracle.StreamInput(sr), oracle.StreamInput(sr)]
    og, oids = oracle.build_graph(desc, sr, spt)
    vm = VideoMixerModel(oracle, lay, 0.5, sr)
    feed = oracle.MonitorFeed(sr, mon[0], mon[1])
    for i in range(steps):
        for s in range(2):
            for k in range(n_frames):
                assert ins[s].write_video(1, Fraction(i * n_frames + k, fps), host_pix[s][k], Fraction(1, fps))
            for off in range(0, T * spt, 1024):
                cnt = min(1024, T * spt - off)
                assert ins[s].write_audio(1, Fraction(i * T * spt + off, sr), host_pcm[s][2 * off:2 * (off + cnt)])
        for k in range(T):
            tick = i * T + k
            vin = [ins[s].run_tick(tick * spt, 2 * spt)[0] for s in range(2)]
            master = og.run_tick(tick, (oids[desc.taps["master"][0]], 0), 2 * spt)
            comp = vm.run_tick(tick * spt, vin)
            feed.run_tick(tick * spt, master, None if comp is None else (comp, lay, Fraction(spt, sr), Fraction(0)))
    assert len(got_jobs) == len(feed.video_out) and len(got_jobs) >= steps * T - 1
    for i, ((pts, dur, blank), pix, (wpts, wdur, wblank, wpix)) in enumerate(zip(got_jobs, got_pix, feed.video_out)):
        assert (pts, dur, blank) == (wpts, wdur, wblank), i
        assert np.array_equal(pix, wpix), i
    assert len(got_frags) == len(feed.audio_out) == (steps * T * spt * 2 - 1) // 2048
    for i, ((dec, dur), frag, (wdec, wdur, wfrag)) in enumerate(zip(got_frags, got_pcm, feed.audio_out)):
        assert Fraction(*dec) == wdec and Fraction(*dur) == wdur, i
        # the master bus is sine-based: packed i16 may differ by one step where a device sine differs from glibc's
        assert np.max(np.abs(frag.astype(np.int32) - wfrag.astype(np.int32))) <= 1, i

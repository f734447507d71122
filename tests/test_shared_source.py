"""Optional shared-source mode (NEW, no reference counterpart): the ingest GPU broadcasts a
source line / frame to the sessions on the other GPUs with NCCL, each of which then runs its own graph.

CPU part: the entry points exist and fail cleanly without a device.  GPU part (needs >= 2 GPUs; run by
`gpurun --gpus 2 -- python -m pytest tests/test_shared_source.py -m gpu`): two processes, one per GPU."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_comm_entry_points_fail_cleanly_without_device():
    import mixlab_b200 as mxl
    ctx = mxl.Context(device=mxl.DEVICE_NONE, sample_rate=48000, samples_per_tick=800)
    with pytest.raises(mxl.MxlError):
        ctx.comm_init(bytes(128), 0, 2)
    ctx.comm_destroy()                                  # no communicator: a no-op
    ctx.close()


def _rank_main(rank, world, uid, q):
    try:
        import mixlab_b200 as mxl
        from mixlab_b200 import workloads as W
        from oracle import pyoracle as po
        po.build()
        ctx = mxl.Context(device=rank, sample_rate=48000, samples_per_tick=800)
        ctx.comm_init(uid, rank, world)
        spt, ticks = 800, 16
        # the shared source exists on the ingest rank only; the others start from garbage
        src = W.uniform_pm1(4242, 2 * spt * ticks) if rank == 0 else np.full(2 * spt * ticks, 7.0, np.float32)
        line = ctx.stereo(src)
        fdata = W.random_bytes(99, W.FRAME_BYTES) if rank == 0 else np.zeros(W.FRAME_BYTES, np.uint8)
        frame = ctx.frame(W.FRAME_W, W.FRAME_H, data=fdata)
        line.broadcast(0)
        frame.broadcast(0)
        # every session then runs its OWN graph on the shared source: Amplifier with a per-rank gain
        amp = ctx.module(mxl.MOD_AMPLIFIER, (0.25 + 0.5 * rank, 0.0))
        out = ctx.line(mxl.LINE_STEREO, spt * ticks)
        amp.run_tick(0, [line, None], [out])
        want_src = W.uniform_pm1(4242, 2 * spt * ticks)
        want = po.amplifier(want_src, None, 0.25 + 0.5 * rank, 0.0)
        ok_audio = bool(np.array_equal(out.download().view(np.uint32), want.view(np.uint32)))
        ok_video = bool(np.array_equal(frame.download_raw(), W.random_bytes(99, W.FRAME_BYTES)))
        # bandwidth of a frame broadcast: device time of 50 broadcasts back to back, after warm-up broadcasts
        # that absorb the skew between the two processes
        for _ in range(5):
            frame.broadcast(0)
        ctx.synchronize()
        ctx.timer_begin()
        for _ in range(50):
            frame.broadcast(0)
        ctx.timer_end()
        ms = ctx.timer_elapsed_ms() / 50
        ctx.comm_destroy()
        ctx.close()
        q.put((rank, ok_audio, ok_video, ms, None))
    except Exception as e:                               # noqa: BLE001 -- reported to the parent
        q.put((rank, False, False, 0.0, repr(e)))


@pytest.mark.gpu
def test_shared_source_broadcast_two_gpus():
    import mixlab_b200 as mxl
    import ctypes as C
    cudart_count = C.c_int(0)
    try:
        C.CDLL("libcudart.so").cudaGetDeviceCount(C.byref(cudart_count))
    except OSError:
        try:
            import torch
            cudart_count.value = torch.cuda.device_count()
        except Exception:                                # noqa: BLE001
            cudart_count.value = 0
    if cudart_count.value < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    uid = mxl.comm_unique_id()
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    procs = [mpctx.Process(target=_rank_main, args=(r, 2, uid, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok_audio, ok_video, ms, err in sorted(res):
        assert err is None, (rank, err)
        assert ok_audio and ok_video, rank
        print("rank %d: 1080p frame broadcast %.3f ms = %.1f GB/s" % (rank, ms, 3110400 / ms / 1e6))

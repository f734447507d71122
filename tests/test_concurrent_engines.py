"""Several engine threads on one GPU, one context each (the reference runs one engine thread per Engine, engine.rs:78;
a host serving several sessions from one process runs several).  Contexts share nothing but the device and the
library's caches-per-context: every thread must get the result a lone engine gets, whatever the interleaving."""
import threading

import numpy as np
import pytest

from mixlab_b200 import workloads as W
from mixlab_b200.session import AVSession

pytestmark = pytest.mark.gpu


def run_session(mxl, seed, ticks_per_call, calls, out):
    try:
        with mxl.Context(0, 48000, 800) as ctx:
            sess = AVSession(ctx, W.config2_graph(), ticks_per_call, video=True, width=560, height=350, fader=0.25,
                             seed=seed, unique_frames=3)
            sess.upload_inputs()
            tick, master, frames = 0, [], []
            for _ in range(calls):
                sess.run_step_host(tick)
                tick += ticks_per_call
                master.append(np.array(sess.host_master.array[:ticks_per_call * 800 * 2]))
                frames.append(np.array(sess.host_out.array[:ticks_per_call * sess.frame_bytes]))
            out[seed] = (np.concatenate(master), np.concatenate(frames))
            sess.close()
    except Exception as e:                                  # surfaced by the assertion in the test body
        out[seed] = e


def test_four_engine_threads_match_lone_engines(mxl):
    seeds = [11, 22, 33, 44]
    alone = {}
    for s in seeds:
        run_session(mxl, s, 4, 12, alone)
        assert not isinstance(alone[s], Exception), alone[s]
    together = {}
    threads = [threading.Thread(target=run_session, args=(mxl, s, 4, 12, together)) for s in seeds]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for s in seeds:
        assert not isinstance(together[s], Exception), together[s]
        assert np.array_equal(alone[s][0].view(np.uint32), together[s][0].view(np.uint32)), s      # audio bit for bit
        assert np.array_equal(alone[s][1], together[s][1]), s                                       # video byte for byte
    # distinct seeds give distinct pictures (the threads did not read each other's frames)
    assert not np.array_equal(together[11][1], together[22][1])

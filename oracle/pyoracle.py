"""ctypes view of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (mixlab_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

WAVE_ON, WAVE_OFF, WAVE_SINE, WAVE_SQUARE, WAVE_TRIANGLE, WAVE_SAW = range(6)
LINE_MONO, LINE_STEREO, LINE_VIDEO = range(3)
(MOD_AMPLIFIER, MOD_ENVELOPE, MOD_EQ_THREE, MOD_FM_SINE, MOD_MIXER, MOD_OSCILLATOR, MOD_PLOTTER,
 MOD_STEREO_PANNER, MOD_STEREO_SPLITTER, MOD_TRIGGER, MOD_METER, MOD_SOURCE_STEREO,
 MOD_SOURCE_MONO) = range(13)


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc -O2 -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("mixlab_oracle.c", "mixlab_oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class EqThreeState(C.Structure):
    _fields_ = [("lo_coef", C.c_double), ("hi_coef", C.c_double), ("lo_poles", C.c_double * 4),
                ("hi_poles", C.c_double * 4), ("history", C.c_double * 3)]


class EnvelopeState(C.Structure):
    _fields_ = [("state", C.c_int), ("seq", C.c_uint64), ("off_amplitude", C.c_double)]


class FrameLayout(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("stride", C.c_uint32 * 3),
                ("plane_h", C.c_uint32 * 3), ("offset", C.c_size_t * 3), ("size", C.c_size_t)]


class ScaleGeometry(C.Structure):
    _fields_ = [("scaled_w", C.c_uint32), ("scaled_h", C.c_uint32), ("letterbox_x", C.c_uint32),
                ("letterbox_y", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_db_to_linear.restype = C.c_double
        _lib.orc_db_to_linear.argtypes = [C.c_double]
        _lib.orc_graph_create.restype = C.c_void_p
        _lib.orc_graph_create.argtypes = [C.c_double, C.c_uint32]
        _lib.orc_graph_destroy.argtypes = [C.c_void_p]
        _lib.orc_graph_add.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _lib.orc_graph_connect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.orc_graph_set_source.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        _lib.orc_graph_run_tick.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_graph_last_order.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib.orc_graph_meter.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_fader_to_u8.restype = C.c_uint8
        _lib.orc_fader_to_u8.argtypes = [C.c_double]
        _lib.orc_clip_detect.restype = C.c_int
        _lib.orc_session_run.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        _lib.orc_resampler_create.restype = C.c_void_p
        _lib.orc_resampler_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        _lib.orc_resampler_destroy.argtypes = [C.c_void_p]
        _lib.orc_resampler_push.restype = C.c_size_t
        _lib.orc_resampler_push.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        _lib.orc_resampler_coef.restype = C.c_double
        _lib.orc_resampler_coef.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        _lib.orc_resampler_ratio.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def db_to_linear(db):
    return lib().orc_db_to_linear(float(db))


def mixer(inputs, gain_db, fader, cue, frames):
    """inputs: list of stereo arrays (2*frames f32) or None.  Returns (master, cue)."""
    n = len(inputs)
    arrs = [None if a is None else _f32(a) for a in inputs]
    ptrs = (C.c_void_p * max(n, 1))(*[None if a is None else a.ctypes.data for a in arrs])
    g = np.ascontiguousarray(gain_db, dtype=np.float64)
    f = np.ascontiguousarray(fader, dtype=np.float64)
    c = np.ascontiguousarray(cue, dtype=np.uint8)
    master = np.empty(2 * frames, np.float32)
    cue_out = np.empty(2 * frames, np.float32)
    lib().orc_mixer(ptrs, _p(g), _p(f), _p(c), C.c_int(n), _p(master), _p(cue_out),
                    C.c_size_t(2 * frames))
    return master, cue_out


def amplifier(inp, mod, amplitude, mod_depth):
    inp = _f32(inp)
    mod = None if mod is None else _f32(mod)
    out = np.empty_like(inp)
    lib().orc_amplifier(_p(inp), _p(mod), C.c_double(amplitude), C.c_double(mod_depth), _p(out),
                        C.c_size_t(inp.size))
    return out


class EqThree:
    def __init__(self, sample_rate):
        self.state = EqThreeState()
        lib().orc_eq_three_create(C.byref(self.state), C.c_double(sample_rate))

    def run(self, gains_db, inp):
        inp = _f32(inp)
        out = np.empty_like(inp)
        lib().orc_eq_three_run(C.byref(self.state), C.c_double(gains_db[0]), C.c_double(gains_db[1]),
                               C.c_double(gains_db[2]), _p(inp), _p(out), C.c_size_t(inp.size))
        return out


def oscillator(t, sample_rate, freq, waveform, n):
    mono = np.empty(n, np.float32)
    stereo = np.empty(2 * n, np.float32)
    lib().orc_oscillator(C.c_uint64(t), C.c_double(sample_rate), C.c_double(freq), C.c_int(waveform),
                         _p(mono), _p(stereo), C.c_size_t(n))
    return mono, stereo


class Envelope:
    def __init__(self):
        self.state = EnvelopeState()
        lib().orc_envelope_create(C.byref(self.state))

    def run(self, t, sample_rate, attack_ms, decay_ms, sustain, release_ms, inp):
        inp = _f32(inp)
        out = np.empty_like(inp)
        lib().orc_envelope_run(C.byref(self.state), C.c_uint64(t), C.c_double(sample_rate),
                               C.c_double(attack_ms), C.c_double(decay_ms), C.c_double(sustain),
                               C.c_double(release_ms), _p(inp), _p(out), C.c_size_t(inp.size))
        return out


def fm_sine(t, sample_rate, freq_lo, freq_hi, inp, n=None):
    inp = None if inp is None else _f32(inp)
    n = inp.size if inp is not None else n
    out = np.empty(2 * n, np.float32)
    lib().orc_fm_sine(C.c_uint64(t), C.c_double(sample_rate), C.c_double(freq_lo), C.c_double(freq_hi),
                      _p(inp), _p(out), C.c_size_t(n))
    return out


def stereo_panner(left, right, n):
    left = None if left is None else _f32(left)
    right = None if right is None else _f32(right)
    out = np.empty(2 * n, np.float32)
    lib().orc_stereo_panner(_p(left), _p(right), _p(out), C.c_size_t(n))
    return out


def stereo_splitter(inp, n):
    inp = None if inp is None else _f32(inp)
    left = np.empty(n, np.float32)
    right = np.empty(n, np.float32)
    lib().orc_stereo_splitter(_p(inp), _p(left), _p(right), C.c_size_t(n))
    return left, right


def trigger(is_open, n):
    out = np.empty(n, np.float32)
    lib().orc_trigger(C.c_int(1 if is_open else 0), _p(out), C.c_size_t(n))
    return out


def pcm_pack_i16(samples):
    samples = _f32(samples)
    out = np.empty(samples.size, np.int16)
    lib().orc_pcm_pack_i16(_p(samples), _p(out), C.c_size_t(samples.size))
    return out


def pcm_unpack_i16(pcm):
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    out = np.empty(pcm.size, np.float32)
    lib().orc_pcm_unpack_i16(_p(pcm), _p(out), C.c_size_t(pcm.size))
    return out


def plotter_tap(stereo):
    stereo = _f32(stereo)
    n = stereo.size // 2
    left = np.empty(n, np.float32)
    right = np.empty(n, np.float32)
    lib().orc_plotter_tap(_p(stereo), _p(left), _p(right), C.c_size_t(n))
    return left, right


def clip_detect(stereo):
    stereo = _f32(stereo)
    return bool(lib().orc_clip_detect(_p(stereo), C.c_size_t(stereo.size)))


def meter(stereo):
    stereo = _f32(stereo)
    peak = (C.c_float * 2)()
    sumsq = (C.c_double * 2)()
    clip = C.c_int(0)
    lib().orc_meter(_p(stereo), C.c_size_t(stereo.size // 2), peak, sumsq, C.byref(clip))
    return (peak[0], peak[1]), (sumsq[0], sumsq[1]), bool(clip.value)


class StreamInput:
    """StreamInput::run_tick (src/module/stream_input.rs:72-147) restated.  MediaTime / MediaDuration are
    Rational64 (util/src/time.rs:9-10,77-78): fractions.Fraction here.  The receiver's two ring buffers
    (src/source.rs:73-78,97-98) are deques fed by write_audio / write_video (source.rs:156-190)."""
    RING_CAPACITY = 65536

    def __init__(self, sample_rate):
        from collections import deque
        self.sample_rate = int(sample_rate)
        self.audio_rx, self.video_rx = deque(), deque()
        self.audio_frame = None          # self.audio_frame: [source_id, source_time, data]
        self.video_frame = None          # self.video_frame: (source_id, source_time, data, duration_hint)
        self.source = None               # SourceTiming (id, epoch)

    def write_audio(self, source_id, source_time, data):
        if len(self.audio_rx) >= self.RING_CAPACITY:
            return False                 # tx.audio.push(frame).map_err(|_| ())
        self.audio_rx.append([source_id, source_time, np.array(data, np.int16)])
        return True

    def write_video(self, source_id, source_time, data, duration_hint):
        if len(self.video_rx) >= self.RING_CAPACITY:
            return False
        self.video_rx.append((source_id, source_time, data, duration_hint))
        return True

    def run_tick(self, engine_time, audio_len):
        """engine_time: absolute sample index (u64); audio_len: f32 in the tick's stereo line.
        Returns (video_out, audio_out): video_out = None or (data, duration_hint, tick_offset)."""
        from fractions import Fraction
        engine_time = Fraction(int(engine_time), self.sample_rate)                  # 73
        audio_out = np.empty(audio_len, np.float32)
        pos = 0
        tick_duration = Fraction(audio_len // 2, self.sample_rate)                   # 80
        video_frame, self.video_frame = self.video_frame, None                      # 82-86
        if video_frame is None and self.video_rx:
            video_frame = self.video_rx.popleft()
        existing_source_id = self.source[0] if self.source is not None else None    # 88
        while pos < audio_len:                                                       # 92
            frame, self.audio_frame = self.audio_frame, None                        # 93-97
            if frame is None and self.audio_rx:
                frame = self.audio_rx.popleft()
            if frame is not None:
                if existing_source_id != frame[0]:                                   # 100-106
                    self.source = (frame[0], engine_time - frame[1])
                n = min(audio_len - pos, frame[2].size)                              # 108
                audio_out[pos:pos + n] = pcm_unpack_i16(frame[2][:n])                # 110-112 convert_sample
                pos += n                                                             # 114
                if n < frame[2].size:                                                # 116-119
                    frame[2] = frame[2][n:]
                    self.audio_frame = frame
            else:
                audio_out[pos:] = 0.0                                                # 121 util::zero
                break
        video_out = None                                                             # 126-143
        if video_frame is not None:
            tick_offset = Fraction(0)
            if self.source is not None:
                off = video_frame[1] + self.source[1] - engine_time
                if off >= 0:
                    tick_offset = off
            if tick_offset > tick_duration:
                self.video_frame = video_frame
            else:
                video_out = (video_frame[2], video_frame[3], tick_offset)
        return video_out, audio_out


def frame_layout(width, height):
    lay = FrameLayout()
    lib().orc_frame_layout_yuv420p(C.c_uint32(width), C.c_uint32(height), C.byref(lay))
    return lay


def frame_blank(lay):
    data = np.empty(lay.size, np.uint8)
    lib().orc_frame_blank(C.byref(lay), _p(data))
    return data


def fader_to_u8(fader):
    return int(lib().orc_fader_to_u8(float(fader)))


def video_crossfade(lay, a, b, fade):
    """a / b: full frame byte arrays in `lay` layout, or None (layer missing)."""
    out = frame_blank(lay)
    a = None if a is None else np.ascontiguousarray(a, dtype=np.uint8)
    b = None if b is None else np.ascontiguousarray(b, dtype=np.uint8)
    lib().orc_video_crossfade(C.byref(lay), _p(a), _p(b), C.c_uint8(fade), _p(out))
    return out


def unify_picture(aw, ah, bw, bh):
    w, h = C.c_uint32(), C.c_uint32()
    lib().orc_unify_picture(C.c_uint32(aw), C.c_uint32(ah), C.c_uint32(bw), C.c_uint32(bh),
                            C.byref(w), C.byref(h))
    return w.value, h.value


def scale_geometry(in_w, in_h, out_w, out_h):
    g = ScaleGeometry()
    lib().orc_scale_geometry_yuv420p(C.c_uint32(in_w), C.c_uint32(in_h), C.c_uint32(out_w),
                                     C.c_uint32(out_h), C.byref(g))
    return g.scaled_w, g.scaled_h, g.letterbox_x, g.letterbox_y


def yuv420p_to_rgba(lay, yuv):
    yuv = np.ascontiguousarray(yuv, dtype=np.uint8)
    out = np.empty(lay.width * lay.height * 4, np.uint8)
    lib().orc_yuv420p_to_rgba(C.byref(lay), _p(yuv), _p(out))
    return out


def rgba_to_yuv420p(lay, rgba, into=None):
    """into: frame bytes whose stride padding is kept (default: a blank frame)."""
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    out = frame_blank(lay) if into is None else np.array(into, np.uint8)
    lib().orc_rgba_to_yuv420p(C.byref(lay), _p(rgba), _p(out))
    return out


def bicubic_plane(src, sw, sh, sstride, dw, dh, dstride):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    dst = np.zeros(dstride * dh, np.uint8)
    lib().orc_bicubic_plane(_p(src), C.c_uint32(sw), C.c_uint32(sh), C.c_uint32(sstride), _p(dst),
                            C.c_uint32(dw), C.c_uint32(dh), C.c_uint32(dstride))
    return dst


def letterbox_scale(data, lay_in, out_w, out_h):
    """DynamicScaler::scale (src/video/encode.rs:338-397) from the pieces above: identity when the sizes agree,
    else a blank target and a per-plane resample into the letterboxed sub-frame."""
    if (lay_in.width, lay_in.height) == (out_w, out_h):
        return np.array(data, np.uint8)
    sw_, sh_, lx, ly = scale_geometry(lay_in.width, lay_in.height, out_w, out_h)
    lay = frame_layout(out_w, out_h)
    want = frame_blank(lay)
    for p in range(3):
        sh = 0 if p == 0 else 1
        sw = lay_in.width if p == 0 else (lay_in.width + 1) // 2
        shh = lay_in.plane_h[p]
        dw, dh = sw_ >> sh, sh_ >> sh
        if dw == 0 or dh == 0:
            continue
        src = data[lay_in.offset[p]:lay_in.offset[p] + lay_in.stride[p] * lay_in.plane_h[p]]
        dst = bicubic_plane(src, sw, shh, lay_in.stride[p], dw, dh, dw)
        plane = want[lay.offset[p]:lay.offset[p] + lay.stride[p] * lay.plane_h[p]].reshape(lay.plane_h[p], lay.stride[p])
        plane[(ly >> sh):(ly >> sh) + dh, (lx >> sh):(lx >> sh) + dw] = dst.reshape(dh, dw)
    return want


def letterbox_scale_fast(data, lay_in, out_w, out_h):
    """orc_letterbox_scale: the same result as letterbox_scale through the CPU-arranged scaler (the timed CPU arm)."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    lay = frame_layout(out_w, out_h)
    out = np.empty(lay.size, np.uint8)
    lib().orc_letterbox_scale(C.byref(lay_in), _p(data), C.byref(lay), _p(out))
    return out


class SessionStruct(C.Structure):
    _fields_ = [("graph", C.c_void_p), ("master_module", C.c_int), ("lay", FrameLayout), ("lay_mon", FrameLayout),
                ("layers_a", C.c_void_p), ("layers_b", C.c_void_p), ("n_layers", C.c_uint32), ("ticks_per_frame", C.c_uint32),
                ("pcm_in", C.c_void_p), ("pcm_in_samples", C.c_size_t), ("fade", C.c_uint8),
                ("composite", C.c_void_p), ("monitor_out", C.c_void_p), ("pcm_out", C.c_void_p), ("scratch", C.c_void_p)]


class Session:
    """orc_session: one live session of bench.py's session variant on the CPU -- per tick two StreamInput unpacks,
    Engine::run_tick over the audio graph, VideoMixer blank + crossfade of the stored layers, the Monitor's scaler
    and PCM pack -- entirely in C (one ctypes call runs n ticks), so that threads of the CPU arm do not meet on the GIL."""

    def __init__(self, graph, master_module, width, height, mon_w, mon_h, layers_a, layers_b, ticks_per_frame, pcm_in, fader):
        self.graph = graph
        self.lay, self.lay_mon = frame_layout(width, height), frame_layout(mon_w, mon_h)
        self.layers_a = np.ascontiguousarray(layers_a, np.uint8)
        self.layers_b = np.ascontiguousarray(layers_b, np.uint8)
        self.pcm_in = np.ascontiguousarray(pcm_in, np.int16)
        n = 2 * graph.spt
        self.composite = np.empty(self.lay.size, np.uint8)
        self.monitor_out = np.empty(self.lay_mon.size, np.uint8)
        self.pcm_out = np.empty(n, np.int16)
        self.scratch = np.empty(3 * n, np.float32)
        self.s = SessionStruct(graph._g, master_module, self.lay, self.lay_mon, _p(self.layers_a), _p(self.layers_b),
                               self.layers_a.size // self.lay.size, ticks_per_frame, _p(self.pcm_in), self.pcm_in.size,
                               fader_to_u8(fader), _p(self.composite), _p(self.monitor_out), _p(self.pcm_out), _p(self.scratch))

    def run(self, tick0, n_ticks):
        lib().orc_session_run(C.byref(self.s), C.c_uint64(tick0), C.c_uint32(n_ticks))


class Resampler:
    """orc_resampler: the definition of the audio sample-rate converter (UNPINNED: the reference has only the TODO,
    src/icecast/mod.rs:94-97).  push() takes interleaved f32 (or int16, converted s / 32768 like StreamInput) and returns
    the output frames that became determined."""

    def __init__(self, in_rate, out_rate, channels):
        self.channels = channels
        self._r = lib().orc_resampler_create(in_rate, out_rate, channels)
        L, M = C.c_uint32(), C.c_uint32()
        lib().orc_resampler_ratio(self._r, C.byref(L), C.byref(M))
        self.L, self.M = L.value, M.value

    def push(self, x):
        x = np.asarray(x)
        if x.dtype == np.int16:
            x = pcm_unpack_i16(x)
        x = _f32(x)
        n = x.size // self.channels
        out = np.empty((n * self.L // self.M + 8) * self.channels, np.float32)
        made = lib().orc_resampler_push(self._r, _p(x), n, _p(out), out.size // self.channels)
        return out[:made * self.channels]

    def coef(self, phase, k):
        return lib().orc_resampler_coef(self._r, phase, k)

    def close(self):
        if self._r:
            lib().orc_resampler_destroy(self._r)
            self._r = None

    def __del__(self):
        self.close()


class OutputDevice:
    """OutputDevice::update's channel assignment (src/module/output_device.rs:152-168) and run_tick (173-207)
    restated; `channels` = stream.config.channels of the opened device (0 = no stream)."""
    RING_CAPACITY = 65536

    def __init__(self, left, right, channels):
        self.scratch = np.zeros(0, np.float32)
        self.left = self.right = None
        self.channels = channels
        self.ring = np.zeros(0, np.float32)
        self.update(left, right)

    def update(self, left, right):
        if self.channels:                                            # `if let Some(stream)` (152)
            if self.left != left or self.right != right:             # 156-160
                self.scratch[:] = 0.0
            self.left = left if (left is not None and left < self.channels) else None       # 164-165
            self.right = right if (right is not None and right < self.channels) else None   # 167-168

    def run_tick(self, inp):
        inp = _f32(inp)
        clip = False                                                 # 176
        if self.channels:                                            # 178
            oc = self.channels
            n = inp.size // 2                                        # 180
            if self.scratch.size < n * oc:                           # 183-185
                self.scratch = np.concatenate([self.scratch, np.zeros(n * oc - self.scratch.size, np.float32)])
            view = self.scratch[:n * oc].reshape(n, oc)
            if self.left is not None:                                # 188-196
                s = inp[0::2]
                clip |= bool(np.any((s < -1.0) | (s > 1.0)))
                view[:, self.left] = s
            if self.right is not None:                               # 198-206 (after left: wins on a shared channel)
                s = inp[1::2]
                clip |= bool(np.any((s < -1.0) | (s > 1.0)))
                view[:, self.right] = s
            room = self.RING_CAPACITY - self.ring.size               # push_slice (207): what fits
            self.ring = np.concatenate([self.ring, self.scratch[:min(n * oc, room)]])
        return clip

    def pop(self, cap):
        out, self.ring = self.ring[:cap], self.ring[cap:]
        return out


class MonitorFeed:
    """Monitor::run_tick (src/module/monitor.rs:112-140), the codec thread's loop body (235-247) and EncodeStream
    (src/video/encode.rs:34-107) with AudioCtx::send_audio (184-221) and VideoCtx::send_frame (279-287), up to
    the two encoder calls.  Collects what aac::Encoder::encode and AvcEncoder::send_frame would be called with."""
    FRAGMENT = 2 * 1024                      # AUDIO_CHANNELS * SAMPLES_PER_CHANNEL_PER_FRAGMENT (encode.rs:20-22)

    def __init__(self, sample_rate, width=560, height=350, time_base=None):
        from fractions import Fraction
        self.F = Fraction
        self.sample_rate = int(sample_rate)
        self.width, self.height = width, height                     # monitor.rs:21-22
        self.time_base = int(time_base if time_base is not None else sample_rate)   # monitor.rs:195
        self.epoch = None
        self.audio_timestamp = Fraction(0)                           # EncodeStream::new (encode.rs:36-44)
        self.video_timestamp = Fraction(0)
        self.pcm_buff = np.empty(0, np.int16)
        self.audio_out = []                  # (decode_timestamp, duration, fragment i16)
        self.video_out = []                  # (pts, duration_in_base, blank?, picture bytes)
        self.blank = frame_blank(frame_layout(width, height))        # VideoCtx::new (encode.rs:273-276)

    @staticmethod
    def _round_to_base(x, base):             # util/src/time.rs:17-19: (ratio * base).to_integer(), toward zero
        y = x * base
        n, d = y.numerator, y.denominator
        return n // d if n >= 0 else -((-n) // d)

    def _encode_video(self, duration, picture, blank):               # encode.rs:86-100
        start = self.video_timestamp
        end = start + duration
        self.video_timestamp = end
        s, e = self._round_to_base(start, self.time_base), self._round_to_base(end, self.time_base)
        self.video_out.append((s, e - s, blank, picture))

    def run_tick(self, time, audio, video):
        """time: absolute sample index; audio: the tick's stereo f32; video: None or
        (data, layout, duration_hint, tick_offset) with Fractions."""
        F = self.F
        absolute = F(int(time), self.sample_rate)                    # monitor.rs:118
        if self.epoch is None:
            self.epoch = absolute                                    # 119
        timestamp = absolute - self.epoch                            # 120
        # encode.send_audio (monitor.rs:236 -> encode.rs:46-59 -> 184-221)
        self.pcm_buff = np.concatenate([self.pcm_buff, pcm_pack_i16(audio)])
        if self.pcm_buff.size > self.FRAGMENT:
            frag = self.pcm_buff[:self.FRAGMENT].copy()
            duration = F(1024, self.sample_rate)
            self.pcm_buff = self.pcm_buff[self.FRAGMENT:]
            self.audio_out.append((self.audio_timestamp, duration, frag))
            self.audio_timestamp += duration
        if video is not None:                                        # monitor.rs:238-243
            data, lay, duration_hint, tick_offset = video
            frame_timestamp = timestamp + tick_offset
            end_timestamp = frame_timestamp + duration_hint          # encode.rs:62
            if not (end_timestamp < self.video_timestamp):           # 64-67
                self._encode_video(end_timestamp - self.video_timestamp,
                                   letterbox_scale(data, lay, self.width, self.height), False)   # 73-75; 279-287
        if self.video_timestamp < timestamp:                         # barrier (monitor.rs:245; encode.rs:78-84)
            self._encode_video(timestamp - self.video_timestamp, self.blank, True)


class Graph:
    """Engine::run_tick walker (engine.rs:400-510) over oracle modules."""

    def __init__(self, sample_rate, samples_per_tick):
        self.spt = samples_per_tick
        self._g = lib().orc_graph_create(C.c_double(sample_rate), C.c_uint32(samples_per_tick))
        self._keep = []
        self.out_types = {}

    def close(self):
        if self._g:
            lib().orc_graph_destroy(self._g)
            self._g = None

    def __del__(self):
        self.close()

    def add(self, kind, params=()):
        p = np.ascontiguousarray(params, dtype=np.float64)
        mid = lib().orc_graph_add(self._g, C.c_int(kind), _p(p) if p.size else None, C.c_int(p.size))
        if mid < 0:
            raise ValueError("unknown module kind %r" % kind)
        return mid

    def connect(self, in_module, in_index, out_module, out_index):
        return lib().orc_graph_connect(self._g, in_module, in_index, out_module, out_index)

    def set_source(self, module, data, width):
        data = _f32(data)
        self._keep.append(data)
        lib().orc_graph_set_source(self._g, module, _p(data), C.c_size_t(data.size // width))

    def run_tick(self, tick, capture=None, capture_len=0):
        buf = None
        if capture is not None:
            buf = np.zeros(capture_len, np.float32)
            lib().orc_graph_run_tick(self._g, C.c_uint64(tick), capture[0], capture[1], _p(buf))
        else:
            lib().orc_graph_run_tick(self._g, C.c_uint64(tick), -1, -1, None)
        return buf

    def last_order(self):
        arr = (C.c_int * 1024)()
        n = lib().orc_graph_last_order(self._g, arr, 1024)
        return list(arr[:n])

    def meter(self, module):
        peak = (C.c_float * 2)()
        sumsq = (C.c_double * 2)()
        clip = C.c_int(0)
        lib().orc_graph_meter(self._g, module, peak, sumsq, C.byref(clip))
        return (peak[0], peak[1]), (sumsq[0], sumsq[1]), bool(clip.value)


# ---- GraphDesc (mixlab_b200.workloads) instantiated on the oracle engine walker ------------------
def oracle_kind(name):
    return {
        "Amplifier": MOD_AMPLIFIER, "Envelope": MOD_ENVELOPE, "EqThree": MOD_EQ_THREE,
        "FmSine": MOD_FM_SINE, "Mixer": MOD_MIXER, "Oscillator": MOD_OSCILLATOR,
        "Plotter": MOD_PLOTTER, "StereoPanner": MOD_STEREO_PANNER,
        "StereoSplitter": MOD_STEREO_SPLITTER, "Trigger": MOD_TRIGGER, "Meter": MOD_METER,
        "SourceStereo": MOD_SOURCE_STEREO, "SourceMono": MOD_SOURCE_MONO,
    }[name]


def oracle_params(name, params):
    if params is None:
        return ()
    if name == "Mixer":
        flat = [float(len(params))]
        for g, f, c in params:
            flat += [float(g), float(f), 1.0 if c else 0.0]
        return flat
    if name == "Oscillator":
        return [float(params[0]), float(params[1])]
    if name == "Trigger":
        return [1.0 if params[0] == 0 else 0.0]      # GATE_OPEN = 0
    return [float(p) for p in params]


def build_graph(desc, sample_rate, spt):
    """desc: an object with .modules [(kind name, params)] and .connections [(im, ii, om, oi)]."""
    g = Graph(float(sample_rate), spt)
    ids = [g.add(oracle_kind(kind), oracle_params(kind, params)) for kind, params in desc.modules]
    for im, ii, om, oi in desc.connections:
        rc = g.connect(ids[im], ii, ids[om], oi)
        if rc != 0:
            raise ValueError("connect(%d,%d,%d,%d) -> %d" % (im, ii, om, oi, rc))
    return g, ids

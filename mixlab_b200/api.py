"""ctypes binding of libmixlab_b200.so -- a thin Python view of include/mixlab_b200.h.

The library IS the product: module classes, graph executor and kernels are C++/CUDA behind the C
ABI.  This file only marshals arguments for tests/ and bench.py and mirrors the reference's
vocabulary (ModuleT: create / params / update / run_tick / inputs / outputs; Workspace.connect;
Engine.run_tick).  There is no CPU fallback anywhere: if the shared library is missing or a compute
entry point is called without a CUDA device, an exception is raised.
"""
import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmixlab_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mixlab_b200.h")

# mxl_status
OK, ERR_INVALID, ERR_LINE_TYPE, ERR_PARAMS, ERR_CUDA, ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5
ERR_NO_INPUT, ERR_NO_OUTPUT, ERR_TYPE_MISMATCH, ERR_OOM, ERR_UNSUPPORTED, ERR_LENGTH = -6, -7, -8, -9, -10, -11
DEVICE_NONE = -1

LINE_MONO, LINE_STEREO, LINE_VIDEO = 0, 1, 2

MOD_AMPLIFIER, MOD_ENVELOPE, MOD_EQ_THREE, MOD_FM_SINE, MOD_MIXER, MOD_MONITOR = 0, 1, 2, 3, 4, 5
MOD_OSCILLATOR, MOD_OUTPUT_DEVICE, MOD_PLOTTER, MOD_STEREO_PANNER, MOD_STEREO_SPLITTER = 6, 7, 8, 9, 10
MOD_STREAM_INPUT, MOD_STREAM_OUTPUT, MOD_TRIGGER, MOD_VIDEO_MIXER, MOD_MEDIA_SOURCE = 11, 12, 13, 14, 15
MOD_METER, MOD_SOURCE_MONO, MOD_SOURCE_STEREO, MOD_SOURCE_VIDEO, MOD_PCM_SINK = 32, 33, 34, 35, 36
STAGE_FUSED_VOICE_MIX = 1000

WAVE_ON, WAVE_OFF, WAVE_SINE, WAVE_SQUARE, WAVE_TRIANGLE, WAVE_SAW = 0, 1, 2, 3, 4, 5
GATE_OPEN, GATE_CLOSED = 0, 1


class MxlError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("mxl status %d: %s" % (status, message))
        self.status = status


class AmplifierParams(C.Structure):
    _fields_ = [("amplitude", C.c_double), ("mod_depth", C.c_double)]


class EnvelopeParams(C.Structure):
    _fields_ = [("attack_ms", C.c_double), ("decay_ms", C.c_double), ("sustain_amplitude", C.c_double),
                ("release_ms", C.c_double)]


class EqThreeParams(C.Structure):
    _fields_ = [("gain_lo_db", C.c_double), ("gain_mid_db", C.c_double), ("gain_hi_db", C.c_double)]


class FmSineParams(C.Structure):
    _fields_ = [("freq_lo", C.c_double), ("freq_hi", C.c_double)]


class MixerChannelParams(C.Structure):
    _fields_ = [("gain_db", C.c_double), ("fader", C.c_double), ("cue", C.c_int32), ("_pad", C.c_int32)]


class MixerParams(C.Structure):
    _fields_ = [("channels", C.POINTER(MixerChannelParams)), ("n_channels", C.c_uint32)]


class OscillatorParams(C.Structure):
    _fields_ = [("freq", C.c_double), ("waveform", C.c_int32), ("_pad", C.c_int32)]


class TriggerParams(C.Structure):
    _fields_ = [("gate", C.c_int32)]


class VideoMixerParams(C.Structure):
    _fields_ = [("a", C.c_int32), ("b", C.c_int32), ("fader", C.c_double)]


class FrameLayout(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("stride", C.c_uint32 * 3),
                ("plane_h", C.c_uint32 * 3), ("offset", C.c_uint64 * 3), ("size", C.c_uint64)]


class ScaleGeometry(C.Structure):
    _fields_ = [("scaled_w", C.c_uint32), ("scaled_h", C.c_uint32), ("letterbox_x", C.c_uint32),
                ("letterbox_y", C.c_uint32)]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_uint32), ("total_ms", C.c_float)]


class StageInfo(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_modules", C.c_int32), ("n_launches", C.c_int32),
                ("last_ms", C.c_float), ("algorithmic_bytes", C.c_uint64), ("host_us", C.c_float), ("_pad", C.c_float)]


class MonitorParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("time_base", C.c_int64)]


class AudioFragment(C.Structure):
    _fields_ = [("decode_num", C.c_int64), ("decode_den", C.c_int64), ("duration_num", C.c_int64),
                ("duration_den", C.c_int64), ("n_samples", C.c_uint32), ("_pad", C.c_uint32)]


class VideoJob(C.Structure):
    _fields_ = [("pts", C.c_int64), ("duration", C.c_int64), ("time_base", C.c_int64), ("blank", C.c_int32),
                ("_pad", C.c_int32), ("frame", C.c_void_p)]


class OutputDeviceParams(C.Structure):
    _fields_ = [("left", C.c_int32), ("right", C.c_int32), ("channels", C.c_uint32), ("_pad", C.c_uint32)]


class PerfAccount(C.Structure):
    _fields_ = [("module_id", C.c_int32), ("kind", C.c_int32), ("last_us", C.c_float), ("host_us", C.c_float)]


class HostRef(C.Structure):
    """mxl_host_ref: one InputRef / OutputRef of the reference with host slices (io.rs:19-34,79-98)."""
    _fields_ = [("type", C.c_int32), ("connected", C.c_int32), ("samples", C.c_void_p), ("len", C.c_uint64),
                ("frame", C.c_void_p), ("duration_num", C.c_int64), ("duration_den", C.c_int64),
                ("offset_num", C.c_int64), ("offset_den", C.c_int64)]


METER_RECORD = np.dtype([("peak", np.float32, 2), ("clip", np.int32), ("_pad", np.int32), ("sumsq", np.float64, 2)])


_PARAM_TYPES = {
    MOD_AMPLIFIER: AmplifierParams, MOD_ENVELOPE: EnvelopeParams, MOD_EQ_THREE: EqThreeParams,
    MOD_FM_SINE: FmSineParams, MOD_MIXER: MixerParams, MOD_OSCILLATOR: OscillatorParams,
    MOD_TRIGGER: TriggerParams, MOD_VIDEO_MIXER: VideoMixerParams, MOD_MONITOR: MonitorParams,
    MOD_STREAM_OUTPUT: MonitorParams, MOD_OUTPUT_DEVICE: OutputDeviceParams,
}

_lib = None


def declared_symbols():
    """Every function include/mixlab_b200.h declares (used by the symbol-export test)."""
    text = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"MXL_API[^;(]*?\b(mxl_[a-z0-9_]+)\s*\(", text)))


def lib():
    """Loads libmixlab_b200.so.  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -m mixlab_b200.build` (or __graft_entry__.build()). "
                          "There is no CPU fallback for the tick path." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32, dbl = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_double
    sig = {
        "mxl_ctx_create": (vp, [i32, u32, u32]),
        "mxl_ctx_create_on_stream": (vp, [i32, u32, u32, vp]),
        "mxl_ctx_destroy": (i32, [vp]),
        "mxl_ctx_synchronize": (i32, [vp]),
        "mxl_ctx_trim_frame_pool": (i32, [vp, u64, u64, C.POINTER(u64)]),
        "mxl_ctx_set_copy_overlap": (i32, [vp, i32]),
        "mxl_ctx_download_fence": (i32, [vp, u32]),
        "mxl_ctx_wait_fence": (i32, [vp, u32]),
        "mxl_ctx_h2d_bytes": (u64, [vp]),
        "mxl_ctx_d2h_bytes": (u64, [vp]),
        "mxl_ctx_stream": (vp, [vp]),
        "mxl_ctx_sample_rate": (u32, [vp]),
        "mxl_ctx_samples_per_tick": (u32, [vp]),
        "mxl_ctx_launch_count": (u64, [vp]),
        "mxl_ctx_timer_begin": (i32, [vp]),
        "mxl_ctx_timer_end": (i32, [vp]),
        "mxl_ctx_timer_elapsed_ms": (i32, [vp, C.POINTER(C.c_float)]),
        "mxl_ctx_flush_l2": (i32, [vp]),
        "mxl_ctx_device_memory": (i32, [vp, C.POINTER(u64), C.POINTER(u64)]),
        "mxl_ctx_set_kernel_timing": (i32, [vp, i32]),
        "mxl_ctx_kernel_times": (i32, [vp, C.POINTER(KernelTime), u32]),
        "mxl_ctx_fused_profile": (i32, [vp, u32, vp, u32]),
        "mxl_ctx_bind_host_to_gpu_node": (i32, [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
        "mxl_last_error": (C.c_char_p, []),
        "mxl_version": (C.c_char_p, []),
        "mxl_db_to_linear": (dbl, [dbl]),
        "mxl_host_alloc": (vp, [C.c_size_t]),
        "mxl_host_free": (i32, [vp]),
        "mxl_line_alloc": (vp, [vp, i32, u64]),
        "mxl_line_free": (i32, [vp]),
        "mxl_line_type_of": (i32, [vp]),
        "mxl_line_frames": (u64, [vp]),
        "mxl_line_len": (u64, [vp]),
        "mxl_line_device_ptr": (vp, [vp]),
        "mxl_line_zero": (i32, [vp]),
        "mxl_line_upload": (i32, [vp, vp, u64]),
        "mxl_line_download": (i32, [vp, vp, u64]),
        "mxl_line_upload_async": (i32, [vp, vp, u64]),
        "mxl_line_download_async": (i32, [vp, vp, u64]),
        "mxl_frame_layout_yuv420p": (i32, [u32, u32, C.POINTER(FrameLayout)]),
        "mxl_unify_picture_settings": (i32, [u32, u32, u32, u32, C.POINTER(u32), C.POINTER(u32)]),
        "mxl_scale_geometry_yuv420p": (i32, [u32, u32, u32, u32, C.POINTER(ScaleGeometry)]),
        "mxl_fader_to_u8": (C.c_uint8, [dbl]),
        "mxl_frame_alloc": (vp, [vp, u32, u32]),
        "mxl_frame_blank": (vp, [vp, u32, u32]),
        "mxl_frame_retain": (vp, [vp]),
        "mxl_frame_release": (i32, [vp]),
        "mxl_frame_get_layout": (i32, [vp, C.POINTER(FrameLayout)]),
        "mxl_frame_device_ptr": (vp, [vp]),
        "mxl_frame_upload": (i32, [vp, C.POINTER(vp), C.POINTER(u32)]),
        "mxl_frame_download": (i32, [vp, C.POINTER(vp), C.POINTER(u32)]),
        "mxl_frame_upload_raw": (i32, [vp, vp, u64]),
        "mxl_frame_download_raw": (i32, [vp, vp, u64]),
        "mxl_frame_upload_raw_async": (i32, [vp, vp, u64]),
        "mxl_frame_download_raw_async": (i32, [vp, vp, u64]),
        "mxl_frames_alloc_batch": (i32, [vp, u32, u32, u32, C.POINTER(vp)]),
        "mxl_frames_upload_raw_async": (i32, [C.POINTER(vp), u32, vp, u64]),
        "mxl_frames_download_raw_async": (i32, [C.POINTER(vp), u32, vp, u64]),
        "mxl_video_line_alloc": (vp, [vp, u32]),
        "mxl_video_line_set": (i32, [vp, u32, vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
        "mxl_video_line_get": (vp, [vp, u32]),
        "mxl_video_line_clear": (i32, [vp]),
        "mxl_module_create": (vp, [vp, i32, vp]),
        "mxl_module_destroy": (None, [vp]),
        "mxl_module_kind_of": (i32, [vp]),
        "mxl_module_update": (i32, [vp, i32, vp]),
        "mxl_module_params": (i32, [vp, vp]),
        "mxl_mixer_params_get": (i32, [vp, C.POINTER(MixerChannelParams), u32]),
        "mxl_module_n_inputs": (u32, [vp]),
        "mxl_module_n_outputs": (u32, [vp]),
        "mxl_module_input_type": (i32, [vp, u32]),
        "mxl_module_output_type": (i32, [vp, u32]),
        "mxl_module_input_label": (C.c_char_p, [vp, u32]),
        "mxl_module_output_label": (C.c_char_p, [vp, u32]),
        "mxl_module_run_tick": (i32, [vp, u64, C.POINTER(vp), u32, C.POINTER(vp), u32]),
        "mxl_module_run_tick_host": (i32, [vp, u64, C.POINTER(HostRef), u32, C.POINTER(HostRef), u32]),
        "mxl_stream_input_write_audio": (i32, [vp, u64, C.c_int64, C.c_int64, vp, u64]),
        "mxl_stream_input_write_video": (i32, [vp, u64, C.c_int64, C.c_int64, vp, C.c_int64, C.c_int64]),
        "mxl_stream_input_pending": (i32, [vp, C.POINTER(u32), C.POINTER(u32)]),
        "mxl_monitor_recv_audio": (i32, [vp, C.POINTER(AudioFragment), vp, u32]),
        "mxl_monitor_recv_video": (i32, [vp, C.POINTER(VideoJob)]),
        "mxl_stream_output_set_live": (i32, [vp, i32]),
        "mxl_output_device_read": (C.c_int64, [vp, vp, u64]),
        "mxl_output_device_clip": (i32, [vp, C.POINTER(C.c_int32)]),
        "mxl_video_line_get_timing": (i32, [vp, u32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
        "mxl_eq_three_state": (i32, [vp, C.POINTER(dbl)]),
        "mxl_envelope_state": (i32, [vp, C.POINTER(C.c_int32), C.POINTER(u64), C.POINTER(dbl)]),
        "mxl_meter_read": (i32, [vp, u32, C.POINTER(C.c_float), C.POINTER(dbl), C.POINTER(C.c_int32)]),
        "mxl_meter_download": (i32, [vp, vp, u32]),
        "mxl_meter_download_async": (i32, [vp, vp, u32]),
        "mxl_plotter_read": (i32, [vp, vp, vp, u32]),
        "mxl_source_set_line": (i32, [vp, vp]),
        "mxl_pcm_sink_download": (i32, [vp, vp, u64]),
        "mxl_pcm_unpack_i16": (i32, [vp, vp, u64, vp]),
        "mxl_pcm_pack_i16": (i32, [vp, vp, vp, u64]),
        "mxl_pcm_unpack_i16_async": (i32, [vp, vp, u64, vp]),
        "mxl_pcm_pack_i16_async": (i32, [vp, vp, vp, u64]),
        "mxl_frame_to_rgba": (i32, [vp, vp]),
        "mxl_frame_scale": (vp, [vp, u32, u32]),
        "mxl_frames_scale": (i32, [vp, C.POINTER(vp), C.POINTER(vp), u32, u32, u32]),
        "mxl_rgba_alloc": (vp, [vp, u32, u32, u32]),
        "mxl_rgba_free": (i32, [vp]),
        "mxl_rgba_device_ptr": (vp, [vp, u32]),
        "mxl_rgba_download": (i32, [vp, u32, u32, vp]),
        "mxl_rgba_download_async": (i32, [vp, u32, u32, vp]),
        "mxl_video_compose_rgba": (i32, [vp, C.POINTER(vp), C.POINTER(vp), u32, dbl, vp, u32]),
        "mxl_frames_to_rgba": (i32, [vp, C.POINTER(vp), u32, vp, u32]),
        "mxl_rgba_upload": (i32, [vp, u32, u32, vp]),
        "mxl_rgba_to_frames": (i32, [vp, vp, u32, u32, C.POINTER(vp)]),
        "mxl_resampler_create": (vp, [vp, u32, u32, u32]),
        "mxl_resampler_destroy": (None, [vp]),
        "mxl_resampler_reset": (i32, [vp]),
        "mxl_resampler_output_frames": (u64, [vp, u64]),
        "mxl_resampler_push_i16": (C.c_int64, [vp, vp, u64, vp]),
        "mxl_resampler_push_line": (C.c_int64, [vp, vp, vp]),
        "mxl_comm_unique_id": (i32, [vp]),
        "mxl_ctx_comm_init": (i32, [vp, vp, i32, i32]),
        "mxl_ctx_comm_destroy": (i32, [vp]),
        "mxl_line_broadcast": (i32, [vp, i32]),
        "mxl_frame_broadcast": (i32, [vp, i32]),
        "mxl_graph_create": (vp, [vp]),
        "mxl_graph_destroy": (None, [vp]),
        "mxl_graph_add_module": (i32, [vp, vp]),
        "mxl_graph_remove_module": (i32, [vp, i32]),
        "mxl_graph_module": (vp, [vp, i32]),
        "mxl_graph_connect": (i32, [vp, i32, u32, i32, u32]),
        "mxl_graph_disconnect": (i32, [vp, i32, u32]),
        "mxl_graph_plan": (i32, [vp, C.POINTER(C.c_int), u32]),
        "mxl_graph_run_ticks": (i32, [vp, u64, u32]),
        "mxl_graph_output": (vp, [vp, i32, u32]),
        "mxl_graph_pin_output": (i32, [vp, i32, u32]),
        "mxl_graph_set_fusion": (i32, [vp, i32]),
        "mxl_graph_set_profiling": (i32, [vp, i32]),
        "mxl_graph_set_stream_split": (i32, [vp, i32]),
        "mxl_graph_stage_count": (i32, [vp]),
        "mxl_graph_stage_info": (i32, [vp, u32, C.POINTER(StageInfo)]),
        "mxl_graph_performance": (i32, [vp, C.POINTER(PerfAccount), u32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    # optional newer entry points are typed lazily by their wrappers
    _lib = L
    return L


def last_error():
    return lib().mxl_last_error().decode("utf-8", "replace")


def check(status):
    if status < 0:
        raise MxlError(status, last_error())
    return status


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def db_to_linear(db):
    return lib().mxl_db_to_linear(float(db))


def fader_to_u8(fader):
    return int(lib().mxl_fader_to_u8(float(fader)))


def frame_layout(width, height):
    lay = FrameLayout()
    check(lib().mxl_frame_layout_yuv420p(width, height, C.byref(lay)))
    return lay


def unify_picture_settings(aw, ah, bw, bh):
    w, h = C.c_uint32(), C.c_uint32()
    check(lib().mxl_unify_picture_settings(aw, ah, bw, bh, C.byref(w), C.byref(h)))
    return w.value, h.value


COMM_ID_BYTES = 128


def comm_unique_id():
    """ncclUniqueId made by rank 0 and handed to every rank's Context.comm_init (any side channel)."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    check(lib().mxl_comm_unique_id(buf))
    return bytes(buf)


def scale_geometry(in_w, in_h, out_w, out_h):
    g = ScaleGeometry()
    check(lib().mxl_scale_geometry_yuv420p(in_w, in_h, out_w, out_h, C.byref(g)))
    return g.scaled_w, g.scaled_h, g.letterbox_x, g.letterbox_y


class PinnedBuffer:
    """Page-locked host memory (mxl_host_alloc) viewed as a numpy array."""

    def __init__(self, nbytes, dtype=np.uint8):
        self.nbytes = int(nbytes)
        self.ptr = lib().mxl_host_alloc(self.nbytes)
        if not self.ptr:
            raise MxlError(ERR_OOM, last_error())
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.uint8).view(dtype)

    def free(self):
        if self.ptr:
            self.array = None
            lib().mxl_host_free(self.ptr)
            self.ptr = None


class Context:
    """One engine thread's device context (src/engine.rs:78-93): one stream, no locking."""

    def __init__(self, device=0, sample_rate=48000, samples_per_tick=800, stream=None):
        L = lib()
        if stream is None:
            self.h = L.mxl_ctx_create(device, sample_rate, samples_per_tick)
        else:
            self.h = L.mxl_ctx_create_on_stream(device, sample_rate, samples_per_tick, stream)
        if not self.h:
            raise MxlError(ERR_NO_DEVICE if device >= 0 else ERR_INVALID, last_error())
        self.device = device
        self.sample_rate = sample_rate
        self.spt = samples_per_tick

    def close(self):
        if self.h:
            lib().mxl_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def synchronize(self):
        check(lib().mxl_ctx_synchronize(self.h))

    def trim_frame_pool(self, keep_bytes=0, new_cap_bytes=0):
        """Frees parked frame buffers down to keep_bytes; returns the bytes given back to the driver."""
        freed = C.c_uint64(0)
        check(lib().mxl_ctx_trim_frame_pool(self.h, keep_bytes, new_cap_bytes, C.byref(freed)))
        return freed.value

    def set_copy_overlap(self, on):
        check(lib().mxl_ctx_set_copy_overlap(self.h, 1 if on else 0))

    def download_fence(self, slot):
        check(lib().mxl_ctx_download_fence(self.h, slot))

    def wait_fence(self, slot):
        check(lib().mxl_ctx_wait_fence(self.h, slot))

    @property
    def h2d_bytes(self):
        return int(lib().mxl_ctx_h2d_bytes(self.h))

    @property
    def d2h_bytes(self):
        return int(lib().mxl_ctx_d2h_bytes(self.h))

    @property
    def launch_count(self):
        return int(lib().mxl_ctx_launch_count(self.h))

    def timer_begin(self):
        check(lib().mxl_ctx_timer_begin(self.h))

    def timer_end(self):
        check(lib().mxl_ctx_timer_end(self.h))

    def timer_elapsed_ms(self):
        ms = C.c_float()
        check(lib().mxl_ctx_timer_elapsed_ms(self.h, C.byref(ms)))
        return ms.value

    def bind_host_to_gpu_node(self):
        """(numa node or -1, cpus in the new affinity mask) -- mxl_ctx_bind_host_to_gpu_node."""
        node, cpus = C.c_int32(), C.c_int32()
        check(lib().mxl_ctx_bind_host_to_gpu_node(self.h, C.byref(node), C.byref(cpus)))
        return node.value, cpus.value

    def fused_profile(self, max_ctas, read=True):
        """mxl_ctx_fused_profile: (n_ctas, 8) SM-clock stamps of the last fused_voice_kernel launch; then re-arms for max_ctas."""
        buf = np.zeros((8192, 8), np.uint64)
        n = check(lib().mxl_ctx_fused_profile(self.h, max_ctas, _ptr(buf) if read else None, buf.shape[0]))
        return buf[:n]

    def flush_l2(self):
        check(lib().mxl_ctx_flush_l2(self.h))

    # ---- lines / frames ----
    def line(self, line_type, frames, data=None):
        ln = Line(self, line_type, frames)
        if data is not None:
            ln.upload(data)
        return ln

    def mono(self, data):
        data = np.ascontiguousarray(data, np.float32)
        return self.line(LINE_MONO, data.size, data)

    def stereo(self, data):
        data = np.ascontiguousarray(data, np.float32)
        return self.line(LINE_STEREO, data.size // 2, data)

    def video_line(self, ticks):
        return VideoLine(self, ticks)

    def frame(self, width, height, data=None, blank=False):
        fr = Frame(self, width, height, blank=blank)
        if data is not None:
            fr.upload_raw(data)
        return fr

    # ---- optional shared-source mode (NCCL broadcast from the ingest GPU) ----
    def comm_init(self, unique_id, rank, world):
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        check(lib().mxl_ctx_comm_init(self.h, buf, rank, world))

    def comm_destroy(self):
        check(lib().mxl_ctx_comm_destroy(self.h))

    def device_memory(self):
        f, t = C.c_uint64(), C.c_uint64()
        check(lib().mxl_ctx_device_memory(self.h, C.byref(f), C.byref(t)))
        return f.value, t.value

    def set_kernel_timing(self, enabled):
        check(lib().mxl_ctx_set_kernel_timing(self.h, 1 if enabled else 0))

    def kernel_times(self):
        """{kernel name: (launches, total ms)} of the launches since the last call (synchronises)."""
        buf = (KernelTime * 64)()
        n = check(lib().mxl_ctx_kernel_times(self.h, buf, 64))
        return {buf[i].name.decode(): (buf[i].launches, buf[i].total_ms) for i in range(n)}

    def frames_batch(self, width, height, n):
        """n frames adjacent in one device allocation (mxl_frames_alloc_batch)."""
        out = (C.c_void_p * n)()
        check(lib().mxl_frames_alloc_batch(self.h, width, height, n, out))
        return [Frame(self, handle=out[i]) for i in range(n)]

    def resampler(self, in_rate, out_rate, channels=2):
        return Resampler(self, in_rate, out_rate, channels)

    def rgba(self, width, height, n_pictures):
        return RgbaPictures(self, width, height, n_pictures)

    def frames_scale(self, frames, out_w, out_h):
        """Batched letterbox scaler: one launch for all planes of all frames (mxl_frames_scale)."""
        n = len(frames)
        src = (C.c_void_p * n)(*[f.h for f in frames])
        dst = (C.c_void_p * n)()
        check(lib().mxl_frames_scale(self.h, src, dst, n, out_w, out_h))
        return [Frame(self, handle=dst[i]) for i in range(n)]

    def compose_rgba(self, layers_a, layers_b, fader, out, first=0):
        """Crossfade + colour conversion in one pass (mxl_video_compose_rgba); None = missing layer."""
        n = len(layers_a)
        a = (C.c_void_p * n)(*[f.h if f is not None else None for f in layers_a])
        b = (C.c_void_p * n)(*[f.h if f is not None else None for f in layers_b])
        check(lib().mxl_video_compose_rgba(self.h, a, b, n, float(fader), out.h, first))

    def frames_to_rgba(self, frames, out, first=0):
        n = len(frames)
        a = (C.c_void_p * n)(*[f.h for f in frames])
        check(lib().mxl_frames_to_rgba(self.h, a, n, out.h, first))

    def module(self, kind, params=None):
        return Module(self, kind, params)

    def graph(self):
        return Graph(self)


class Line:
    """Output::{Mono,Stereo} of src/engine/io.rs:64-77, resident in HBM."""

    def __init__(self, ctx, line_type, frames, handle=None, owned=True):
        self.ctx = ctx
        self.owned = owned
        if handle is None:
            handle = lib().mxl_line_alloc(ctx.h, line_type, frames)
            if not handle:
                raise MxlError(ERR_INVALID, last_error())
        self.h = handle

    @property
    def type(self):
        return lib().mxl_line_type_of(self.h)

    @property
    def frames(self):
        return int(lib().mxl_line_frames(self.h))

    def __len__(self):
        return int(lib().mxl_line_len(self.h))

    def upload(self, data):
        data = np.ascontiguousarray(data, np.float32)
        check(lib().mxl_line_upload(self.h, _ptr(data), data.size))

    def download(self, n=None):
        n = len(self) if n is None else n
        out = np.empty(n, np.float32)
        check(lib().mxl_line_download(self.h, _ptr(out), n))
        return out

    def zero(self):
        check(lib().mxl_line_zero(self.h))

    def broadcast(self, root):
        """Shared-source mode: the root rank's samples replace this line on every rank (mxl_line_broadcast)."""
        check(lib().mxl_line_broadcast(self.h, root))

    def free(self):
        if self.h and self.owned:
            lib().mxl_line_free(self.h)
        self.h = None


class Resampler:
    """mxl_resampler: audio sample-rate converter of the ingest side (new; the reference has only the TODO)."""

    def __init__(self, ctx, in_rate, out_rate, channels=2):
        self.ctx, self.channels = ctx, channels
        self.h = lib().mxl_resampler_create(ctx.h, in_rate, out_rate, channels)
        if not self.h:
            raise MxlError(ERR_INVALID, last_error())
        self.out = ctx.line(LINE_STEREO if channels == 2 else LINE_MONO, 0)

    def push_i16(self, pcm):
        """interleaved int16 from the host -> the newly determined output frames (float32, interleaved)."""
        pcm = np.ascontiguousarray(pcm, np.int16)
        n = check(lib().mxl_resampler_push_i16(self.h, _ptr(pcm), pcm.size // self.channels, self.out.h))
        return self.out.download(n * self.channels) if n else np.empty(0, np.float32)

    def push_line(self, line):
        n = check(lib().mxl_resampler_push_line(self.h, line.h, self.out.h))
        return self.out.download(n * self.channels) if n else np.empty(0, np.float32)

    def output_frames(self, in_frames):
        return int(lib().mxl_resampler_output_frames(self.h, in_frames))

    def reset(self):
        check(lib().mxl_resampler_reset(self.h))

    def close(self):
        if self.h:
            lib().mxl_resampler_destroy(self.h)
            self.h = None
        self.out.free()


class RgbaPictures:
    """n device-resident RGBA8 pictures (mxl_rgba)."""

    def __init__(self, ctx, width, height, n):
        self.ctx, self.width, self.height, self.n = ctx, width, height, n
        self.h = lib().mxl_rgba_alloc(ctx.h, width, height, n)
        if not self.h:
            raise MxlError(ERR_INVALID, last_error())

    @property
    def picture_bytes(self):
        return self.width * self.height * 4

    def download(self, first=0, count=None):
        count = self.n - first if count is None else count
        out = np.empty(count * self.picture_bytes, np.uint8)
        check(lib().mxl_rgba_download(self.h, first, count, _ptr(out)))
        return out.reshape(count, self.picture_bytes)

    def upload(self, data, first=0):
        data = np.ascontiguousarray(data, np.uint8)
        count = data.size // self.picture_bytes
        check(lib().mxl_rgba_upload(self.h, first, count, _ptr(data)))

    def to_frames(self, frames, first=0):
        """RGBA8 -> yuv420p into existing frames of the pictures' size (mxl_rgba_to_frames), one launch."""
        n = len(frames)
        arr = (C.c_void_p * n)(*[f.h for f in frames])
        check(lib().mxl_rgba_to_frames(self.ctx.h, self.h, first, n, arr))

    def download_async(self, host_ptr, first=0, count=None):
        count = self.n - first if count is None else count
        check(lib().mxl_rgba_download_async(self.h, first, count, host_ptr))

    def free(self):
        if self.h:
            lib().mxl_rgba_free(self.h)
            self.h = None


class Frame:
    """AvFrame<Video> in yuv420p (codec/src/ffmpeg/frame.rs:76-138), refcounted, device-resident."""

    def __init__(self, ctx, width=0, height=0, blank=False, handle=None):
        self.ctx = ctx
        if handle is None:
            fn = lib().mxl_frame_blank if blank else lib().mxl_frame_alloc
            handle = fn(ctx.h, width, height)
            if not handle:
                raise MxlError(ERR_INVALID, last_error())
        self.h = handle

    @property
    def layout(self):
        lay = FrameLayout()
        check(lib().mxl_frame_get_layout(self.h, C.byref(lay)))
        return lay

    def upload_raw(self, data):
        data = np.ascontiguousarray(data, np.uint8)
        check(lib().mxl_frame_upload_raw(self.h, _ptr(data), data.size))

    def broadcast(self, root):
        check(lib().mxl_frame_broadcast(self.h, root))

    def download_raw(self):
        out = np.empty(self.layout.size, np.uint8)
        check(lib().mxl_frame_download_raw(self.h, _ptr(out), out.size))
        return out

    def upload_planes(self, planes, strides):
        arrs = [np.ascontiguousarray(p, np.uint8) for p in planes]
        ptrs = (C.c_void_p * 3)(*[a.ctypes.data for a in arrs])
        st = (C.c_uint32 * 3)(*strides)
        check(lib().mxl_frame_upload(self.h, ptrs, st))

    def download_planes(self, strides):
        lay = self.layout
        arrs = [np.zeros(strides[p] * lay.plane_h[p], np.uint8) for p in range(3)]
        ptrs = (C.c_void_p * 3)(*[a.ctypes.data for a in arrs])
        st = (C.c_uint32 * 3)(*strides)
        check(lib().mxl_frame_download(self.h, ptrs, st))
        return arrs

    def to_rgba(self):
        lay = self.layout
        out = np.empty(lay.width * lay.height * 4, np.uint8)
        check(lib().mxl_frame_to_rgba(self.h, _ptr(out)))
        return out

    def scale(self, out_w, out_h):
        h = lib().mxl_frame_scale(self.h, out_w, out_h)
        if not h:
            raise MxlError(ERR_CUDA, last_error())
        return Frame(self.ctx, handle=h)

    def retain(self):
        lib().mxl_frame_retain(self.h)
        return self

    def release(self):
        if self.h:
            lib().mxl_frame_release(self.h)
            self.h = None


class VideoLine:
    """Output::Video(Option<VideoFrame>) per tick (io.rs:8-17,64-69): one slot per tick of a call."""

    def __init__(self, ctx, ticks=0, handle=None, owned=True):
        self.ctx = ctx
        self.owned = owned
        if handle is None:
            handle = lib().mxl_video_line_alloc(ctx.h, ticks)
            if not handle:
                raise MxlError(ERR_INVALID, last_error())
        self.h = handle

    @property
    def slots(self):
        return int(lib().mxl_line_frames(self.h))

    def set(self, slot, frame, duration=(1, 60), offset=(0, 1)):
        check(lib().mxl_video_line_set(self.h, slot, frame.h if frame is not None else None,
                                       duration[0], duration[1], offset[0], offset[1]))

    def get(self, slot):
        h = lib().mxl_video_line_get(self.h, slot)
        return Frame(self.ctx, handle=h) if h else None      # borrowed: do not release

    def timing(self, slot):
        """((duration_num, duration_den), (offset_num, offset_den)) of the slot's VideoFrame."""
        d, o = (C.c_int64 * 2)(), (C.c_int64 * 2)()
        check(lib().mxl_video_line_get_timing(self.h, slot, d, o))
        return (d[0], d[1]), (o[0], o[1])

    def clear(self):
        check(lib().mxl_video_line_clear(self.h))

    def free(self):
        if self.h and self.owned:
            lib().mxl_line_free(self.h)
        self.h = None


def _mixer_params(channels):
    arr = (MixerChannelParams * max(len(channels), 1))()
    for i, ch in enumerate(channels):
        arr[i].gain_db, arr[i].fader, arr[i].cue = float(ch[0]), float(ch[1]), int(bool(ch[2]))
    return MixerParams(C.cast(arr, C.POINTER(MixerChannelParams)), len(channels)), arr


def make_params(kind, params):
    """params: a ctypes struct of the kind, a tuple of its fields, or for the mixer a list of
    (gain_db, fader, cue)."""
    if params is None:
        return None, None
    if isinstance(params, C.Structure):
        return params, None
    if kind == MOD_MIXER:
        return _mixer_params(list(params))
    return _PARAM_TYPES[kind](*params), None


class Module:
    """trait ModuleT (src/module/mod.rs:7-19) through DynModuleHostT (src/engine/module.rs:88-119)."""

    def __init__(self, ctx, kind, params=None, handle=None, owned=True):
        self.ctx = ctx
        self.kind = kind
        self.owned = owned
        if handle is None:
            p, keep = make_params(kind, params)
            handle = lib().mxl_module_create(ctx.h, kind, C.byref(p) if p is not None else None)
            if not handle:
                raise MxlError(ERR_PARAMS, last_error())
        self.h = handle

    def update(self, params, kind=None):
        kind = self.kind if kind is None else kind
        p, keep = make_params(kind, params)
        check(lib().mxl_module_update(self.h, kind, C.byref(p) if p is not None else None))

    def params(self):
        if self.kind == MOD_MIXER:
            n = check(lib().mxl_mixer_params_get(self.h, None, 0))
            arr = (MixerChannelParams * max(n, 1))()
            check(lib().mxl_mixer_params_get(self.h, arr, n))
            return [(arr[i].gain_db, arr[i].fader, bool(arr[i].cue)) for i in range(n)]
        t = _PARAM_TYPES.get(self.kind)
        if t is None:
            check(lib().mxl_module_params(self.h, None))
            return None
        p = t()
        check(lib().mxl_module_params(self.h, C.byref(p)))
        return p

    def inputs(self):
        L = lib()
        out = []
        for i in range(L.mxl_module_n_inputs(self.h)):
            lab = L.mxl_module_input_label(self.h, i)
            out.append((lab.decode() if lab is not None else None, L.mxl_module_input_type(self.h, i)))
        return out

    def outputs(self):
        L = lib()
        out = []
        for i in range(L.mxl_module_n_outputs(self.h)):
            lab = L.mxl_module_output_label(self.h, i)
            out.append((lab.decode() if lab is not None else None, L.mxl_module_output_type(self.h, i)))
        return out

    def run_tick(self, t, inputs, outputs):
        """inputs: lines or None (InputRef::Disconnected); outputs: lines.  Asynchronous."""
        ins = (C.c_void_p * max(len(inputs), 1))(*[(x.h if x is not None else None) for x in inputs])
        outs = (C.c_void_p * max(len(outputs), 1))(*[x.h for x in outputs])
        check(lib().mxl_module_run_tick(self.h, t, ins, len(inputs), outs, len(outputs)))

    def run_tick_host(self, t, inputs, outputs):
        """ModuleT::run_tick with host slices (mxl_module_run_tick_host), synchronous.
        inputs: per terminal None (InputRef::Disconnected), a float32 numpy array (Mono / Stereo slice) or, for a
        video terminal, ("video", Frame-or-None, (dur_num, dur_den), (off_num, off_den)).
        outputs: float32 numpy arrays written in place, or the string "video".  Returns, per video output, the
        Frame handed back (owned) or None."""
        kinds_in = [ty for _, ty in self.inputs()]
        kinds_out = [ty for _, ty in self.outputs()]
        ins = (HostRef * max(len(inputs), 1))()
        outs = (HostRef * max(len(outputs), 1))()
        for i, x in enumerate(inputs):
            ty = kinds_in[i] if i < len(kinds_in) else LINE_MONO
            ins[i].type = ty
            if x is None:
                continue
            ins[i].connected = 1
            if isinstance(x, tuple) and x[0] == "video":
                _, fr, dur, off = x
                ins[i].type = LINE_VIDEO
                ins[i].frame = fr.h if fr is not None else None
                ins[i].duration_num, ins[i].duration_den = dur
                ins[i].offset_num, ins[i].offset_den = off
            else:
                if isinstance(x, tuple):                      # (line_type, array): claim another line type
                    ins[i].type, x = x
                assert x.dtype == np.float32 and x.flags.c_contiguous
                ins[i].samples = x.ctypes.data
                ins[i].len = x.size
        for i, y in enumerate(outputs):
            ty = kinds_out[i] if i < len(kinds_out) else LINE_MONO
            outs[i].type = ty
            if isinstance(y, str) and y == "video":
                outs[i].type = LINE_VIDEO
            else:
                if isinstance(y, tuple):
                    outs[i].type, y = y
                assert y.dtype == np.float32 and y.flags.c_contiguous
                outs[i].samples = y.ctypes.data
                outs[i].len = y.size
        check(lib().mxl_module_run_tick_host(self.h, t, ins, len(inputs), outs, len(outputs)))
        frames = []
        for i, y in enumerate(outputs):
            if isinstance(y, str) and y == "video":
                frames.append(Frame(self.ctx, handle=outs[i].frame) if outs[i].frame else None)
        return frames

    # StreamInput: what the receiver pushes (SourceSend::write_audio / write_video, src/source.rs:156-190)
    def stream_write_audio(self, source_id, source_time, samples):
        """source_time: (num, den) seconds; samples: interleaved int16."""
        samples = np.ascontiguousarray(samples, np.int16)
        check(lib().mxl_stream_input_write_audio(self.h, source_id, source_time[0], source_time[1], _ptr(samples), samples.size))

    def stream_write_video(self, source_id, source_time, frame, duration):
        check(lib().mxl_stream_input_write_video(self.h, source_id, source_time[0], source_time[1], frame.h, duration[0], duration[1]))

    def stream_pending(self):
        a, v = C.c_uint32(), C.c_uint32()
        check(lib().mxl_stream_input_pending(self.h, C.byref(a), C.byref(v)))
        return a.value, v.value

    # Monitor: what the encoders are fed (AudioCtx::send_audio -> aac encode, VideoCtx::send_frame -> x264)
    def monitor_recv_audio(self):
        """None, or (decode_timestamp (num, den), duration (num, den), int16 fragment)."""
        info, buf = AudioFragment(), np.empty(2048, np.int16)
        if check(lib().mxl_monitor_recv_audio(self.h, C.byref(info), _ptr(buf), buf.size)) == 0:
            return None
        return (info.decode_num, info.decode_den), (info.duration_num, info.duration_den), buf[:info.n_samples]

    def monitor_recv_video(self):
        """None, or (pts, duration, time_base, blank, Frame owned by the caller)."""
        job = VideoJob()
        if check(lib().mxl_monitor_recv_video(self.h, C.byref(job))) == 0:
            return None
        return job.pts, job.duration, job.time_base, bool(job.blank), Frame(self.ctx, handle=job.frame)

    def output_device_read(self, cap):
        out = np.empty(cap, np.float32)
        n = check(lib().mxl_output_device_read(self.h, _ptr(out), cap))
        return out[:n]

    def output_device_clip(self):
        c = C.c_int32()
        check(lib().mxl_output_device_clip(self.h, C.byref(c)))
        return bool(c.value)

    def stream_output_set_live(self, live):
        check(lib().mxl_stream_output_set_live(self.h, 1 if live else 0))

    # kind-specific read-backs
    def eq_three_state(self):
        st = (C.c_double * 11)()
        check(lib().mxl_eq_three_state(self.h, st))
        return np.array(st[:], np.float64)

    def envelope_state(self):
        s, q, a = C.c_int32(), C.c_uint64(), C.c_double()
        check(lib().mxl_envelope_state(self.h, C.byref(s), C.byref(q), C.byref(a)))
        return s.value, q.value, a.value

    def meter_read(self, slot=0):
        pk, sq, cl = (C.c_float * 2)(), (C.c_double * 2)(), C.c_int32()
        check(lib().mxl_meter_read(self.h, slot, pk, sq, C.byref(cl)))
        return (pk[0], pk[1]), (sq[0], sq[1]), bool(cl.value)

    def meter_download(self, n_slots, out=None):
        out = np.empty(n_slots, METER_RECORD) if out is None else out
        n = check(lib().mxl_meter_download(self.h, _ptr(out), n_slots))
        return out[:n]

    def plotter_read(self, cap):
        left, right = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
        n = check(lib().mxl_plotter_read(self.h, _ptr(left), _ptr(right), cap))
        return left[:n], right[:n]

    def set_source_line(self, line):
        check(lib().mxl_source_set_line(self.h, line.h if line is not None else None))

    def pcm_download(self, n):
        out = np.empty(n, np.int16)
        check(lib().mxl_pcm_sink_download(self.h, _ptr(out), n))
        return out

    def destroy(self):
        if self.h and self.owned:
            lib().mxl_module_destroy(self.h)
        self.h = None


class Graph:
    """Workspace (src/engine/workspace.rs) + Engine::run_tick (src/engine.rs:400-510)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.h = lib().mxl_graph_create(ctx.h)
        if not self.h:
            raise MxlError(ERR_INVALID, last_error())
        self.modules = {}

    def add(self, kind, params=None):
        m = Module(self.ctx, kind, params)
        mid = check(lib().mxl_graph_add_module(self.h, m.h))
        m.owned = False           # the graph owns it now
        self.modules[mid] = m
        return mid

    def module(self, mid):
        return self.modules[mid]

    def remove(self, mid):
        check(lib().mxl_graph_remove_module(self.h, mid))
        self.modules.pop(mid).h = None

    def connect(self, in_module, in_index, out_module, out_index):
        """Workspace::connect(InputId, OutputId) (workspace.rs:97-114)."""
        check(lib().mxl_graph_connect(self.h, in_module, in_index, out_module, out_index))

    def disconnect(self, in_module, in_index):
        check(lib().mxl_graph_disconnect(self.h, in_module, in_index))

    def plan(self):
        arr = (C.c_int * 4096)()
        n = check(lib().mxl_graph_plan(self.h, arr, 4096))
        return list(arr[:n])

    def run_ticks(self, tick0, n_ticks):
        check(lib().mxl_graph_run_ticks(self.h, tick0, n_ticks))

    def output(self, mid, out_index):
        h = lib().mxl_graph_output(self.h, mid, out_index)
        if not h:
            err = last_error()
            if "fused voice group" in err:
                raise MxlError(ERR_INVALID, err)
            return None
        if lib().mxl_line_type_of(h) == LINE_VIDEO:
            return VideoLine(self.ctx, handle=h, owned=False)
        return Line(self.ctx, 0, 0, handle=h, owned=False)

    def pin_output(self, mid, out_index):
        """Keep a line inside a fused voice group observable (mxl_graph_pin_output)."""
        check(lib().mxl_graph_pin_output(self.h, mid, out_index))

    def set_fusion(self, on):
        check(lib().mxl_graph_set_fusion(self.h, 1 if on else 0))

    def set_profiling(self, on):
        check(lib().mxl_graph_set_profiling(self.h, 1 if on else 0))

    def set_stream_split(self, on):
        check(lib().mxl_graph_set_stream_split(self.h, 1 if on else 0))

    def stages(self):
        n = check(lib().mxl_graph_stage_count(self.h))
        out = []
        for i in range(n):
            s = StageInfo()
            check(lib().mxl_graph_stage_info(self.h, i, C.byref(s)))
            out.append(dict(kind=s.kind, n_modules=s.n_modules, n_launches=s.n_launches,
                            last_ms=s.last_ms, algorithmic_bytes=int(s.algorithmic_bytes), host_us=s.host_us))
        return out

    def performance(self):
        """PerformanceInfo.accounts: [(module_id or -1 for the Engine account, kind, last_us per tick, host_us per tick)]."""
        arr = (PerfAccount * (len(self.modules) + 1))()
        n = check(lib().mxl_graph_performance(self.h, arr, len(arr)))
        return [(a.module_id, a.kind, a.last_us, a.host_us) for a in arr[:n]]

    def destroy(self):
        if self.h:
            lib().mxl_graph_destroy(self.h)
            self.h = None
            for m in self.modules.values():
                m.h = None

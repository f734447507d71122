"""Synthetic workloads of BASELINE.json / SURVEY.md §8(d), as plain data.

A graph description is backend-neutral: `modules` is a list of (kind name, params) and
`connections` a list of (in_module, in_index, out_module, out_index) -- the arguments of
Workspace::connect (src/engine/workspace.rs:97-114).  `build_graph` instantiates it on the device
through the C ABI; tests and bench.py instantiate the same description on the CPU oracle.
"""
import numpy as np

from . import api

MASK64 = (1 << 64) - 1


def splitmix64(seed, n):
    """n outputs of splitmix64 seeded with `seed` (numpy uint64)."""
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & MASK64) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform_pm1(seed, n):
    """uniform [-1, 1) f32 with 24 random bits (exactly representable)."""
    r = (splitmix64(seed, n) >> np.uint64(40)).astype(np.float64)
    return (r / float(1 << 23) - 1.0).astype(np.float32)


def uniform_01(seed, n):
    r = (splitmix64(seed, n) >> np.uint64(40)).astype(np.float64)
    return (r / float(1 << 24)).astype(np.float32)


def random_bytes(seed, n):
    """low 8 bits of splitmix64."""
    return (splitmix64(seed, n) & np.uint64(0xFF)).astype(np.uint8)


KIND = {
    "Amplifier": api.MOD_AMPLIFIER, "Envelope": api.MOD_ENVELOPE, "EqThree": api.MOD_EQ_THREE,
    "FmSine": api.MOD_FM_SINE, "Mixer": api.MOD_MIXER, "Oscillator": api.MOD_OSCILLATOR,
    "Plotter": api.MOD_PLOTTER, "StereoPanner": api.MOD_STEREO_PANNER,
    "StereoSplitter": api.MOD_STEREO_SPLITTER, "Trigger": api.MOD_TRIGGER,
    "VideoMixer": api.MOD_VIDEO_MIXER, "Meter": api.MOD_METER, "SourceMono": api.MOD_SOURCE_MONO,
    "SourceStereo": api.MOD_SOURCE_STEREO, "SourceVideo": api.MOD_SOURCE_VIDEO, "PcmSink": api.MOD_PCM_SINK,
    "Monitor": api.MOD_MONITOR, "StreamInput": api.MOD_STREAM_INPUT, "StreamOutput": api.MOD_STREAM_OUTPUT,
    "OutputDevice": api.MOD_OUTPUT_DEVICE,
}
# stage kinds that are not module kinds (mxl_stage_info.kind)
STAGE_KIND = dict(KIND, FusedVoiceMix=api.STAGE_FUSED_VOICE_MIX)

_WAVES = [api.WAVE_SINE, api.WAVE_SAW, api.WAVE_TRIANGLE, api.WAVE_SQUARE]
_EQ_PATTERNS = [(-6.0, 0.0, 4.0), (0.0, 4.0, -6.0), (4.0, -6.0, 0.0), (4.0, 0.0, 4.0)]


class GraphDesc:
    def __init__(self, name):
        self.name = name
        self.modules = []          # (kind name, params)
        self.connections = []      # (in_module, in_index, out_module, out_index)
        self.taps = {}             # label -> (module, output index)

    def add(self, kind, params=None):
        self.modules.append((kind, params))
        return len(self.modules) - 1

    def connect(self, in_module, in_index, out_module, out_index):
        self.connections.append((in_module, in_index, out_module, out_index))

    def n_modules(self):
        return len(self.modules)


def osc_eq_pan_mixer(n_voices, with_meter=True, name=None):
    """n x [Oscillator(mono) -> EqThree -> StereoPanner(L=R)] -> Mixer(n) [-> Meter].

    n = 10 with the meter is BASELINE config 2 (32 modules, 464*S algorithmic bytes per tick);
    n = 5 without it is the audio half of config 4 (16 modules, 236*S)."""
    d = GraphDesc(name or "osc_eq_pan_mixer%d" % n_voices)
    pans = []
    for k in range(n_voices):
        osc = d.add("Oscillator", (55.0 * 2.0 ** (k / 3.0), _WAVES[k % 4], 0))
        eq = d.add("EqThree", _EQ_PATTERNS[k % 4])
        pan = d.add("StereoPanner")
        d.connect(eq, 0, osc, 0)
        d.connect(pan, 0, eq, 0)
        d.connect(pan, 1, eq, 0)
        pans.append(pan)
    chans = []
    for k in range(n_voices):
        frac = k / (n_voices - 1) if n_voices > 1 else 1.0
        chans.append((-24.0 + 30.0 * frac, 0.1 + 0.9 * frac, k % 2 == 0))
    mixer = d.add("Mixer", chans)
    for k, pan in enumerate(pans):
        d.connect(mixer, k, pan, 0)
    d.taps["master"] = (mixer, 0)
    d.taps["cue"] = (mixer, 1)
    if with_meter:
        meter = d.add("Meter")
        d.connect(meter, 0, mixer, 0)
        d.taps["meter"] = (meter, None)
    return d


def config2_graph():
    d = osc_eq_pan_mixer(10, True, "config2_32mod")
    assert d.n_modules() == 32
    return d


def config4_audio_graph():
    d = osc_eq_pan_mixer(5, False, "config4_16mod")
    assert d.n_modules() == 16
    return d


def config1_graph():
    """4 stereo sources -> Mixer(4) -> Amplifier(control = mono source)."""
    d = GraphDesc("config1_mixer_amp")
    srcs = [d.add("SourceStereo") for _ in range(4)]
    ctl = d.add("SourceMono")
    gains = [0.0, -6.0, 3.0, -12.0]
    faders = [1.0, 0.8, 0.5, 0.25]
    cues = [False, True, False, True]
    mixer = d.add("Mixer", list(zip(gains, faders, cues)))
    amp = d.add("Amplifier", (0.9, 0.5))
    for k, s in enumerate(srcs):
        d.connect(mixer, k, s, 0)
    d.connect(amp, 0, mixer, 0)
    d.connect(amp, 1, ctl, 0)
    d.taps["out"] = (amp, 0)
    d.taps["cue"] = (mixer, 1)
    d.sources = {srcs[k]: ("stereo", k + 1) for k in range(4)}
    d.sources[ctl] = ("mono01", 5)
    return d


def algorithmic_bytes_per_tick(desc, spt):
    """API-level line bytes (inputs read + outputs written) per tick, SURVEY.md §8(d)."""
    conn = {(c[0], c[1]) for c in desc.connections}
    total = 0
    for mid, (kind, params) in enumerate(desc.modules):
        def ins(i, w):
            return w * spt if (mid, i) in conn else 0
        if kind == "Oscillator":
            total += 12 * spt
        elif kind == "EqThree" or kind == "Envelope":
            total += ins(0, 4) + 4 * spt
        elif kind == "StereoPanner":
            total += ins(0, 4) + ins(1, 4) + 8 * spt
        elif kind == "StereoSplitter":
            total += ins(0, 8) + 8 * spt
        elif kind == "Mixer":
            total += sum(ins(i, 8) for i in range(len(params))) + 16 * spt
        elif kind == "Amplifier":
            total += ins(0, 8) + ins(1, 4) + 8 * spt
        elif kind == "FmSine":
            total += ins(0, 4) + 8 * spt
        elif kind == "Trigger":
            total += 4 * spt
        elif kind in ("Meter", "PcmSink"):
            total += ins(0, 8) + (4 * spt if kind == "PcmSink" else 0)
    return total


def build_graph(ctx, desc):
    """Instantiates a GraphDesc on the device.  Returns (Graph, [module ids])."""
    g = ctx.graph()
    ids = [g.add(KIND[kind], params) for kind, params in desc.modules]
    for im, ii, om, oi in desc.connections:
        g.connect(ids[im], ii, ids[om], oi)
    return g, ids


FRAME_W, FRAME_H = 1920, 1080
FRAME_BYTES = 3110400
CROSSFADE_BYTES_PER_FRAME = 3 * FRAME_BYTES       # two layers read + one frame written

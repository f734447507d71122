//! `impl ModuleT` (src/module/mod.rs:7-19) for every module of the tick hot path, integration level (a) of
//! INTEGRATION.md: `run_tick` hands the engine's own host slices to `mxl_module_run_tick_host`.  UNCOMPILED (no rustc
//! in the image this repository is built in); written against the reference's crates at d73346d.
//!
//! Each type keeps the reference's name, `Params`, `Indication`, terminals and labels, so `enumerate_modules!` can point
//! at these instead of the CPU modules and nothing else in the engine changes.
use mixlab_protocol::{AmplifierParams, EnvelopeParams, EqThreeParams, FmSineParams, GateState, LineType, MixerParams,
                      OscillatorParams, PlotterIndication, Terminal, VideoMixerParams, Waveform, VIDEO_MIXER_CHANNELS};

use crate::engine::{InputRef, ModuleCtx, OutputRef};
use crate::host_refs;
use crate::module::ModuleT;
use crate::{check, gpu_ctx, last_error, sys};

/// Owns one `mxl_module`; dropped with the Rust module (the reference drops a module on DeleteModule, engine.rs:321-352).
pub struct Raw(pub *mut sys::mxl_module);
impl Drop for Raw {
    fn drop(&mut self) { unsafe { sys::mxl_module_destroy(self.0) } }
}

fn create_raw(kind: sys::mxl_module_kind, pod: *const std::os::raw::c_void) -> Raw {
    let raw = unsafe { sys::mxl_module_create(gpu_ctx(), kind, pod) };
    assert!(!raw.is_null(), "{}", last_error());
    Raw(raw)
}

/// `run_tick(t, &[InputRef], &mut [OutputRef])` through the host-slice entry point.  `in_types` = the terminals' line
/// types (a Disconnected input still names its terminal).  Returns the `mxl_host_ref`s of the outputs (video frames).
fn run_host(raw: &Raw, t: u64, inputs: &[InputRef], in_types: &[i32], outputs: &mut [OutputRef]) -> Vec<sys::mxl_host_ref> {
    let ins: Vec<sys::mxl_host_ref> = inputs.iter().zip(in_types).map(|(i, ty)| host_refs::input(i, *ty)).collect();
    let mut outs: Vec<sys::mxl_host_ref> = outputs.iter_mut().map(host_refs::output).collect();
    check(unsafe { sys::mxl_module_run_tick_host(raw.0, t, ins.as_ptr(), ins.len() as u32, outs.as_mut_ptr(), outs.len() as u32) });
    outs                                                         // audio slices are complete when the call returns
}

fn line(ty: LineType) -> i32 {
    match ty { LineType::Mono => sys::MXL_LINE_MONO, LineType::Stereo => sys::MXL_LINE_STEREO, LineType::Video => sys::MXL_LINE_VIDEO }
}

/// Modules whose parameters are a plain struct of f64: one macro instead of five copies.
macro_rules! pod_module {
    ($name:ident, $params:ty, $kind:expr, $pod:ident, |$p:ident| $to_pod:expr, inputs: $ins:expr, outputs: $outs:expr) => {
        pub struct $name { raw: Raw, params: $params, inputs: Vec<Terminal>, outputs: Vec<Terminal> }

        impl ModuleT for $name {
            type Params = $params;
            type Indication = ();
            type Event = ();

            fn create(params: Self::Params, _: ModuleCtx<Self>) -> (Self, ()) {
                let pod: sys::$pod = { let $p = &params; $to_pod };
                let raw = create_raw($kind, &pod as *const _ as *const _);
                ($name { raw, params, inputs: $ins, outputs: $outs }, ())
            }
            fn params(&self) -> Self::Params { self.params.clone() }
            fn update(&mut self, new_params: Self::Params) -> Option<()> {
                let pod: sys::$pod = { let $p = &new_params; $to_pod };
                check(unsafe { sys::mxl_module_update(self.raw.0, $kind, &pod as *const _ as *const _) });
                self.params = new_params;
                None
            }
            fn run_tick(&mut self, t: u64, inputs: &[InputRef], outputs: &mut [OutputRef]) -> Option<()> {
                let types: Vec<i32> = self.inputs.iter().map(|term| line(term.line_type())).collect();
                run_host(&self.raw, t, inputs, &types, outputs);
                None
            }
            fn inputs(&self) -> &[Terminal] { &self.inputs }
            fn outputs(&self) -> &[Terminal] { &self.outputs }
        }
    };
}

// src/module/amplifier.rs:21-27 -- Stereo "Input", Mono "Control" -> Stereo
pod_module!(Amplifier, AmplifierParams, sys::MXL_MOD_AMPLIFIER, mxl_amplifier_params,
    |p| sys::mxl_amplifier_params { amplitude: p.amplitude, mod_depth: p.mod_depth },
    inputs: vec![LineType::Stereo.labeled("Input"), LineType::Mono.labeled("Control")],
    outputs: vec![LineType::Stereo.unlabeled()]);

// src/module/envelope.rs:72-80 -- Mono gate -> Mono
pod_module!(Envelope, EnvelopeParams, sys::MXL_MOD_ENVELOPE, mxl_envelope_params,
    |p| sys::mxl_envelope_params { attack_ms: p.attack_ms, decay_ms: p.decay_ms, sustain_amplitude: p.sustain_amplitude, release_ms: p.release_ms },
    inputs: vec![LineType::Mono.unlabeled()],
    outputs: vec![LineType::Mono.unlabeled()]);

// src/module/eq_three.rs:33-56 -- Mono -> Mono; the pole state lives in the device module and survives update()
pod_module!(EqThree, EqThreeParams, sys::MXL_MOD_EQ_THREE, mxl_eq_three_params,
    |p| sys::mxl_eq_three_params { gain_lo_db: p.gain_lo.0, gain_mid_db: p.gain_mid.0, gain_hi_db: p.gain_hi.0 },
    inputs: vec![LineType::Mono.unlabeled()],
    outputs: vec![LineType::Mono.unlabeled()]);

// src/module/fm_sine.rs:18-30 -- Mono -> Stereo
pod_module!(FmSine, FmSineParams, sys::MXL_MOD_FM_SINE, mxl_fm_sine_params,
    |p| sys::mxl_fm_sine_params { freq_lo: p.freq_lo, freq_hi: p.freq_hi },
    inputs: vec![LineType::Mono.unlabeled()],
    outputs: vec![LineType::Stereo.unlabeled()]);

// src/module/oscillator.rs:40-58 -- -> Mono "Mono", Stereo "Stereo"; Waveform in declaration order (protocol lib.rs:233-241)
pod_module!(Oscillator, OscillatorParams, sys::MXL_MOD_OSCILLATOR, mxl_oscillator_params,
    |p| sys::mxl_oscillator_params { freq: p.freq, _pad: 0, waveform: match p.waveform {
        Waveform::On => sys::MXL_WAVE_ON, Waveform::Off => sys::MXL_WAVE_OFF, Waveform::Sine => sys::MXL_WAVE_SINE,
        Waveform::Square => sys::MXL_WAVE_SQUARE, Waveform::Triangle => sys::MXL_WAVE_TRIANGLE, Waveform::Saw => sys::MXL_WAVE_SAW } },
    inputs: vec![],
    outputs: vec![LineType::Mono.labeled("Mono"), LineType::Stereo.labeled("Stereo")]);

// src/module/trigger.rs:18-33 -- -> Mono
pod_module!(Trigger, GateState, sys::MXL_MOD_TRIGGER, mxl_trigger_params,
    |p| sys::mxl_trigger_params { gate: match p { GateState::Open => sys::MXL_GATE_OPEN, GateState::Closed => sys::MXL_GATE_CLOSED } },
    inputs: vec![],
    outputs: vec![LineType::Mono.unlabeled()]);

/// Modules with `Params = ()`.
macro_rules! unit_module {
    ($name:ident, $kind:expr, inputs: $ins:expr, outputs: $outs:expr) => {
        pub struct $name { raw: Raw, inputs: Vec<Terminal>, outputs: Vec<Terminal> }

        impl ModuleT for $name {
            type Params = ();
            type Indication = ();
            type Event = ();

            fn create(_: (), _: ModuleCtx<Self>) -> (Self, ()) {
                ($name { raw: create_raw($kind, std::ptr::null()), inputs: $ins, outputs: $outs }, ())
            }
            fn params(&self) {}
            fn update(&mut self, _: ()) -> Option<()> { None }
            fn run_tick(&mut self, t: u64, inputs: &[InputRef], outputs: &mut [OutputRef]) -> Option<()> {
                let types: Vec<i32> = self.inputs.iter().map(|term| line(term.line_type())).collect();
                run_host(&self.raw, t, inputs, &types, outputs);
                None
            }
            fn inputs(&self) -> &[Terminal] { &self.inputs }
            fn outputs(&self) -> &[Terminal] { &self.outputs }
        }
    };
}

// src/module/stereo_panner.rs:15-28 / stereo_splitter.rs:15-31
unit_module!(StereoPanner, sys::MXL_MOD_STEREO_PANNER,
    inputs: vec![LineType::Mono.labeled("L"), LineType::Mono.labeled("R")], outputs: vec![LineType::Stereo.unlabeled()]);
unit_module!(StereoSplitter, sys::MXL_MOD_STEREO_SPLITTER,
    inputs: vec![LineType::Stereo.unlabeled()], outputs: vec![LineType::Mono.labeled("L"), LineType::Mono.labeled("R")]);

/// src/module/mixer.rs -- N x Stereo "1".."N" -> Stereo "Master", Stereo "Cue".  `update` re-creates the terminals when
/// the channel count changes, as the reference does (mixer.rs:40-44).
pub struct Mixer { raw: Raw, params: MixerParams, inputs: Vec<Terminal>, outputs: Vec<Terminal> }

fn mixer_pods(p: &MixerParams) -> Vec<sys::mxl_mixer_channel_params> {
    p.channels.iter().map(|c| sys::mxl_mixer_channel_params { gain_db: c.gain.0, fader: c.fader, cue: c.cue as i32, _pad: 0 }).collect()
}
fn mixer_terminals(n: usize) -> Vec<Terminal> {
    (0..n).map(|i| LineType::Stereo.labeled(&(i + 1).to_string())).collect()      // mixer.rs:23-25
}

impl ModuleT for Mixer {
    type Params = MixerParams;
    type Indication = ();
    type Event = ();

    fn create(params: MixerParams, _: ModuleCtx<Self>) -> (Self, ()) {
        let chans = mixer_pods(&params);
        let pod = sys::mxl_mixer_params { channels: chans.as_ptr(), n_channels: chans.len() as u32 };
        let raw = create_raw(sys::MXL_MOD_MIXER, &pod as *const _ as *const _);
        let n = params.channels.len();
        (Mixer { raw, params, inputs: mixer_terminals(n),
                 outputs: vec![LineType::Stereo.labeled("Master"), LineType::Stereo.labeled("Cue")] }, ())
    }
    fn params(&self) -> MixerParams { self.params.clone() }
    fn update(&mut self, new_params: MixerParams) -> Option<()> {
        let chans = mixer_pods(&new_params);
        let pod = sys::mxl_mixer_params { channels: chans.as_ptr(), n_channels: chans.len() as u32 };
        check(unsafe { sys::mxl_module_update(self.raw.0, sys::MXL_MOD_MIXER, &pod as *const _ as *const _) });
        self.inputs = mixer_terminals(new_params.channels.len());
        self.params = new_params;
        None
    }
    fn run_tick(&mut self, t: u64, inputs: &[InputRef], outputs: &mut [OutputRef]) -> Option<()> {
        let types = vec![sys::MXL_LINE_STEREO; inputs.len()];
        run_host(&self.raw, t, inputs, &types, outputs);
        None
    }
    fn inputs(&self) -> &[Terminal] { &self.inputs }
    fn outputs(&self) -> &[Terminal] { &self.outputs }
}

/// src/module/plotter.rs:37-56 -- Stereo -> indication every 6th tick (de-interleaved tap).
pub struct Plotter { raw: Raw, inputs: Vec<Terminal> }

impl ModuleT for Plotter {
    type Params = ();
    type Indication = PlotterIndication;
    type Event = ();

    fn create(_: (), _: ModuleCtx<Self>) -> (Self, PlotterIndication) {
        (Plotter { raw: create_raw(sys::MXL_MOD_PLOTTER, std::ptr::null()), inputs: vec![LineType::Stereo.unlabeled()] },
         PlotterIndication { inputs: vec![vec![], vec![]] })
    }
    fn params(&self) {}
    fn update(&mut self, _: ()) -> Option<PlotterIndication> { None }
    fn run_tick(&mut self, t: u64, inputs: &[InputRef], outputs: &mut [OutputRef]) -> Option<PlotterIndication> {
        run_host(&self.raw, t, inputs, &[sys::MXL_LINE_STEREO], outputs);
        let n = crate::SAMPLES_PER_TICK as usize;
        let (mut left, mut right) = (vec![0f32; n], vec![0f32; n]);
        let got = check(unsafe { sys::mxl_plotter_read(self.raw.0, left.as_mut_ptr(), right.as_mut_ptr(), n as u32) }) as usize;
        if got == 0 { return None; }                             // not a 6th tick, or input disconnected (plotter.rs:40)
        left.truncate(got);
        right.truncate(got);
        Some(PlotterIndication { inputs: vec![left, right] })
    }
    fn inputs(&self) -> &[Terminal] { &self.inputs }
    fn outputs(&self) -> &[Terminal] { &[] }
}

/// src/module/video_mixer.rs -- 4 x Video "1".."4" -> Video "Output", "A", "B".
pub struct VideoMixer { raw: Raw, params: VideoMixerParams, inputs: Vec<Terminal>, outputs: Vec<Terminal> }

fn video_mixer_pod(p: &VideoMixerParams) -> sys::mxl_video_mixer_params {
    // Option<usize> -> -1 for None (protocol lib.rs:405-420)
    sys::mxl_video_mixer_params { a: p.a.map(|x| x as i32).unwrap_or(-1), b: p.b.map(|x| x as i32).unwrap_or(-1), fader: p.fader }
}

impl ModuleT for VideoMixer {
    type Params = VideoMixerParams;
    type Indication = ();
    type Event = ();

    fn create(params: VideoMixerParams, _: ModuleCtx<Self>) -> (Self, ()) {
        let pod = video_mixer_pod(&params);
        let raw = create_raw(sys::MXL_MOD_VIDEO_MIXER, &pod as *const _ as *const _);
        let inputs = (0..VIDEO_MIXER_CHANNELS).map(|i| LineType::Video.labeled(&(i + 1).to_string())).collect();   // video_mixer.rs:29-31
        let outputs = vec![LineType::Video.labeled("Output"), LineType::Video.labeled("A"), LineType::Video.labeled("B")];
        (VideoMixer { raw, params, inputs, outputs }, ())
    }
    fn params(&self) -> VideoMixerParams { self.params.clone() }
    fn update(&mut self, new_params: VideoMixerParams) -> Option<()> {
        let pod = video_mixer_pod(&new_params);
        check(unsafe { sys::mxl_module_update(self.raw.0, sys::MXL_MOD_VIDEO_MIXER, &pod as *const _ as *const _) });
        self.params = new_params;
        None
    }
    fn run_tick(&mut self, t: u64, inputs: &[InputRef], outputs: &mut [OutputRef]) -> Option<()> {
        let types = vec![sys::MXL_LINE_VIDEO; inputs.len()];
        let outs = run_host(&self.raw, t, inputs, &types, outputs);
        // a video output comes back as a device frame the caller owns one reference of (NULL = None)
        for (o, r) in outputs.iter_mut().zip(outs.iter()) {
            if let OutputRef::Video(slot) = o {
                **slot = if r.frame.is_null() { None } else { Some(video::host_frame(r)) };
            }
        }
        None
    }
    fn inputs(&self) -> &[Terminal] { &self.inputs }
    fn outputs(&self) -> &[Terminal] { &self.outputs }
}

/// Pictures across the boundary.
pub mod video {
    use super::*;
    use crate::engine::VideoFrame;
    use mixlab_util::time::MediaDuration;

    /// The device copy of a decoded picture: uploaded once per `Arc<AvFrame>` and cached in a side table keyed by the
    /// frame's data pointer (the reference clones frames by refcount, codec/src/ffmpeg/frame.rs:351-361, so the same
    /// picture reaches several ticks and modules).  Entries leave with the last reference (a `Drop` hook on the cache's
    /// weak handle); elided here.
    pub fn device_frame(vf: &VideoFrame) -> *mut sys::mxl_frame {
        let pic = vf.data.decoded.picture_settings();
        let frame = unsafe { sys::mxl_frame_alloc(gpu_ctx(), pic.width as u32, pic.height as u32) };
        assert!(!frame.is_null(), "{}", last_error());
        let data = vf.data.decoded.frame_data();                 // planes + linesizes (frame.rs:188-197)
        let planes = [data.planes[0].as_ptr(), data.planes[1].as_ptr(), data.planes[2].as_ptr()];
        let strides = [data.strides[0] as u32, data.strides[1] as u32, data.strides[2] as u32];
        check(unsafe { sys::mxl_frame_upload(frame, planes.as_ptr(), strides.as_ptr()) });
        frame
    }

    /// A composited picture back into an `AvFrame` for the sinks that stay on the host (Monitor / StreamOutput encoders).
    pub fn host_frame(r: &sys::mxl_host_ref) -> VideoFrame {
        let mut lay: sys::mxl_frame_layout = unsafe { std::mem::zeroed() };
        check(unsafe { sys::mxl_frame_get_layout(r.frame, &mut lay) });
        let mut av = mixlab_codec::ffmpeg::AvFrame::blank(&mixlab_codec::ffmpeg::PictureSettings::yuv420p(lay.width as usize, lay.height as usize));
        {
            let mut data = av.frame_data_mut();
            let planes = [data.planes[0].as_mut_ptr(), data.planes[1].as_mut_ptr(), data.planes[2].as_mut_ptr()];
            let strides = [data.strides[0] as u32, data.strides[1] as u32, data.strides[2] as u32];
            check(unsafe { sys::mxl_frame_download(r.frame, planes.as_ptr(), strides.as_ptr()) });
        }
        unsafe { sys::mxl_frame_release(r.frame) };
        VideoFrame {
            data: std::sync::Arc::new(crate::video::Frame { decoded: av, duration_hint: MediaDuration::new(r.duration_num, r.duration_den) }),
            tick_offset: MediaDuration::new(r.offset_num, r.offset_den),
        }
    }
}

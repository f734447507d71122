//! Integration level (c) of INTEGRATION.md: `Engine::run_tick` (src/engine.rs:400-510) behind `mxl_graph_run_ticks`.
//! UNCOMPILED (no rustc in the image this repository is built in).
//!
//! The workspace (src/engine/workspace.rs) stays the source of truth for the UI; every edit `client_update` applies
//! (engine.rs:277-398) is mirrored into the device graph on the same thread, between ticks -- so, as in the reference, a
//! module's `update` and `run_tick` never race.  Per tick the engine thread then makes ONE call; sub-graphs such as
//! Oscillator -> EqThree -> StereoPanner -> Mixer -> Meter run as two launches instead of one per module.
use std::collections::HashMap;

use mixlab_protocol::{InputId, ModuleId, ModuleParams, OutputId};

use crate::{check, gpu_ctx, last_error, sys};

pub struct GpuGraph {
    raw: *mut sys::mxl_graph,
    ids: HashMap<ModuleId, i32>,             // workspace ModuleId -> device graph module id
}

impl GpuGraph {
    pub fn new() -> GpuGraph {
        let raw = unsafe { sys::mxl_graph_create(gpu_ctx()) };
        assert!(!raw.is_null(), "{}", last_error());
        GpuGraph { raw, ids: HashMap::new() }
    }

    /// ClientMessage::CreateModule (engine.rs:283-305): `module` was created with mxl_module_create by the params mapping
    /// of modules.rs; the graph takes ownership.
    pub fn create_module(&mut self, id: ModuleId, module: *mut sys::mxl_module) {
        let dev_id = check(unsafe { sys::mxl_graph_add_module(self.raw, module) });
        self.ids.insert(id, dev_id);
    }

    /// ClientMessage::UpdateModuleParams (engine.rs:307-319)
    pub fn update_module(&mut self, id: ModuleId, kind: sys::mxl_module_kind, pod: *const std::os::raw::c_void) {
        let m = unsafe { sys::mxl_graph_module(self.raw, self.ids[&id]) };
        check(unsafe { sys::mxl_module_update(m, kind, pod) });
    }

    /// ClientMessage::DeleteModule (engine.rs:321-352): connections touching the module go with it.
    pub fn delete_module(&mut self, id: ModuleId) {
        if let Some(dev_id) = self.ids.remove(&id) {
            check(unsafe { sys::mxl_graph_remove_module(self.raw, dev_id) });
        }
    }

    /// ClientMessage::CreateConnection -> Workspace::connect (workspace.rs:97-114): the same three errors.
    pub fn connect(&mut self, input: InputId, output: OutputId) -> Result<(), sys::mxl_status> {
        let st = unsafe { sys::mxl_graph_connect(self.raw, self.ids[&input.module_id()], input.index() as u32,
                                                 self.ids[&output.module_id()], output.index() as u32) };
        if st < 0 { Err(st) } else { Ok(()) }       // MXL_ERR_NO_INPUT / _NO_OUTPUT / _TYPE_MISMATCH = ConnectError::*
    }

    /// ClientMessage::DeleteConnection -> Workspace::disconnect (workspace.rs:116-118)
    pub fn disconnect(&mut self, input: InputId) {
        check(unsafe { sys::mxl_graph_disconnect(self.raw, self.ids[&input.module_id()], input.index() as u32) });
    }

    /// Engine::run_tick: `t = tick * SAMPLES_PER_TICK` is formed inside (engine.rs:490).  `n_ticks` > 1 when the caller
    /// tolerates the latency (an offline render); the live engine passes 1.
    pub fn run_ticks(&mut self, tick: u64, n_ticks: u32) {
        check(unsafe { sys::mxl_graph_run_ticks(self.raw, tick, n_ticks) });
    }

    /// What a sink that stays on the host reads after the tick (OutputDevice's cpal ring, the encoders' feeds have their
    /// own entry points: mxl_output_device_read, mxl_monitor_recv_audio / _video).
    pub fn download(&mut self, output: OutputId, dst: &mut [f32]) {
        let line = unsafe { sys::mxl_graph_output(self.raw, self.ids[&output.module_id()], output.index() as u32) };
        assert!(!line.is_null(), "{}", last_error());
        check(unsafe { sys::mxl_line_download(line, dst.as_mut_ptr(), dst.len() as u64) });
    }

    /// EngineStat::report (src/engine/timing.rs:45-60): µs per tick per module, Engine account first (module_id = -1).
    pub fn performance(&mut self) -> Vec<sys::mxl_perf_account> {
        let mut acc = vec![sys::mxl_perf_account { module_id: 0, kind: 0, last_us: 0.0, host_us: 0.0 }; self.ids.len() + 1];
        let n = check(unsafe { sys::mxl_graph_performance(self.raw, acc.as_mut_ptr(), acc.len() as u32) }) as usize;
        acc.truncate(n);
        acc
    }
}

impl Drop for GpuGraph {
    fn drop(&mut self) { unsafe { sys::mxl_graph_destroy(self.raw) } }
}

/// ModuleParams (protocol/src/lib.rs:188-207) -> (kind, POD) for module creation at workspace load (workspace.rs:28-33).
pub fn kind_of(params: &ModuleParams) -> sys::mxl_module_kind {
    match params {
        ModuleParams::Amplifier(_) => sys::MXL_MOD_AMPLIFIER,
        ModuleParams::Envelope(_) => sys::MXL_MOD_ENVELOPE,
        ModuleParams::EqThree(_) => sys::MXL_MOD_EQ_THREE,
        ModuleParams::FmSine(_) => sys::MXL_MOD_FM_SINE,
        ModuleParams::Mixer(_) => sys::MXL_MOD_MIXER,
        ModuleParams::Monitor => sys::MXL_MOD_MONITOR,
        ModuleParams::Oscillator(_) => sys::MXL_MOD_OSCILLATOR,
        ModuleParams::OutputDevice(_) => sys::MXL_MOD_OUTPUT_DEVICE,
        ModuleParams::Plotter(_) => sys::MXL_MOD_PLOTTER,
        ModuleParams::StereoPanner(_) => sys::MXL_MOD_STEREO_PANNER,
        ModuleParams::StereoSplitter(_) => sys::MXL_MOD_STEREO_SPLITTER,
        ModuleParams::StreamInput(_) => sys::MXL_MOD_STREAM_INPUT,
        ModuleParams::StreamOutput(_) => sys::MXL_MOD_STREAM_OUTPUT,
        ModuleParams::Trigger(_) => sys::MXL_MOD_TRIGGER,
        ModuleParams::VideoMixer(_) => sys::MXL_MOD_VIDEO_MIXER,
        ModuleParams::MediaSource(_) => sys::MXL_MOD_MEDIA_SOURCE,      // refused by mxl_module_create: stays on the host
    }
}

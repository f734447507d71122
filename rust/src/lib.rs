//! mixlab-b200: the B200 back end of Mixlab's tick hot path behind `trait ModuleT`.
//! UNCOMPILED in this repository (no rustc in the image) -- see ../README.md.
//!
//! Crate-internal names follow the reference: `crate::engine::{InputRef, OutputRef, VideoFrame, ModuleCtx}` and
//! `crate::module::ModuleT` are the reference's own items (src/engine/io.rs, src/engine/module.rs, src/module/mod.rs);
//! when this crate is vendored into the reference tree they are imported from there (`use mixlab::...`).
pub mod sys;
pub mod modules;
pub mod engine_graph;

use std::cell::Cell;
use std::ffi::CStr;

/// SAMPLE_RATE / SAMPLES_PER_TICK of the reference are compile-time constants (src/engine.rs:52-55); the back end
/// takes them at context creation.
pub const SAMPLE_RATE: u32 = 44100;
pub const SAMPLES_PER_TICK: u32 = 735;

thread_local! {
    /// One context per engine thread (src/engine.rs:78-93: a single thread owns every module, `&mut self`).
    static CTX: Cell<*mut sys::mxl_ctx> = Cell::new(std::ptr::null_mut());
}

/// The engine thread's context, created on first use on GPU 0 (a session -> GPU map would pick the device here).
pub fn gpu_ctx() -> *mut sys::mxl_ctx {
    CTX.with(|c| {
        if c.get().is_null() {
            let ctx = unsafe { sys::mxl_ctx_create(0, SAMPLE_RATE, SAMPLES_PER_TICK) };
            assert!(!ctx.is_null(), "{}", last_error());
            c.set(ctx);
        }
        c.get()
    })
}

pub fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::mxl_last_error()) }.to_string_lossy().into_owned()
}

/// A negative status is re-raised as the panic the CPU module would have hit (line-type mismatch: io.rs:40-41,49-50;
/// params-variant mismatch: module.rs:108).  Nothing unwinds across the C boundary itself.
pub fn check(status: i32) -> i32 {
    if status < 0 {
        panic!("{}", last_error());
    }
    status
}

/// `InputRef` / `OutputRef` (src/engine/io.rs:19-34,79-98) as the `mxl_host_ref` the C ABI takes.
pub mod host_refs {
    use super::sys;
    use crate::engine::{InputRef, OutputRef};

    pub fn zeroed() -> sys::mxl_host_ref {
        sys::mxl_host_ref { type_: 0, connected: 0, samples: std::ptr::null_mut(), len: 0, frame: std::ptr::null_mut(),
                            duration_num: 0, duration_den: 1, offset_num: 0, offset_den: 1 }
    }

    /// `line_type` = the terminal's type: a Disconnected input still says which terminal it sits on.
    pub fn input(i: &InputRef, line_type: i32) -> sys::mxl_host_ref {
        let mut r = zeroed();
        r.type_ = line_type;
        match i {
            InputRef::Disconnected => {}
            InputRef::Mono(s) => { r.type_ = sys::MXL_LINE_MONO; r.connected = 1; r.samples = s.as_ptr() as *mut f32; r.len = s.len() as u64; }
            InputRef::Stereo(s) => { r.type_ = sys::MXL_LINE_STEREO; r.connected = 1; r.samples = s.as_ptr() as *mut f32; r.len = s.len() as u64; }
            InputRef::Video(v) => {
                r.type_ = sys::MXL_LINE_VIDEO;
                r.connected = 1;
                if let Some(vf) = v {
                    // the device copy of the picture is made once per decoded frame and cached next to the Arc<AvFrame>
                    r.frame = crate::modules::video::device_frame(vf);
                    let d = vf.data.duration_hint;      // MediaDuration = Rational64 (util/src/time.rs:77-78)
                    r.duration_num = *d.0.numer(); r.duration_den = *d.0.denom();
                    let o = vf.tick_offset;
                    r.offset_num = *o.0.numer(); r.offset_den = *o.0.denom();
                }
            }
        }
        r
    }

    pub fn output(o: &mut OutputRef) -> sys::mxl_host_ref {
        let mut r = zeroed();
        r.connected = 1;
        match o {
            OutputRef::Mono(s) => { r.type_ = sys::MXL_LINE_MONO; r.samples = s.as_mut_ptr(); r.len = s.len() as u64; }
            OutputRef::Stereo(s) => { r.type_ = sys::MXL_LINE_STEREO; r.samples = s.as_mut_ptr(); r.len = s.len() as u64; }
            OutputRef::Video(_) => { r.type_ = sys::MXL_LINE_VIDEO; }          // written by the call
        }
        r
    }
}

//! `GpuPipeline`: infur's `Scale -> Model -> ColorCode` section (infur/src/app.rs:107-153) as ONE `Processor` backed by
//! libinfur_b200.so (hand-written sm_100a kernels; no CPU fallback -- `GpuPipeline::new` fails without a B200).
//!
//! Drop-in: `ProcessingApp` (app.rs:53-62) holds one `GpuPipeline` in place of its `scale`, `model` and `decoder` fields;
//! `control` forwards `AppCmd::Scale` / `AppCmd::Model` (app.rs:91-105); `advance` becomes
//! `vid.advance -> gpu.advance -> GUIFrame { id, buffer, decoded_buffer }`.  Like the ORT session today (main.rs:38-40) the
//! pipeline must be created on the "Proc" thread: a handle has one owner thread.
//!
//! The trait below is a copy of the DEFINITION in infur/src/processing.rs:23-60 so that this crate builds on its own; inside
//! infur's workspace replace it by `use infur::processing::{Frame, Processor};`.
mod ffi;

use epaint::{Color32, ColorImage};
use std::{
    ffi::{CStr, CString},
    ptr,
};

/// infur/src/processing.rs:23-60
pub trait Processor {
    type Command;
    type ControlError;
    type Input;
    type Output;
    type ProcessResult;
    fn control(&mut self, cmd: Self::Command) -> Result<&mut Self, Self::ControlError>;
    fn advance(&mut self, input: &Self::Input, out: &mut Self::Output) -> Self::ProcessResult;
    fn is_dirty(&self) -> bool;
}

/// infur/src/processing.rs:9-18 with the image as a tight HWC B,G,R byte buffer (what `BgrImage` derefs to,
/// image-ext/src/image_bgr.rs:7-11).
pub struct Frame {
    pub id: u64,
    pub width: u32,
    pub height: u32,
    pub bgr: Vec<u8>,
}

/// The `AppCmd` subset the GPU path handles (app.rs:39-51).
pub enum GpuCmd {
    Scale(f32),
    Model(String),
}

#[derive(thiserror::Error, Debug)]
#[error("{msg}")]
pub struct GpuError {
    /// `INFUR_E_*` of include/infur_b200.h; maps 1:1 onto ValidScaleError / ScaleProcError / ModelCmdError / ModelProcError
    pub code: i32,
    pub msg: String,
}

/// What app.rs:132-149 builds per frame.
pub struct GpuFrame {
    pub id: u64,
    pub buffer: ColorImage,
    pub decoded_buffer: Option<ColorImage>,
}

pub struct GpuPipeline {
    h: *mut ffi::Handle,
    frame: Vec<Color32>,
    decoded: Vec<Color32>,
}

// The handle is used from the thread that owns the pipeline only (one owner thread, like the ORT session).
impl GpuPipeline {
    /// One GPU (`devices = [ordinal]`) or several GPUs of the box behind one handle (frames are routed by id, weights are
    /// broadcast with NCCL inside `control(GpuCmd::Model(..))`).
    pub fn new(devices: &[i32]) -> Result<Self, GpuError> {
        assert!(unsafe { ffi::infur_b200_abi_version() } == ffi::ABI_VERSION, "libinfur_b200.so ABI mismatch");
        let mut cfg: ffi::Config = unsafe { std::mem::zeroed() };
        unsafe { ffi::infur_b200_default_config(&mut cfg) };
        if devices.is_empty() || devices.len() > ffi::MAX_DEVICES {
            return Err(GpuError { code: ffi::E_INVALID_ARG, msg: "1..8 devices".into() });
        }
        cfg.device = devices[0];
        cfg.num_devices = devices.len() as i32;
        cfg.devices[..devices.len()].copy_from_slice(devices);
        let mut h = ptr::null_mut();
        let rc = unsafe { ffi::infur_b200_create(&cfg, &mut h) };
        if rc != ffi::OK {
            return Err(Self::err(ptr::null(), rc));
        }
        Ok(Self { h, frame: vec![], decoded: vec![] })
    }

    fn err(h: *const ffi::Handle, code: i32) -> GpuError {
        let msg = unsafe { CStr::from_ptr(ffi::infur_b200_last_error(h)) }.to_string_lossy().into_owned();
        GpuError { code, msg }
    }

    /// `Model::get_info` (predict_onnx.rs:341-345): "input\tdtype\tout,aux", None without a model.
    pub fn model_info(&self) -> Option<String> {
        self.text(ffi::infur_b200_model_info)
    }

    /// Class captions (README.md:77): one "index\tlabel\tr,g,b" line per class of the loaded model.
    pub fn class_legend(&self) -> Option<String> {
        self.text(ffi::infur_b200_class_legend)
    }

    fn text(&self, f: unsafe extern "C" fn(*const ffi::Handle, *mut std::os::raw::c_char, usize, *mut usize) -> i32) -> Option<String> {
        let mut need = 0usize;
        let rc = unsafe { f(self.h, ptr::null_mut(), 0, &mut need) };
        if rc == ffi::E_INVALID_ARG || need == 0 {
            return None;
        }
        let mut buf = vec![0u8; need];
        let rc = unsafe { f(self.h, buf.as_mut_ptr() as *mut _, need, &mut need) };
        (rc == ffi::OK).then(|| String::from_utf8_lossy(&buf[..need - 1]).into_owned())
    }
}

impl Drop for GpuPipeline {
    fn drop(&mut self) {
        unsafe { ffi::infur_b200_destroy(self.h) }
    }
}

fn color_image(size: [usize; 2], px: &[Color32]) -> ColorImage {
    ColorImage { size, pixels: px.to_vec() }
}

impl Processor for GpuPipeline {
    type Command = GpuCmd;
    type ControlError = GpuError;
    type Input = Option<Frame>;
    type Output = Option<GpuFrame>;
    type ProcessResult = Result<(), GpuError>;

    /// `Scale::control` (processing.rs:220-226) / `Model::control` (predict_onnx.rs:283-315); a failed load keeps the previous model.
    fn control(&mut self, cmd: GpuCmd) -> Result<&mut Self, GpuError> {
        let rc = match cmd {
            GpuCmd::Scale(f) => unsafe { ffi::infur_b200_scale_control(self.h, f) },
            GpuCmd::Model(p) => {
                let c = CString::new(p).map_err(|_| GpuError { code: ffi::E_INVALID_ARG, msg: "path contains NUL".into() })?;
                unsafe { ffi::infur_b200_model_load(self.h, c.as_ptr()) }
            }
        };
        if rc != ffi::OK {
            Err(Self::err(self.h, rc))
        } else {
            Ok(self)
        }
    }

    /// `scale.advance -> model.advance -> decoder.advance` + the display buffer (app.rs:109-149) in one GPU pass.
    fn advance(&mut self, inp: &Option<Frame>, out: &mut Self::Output) -> Self::ProcessResult {
        let Some(frame) = inp else { return Ok(()) };
        assert_eq!(frame.bgr.len(), frame.width as usize * frame.height as usize * 3);
        let mut o: ffi::Out = unsafe { std::mem::zeroed() };
        o.struct_size = std::mem::size_of::<ffi::Out>() as u32;
        // size query (no buffers), then the real call into buffers re-used across frames
        let rc = unsafe { ffi::infur_b200_advance(self.h, frame.bgr.as_ptr(), frame.width, frame.height, frame.id, &mut o) };
        if rc != ffi::OK {
            return Err(Self::err(self.h, rc));
        }
        let px = (o.out_w * o.out_h) as usize;
        self.frame.resize(px, Color32::BLACK); // Color32 is #[repr(C)] [u8; 4], premultiplied RGBA
        o.frame_rgba = self.frame.as_mut_ptr() as *mut u8;
        o.frame_rgba_cap = px * 4;
        if o.has_decoded != 0 {
            self.decoded.resize(px, Color32::BLACK);
            o.decoded_rgba = self.decoded.as_mut_ptr() as *mut u8;
            o.decoded_rgba_cap = px * 4;
        }
        let rc = unsafe { ffi::infur_b200_advance(self.h, frame.bgr.as_ptr(), frame.width, frame.height, frame.id, &mut o) };
        if rc != ffi::OK {
            return Err(Self::err(self.h, rc));
        }
        let size = [o.out_w as usize, o.out_h as usize]; // ColorImage.size = [w, h] (app.rs:140-144)
        *out = Some(GpuFrame {
            id: o.id,
            buffer: color_image(size, &self.frame),
            decoded_buffer: (o.has_decoded != 0).then(|| color_image(size, &self.decoded)), // None without a model, app.rs:127-129
        });
        Ok(())
    }

    fn is_dirty(&self) -> bool {
        unsafe { ffi::infur_b200_is_dirty(self.h) != 0 }
    }
}

/// Streaming use (configs 3-5): `submit` copies a frame into the pinned ring slot of GPU `(id - 1) % n`, `wait` returns results
/// in submission order.  Pointers inside the result stay valid until the second-next `wait`.
impl GpuPipeline {
    pub fn submit(&mut self, frame: &Frame) -> Result<u64, GpuError> {
        let mut t = 0u64;
        let rc = unsafe { ffi::infur_b200_submit(self.h, frame.bgr.as_ptr(), frame.width, frame.height, frame.id, &mut t) };
        if rc != ffi::OK {
            Err(Self::err(self.h, rc))
        } else {
            Ok(t)
        }
    }

    pub fn wait(&mut self, ticket: u64) -> Result<GpuFrame, GpuError> {
        let mut r: ffi::Result_ = unsafe { std::mem::zeroed() };
        let rc = unsafe { ffi::infur_b200_wait(self.h, ticket, &mut r) };
        if rc != ffi::OK {
            return Err(Self::err(self.h, rc));
        }
        let px = (r.out_w * r.out_h) as usize;
        let size = [r.out_w as usize, r.out_h as usize];
        let view = |p: *const u8| unsafe { std::slice::from_raw_parts(p as *const Color32, px) };
        Ok(GpuFrame {
            id: r.id,
            buffer: if r.frame_rgba.is_null() { ColorImage::new(size, Color32::BLACK) } else { color_image(size, view(r.frame_rgba)) },
            decoded_buffer: (r.has_decoded != 0 && !r.decoded_rgba.is_null()).then(|| color_image(size, view(r.decoded_rgba))),
        })
    }

    pub fn flush(&mut self) -> Result<(), GpuError> {
        let rc = unsafe { ffi::infur_b200_flush(self.h) };
        if rc != ffi::OK {
            Err(Self::err(self.h, rc))
        } else {
            Ok(())
        }
    }
}

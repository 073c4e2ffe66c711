//! Raw declarations of include/infur_b200.h (ABI version 2).  Field order and types mirror the C structs exactly.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::c_char;

pub const ABI_VERSION: i32 = 2;
pub const MAX_DEVICES: usize = 8;

pub const OK: i32 = 0;
pub const E_INVALID_ARG: i32 = 1;
pub const E_SCALE_NONPOSITIVE: i32 = 2; // ValidScaleError, processing.rs:159-168
pub const E_ZERO_SIZE_IN: i32 = 3; // ScaleProcError::ZeroSizeIn, processing.rs:203-204
pub const E_ZERO_SIZE_OUT: i32 = 4; // ScaleProcError::ZeroSizeOut, processing.rs:205-206
pub const E_MODEL_LOAD: i32 = 5; // ModelCmdError::OrtError, predict_onnx.rs:43-46
pub const E_MODEL_INPUT_FORMAT: i32 = 6; // ModelInputFormatError, predict_onnx.rs:50-54
pub const E_SHAPE: i32 = 7; // ModelProcError::ShapeError, predict_onnx.rs:35-36
pub const E_RUNTIME: i32 = 8; // ModelProcError::RuntimeError, predict_onnx.rs:37-38
pub const E_BUFFER_TOO_SMALL: i32 = 9;
pub const E_NO_DEVICE: i32 = 10;
pub const E_UNSUPPORTED: i32 = 11;
pub const E_TICKET: i32 = 12;
pub const E_STREAM_END: i32 = 13;

#[repr(C)]
pub struct Handle {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct Config {
    pub struct_size: u32,
    pub device: i32,
    pub max_batch: i32,
    pub ring_depth: i32,
    pub resize_mode: i32,
    pub compute_aux: i32,
    pub blend: i32,
    pub conv_impl: i32,
    pub use_cuda_graph: i32,
    pub autotune: i32,
    pub num_devices: i32,
    pub devices: [i32; MAX_DEVICES],
    pub frame_rgba: i32,
    pub confidence: i32,
}

#[repr(C)]
pub struct Out {
    pub struct_size: u32,
    pub scaled_bgr: *mut u8,
    pub scaled_bgr_cap: usize,
    pub frame_rgba: *mut u8,
    pub frame_rgba_cap: usize,
    pub class_map: *mut u8,
    pub class_map_cap: usize,
    pub decoded_rgba: *mut u8,
    pub decoded_rgba_cap: usize,
    pub blended_rgba: *mut u8,
    pub blended_rgba_cap: usize,
    pub logits_f32: *mut f32,
    pub logits_cap: usize,
    pub aux_logits_f32: *mut f32,
    pub aux_logits_cap: usize,
    pub out_w: u32,
    pub out_h: u32,
    pub num_classes: u32,
    pub has_decoded: i32,
    pub id: u64,
    pub required: [usize; 7],
}

#[repr(C)]
pub struct Result_ {
    pub ticket: u64,
    pub id: u64,
    pub out_w: u32,
    pub out_h: u32,
    pub num_classes: u32,
    pub has_decoded: i32,
    pub device: i32,
    pub class_map: *const u8,
    pub decoded_rgba: *const u8,
    pub blended_rgba: *const u8,
    pub frame_rgba: *const u8,
}

extern "C" {
    pub fn infur_b200_abi_version() -> i32;
    pub fn infur_b200_default_config(cfg: *mut Config);
    pub fn infur_b200_create(cfg: *const Config, out: *mut *mut Handle) -> i32;
    pub fn infur_b200_destroy(h: *mut Handle);
    pub fn infur_b200_last_error(h: *const Handle) -> *const c_char;
    pub fn infur_b200_scale_control(h: *mut Handle, factor: f32) -> i32;
    pub fn infur_b200_model_load(h: *mut Handle, utf8_path: *const c_char) -> i32;
    pub fn infur_b200_model_info(h: *const Handle, buf: *mut c_char, cap: usize, required: *mut usize) -> i32;
    pub fn infur_b200_class_legend(h: *const Handle, buf: *mut c_char, cap: usize, required: *mut usize) -> i32;
    pub fn infur_b200_is_dirty(h: *const Handle) -> i32;
    pub fn infur_b200_advance(h: *mut Handle, bgr: *const u8, w: u32, hgt: u32, id: u64, out: *mut Out) -> i32;
    pub fn infur_b200_submit(h: *mut Handle, bgr: *const u8, w: u32, hgt: u32, id: u64, ticket: *mut u64) -> i32;
    pub fn infur_b200_flush(h: *mut Handle) -> i32;
    pub fn infur_b200_wait(h: *mut Handle, ticket: u64, out: *mut Result_) -> i32;
    pub fn infur_b200_num_devices(h: *const Handle) -> i32;
}

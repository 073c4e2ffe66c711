/*
 * infur_b200 -- C ABI of the B200-native `Scale -> Model -> ColorCode` hot path of ahirner/infur.
 *
 * The reference has no FFI today: its plug-in surface is the Rust trait `Processor`
 * (infur/src/processing.rs:23-60: control / advance / is_dirty / generate) and the three
 * implementations sequenced by `ProcessingApp::advance` (infur/src/app.rs:107-153).  Every entry
 * point below names the reference item it replaces; INTEGRATION.md shows the Rust binding
 * (`impl Processor for GpuPipeline`) a maintainer would add.
 *
 * Conventions: C99, `int32_t` status (0 = OK; nothing throws or aborts across the boundary),
 * opaque handle, caller-owned buffers with explicit capacities, no callbacks.  One owner thread per
 * handle (the reference creates its ORT session on the "Proc" thread because it cannot be sent,
 * infur/src/main.rs:38-40).  A handle drives one CUDA device, or -- cfg.num_devices > 1 -- the GPUs of one box from that
 * single owner thread: one worker thread + stream set + pinned ring per GPU inside the library, the packed weights broadcast
 * with NCCL at model_load, frames / ring slots routed round-robin by id, results handed back in submission order
 * (DESIGN.md "Multi-GPU").  One process per GPU with one single-device handle each remains possible (torchrun).
 *
 * There is NO CPU fallback: without a CUDA device create() returns INFUR_E_NO_DEVICE.
 */
#ifndef INFUR_B200_H
#define INFUR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INFUR_B200_ABI_VERSION 2
#define INFUR_B200_MAX_DEVICES 8

typedef struct infur_b200_handle infur_b200_handle;

/* ---- status codes (mirror the reference's error enums) ---------------------------------- */
enum {
  INFUR_OK = 0,
  INFUR_E_INVALID_ARG = 1,
  INFUR_E_SCALE_NONPOSITIVE = 2,  /* ValidScaleError "Cannot scale by negative number" processing.rs:159-168 */
  INFUR_E_ZERO_SIZE_IN = 3,       /* ScaleProcError::ZeroSizeIn   processing.rs:203-204 */
  INFUR_E_ZERO_SIZE_OUT = 4,      /* ScaleProcError::ZeroSizeOut  processing.rs:205-206 */
  INFUR_E_MODEL_LOAD = 5,         /* ModelCmdError::OrtError      predict_onnx.rs:43-46 (unreadable / unparsable / unsupported file) */
  INFUR_E_MODEL_INPUT_FORMAT = 6, /* ModelInputFormatError::Infer predict_onnx.rs:50-54,223-265 */
  INFUR_E_SHAPE = 7,              /* ModelProcError::ShapeError   predict_onnx.rs:35-36 */
  INFUR_E_RUNTIME = 8,            /* ModelProcError::RuntimeError predict_onnx.rs:37-38 (here: CUDA errors) */
  INFUR_E_BUFFER_TOO_SMALL = 9,   /* C analogue of "stage re-allocates Out on size change" (processing.rs:260-268) */
  INFUR_E_NO_DEVICE = 10,
  INFUR_E_UNSUPPORTED = 11,
  INFUR_E_TICKET = 12,
  INFUR_E_STREAM_END = 13         /* read_exact hit EOF (ff-video/src/decoder.rs:156-162): the caller classifies it as FinishedNormally /
                                   * ExactReadError from the decoder process' exit status, exactly as the reference does */
};

enum { INFUR_RESIZE_NEAREST = 0 /* fr::ResizeAlg::Nearest, processing.rs:189 (the reference's only mode; the default) */,
       INFUR_RESIZE_BILINEAR = 1 /* opt-in extension (README.md:74 TODO): half-pixel bilinear, un-fused f32, round-half-up to u8 */ };
enum { INFUR_CONF_RAW = 0, INFUR_CONF_SOFTMAX = 1 };
enum { INFUR_CONV_TCGEN05 = 0, INFUR_CONV_VALIDATE = 1 /* slow CUDA-core kernel, validation only; never selected implicitly */,
       INFUR_CONV_TCGEN05_PAIR = 2 /* conv_test only: force the CTA-pair (cta_group::2) variant of the tcgen05 kernel */,
       INFUR_CONV_TCGEN05_HALO = 3, /* conv_test only: force the halo-patch variant (3x3 / stride 1 convolutions) */
       INFUR_CONV_TCGEN05_I8 = 4, /* conv_test only: the layer as an int8 plan runs it (u8 tensors, tcgen05.mma.kind::i8) */
       INFUR_CONV_TCGEN05_I8_PAIR = 5 /* conv_test only: the same through the CTA-pair kernel */,
       INFUR_CONV_TCGEN05_PAIR_DEEP = 6 /* conv_test only: CTA pair with eight epilogue chunk buffers (layers with a residual) */,
       INFUR_CONV_TCGEN05_I8_PAIR_DEEP = 7 };

typedef struct infur_b200_config {
  uint32_t struct_size;  /* sizeof(infur_b200_config) */
  int32_t device;        /* CUDA ordinal */
  int32_t max_batch;     /* frames per ring slot (default 8) */
  int32_t ring_depth;    /* pinned ring slots (default 3) */
  int32_t resize_mode;   /* INFUR_RESIZE_NEAREST (parity mode) or INFUR_RESIZE_BILINEAR */
  int32_t compute_aux;   /* also evaluate the `aux` head (the reference's caller discards it, app.rs:116) */
  int32_t blend;         /* also produce blended_rgba (new feature; gui.rs:324-329 "todo: blend somehow?") */
  int32_t conv_impl;     /* INFUR_CONV_TCGEN05 */
  int32_t use_cuda_graph;/* 1: the kernel sequence of a plan is captured once into a CUDA graph and replayed with one launch per
                          * step (host-buffer and ring entry points; default 1).  0: plain stream launches */
  int32_t autotune;      /* time the N-tile candidates of every convolution once per layer-shape class and keep the fastest
                          * (default 1; results are bit-identical for every choice).  Decisions are cached per handle and across
                          * plans, so a new Scale factor re-tunes nothing whose tile grid it does not change */
  int32_t num_devices;   /* 0 or 1: one GPU, `device`.  2..INFUR_B200_MAX_DEVICES: the GPUs devices[0..num_devices) of this box behind
                          * ONE handle: weights are packed on devices[0] and ncclBroadcast to the others inside model_load; frame id
                          * (1-based, ff-video/src/decoder.rs:163-164) goes to devices[(id - 1) % n], ring ticket t to devices[(t - 1) % n] */
  int32_t devices[INFUR_B200_MAX_DEVICES];
  int32_t frame_rgba;    /* ring / submit path: also return GUIFrame.buffer (app.rs:132-144), the scaled frame as (r,g,b,255); default 1 */
  int32_t confidence;    /* INFUR_CONF_RAW (the reference: raw logits as confidence, decode_predict.rs:67-77) or INFUR_CONF_SOFTMAX
                          * (README.md:76 TODO: alpha from the softmax probability of the winning class) */
} infur_b200_config;

/* Fills *cfg with the defaults of the three stages: factor 1.0, dirty, no model
 * (processing.rs:185-193, predict_onnx.rs:151-155). */
void infur_b200_default_config(infur_b200_config* cfg);

int32_t infur_b200_abi_version(void);

/* `Default::default()` of Scale + Model + ColorCode and the process-wide ENVIRONMENT
 * (predict_onnx.rs:19-30). */
int32_t infur_b200_create(const infur_b200_config* cfg, infur_b200_handle** out);
void infur_b200_destroy(infur_b200_handle* h);

/* thiserror Display strings (processing.rs:203-210, predict_onnx.rs:33-54); handle-local, valid until
 * the next call on the handle.  h == NULL returns the message of the last failed create(). */
const char* infur_b200_last_error(const infur_b200_handle* h);

/* ---- control ------------------------------------------------------------------------------ */

/* Scale::control (processing.rs:220-226): f <= 0 -> INFUR_E_SCALE_NONPOSITIVE, state unchanged;
 * else dirty = (f != old factor). NaN passes, as in the reference. */
int32_t infur_b200_scale_control(infur_b200_handle* h, float factor);

/* Model::control(ModelCmd::Load(path)) (predict_onnx.rs:283-315): "" unloads.  On any failure the
 * previously loaded model stays active (:289-308).  Accepted graphs: float FCN-style (Conv / Relu / Add / MaxPool / Resize,
 * computed in fp16 with f32 accumulation) and QOperator-quantised ones (QuantizeLinear / QLinearConv / QLinearAdd /
 * DequantizeLinear -- the operator set of fcn-resnet50-12-int8.onnx, the file the reference's tests load, :350-381), which
 * run natively in int8 (u8 x s8 -> s32 tensor-core MMA) when every tensor is u8 with zero-point-0 convolution inputs, else
 * with the integers carried exactly in fp16; both forms reproduce the integer arithmetic of those operators bit for bit. */
int32_t infur_b200_model_load(infur_b200_handle* h, const char* utf8_path);

/* Same, from an in-memory copy of the .onnx file (used after a rank-0 broadcast of the file). */
int32_t infur_b200_model_load_bytes(infur_b200_handle* h, const void* onnx, size_t size);

/* ModelInfo / Model::get_info (predict_onnx.rs:56-62,341-345).  Writes a NUL-terminated text
 * "input0_name\tinput0_dtype\tout0,out1,...".  Returns INFUR_E_INVALID_ARG when no model is loaded
 * (get_info() == None), INFUR_E_BUFFER_TOO_SMALL (with *required set) when cap is too small. */
int32_t infur_b200_model_info(const infur_b200_handle* h, char* buf, size_t cap, size_t* required);

/* Multi-GPU initialisation ("NCCL broadcast of weights at init only").
 * (a) One handle, several GPUs (cfg.num_devices > 1): model_load parses the file once, every device builds its plan structures,
 *     ONLY devices[0] packs and uploads the weight arena, and the arena reaches the other devices with ncclBroadcast over the
 *     library's own communicator (ncclCommInitAll at create; libnccl.so.2 is loaded at run time).  The load is atomic across
 *     devices: if any device fails, every device keeps its previous model (predict_onnx.rs:289-308).
 * (b) One process per GPU (torchrun): every rank parses the file for the graph structure, but only the root packs and uploads
 *     weights; the others pass INFUR_LOAD_SKIP_WEIGHTS (arena allocated, zero-filled), then receive the root's packed arena through
 *     model_weights_export -> the caller's broadcast -> model_weights_import.  Both copy device-to-device between the library's
 *     arena and a caller-owned DEVICE buffer of *bytes (single-device handles only). */
enum { INFUR_LOAD_DEFAULT = 0, INFUR_LOAD_SKIP_WEIGHTS = 1 };
int32_t infur_b200_model_load_opts(infur_b200_handle* h, const char* utf8_path, int32_t flags);
int32_t infur_b200_model_weights_size(const infur_b200_handle* h, size_t* bytes);
int32_t infur_b200_model_weights_export(infur_b200_handle* h, void* d_dst, size_t bytes);
int32_t infur_b200_model_weights_import(infur_b200_handle* h, const void* d_src, size_t bytes);
/* Diagnostics (multi-device handles): 64-bit FNV-1a checksum of device `index`'s weight arena, to check the broadcast. */
int32_t infur_b200_model_weights_checksum(infur_b200_handle* h, int32_t index, uint64_t* sum);
/* Number of GPUs behind the handle (1 for a single-device handle). */
int32_t infur_b200_num_devices(const infur_b200_handle* h);

/* Class-label captions (README.md:77 TODO "class-label captions"; the reference only colour-codes, decode_predict.rs:9-36).
 * Writes a NUL-terminated text, one line per class 0..K-1 of the loaded model: "index\tlabel\tr,g,b" with the palette colour
 * COLORS_PALETTE[index % 20] (decode_predict.rs:9-30,34).  Labels: the 21 Pascal-VOC names torchvision's fcn_resnet50 is
 * trained on when K == 21 ("__background__", "aeroplane", ... "tvmonitor"), else "class <index>".  No model: INFUR_E_INVALID_ARG. */
int32_t infur_b200_class_legend(const infur_b200_handle* h, char* buf, size_t cap, size_t* required);
/* Scale::is_dirty (processing.rs:228-230); Model and ColorCode are never dirty
 * (predict_onnx.rs:336-338, decode_predict.rs:81-83). */
int32_t infur_b200_is_dirty(const infur_b200_handle* h);

/* ---- advance ------------------------------------------------------------------------------- */

typedef struct infur_b200_out {
  uint32_t struct_size; /* sizeof(infur_b200_out) */
  /* caller-owned HOST buffers; NULL = output not wanted; *_cap in bytes */
  uint8_t* scaled_bgr;   size_t scaled_bgr_cap;   /* Scale output, [out_h][out_w][3] u8 */
  uint8_t* frame_rgba;   size_t frame_rgba_cap;   /* GUIFrame.buffer, app.rs:132-144: (r,g,b,255) */
  uint8_t* class_map;    size_t class_map_cap;    /* k_max per pixel, decode_predict.rs:67-77, u8 */
  uint8_t* decoded_rgba; size_t decoded_rgba_cap; /* GUIFrame.decoded_buffer: premultiplied Color32 */
  uint8_t* blended_rgba; size_t blended_rgba_cap; /* mask "over" frame (cfg.blend) */
  float* logits_f32;     size_t logits_cap;       /* Model::advance out[0]: [K][out_h][out_w] f32 (debug path) */
  float* aux_logits_f32; size_t aux_logits_cap;   /* Model::advance out[1] (cfg.compute_aux) */
  /* written by the library */
  uint32_t out_w, out_h;     /* (w as f32 * factor) as u32, processing.rs:253-254 */
  uint32_t num_classes;      /* K */
  int32_t has_decoded;       /* 0 when no model is loaded: decoded_img = None, app.rs:127-129 */
  uint64_t id;               /* Frame.id passes through, processing.rs:239,264 */
  size_t required[7];        /* bytes needed for each buffer above, in declaration order */
} infur_b200_out;

/* Scale::advance -> Model::advance -> ColorCode::advance as sequenced in app.rs:109-149.
 * Synchronous.  `bgr` is a caller-owned tight HWC B,G,R u8 image of w*h*3 bytes
 * (image-ext/src/image_bgr.rs:7-11), only read during the call.  Clears the dirty flag
 * (processing.rs:233).  Errors: INFUR_E_ZERO_SIZE_IN / _OUT, INFUR_E_BUFFER_TOO_SMALL (required[]
 * filled, nothing written), INFUR_E_RUNTIME.  A call with every buffer pointer NULL is a size query:
 * it fills out_w / out_h / num_classes / has_decoded / required[] and runs nothing. */
int32_t infur_b200_advance(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt, uint64_t id,
                           infur_b200_out* out);

/* Page-locked host memory for the buffers handed to infur_b200_advance / _advance_batch (frames in, results out).  Any host
 * pointer works; with pinned ones the copies are direct DMA transfers (6.2 MB in + 10.4 MB out for one 1080p frame: ~0.3 ms
 * instead of ~0.8 ms through the driver's staging buffers), which is most of what a single-frame call (configs[1]) can save. */
int32_t infur_b200_host_alloc(size_t bytes, void** out);
void infur_b200_host_free(void* p);

/* Batched synchronous variant: n frames of identical size, stored back to back. outs[i] as above. */
int32_t infur_b200_advance_batch(infur_b200_handle* h, const uint8_t* bgr, uint32_t n, uint32_t w, uint32_t hgt,
                                 const uint64_t* ids, infur_b200_out* outs);

/* ---- pinned ring: asynchronous batches (new; configs "batch=8, pinned ring buffer") ------- */

typedef struct infur_b200_slot {
  uint64_t ticket;
  uint32_t n, w, h;          /* frames in the slot, input size */
  uint32_t out_w, out_h;     /* valid after ring_wait */
  uint32_t num_classes;
  int32_t has_decoded;
  int32_t device;            /* CUDA ordinal the slot runs on: devices[(ticket - 1) % num_devices] */
  uint8_t* bgr_in;           /* PINNED host memory: the frame source writes n*w*h*3 bytes here */
  const uint8_t* class_map;  /* PINNED, [n][out_h][out_w]      valid after ring_wait until the slot is re-acquired */
  const uint8_t* decoded_rgba; /* PINNED, [n][out_h][out_w][4]   GUIFrame.decoded_buffer */
  const uint8_t* blended_rgba; /* PINNED or NULL (cfg.blend) */
  const uint8_t* frame_rgba;   /* PINNED or NULL (cfg.frame_rgba): GUIFrame.buffer, app.rs:132-144 */
} infur_b200_slot;

/* Next free slot sized for n frames of w x h; INFUR_E_TICKET when all ring_depth slots (per device) are in flight.
 * Tickets are 1-based and consecutive; with several devices ticket t lives on devices[(t - 1) % num_devices]. */
int32_t infur_b200_ring_acquire(infur_b200_handle* h, uint32_t n, uint32_t w, uint32_t hgt, infur_b200_slot* slot);
/* Frame source -> pinned memory without a staging copy (FFMpegDecoder::read_frame, ff-video/src/decoder.rs:150-165, for a
 * whole slot): reads up to the slot's n frames of w*h*3 bytes each from file descriptor `fd` (the rawvideo bgr24 pipe of
 * decoder.rs:51-75) straight into the slot's pinned input buffer, `read_exact` per frame (short reads and EINTR are retried).
 * *frames_read = complete frames now in the slot; the slot's frame count shrinks to it (a slot left with 0 frames is released).
 * Returns INFUR_OK when all n frames arrived, INFUR_E_STREAM_END when EOF came first (*partial_bytes = bytes of an incomplete
 * last frame, 0 = EOF exactly on a frame boundary), INFUR_E_RUNTIME on an I/O error.  Frame ids stay with the caller
 * (1-based counter, decoder.rs:163-164). */
int32_t infur_b200_ring_read(infur_b200_handle* h, uint64_t ticket, int32_t fd, uint32_t* frames_read, size_t* partial_bytes);

/* Enqueue H2D copy, the whole path (with the Scale factor current at this call; the slot's output buffers grow if a factor
 * raised since ring_acquire needs it), and the D2H copies of the slot; returns immediately.  On failure the slot is released. */
int32_t infur_b200_ring_submit(infur_b200_handle* h, uint64_t ticket);
/* Block until the slot's results are in pinned memory.  Tickets may be waited in any order; each device completes its own
 * tickets in submission order, so waiting in submission order never blocks longer than necessary. */
int32_t infur_b200_ring_wait(infur_b200_handle* h, uint64_t ticket, infur_b200_slot* slot);
/* Give an acquired (not yet submitted) slot back to the ring. */
int32_t infur_b200_ring_release(infur_b200_handle* h, uint64_t ticket);

/* ---- frame-level asynchronous API on top of the ring (what a "Proc" thread calls per frame) ------------------------------
 * submit: copies the frame (tight HWC B,G,R u8) into the open pinned slot of devices[(id - 1) % num_devices] (id = 0 is treated
 * as the submission counter); when that slot holds cfg.max_batch frames it is enqueued.  wait: results of one frame; frames may
 * be waited in any order, a frame whose slot is still open is flushed first.  Result pointers are PINNED host memory owned by
 * the library, valid until the second-next infur_b200_wait / infur_b200_flush call on the handle (the slot is recycled only
 * after every frame of it has been waited and one more wait has passed).  All frames between two flushes must share w x h. */
typedef struct infur_b200_result {
  uint64_t ticket, id;
  uint32_t out_w, out_h, num_classes;
  int32_t has_decoded;       /* 0 when no model is loaded: decoded_img = None, app.rs:127-129 */
  int32_t device;
  const uint8_t* class_map;    /* [out_h][out_w] */
  const uint8_t* decoded_rgba; /* [out_h][out_w][4] premultiplied (GUIFrame.decoded_buffer) */
  const uint8_t* blended_rgba; /* or NULL */
  const uint8_t* frame_rgba;   /* or NULL (GUIFrame.buffer) */
} infur_b200_result;
int32_t infur_b200_submit(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt, uint64_t id, uint64_t* ticket);
int32_t infur_b200_flush(infur_b200_handle* h);
int32_t infur_b200_wait(infur_b200_handle* h, uint64_t ticket, infur_b200_result* out);

/* ---- device-resident path (bench `value`: inputs already in HBM) --------------------------- */

/* d_bgr: device pointer to n tight BGR frames (n <= cfg.max_batch).  Outputs are caller-owned DEVICE buffers with capacities in
 * bytes: class map [n][out_h][out_w], decoded / blended RGBA [n][out_h][out_w][4].  The output size follows the current Scale
 * factor; a call with every buffer NULL is a size query (fills out_w / out_h / num_classes / has_decoded / required[], runs
 * nothing); a buffer that is too small returns INFUR_E_BUFFER_TOO_SMALL with required[] filled and nothing enqueued.  Without
 * a loaded model has_decoded = 0 and the buffers are left untouched (app.rs:127-129).  Enqueued on the handle's compute stream
 * (device 0 of a multi-device handle); `sync` != 0 waits for completion. */
typedef struct infur_b200_device_out {
  uint32_t struct_size;      /* sizeof(infur_b200_device_out) */
  uint8_t* d_class_map;    size_t class_map_cap;
  uint8_t* d_decoded_rgba; size_t decoded_rgba_cap;
  uint8_t* d_blended_rgba; size_t blended_rgba_cap;   /* NULL unless cfg.blend */
  /* written by the library */
  uint32_t out_w, out_h, num_classes;
  int32_t has_decoded;
  size_t required[3];        /* bytes needed for the three buffers above */
} infur_b200_device_out;
int32_t infur_b200_advance_device(infur_b200_handle* h, const uint8_t* d_bgr, uint32_t n, uint32_t w, uint32_t hgt,
                                  infur_b200_device_out* out, int32_t sync);

/* CUDA stream (cudaStream_t) the device-resident path launches on, for event timing by the caller. */
void* infur_b200_compute_stream(const infur_b200_handle* h);

/* Kernels this library has launched on the handle since create() (bench `gpu_launches`). */
uint64_t infur_b200_launch_count(const infur_b200_handle* h);

/* ---- single-stage entry points (each Processor on its own; used by the parity tests) ------- */

/* Scale::advance alone (processing.rs:232-281) with the handle's current factor: host in, host out.
 * All pointers NULL == advance(&None, ..): clears the dirty flag and returns OK (:233-237). */
int32_t infur_b200_scale_advance(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt,
                                 uint8_t* out_bgr, size_t out_cap, uint32_t* out_w, uint32_t* out_h);

/* Model::advance alone (predict_onnx.rs:317-334): the image is fed as is (no Scale), the outputs are the
 * network's `out` (and, with cfg.compute_aux, `aux`) tensors with the batch dimension stripped,
 * [K][hgt][w] f32 each -- the reference's Vec<ArrayD<f32>>.  Without a loaded model it returns OK with
 * *has_model = 0 and leaves the buffers untouched (:321-323).  Does not touch the Scale state.
 * Debug / parity path: it materialises full-resolution logits, which the fused path never does. */
int32_t infur_b200_model_advance(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt, float* logits_f32, size_t logits_cap,
                                 float* aux_logits_f32, size_t aux_cap, uint32_t* num_classes, int32_t* has_model);

/* Diagnostics: the `out` head's logits BEFORE the final Resize, [K][lh][lw] f32 (lh = hgt / 8, lw = w / 8 for FCN), of
 * one image fed as is (no Scale).  This is the tensor the fused post-kernel up-samples; the parity tests compare it with
 * the oracle directly (bit-exactly for quantised models).  Sizes are written even when the buffer is too small
 * (INFUR_E_BUFFER_TOO_SMALL); without a model *k = 0 and the call returns OK. */
int32_t infur_b200_model_lowres(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt, float* lowres, size_t cap_floats,
                                uint32_t* k, uint32_t* lw, uint32_t* lh);

/* Pre-processing of ImageSession::forward alone (predict_onnx.rs:103-137): [h][w][3] u8 BGR ->
 * [3][h][w] f32 RGB-normalised. */
int32_t infur_b200_preprocess(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt, float* out_nchw,
                              size_t out_cap_bytes);

/* ColorCode::advance alone (decode_predict.rs:53-79) on a [k][hgt][w] f32 confidence map; any k >= 1.
 * rgba: [hgt][w][4] premultiplied; class_map (may be NULL): [hgt][w] u8 (k_max & 0xff). */
int32_t infur_b200_color_code(infur_b200_handle* h, const float* hm, uint32_t k, uint32_t w, uint32_t hgt,
                              uint8_t* rgba, uint8_t* class_map);

/* Fused post-kernel alone: low-res logits [k][lh][lw] f32 -> bilinear half-pixel upsample to
 * out_w x out_h -> argmax/confidence colour (the network's final Resize, executed inside
 * session.run predict_onnx.rs:138, fused with decode_predict.rs:53-79).  frame_bgr (may be NULL) is the
 * [out_h][out_w][3] frame to blend onto when blended_rgba != NULL. */
int32_t infur_b200_upsample_color(infur_b200_handle* h, const float* lowres, uint32_t k, uint32_t lw, uint32_t lh,
                                  uint32_t out_w, uint32_t out_h, const uint8_t* frame_bgr, uint8_t* class_map,
                                  uint8_t* decoded_rgba, uint8_t* blended_rgba, float* logits_f32);

/* The 20 x 256 x 4 premultiplied colour table (decode_predict.rs:9-36 through epaint's
 * Color32::from_rgba_unmultiplied) the post-kernel indexes; 20480 bytes. */
int32_t infur_b200_color_lut(const infur_b200_handle* h, uint8_t* lut, size_t cap);

/* ---- diagnostics --------------------------------------------------------------------------- */

/* One convolution through the selected implementation, host tensors in NHWC fp16 (x, residual, y),
 * weights [cout][kh][kw][cin] fp16, bias f32.  y_f32 != NULL asks for an f32 NHWC output instead.
 * Returns INFUR_E_UNSUPPORTED for a shape the implementation cannot run. */
typedef struct infur_b200_conv_desc {
  uint32_t n, h, w, cin, cout, kh, kw, stride, pad, dil;
  int32_t relu;
  int32_t impl; /* INFUR_CONV_TCGEN05 / INFUR_CONV_VALIDATE */
  /* Quantised layer (qmul != NULL; tcgen05 implementations only): x, wgt, residual hold integers (q - zero point) in
   * fp16, bias the int32 bias as f32, and the epilogue requantises like QLinearConv [+ QLinearAdd]:
   *   r = clamp(rne((acc + bias) * qmul[c]), q_lo, q_hi);  with residual: r = clamp(rne(r * q_ra + res * q_rb), q_lo2, q_hi2);
   * y holds r (fp16), y_f32 holds r * q_deq. */
  const float* qmul; /* [cout] or NULL */
  float q_lo, q_hi, q_ra, q_rb, q_lo2, q_hi2, q_deq;
  /* INFUR_CONV_TCGEN05_I8: the layer as an int8 plan runs it (u8 activation tensors, native int8 MMA; x must be >= 0,
   * i.e. input zero point 0): zero points of the residual and of the output tensor.  Host buffers keep the centred form. */
  int32_t q_zres, q_zout;
} infur_b200_conv_desc;
int32_t infur_b200_conv_test(infur_b200_handle* h, const infur_b200_conv_desc* d, const uint16_t* x,
                             const uint16_t* wgt, const float* bias, const uint16_t* residual, uint16_t* y,
                             float* y_f32, float* elapsed_ms);

/* Human-readable execution plan of the loaded model for an input of w x hgt (after Scale), n frames:
 * one line per kernel with shapes, tile choice, algorithmic FLOPs and bytes. */
int32_t infur_b200_plan_text(infur_b200_handle* h, uint32_t n, uint32_t w, uint32_t hgt, char* buf, size_t cap,
                             size_t* required);

/* Per-kernel CUDA-event timing of one forward of the current plan: ms per op in plan_text order, followed
 * by two more entries, the pre-kernel (Scale + normalise) and the post-kernel (upsample + ColorCode).
 * *count = number of entries written (ops + 2). */
int32_t infur_b200_profile_ops(infur_b200_handle* h, const uint8_t* d_bgr, uint32_t n, uint32_t w, uint32_t hgt,
                               int32_t iters, float* ms, int32_t cap, int32_t* count);

/* The same timing without stalling the pipeline, for use INSIDE a sustained loop: profile_step enqueues one device-resident
 * step (like advance_device into the plan's own output buffers) with a CUDA event after every kernel and returns
 * immediately; profile_collect waits for the stream, averages every step recorded since the last collect into ms[] (same
 * layout as profile_ops) and writes the number of steps averaged to *steps. */
int32_t infur_b200_profile_step(infur_b200_handle* h, const uint8_t* d_bgr, uint32_t n, uint32_t w, uint32_t hgt);
int32_t infur_b200_profile_collect(infur_b200_handle* h, float* ms, int32_t cap, int32_t* count, int32_t* steps);

/* Milliseconds the most recent plan build took (tensor maps, buffers, autotune of tile shapes not seen before), and how many
 * convolutions it autotuned (0 when every layer-shape class was known). */
int32_t infur_b200_plan_build_stats(const infur_b200_handle* h, float* ms, int32_t* tuned_convs);

/* Autotune decisions as text, so that a host can persist them next to the model file and skip the measurements in the next process
 * (one line per layer-shape class: "cin cout k stride dilation mode has_res cin2 bucket block_n variant").  export: every decision of
 * the handle (device 0 of a multi-device handle); import: adds decisions (all devices); unknown / malformed lines are INFUR_E_INVALID_ARG.
 * Results never depend on these choices (every variant is bit-identical), only plan-build time does. */
int32_t infur_b200_tune_export(infur_b200_handle* h, char* buf, size_t cap, size_t* required);
int32_t infur_b200_tune_import(infur_b200_handle* h, const char* text);

/* Parses an .onnx file on the CPU only (no device needed) and writes the fused op list as text;
 * lets the loader be tested without a GPU. */
int32_t infur_b200_onnx_describe(const char* utf8_path, char* buf, size_t cap, size_t* required);

#ifdef __cplusplus
}
#endif
#endif /* INFUR_B200_H */

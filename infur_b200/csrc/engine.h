// Engine internals behind the C ABI (include/infur_b200.h): device model, execution plans, forward.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/infur_b200.h"
#include "conv_tc.h"
#include "kernels.h"
#include "onnx_reader.h"

namespace infur {

struct Status {
  int code = INFUR_OK;
  std::string msg;
  bool ok() const { return code == INFUR_OK; }
  static Status error(int c, std::string m) { Status s; s.code = c; s.msg = std::move(m); return s; }
};

// One convolution of the lowered model with its packed device weights.
struct DevConv {
  int cin = 0, cout = 0, kh = 1, kw = 1, stride = 1, pad = 0, dil = 1;
  bool relu = false;
  bool stem = false;       // 7x7/s2 RGB stem: reads the padded NHWC4 buffer through 16-pixel windows
  bool tc_ok = false;      // expressible by the tcgen05 kernel
  std::string why_not;     // reason when !tc_ok
  int block_n = 0, cout_pad = 0, taps = 0, cchunks = 0, kdim = 0;
  int cin2 = 0, stride2 = 1;  // fused projection shortcut (ConvOp::in2): one extra tap of cin2 channels
  size_t w_off = 0;        // fp16 [cout_pad][kdim] for the tcgen05 kernel (byte offset into the arena)
  size_t wv_off = 0;       // fp16 [cout][kh][kw][cin] for the validation kernel (== w_off unless stem)
  size_t b_off = 0;        // f32 [cout_pad]
  // quantised layer: requantisation multipliers f32 [cout_pad] at q_off, scalars as in ConvOp
  bool quant = false;
  size_t q_off = 0;
  float q_lo = 0.f, q_hi = 0.f, q_ra = 0.f, q_rb = 0.f, q_lo2 = 0.f, q_hi2 = 0.f, q_deq = 0.f;
  // int8 plan (DeviceModel::i8): mode 2 = fp16-carried operands with u8 output (stem), 3 = native int8 (u8 x s8 -> s32);
  // 0 = follow `quant` with fp16 tensors.  u8 tensors hold the raw q: zero points of the residual and of the output.
  int mode = 0;
  int res_zp = 0, out_zp = 0;
  size_t bi_off = 0;       // mode 3: int32 bias [cout_pad]
  bool small_acc = false;  // mode 3: 255 * sum|w| + |bias| < 2^22 for every output channel (ConvTcGeom::q_tail)
  float qmul_max = 0.f;    // quantised layer: largest requantisation multiplier
};

struct DeviceModel {
  LoweredModel lm;               // weights released after upload
  std::vector<DevConv> convs;    // parallel to lm.ops (unused entries for MaxPool)
  std::vector<char> needed;      // op is on the path of a computed head
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  int out_head = -1, aux_head = -1;
  bool i8 = false;               // quantised model run as an int8 plan: u8 activation tensors, native int8 convolutions
  size_t lut_q_off = 0;          // quantised models: fp16 [3][256] pre-kernel table of QuantizeLinear(input) - zero point
  ~DeviceModel();
};

struct TensorInfo {
  int h = 0, w = 0, c = 0, ld = 0;   // ld = channel stride (>= c)
  bool f32 = false;
  size_t bytes = 0;
  void* ptr = nullptr;
};

struct PlanOp {
  int op = -1;                 // index into model ops
  int block_n = 0;             // N tile of the tcgen05 kernel chosen for this op
  bool pair = false;           // CTA-pair (cta_group::2) variant
  int variant = 0;             // 0 plain, 1 CTA pair, 2 halo patch (3x3)
  bool is_conv = false;
  bool skip = false;           // this convolution runs inside the previous op's fused kernel (conv_b2b_kernel)
  bool b2b = false;            // this op is a fused 3x3 -> 1x1 (+residual) pair
  ConvTcMaps maps;
  ConvTcGeom geom;
  DirectConvArgs direct;       // validation path
  double flops = 0, bytes = 0; // algorithmic, for plan_text / roofline
  std::string text;
};

// CUDA graph of one step for one set of buffers (cfg.use_cuda_graph)
struct GraphKey {
  const void *in, *cls, *dec, *bl, *fr, *lg, *aux;
  bool operator<(const GraphKey& o) const { return std::tie(in, cls, dec, bl, fr, lg, aux) < std::tie(o.in, o.cls, o.dec, o.bl, o.fr, o.lg, o.aux); }
};
struct PlanGraph { cudaGraphExec_t exec = nullptr; uint64_t kernels = 0; };

struct Plan {
  int n = 0, w = 0, h = 0;       // input frames
  int ow = 0, oh = 0;            // after Scale
  float factor = 1.f;
  bool has_model = false;
  int lh = 0, lw = 0, k = 0, ldk = 0;
  std::vector<TensorInfo> tensors;
  std::vector<PlanOp> ops;
  std::vector<void*> owned;      // device allocations
  // device buffers
  int32_t *xmap = nullptr, *ymap = nullptr;
  int32_t *sbx0 = nullptr, *sbx1 = nullptr, *sby0 = nullptr, *sby1 = nullptr;   // bilinear Scale taps
  float *sblx0 = nullptr, *sblx1 = nullptr, *sbly0 = nullptr, *sbly1 = nullptr;
  int32_t *y0 = nullptr, *y1 = nullptr, *x0 = nullptr, *x1 = nullptr;
  float *ly0 = nullptr, *ly1 = nullptr, *lx0 = nullptr, *lx1 = nullptr;
  int32_t *cell_xs = nullptr, *cell_ys = nullptr;   // first output column / row of every low-res cell (post_cell_kernel)
  __half* stem_in = nullptr;
  uint8_t* scaled = nullptr;     // Scale output when factor != 1
  float* lowres = nullptr;       // [n][lh][lw][ldk]
  float* aux_lowres = nullptr;
  int32_t* top_code = nullptr;   // post-kernel fast-path scratch [n][lh][lw]
  int max_lr = 0, max_lc = 0;
  // staging for the host-buffer entry points
  uint8_t* d_in = nullptr;
  uint8_t* d_class = nullptr;
  uint32_t *d_decoded = nullptr, *d_blended = nullptr, *d_frame_rgba = nullptr;
  float* d_logits = nullptr;
  size_t act_bytes = 0;
  std::map<GraphKey, PlanGraph> graphs;   // captured steps, by buffer set
  bool graphs_off = false;                // capture failed once: plain launches from then on
  ~Plan();
};

struct OutPtrs {
  uint8_t* class_map = nullptr;
  uint32_t* decoded = nullptr;
  uint32_t* blended = nullptr;
  uint32_t* frame_rgba = nullptr;
  float* logits = nullptr;      // out head, full resolution (debug)
  float* aux_logits = nullptr;
};

struct RingSlot {
  int state = 0;  // 0 free, 1 acquired, 2 submitted, 3 waited but still lent to the frame-level API (infur_b200_wait)
  uint64_t ticket = 0;
  uint32_t n = 0, w = 0, h = 0, ow = 0, oh = 0;
  size_t in_cap = 0, out_cap_px = 0;
  bool out_blend = false, out_frame = false;   // which optional output buffers exist
  uint8_t* h_in = nullptr;      // pinned
  uint8_t* h_class = nullptr;
  uint8_t* h_decoded = nullptr;
  uint8_t* h_blended = nullptr;
  uint8_t* h_frame = nullptr;
  uint8_t* d_in = nullptr;
  uint8_t* d_class = nullptr;
  uint32_t* d_decoded = nullptr;
  uint32_t* d_blended = nullptr;
  uint32_t* d_frame = nullptr;
  cudaEvent_t ev_h2d = nullptr, ev_done = nullptr, ev_out = nullptr;
  int has_decoded = 0;
  uint32_t k = 0;
  // multi-device handles: ring_submit is executed by the device's worker thread; ring_wait blocks on `issued`
  std::shared_ptr<struct SubmitState> sub;
};

struct SubmitState {
  std::mutex m;
  std::condition_variable cv;
  bool done = false;
  int32_t rc = INFUR_OK;
  std::string err;
};

// One worker thread per GPU of a multi-device handle: every CUDA call for that GPU is made by its worker, in FIFO order.
struct Worker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::deque<std::function<void()>> q;
  bool stop = false;
  void start();
  void post(std::function<void()> f);
  void run_sync(const std::function<void()>& f);   // post + wait for completion
  void shutdown();
};

// Autotune decision per layer-shape class: the layer's parameters and the amount of work (128-pixel M tiles of the whole batch,
// in half-octave buckets) -- not the exact image size or batch
struct TuneKey {
  int cin, cout, kh, stride, dil, mode, has_res, cin2, bucket;
  bool operator<(const TuneKey& o) const {
    return std::tie(cin, cout, kh, stride, dil, mode, has_res, cin2, bucket) <
           std::tie(o.cin, o.cout, o.kh, o.stride, o.dil, o.mode, o.has_res, o.cin2, o.bucket);
  }
};
struct TuneChoice { int block_n, variant; };

// frame-level API (infur_b200_submit / wait): where a frame ticket lives
struct FrameRef { int dev = 0; uint64_t slot_ticket = 0; uint32_t index = 0; uint64_t id = 0; };
struct OpenSlot { bool open = false; uint64_t slot_ticket = 0; uint32_t count = 0, w = 0, h = 0; uint8_t* bgr_in = nullptr; };

}  // namespace infur

struct infur_b200_handle {
  infur_b200_config cfg;
  int num_sms = 0;
  cudaStream_t stream = nullptr, h2d = nullptr, d2h = nullptr;
  std::string last_error;
  float factor = 1.0f;
  bool dirty = true;
  float* d_lut_f = nullptr;
  __half* d_lut_h = nullptr;
  __half* d_lut_u8 = nullptr;    // identity table for Uint8-input models (predict_onnx.rs:117-122: raw bytes, B,G,R order)
  uint32_t* d_color_lut = nullptr;
  std::vector<uint8_t> color_lut;
  std::unique_ptr<infur::DeviceModel> model;
  uint64_t model_gen = 0;
  std::map<std::tuple<int, int, int, uint32_t, uint64_t>, std::unique_ptr<infur::Plan>> plans;
  std::map<std::tuple<int, int, int, uint32_t, uint64_t>, uint64_t> plan_used;   // LRU stamps
  uint64_t plan_clock = 0;
  std::map<infur::TuneKey, infur::TuneChoice> tune_cache;   // survives plan eviction; cleared with the model
  float last_build_ms = 0.f;
  int last_build_tuned = 0;
  std::atomic<uint64_t> launches{0};
  std::vector<infur::RingSlot> ring;
  uint64_t next_ticket = 1;
  // in-loop profiling (infur_b200_profile_step / _collect): one event array per recorded step
  std::vector<std::vector<cudaEvent_t>> prof_sets;
  // ---- multi-device handles (cfg.num_devices > 1): this handle is device 0's context and the root of the group
  int index = 0;                                  // position in the device list
  infur_b200_handle* root = nullptr;              // children point to the root; nullptr on the root / a single-device handle
  std::vector<infur_b200_handle*> devs;           // root only: [this, child 1, ...]; empty for a single-device handle
  std::vector<std::unique_ptr<infur::Worker>> workers;   // root only, parallel to devs
  void* nccl_comms[INFUR_B200_MAX_DEVICES] = {};
  bool dup_devices = false;                       // test hook: one GPU listed several times (no NCCL communicator)
  // ---- frame-level API state (root / single-device handle)
  std::map<uint64_t, infur::FrameRef> frames;     // frame ticket -> slot
  std::map<uint64_t, uint32_t> slot_waits;        // slot ticket (submitted) -> frames of it not yet waited
  std::vector<infur::OpenSlot> open_slots;        // per device
  std::vector<std::pair<int, uint64_t>> lent, lent_prev;   // slots fully waited: recycled two waits later
  uint64_t next_frame_ticket = 1, submit_counter = 0;
  std::map<uint64_t, int> ticket_dev;             // ring ticket -> index into the device list
};

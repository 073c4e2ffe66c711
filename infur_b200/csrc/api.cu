// extern "C" surface of libinfur_b200.so -- see include/infur_b200.h for the contract of each entry point
// and the reference item it mirrors.
//
// Structure: every entry point has a `*_impl(ctx, ...)` that works on ONE device context (an infur_b200_handle that owns
// streams, plans, a model and a pinned ring).  A single-device handle IS such a context and calls the impl inline on the
// owner thread.  A multi-device handle (cfg.num_devices > 1) is a root object without CUDA resources: it owns one context +
// one worker thread per GPU, and the public functions route to them (group.cu has the worker and the NCCL loader).
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <sstream>

#include "engine.h"
#include "group.h"
#include "tables.h"

namespace infur {
Status build_device_model(LoweredModel&& lm, const infur_b200_config& cfg, bool skip_weights, std::unique_ptr<DeviceModel>& out);
Status get_plan(infur_b200_handle* H, int n, int w, int h, Plan** out);
Status run_forward(infur_b200_handle* H, Plan& p, const uint8_t* d_bgr, const OutPtrs& o, cudaStream_t s, float* op_ms, cudaEvent_t* evs);
void fill_pre_args(infur_b200_handle* H, Plan& p, const uint8_t* d_bgr, PreArgs& pa);
Status conv_test_impl(infur_b200_handle* H, const infur_b200_conv_desc* d, const uint16_t* x, const uint16_t* wgt, const float* bias,
                      const uint16_t* residual, uint16_t* y, float* y_f32, float* elapsed_ms);
}  // namespace infur

using namespace infur;

static thread_local std::string g_create_error;

static int32_t fail(infur_b200_handle* h, int code, const std::string& msg) {
  if (h) h->last_error = msg; else g_create_error = msg;
  return code;
}
static int32_t fail(infur_b200_handle* h, const Status& st) { return fail(h, st.code, st.msg); }

#define API_CU(h, expr)                                                                                      \
  do {                                                                                                       \
    cudaError_t e__ = (expr);                                                                                \
    if (e__ != cudaSuccess) return fail(h, INFUR_E_RUNTIME, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

// ---- routing helpers ---------------------------------------------------------------------------
static inline bool is_group(const infur_b200_handle* h) { return h && !h->devs.empty(); }
static inline int num_devs(const infur_b200_handle* h) { return is_group(h) ? (int)h->devs.size() : 1; }
static inline infur_b200_handle* ctx_of(infur_b200_handle* h, int idx) { return is_group(h) ? h->devs[(size_t)idx] : h; }

// Run f(ctx) for device `idx`: inline for a single-device handle, on the device's worker thread (FIFO with everything else
// queued for that GPU) for a multi-device handle.  The context's error text is mirrored into the root.
template <class F>
static int32_t run_on(infur_b200_handle* h, int idx, F f) {
  if (!is_group(h)) return f(h);
  infur_b200_handle* c = h->devs[(size_t)idx];
  int32_t rc = INFUR_OK;
  h->workers[(size_t)idx]->run_sync([&] { rc = f(c); });
  if (rc != INFUR_OK) h->last_error = c->last_error;
  return rc;
}

// f(ctx, idx) on every device in parallel; first failure wins.
template <class F>
static int32_t run_all(infur_b200_handle* h, F f) {
  if (!is_group(h)) return f(h, 0);
  const size_t n = h->devs.size();
  std::vector<int32_t> rc(n, INFUR_OK);
  Latch latch((int)n);
  for (size_t i = 0; i < n; ++i)
    h->workers[i]->post([&, i] { rc[i] = f(h->devs[i], (int)i); latch.count_down(); });
  latch.wait();
  for (size_t i = 0; i < n; ++i)
    if (rc[i] != INFUR_OK) { h->last_error = h->devs[i]->last_error; return rc[i]; }
  return INFUR_OK;
}

extern "C" {

void infur_b200_default_config(infur_b200_config* cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof(*cfg));
  cfg->struct_size = sizeof(*cfg);
  cfg->device = 0; cfg->max_batch = 8; cfg->ring_depth = 3; cfg->resize_mode = INFUR_RESIZE_NEAREST;
  cfg->compute_aux = 0; cfg->blend = 0; cfg->conv_impl = INFUR_CONV_TCGEN05; cfg->use_cuda_graph = 1; cfg->autotune = 1;
  cfg->num_devices = 0; cfg->frame_rgba = 1; cfg->confidence = INFUR_CONF_RAW;
}

int32_t infur_b200_abi_version(void) { return INFUR_B200_ABI_VERSION; }

}  // extern "C"

static void destroy_ctx(infur_b200_handle* h);

// One device context: streams, lookup tables, the empty ring.
static int32_t create_ctx(const infur_b200_config& c, int device, infur_b200_handle** out) {
  API_CU(nullptr, cudaSetDevice(device));
  cudaDeviceProp prop;
  API_CU(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(nullptr, INFUR_E_NO_DEVICE, std::string("device '") + prop.name + "' is not sm_100 (Blackwell B200); this library is built for sm_100a only");
  auto* h = new infur_b200_handle();
  h->cfg = c;
  h->cfg.device = device;
  h->num_sms = prop.multiProcessorCount;
  auto bail = [&](const std::string& m) { std::string mm = m; destroy_ctx(h); return fail(nullptr, INFUR_E_RUNTIME, mm); };
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&h->h2d, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->d2h, cudaStreamNonBlocking) != cudaSuccess)
    return bail("cudaStreamCreate failed");
  float lut_f[768];
  build_norm_lut(lut_f);
  __half lut_h[768];
  for (int i = 0; i < 768; ++i) lut_h[i] = __float2half_rn(lut_f[i]);
  h->color_lut.resize(20 * 256 * 4);
  build_color_lut(h->color_lut.data());
  __half lut_u8[768];
  for (int i = 0; i < 768; ++i) lut_u8[i] = __float2half_rn((float)(i & 255));   // 0..255 are exact in fp16
  if (cudaMalloc(&h->d_lut_u8, sizeof(lut_u8)) != cudaSuccess || cudaMalloc(&h->d_lut_f, sizeof(lut_f)) != cudaSuccess ||
      cudaMalloc(&h->d_lut_h, sizeof(lut_h)) != cudaSuccess || cudaMalloc(&h->d_color_lut, h->color_lut.size()) != cudaSuccess)
    return bail("cudaMalloc of lookup tables failed");
  // stream-ordered uploads followed by a stream sync: nothing is left on the legacy stream, which does not order with h->stream
  cudaMemcpyAsync(h->d_lut_u8, lut_u8, sizeof(lut_u8), cudaMemcpyHostToDevice, h->stream);
  cudaMemcpyAsync(h->d_lut_f, lut_f, sizeof(lut_f), cudaMemcpyHostToDevice, h->stream);
  cudaMemcpyAsync(h->d_lut_h, lut_h, sizeof(lut_h), cudaMemcpyHostToDevice, h->stream);
  cudaMemcpyAsync(h->d_color_lut, h->color_lut.data(), h->color_lut.size(), cudaMemcpyHostToDevice, h->stream);
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) return bail("upload of lookup tables failed");
  cudaError_t e = conv_tc_init();
  if (e != cudaSuccess) return bail(std::string("conv_tc_init: ") + cudaGetErrorString(e));
  h->ring.resize((size_t)c.ring_depth);
  h->open_slots.resize(1);
  *out = h;
  return INFUR_OK;
}

static void free_slot(RingSlot& s) {
  if (s.h_in) cudaFreeHost(s.h_in);
  if (s.h_class) cudaFreeHost(s.h_class);
  if (s.h_decoded) cudaFreeHost(s.h_decoded);
  if (s.h_blended) cudaFreeHost(s.h_blended);
  if (s.h_frame) cudaFreeHost(s.h_frame);
  if (s.d_in) cudaFree(s.d_in);
  if (s.d_class) cudaFree(s.d_class);
  if (s.d_decoded) cudaFree(s.d_decoded);
  if (s.d_blended) cudaFree(s.d_blended);
  if (s.d_frame) cudaFree(s.d_frame);
  if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
  if (s.ev_done) cudaEventDestroy(s.ev_done);
  if (s.ev_out) cudaEventDestroy(s.ev_out);
  s = RingSlot();
}

static void destroy_ctx(infur_b200_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  for (auto& set : h->prof_sets) for (auto& e : set) cudaEventDestroy(e);
  for (auto& s : h->ring) free_slot(s);
  h->plans.clear();
  h->model.reset();
  if (h->d_lut_f) cudaFree(h->d_lut_f);
  if (h->d_lut_h) cudaFree(h->d_lut_h);
  if (h->d_lut_u8) cudaFree(h->d_lut_u8);
  if (h->d_color_lut) cudaFree(h->d_color_lut);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->h2d) cudaStreamDestroy(h->h2d);
  if (h->d2h) cudaStreamDestroy(h->d2h);
  delete h;
}

extern "C" {

int32_t infur_b200_create(const infur_b200_config* cfg, infur_b200_handle** out) {
  if (!out) return fail(nullptr, INFUR_E_INVALID_ARG, "create: out is NULL");
  *out = nullptr;
  infur_b200_config c;
  infur_b200_default_config(&c);
  if (cfg) {
    if (cfg->struct_size != sizeof(infur_b200_config)) return fail(nullptr, INFUR_E_INVALID_ARG, "create: config struct_size mismatch");
    c = *cfg;
  }
  if (c.resize_mode != INFUR_RESIZE_NEAREST && c.resize_mode != INFUR_RESIZE_BILINEAR) return fail(nullptr, INFUR_E_INVALID_ARG, "create: unknown resize_mode");
  if (c.conv_impl != INFUR_CONV_TCGEN05 && c.conv_impl != INFUR_CONV_VALIDATE) return fail(nullptr, INFUR_E_INVALID_ARG, "create: unknown conv_impl");
  if (c.use_cuda_graph != 0 && c.use_cuda_graph != 1) return fail(nullptr, INFUR_E_INVALID_ARG, "create: use_cuda_graph must be 0 or 1");
  if (c.confidence != INFUR_CONF_RAW && c.confidence != INFUR_CONF_SOFTMAX) return fail(nullptr, INFUR_E_INVALID_ARG, "create: unknown confidence mode");
  if (c.max_batch < 1 || c.max_batch > 64 || c.ring_depth < 1 || c.ring_depth > 16) return fail(nullptr, INFUR_E_INVALID_ARG, "create: max_batch must be 1..64, ring_depth 1..16");
  if (c.num_devices < 0 || c.num_devices > INFUR_B200_MAX_DEVICES) return fail(nullptr, INFUR_E_INVALID_ARG, "create: num_devices must be 0..8");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return fail(nullptr, INFUR_E_NO_DEVICE, "no CUDA device: this library has no CPU fallback"); }
  bool dup = false;
  if (c.num_devices <= 1) {
    const int device = c.num_devices == 1 ? c.devices[0] : c.device;
    if (device < 0 || device >= ndev) return fail(nullptr, INFUR_E_NO_DEVICE, "create: device ordinal out of range");
    c.num_devices = 0;
    return create_ctx(c, device, out);
  }
  for (int i = 0; i < c.num_devices; ++i) {
    if (c.devices[i] < 0 || c.devices[i] >= ndev) return fail(nullptr, INFUR_E_NO_DEVICE, "create: device ordinal out of range in devices[]");
    // Test hook: INFUR_B200_ALLOW_DUP_DEVICES=1 lets one GPU appear several times, so the routing / worker / ordering logic of a
    // multi-device handle can be exercised on a one-GPU box.  NCCL refuses duplicate GPUs in a communicator, so such a handle
    // copies the weight arena device-to-device instead of broadcasting it.
    for (int j = 0; j < i; ++j)
      if (c.devices[j] == c.devices[i]) {
        const char* e = getenv("INFUR_B200_ALLOW_DUP_DEVICES");
        if (!(e && e[0] == '1')) return fail(nullptr, INFUR_E_INVALID_ARG, "create: devices[] lists a GPU twice");
        dup = true;
      }
  }
  // multi-device root: no CUDA resources of its own
  auto* root = new infur_b200_handle();
  root->cfg = c;
  root->cfg.device = c.devices[0];
  for (int i = 0; i < c.num_devices; ++i) {
    infur_b200_handle* child = nullptr;
    const int32_t rc = create_ctx(c, c.devices[i], &child);
    if (rc != INFUR_OK) { std::string m = g_create_error; infur_b200_destroy(root); return fail(nullptr, rc, m); }
    child->root = root; child->index = i;
    root->devs.push_back(child);
  }
  for (int i = 0; i < c.num_devices; ++i) {
    root->workers.emplace_back(new Worker());
    root->workers.back()->start();
  }
  root->open_slots.resize((size_t)c.num_devices);
  // the library's own communicator for the weight broadcast of model_load ("NCCL broadcast of weights at init only")
  std::string err;
  root->dup_devices = dup;
  if (!dup && !nccl_comm_init_all(root->nccl_comms, c.num_devices, c.devices, &err)) { infur_b200_destroy(root); return fail(nullptr, INFUR_E_RUNTIME, "create: " + err); }
  *out = root;
  return INFUR_OK;
}

void infur_b200_destroy(infur_b200_handle* h) {
  if (!h) return;
  if (!is_group(h) && h->workers.empty()) { destroy_ctx(h); return; }
  for (auto& w : h->workers) w->shutdown();
  nccl_comm_destroy_all(h->nccl_comms, (int)h->devs.size());
  for (auto* c : h->devs) destroy_ctx(c);
  delete h;
}

const char* infur_b200_last_error(const infur_b200_handle* h) { return h ? h->last_error.c_str() : g_create_error.c_str(); }

int32_t infur_b200_num_devices(const infur_b200_handle* h) { return h ? num_devs(h) : 0; }

int32_t infur_b200_scale_control(infur_b200_handle* h, float factor) {
  if (!h) return INFUR_E_INVALID_ARG;
  if (factor <= 0.0f) return fail(h, INFUR_E_SCALE_NONPOSITIVE, "Cannot scale by negative number");
  h->dirty = factor != h->factor;   // NaN != anything: dirty, as in the reference
  h->factor = factor;
  if (is_group(h))   // queued behind everything already submitted: frames submitted before this call keep the old factor
    for (size_t i = 0; i < h->devs.size(); ++i) {
      infur_b200_handle* c = h->devs[i];
      h->workers[i]->post([c, factor] { c->dirty = factor != c->factor; c->factor = factor; });
    }
  return INFUR_OK;
}

int32_t infur_b200_is_dirty(const infur_b200_handle* h) { return h && h->dirty ? 1 : 0; }

}  // extern "C"

// ---- model load --------------------------------------------------------------------------------

// Parse + lower on the calling thread (no device needed).
static int32_t parse_model(infur_b200_handle* h, std::vector<uint8_t>&& bytes, LoweredModel& lm) {
  try {
    OnnxGraph g;
    parse_onnx(std::move(bytes), g);
    lower_model(g, lm);
  } catch (const ModelError& e) {
    return fail(h, e.code, e.msg);
  } catch (const std::exception& e) {
    return fail(h, INFUR_E_MODEL_LOAD, std::string("Failed to load model: ") + e.what());
  }
  return INFUR_OK;
}

static int32_t stage_model(infur_b200_handle* c, LoweredModel&& lm, bool skip_weights, std::unique_ptr<DeviceModel>& dm) {
  cudaSetDevice(c->cfg.device);
  try {
    Status st = build_device_model(std::move(lm), c->cfg, skip_weights, dm);
    if (!st.ok()) return fail(c, st);
  } catch (const std::exception& e) {
    return fail(c, INFUR_E_MODEL_LOAD, std::string("Failed to load model: ") + e.what());
  }
  return INFUR_OK;
}

// only now: a failed load keeps the previous model (predict_onnx.rs:289-308)
static void commit_model(infur_b200_handle* c, std::unique_ptr<DeviceModel>&& dm) {
  cudaSetDevice(c->cfg.device);
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->d2h);
  c->plans.clear();
  c->plan_used.clear();
  c->model = std::move(dm);
  c->model_gen++;
}

static int32_t load_from_bytes(infur_b200_handle* h, std::vector<uint8_t>&& bytes, int32_t flags) {
  LoweredModel lm;
  int32_t rc = parse_model(h, std::move(bytes), lm);
  if (rc != INFUR_OK) return rc;
  if (!is_group(h)) {
    std::unique_ptr<DeviceModel> dm;
    rc = stage_model(h, std::move(lm), (flags & INFUR_LOAD_SKIP_WEIGHTS) != 0, dm);
    if (rc != INFUR_OK) return rc;
    commit_model(h, std::move(dm));
    return INFUR_OK;
  }
  if (flags & INFUR_LOAD_SKIP_WEIGHTS) return fail(h, INFUR_E_INVALID_ARG, "model_load: INFUR_LOAD_SKIP_WEIGHTS is for single-device handles (a multi-device handle broadcasts by itself)");
  // every device builds its structures in parallel; only devices[0] packs and uploads the weights
  const size_t n = h->devs.size();
  std::vector<std::unique_ptr<DeviceModel>> staged(n);
  std::vector<LoweredModel> copies(n);
  for (size_t i = 1; i < n; ++i) copies[i] = lm;
  copies[0] = std::move(lm);
  rc = run_all(h, [&](infur_b200_handle* c, int i) { return stage_model(c, std::move(copies[(size_t)i]), i != 0, staged[(size_t)i]); });
  if (rc != INFUR_OK) return rc;   // staged models are dropped: every device keeps its previous model
  const size_t bytes_arena = staged[0]->arena_bytes;
  for (size_t i = 1; i < n; ++i)
    if (staged[i]->arena_bytes != bytes_arena) return fail(h, INFUR_E_RUNTIME, "model_load: devices disagree on the weight arena size");
  std::vector<void*> bufs(n);
  std::vector<cudaStream_t> streams(n);
  std::vector<int> ords(n);
  for (size_t i = 0; i < n; ++i) { bufs[i] = staged[i]->arena; streams[i] = h->devs[i]->stream; ords[i] = h->devs[i]->cfg.device; }
  std::string err;
  if (h->dup_devices) {   // test hook (see create): same GPU several times, plain device-to-device copies
    for (size_t i = 1; i < n; ++i) {
      cudaSetDevice(ords[i]);
      if (cudaMemcpyAsync(bufs[i], bufs[0], bytes_arena, cudaMemcpyDeviceToDevice, streams[i]) != cudaSuccess || cudaStreamSynchronize(streams[i]) != cudaSuccess)
        return fail(h, INFUR_E_RUNTIME, "model_load: weight copy failed");
    }
  } else if (!nccl_broadcast_all(h->nccl_comms, (int)n, ords.data(), bufs.data(), bytes_arena, streams.data(), &err))
    return fail(h, INFUR_E_RUNTIME, "model_load: weight broadcast failed: " + err);
  run_all(h, [&](infur_b200_handle* c, int i) { commit_model(c, std::move(staged[(size_t)i])); return (int32_t)INFUR_OK; });
  return INFUR_OK;
}

extern "C" {

int32_t infur_b200_model_load_opts(infur_b200_handle* h, const char* utf8_path, int32_t flags) {
  if (!h || !utf8_path) return fail(h, INFUR_E_INVALID_ARG, "model_load: NULL argument");
  if (utf8_path[0] == '\0') {   // ModelCmd::Load("") unloads (predict_onnx.rs:310-312)
    return run_all(h, [](infur_b200_handle* c, int) { commit_model(c, nullptr); return (int32_t)INFUR_OK; });
  }
  std::vector<uint8_t> bytes;
  try { read_file(utf8_path, bytes); }
  catch (const std::exception& e) { return fail(h, INFUR_E_MODEL_LOAD, std::string("Failed to load model: ") + e.what()); }
  return load_from_bytes(h, std::move(bytes), flags);
}

int32_t infur_b200_model_load(infur_b200_handle* h, const char* utf8_path) { return infur_b200_model_load_opts(h, utf8_path, INFUR_LOAD_DEFAULT); }

int32_t infur_b200_model_load_bytes(infur_b200_handle* h, const void* onnx, size_t size) {
  if (!h || (!onnx && size)) return fail(h, INFUR_E_INVALID_ARG, "model_load_bytes: NULL argument");
  std::vector<uint8_t> bytes((const uint8_t*)onnx, (const uint8_t*)onnx + size);
  return load_from_bytes(h, std::move(bytes), INFUR_LOAD_DEFAULT);
}

// model structures only change inside model_load (synchronous on the owner thread), so reading them here is safe
static const DeviceModel* model_of(const infur_b200_handle* h) { return is_group(h) ? h->devs[0]->model.get() : h->model.get(); }

static int32_t copy_text(const std::string& s, char* buf, size_t cap, size_t* required) {
  if (required) *required = s.size() + 1;
  if (!buf || cap < s.size() + 1) return INFUR_E_BUFFER_TOO_SMALL;
  memcpy(buf, s.c_str(), s.size() + 1);
  return INFUR_OK;
}

int32_t infur_b200_model_info(const infur_b200_handle* h, char* buf, size_t cap, size_t* required) {
  if (!h) return INFUR_E_INVALID_ARG;
  const DeviceModel* m = model_of(h);
  if (!m) return INFUR_E_INVALID_ARG;
  const ModelIO& io = m->lm.io;
  std::string s = (io.input_names.empty() ? std::string() : io.input_names[0]) + "\t" + io.input0_dtype + "\t";
  for (size_t i = 0; i < io.output_names.size(); ++i) s += (i ? "," : "") + io.output_names[i];
  return copy_text(s, buf, cap, required);
}

int32_t infur_b200_class_legend(const infur_b200_handle* h, char* buf, size_t cap, size_t* required) {
  if (!h) return INFUR_E_INVALID_ARG;
  const DeviceModel* m = model_of(h);
  if (!m || m->out_head < 0) return INFUR_E_INVALID_ARG;
  // torchvision's fcn_resnet50 / the ONNX zoo's fcn-resnet50-12 are trained on the 20 Pascal-VOC categories + background
  static const char* voc[21] = {"__background__", "aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow",
                                "diningtable", "dog", "horse", "motorbike", "person", "pottedplant", "sheep", "sofa", "train", "tvmonitor"};
  const int k = m->lm.heads[(size_t)m->out_head].num_classes;
  const std::vector<uint8_t>& lut = is_group(h) ? h->devs[0]->color_lut : h->color_lut;
  std::ostringstream os;
  for (int c = 0; c < k; ++c) {
    const uint8_t* rgba = lut.data() + ((size_t)(c % 20) * 256 + 255) * 4;   // alpha 255: the palette colour itself (decode_predict.rs:34)
    os << c << "\t";
    if (k == 21) os << voc[c]; else os << "class " << c;
    os << "\t" << (int)rgba[0] << "," << (int)rgba[1] << "," << (int)rgba[2] << "\n";
  }
  return copy_text(os.str(), buf, cap, required);
}

int32_t infur_b200_model_weights_size(const infur_b200_handle* h, size_t* bytes) {
  if (!h || !bytes || !model_of(h)) return INFUR_E_INVALID_ARG;
  *bytes = model_of(h)->arena_bytes;
  return INFUR_OK;
}
int32_t infur_b200_model_weights_export(infur_b200_handle* h, void* d_dst, size_t bytes) {
  if (!h || is_group(h) || !d_dst || !h->model || bytes != h->model->arena_bytes) return fail(h, INFUR_E_INVALID_ARG, "weights_export: no model, size mismatch, or a multi-device handle");
  cudaSetDevice(h->cfg.device);
  API_CU(h, cudaMemcpyAsync(d_dst, h->model->arena, bytes, cudaMemcpyDeviceToDevice, h->stream));
  API_CU(h, cudaStreamSynchronize(h->stream));
  return INFUR_OK;
}
int32_t infur_b200_model_weights_import(infur_b200_handle* h, const void* d_src, size_t bytes) {
  if (!h || is_group(h) || !d_src || !h->model || bytes != h->model->arena_bytes) return fail(h, INFUR_E_INVALID_ARG, "weights_import: no model, size mismatch, or a multi-device handle");
  cudaSetDevice(h->cfg.device);
  API_CU(h, cudaMemcpyAsync(h->model->arena, d_src, bytes, cudaMemcpyDeviceToDevice, h->stream));
  API_CU(h, cudaStreamSynchronize(h->stream));   // the copy is complete (not merely enqueued) before any forward can start
  return INFUR_OK;
}

int32_t infur_b200_model_weights_checksum(infur_b200_handle* h, int32_t index, uint64_t* sum) {
  if (!h || !sum || index < 0 || index >= num_devs(h)) return fail(h, INFUR_E_INVALID_ARG, "weights_checksum: bad argument");
  return run_on(h, index, [&](infur_b200_handle* c) -> int32_t {
    if (!c->model) return fail(c, INFUR_E_INVALID_ARG, "weights_checksum: no model loaded");
    cudaSetDevice(c->cfg.device);
    std::vector<uint8_t> host(c->model->arena_bytes);
    API_CU(c, cudaMemcpyAsync(host.data(), c->model->arena, host.size(), cudaMemcpyDeviceToHost, c->stream));
    API_CU(c, cudaStreamSynchronize(c->stream));
    uint64_t hsh = 1469598103934665603ull;
    for (uint8_t b : host) { hsh ^= b; hsh *= 1099511628211ull; }
    *sum = hsh;
    return INFUR_OK;
  });
}

}  // extern "C"

// ---- advance -----------------------------------------------------------------------------------

static void fill_required(const Plan& p, int k, size_t req[7]) {
  const size_t px = (size_t)p.ow * p.oh;
  req[0] = px * 3; req[1] = px * 4; req[2] = px; req[3] = px * 4; req[4] = px * 4; req[5] = px * 4 * (size_t)k; req[6] = req[5];
}

// n frames given by pointer (they need not be contiguous: a multi-device handle hands each GPU every n-th frame of a batch)
static int32_t advance_impl(infur_b200_handle* h, const uint8_t* const* frames, uint32_t n, uint32_t w, uint32_t hgt, const uint64_t* ids,
                            infur_b200_out* const* outs) {
  if ((int)n > h->cfg.max_batch) return fail(h, INFUR_E_INVALID_ARG, "advance: batch larger than max_batch");
  cudaSetDevice(h->cfg.device);
  h->dirty = false;   // Scale::advance clears dirty first (processing.rs:233)
  Plan* pp = nullptr;
  Status st = get_plan(h, (int)n, (int)w, (int)hgt, &pp);
  if (!st.ok()) return fail(h, st);
  Plan& p = *pp;
  const int k = p.has_model ? p.k : 0;
  size_t req[7];
  fill_required(p, k, req);
  bool too_small = false, want_logits = false, want_aux = false, any_buffer = false, want_frame = false;
  for (uint32_t i = 0; i < n; ++i) {
    infur_b200_out& o = *outs[i];
    any_buffer |= o.scaled_bgr || o.frame_rgba || o.class_map || o.decoded_rgba || o.blended_rgba || o.logits_f32 || o.aux_logits_f32;
    o.out_w = (uint32_t)p.ow; o.out_h = (uint32_t)p.oh; o.num_classes = (uint32_t)k; o.has_decoded = p.has_model ? 1 : 0;
    o.id = ids ? ids[i] : 0;
    memcpy(o.required, req, sizeof(req));
    if (!p.has_model) { o.required[2] = o.required[3] = o.required[4] = o.required[5] = o.required[6] = 0; }
    if (o.scaled_bgr && o.scaled_bgr_cap < req[0]) too_small = true;
    if (o.frame_rgba && o.frame_rgba_cap < req[1]) too_small = true;
    want_frame |= o.frame_rgba != nullptr;
    if (p.has_model) {
      if (o.class_map && o.class_map_cap < req[2]) too_small = true;
      if (o.decoded_rgba && o.decoded_rgba_cap < req[3]) too_small = true;
      if (o.blended_rgba && o.blended_rgba_cap < req[4]) too_small = true;
      if (o.logits_f32 && o.logits_cap < req[5]) too_small = true;
      if (o.aux_logits_f32 && o.aux_logits_cap < req[6]) too_small = true;
      if (o.blended_rgba && !h->cfg.blend) return fail(h, INFUR_E_INVALID_ARG, "advance: blended_rgba requested but cfg.blend == 0");
      if (o.aux_logits_f32 && !(h->cfg.compute_aux && p.aux_lowres)) return fail(h, INFUR_E_INVALID_ARG, "advance: aux logits requested but cfg.compute_aux == 0 or the model has no aux head");
      want_logits |= o.logits_f32 != nullptr;
      want_aux |= o.aux_logits_f32 != nullptr;
    }
  }
  if (too_small) return fail(h, INFUR_E_BUFFER_TOO_SMALL, "advance: an output buffer is too small (see required[])");
  const size_t frame_bytes = (size_t)w * hgt * 3, px = (size_t)p.ow * p.oh;
  if (px == 0 || !any_buffer) return INFUR_OK;   // no buffer at all = size query: out_w/out_h/num_classes/required[] only
  for (uint32_t i = 0; i < n; ++i)
    if (!frames[i]) return fail(h, INFUR_E_INVALID_ARG, "advance: bgr is NULL");
  bool contiguous = true;
  for (uint32_t i = 1; i < n; ++i) contiguous &= frames[i] == frames[0] + (size_t)i * frame_bytes;
  if (contiguous) API_CU(h, cudaMemcpyAsync(p.d_in, frames[0], frame_bytes * n, cudaMemcpyHostToDevice, h->stream));
  else for (uint32_t i = 0; i < n; ++i) API_CU(h, cudaMemcpyAsync(p.d_in + (size_t)i * frame_bytes, frames[i], frame_bytes, cudaMemcpyHostToDevice, h->stream));
  float *d_logits = nullptr, *d_aux = nullptr;
  if (want_logits) API_CU(h, cudaMalloc(&d_logits, px * n * k * 4));
  if (want_aux && cudaMalloc(&d_aux, px * n * k * 4) != cudaSuccess) { cudaFree(d_logits); return fail(h, INFUR_E_RUNTIME, "advance: cudaMalloc of the aux logits failed"); }
  OutPtrs o;
  o.class_map = p.d_class; o.decoded = p.d_decoded; o.blended = p.d_blended; o.frame_rgba = want_frame ? p.d_frame_rgba : nullptr;
  o.logits = d_logits; o.aux_logits = d_aux;
  st = run_forward(h, p, p.d_in, o, h->stream, nullptr, nullptr);
  if (st.ok()) {
    const uint8_t* scaled = p.factor == 1.0f ? p.d_in : p.scaled;
    for (uint32_t i = 0; i < n && st.ok(); ++i) {
      infur_b200_out& u = *outs[i];
      auto d2h = [&](void* dst, const void* src, size_t bytes) {
        if (dst && st.ok()) { cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream); if (e != cudaSuccess) st = Status::error(INFUR_E_RUNTIME, cudaGetErrorString(e)); }
      };
      d2h(u.scaled_bgr, scaled + i * px * 3, px * 3);
      d2h(u.frame_rgba, p.d_frame_rgba + i * px, px * 4);
      if (p.has_model) {
        d2h(u.class_map, p.d_class + i * px, px);
        d2h(u.decoded_rgba, p.d_decoded + i * px, px * 4);
        if (p.d_blended) d2h(u.blended_rgba, p.d_blended + i * px, px * 4);
        if (d_logits) d2h(u.logits_f32, d_logits + i * px * k, px * 4 * k);
        if (d_aux) d2h(u.aux_logits_f32, d_aux + i * px * k, px * 4 * k);
      }
    }
  }
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (d_logits) cudaFree(d_logits);
  if (d_aux) cudaFree(d_aux);
  if (!st.ok()) return fail(h, st);
  if (e != cudaSuccess) return fail(h, INFUR_E_RUNTIME, std::string("advance: ") + cudaGetErrorString(e));
  return INFUR_OK;
}

extern "C" {

int32_t infur_b200_advance_batch(infur_b200_handle* h, const uint8_t* bgr, uint32_t n, uint32_t w, uint32_t hgt, const uint64_t* ids,
                                 infur_b200_out* outs) {
  if (!h || !outs || n == 0) return fail(h, INFUR_E_INVALID_ARG, "advance: NULL argument or empty batch");
  for (uint32_t i = 0; i < n; ++i)
    if (outs[i].struct_size != sizeof(infur_b200_out)) return fail(h, INFUR_E_INVALID_ARG, "advance: out struct_size mismatch");
  if (!bgr && (size_t)w * hgt != 0) return fail(h, INFUR_E_INVALID_ARG, "advance: bgr is NULL");
  const size_t frame_bytes = (size_t)w * hgt * 3;
  const int nd = num_devs(h);
  // frame id -> devices[(id - 1) % n] (ids are 1-based, ff-video/src/decoder.rs:163-164); without ids, position in the batch
  std::vector<std::vector<const uint8_t*>> fr((size_t)nd);
  std::vector<std::vector<uint64_t>> fid((size_t)nd);
  std::vector<std::vector<infur_b200_out*>> fo((size_t)nd);
  for (uint32_t i = 0; i < n; ++i) {
    const uint64_t id = ids ? ids[i] : (uint64_t)i + 1;
    const size_t d = (size_t)((id + (uint64_t)nd - 1) % (uint64_t)nd);
    fr[d].push_back(bgr ? bgr + (size_t)i * frame_bytes : nullptr);
    fid[d].push_back(ids ? ids[i] : 0);
    fo[d].push_back(&outs[i]);
  }
  h->dirty = false;
  if (!is_group(h)) return advance_impl(h, fr[0].data(), n, w, hgt, ids ? fid[0].data() : nullptr, fo[0].data());
  return run_all(h, [&](infur_b200_handle* c, int i) -> int32_t {
    const size_t d = (size_t)i;
    if (fr[d].empty()) return INFUR_OK;
    return advance_impl(c, fr[d].data(), (uint32_t)fr[d].size(), w, hgt, ids ? fid[d].data() : nullptr, fo[d].data());
  });
}

int32_t infur_b200_host_alloc(size_t bytes, void** out) {
  if (!out) return INFUR_E_INVALID_ARG;
  *out = nullptr;
  const cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) { cudaGetLastError(); *out = nullptr; return fail(nullptr, INFUR_E_RUNTIME, std::string("host_alloc: ") + cudaGetErrorString(e)); }
  return INFUR_OK;
}
void infur_b200_host_free(void* p) { if (p) cudaFreeHost(p); }

int32_t infur_b200_advance(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt, uint64_t id, infur_b200_out* out) {
  return infur_b200_advance_batch(h, bgr, 1, w, hgt, &id, out);
}

}  // extern "C"

static int32_t advance_device_impl(infur_b200_handle* h, const uint8_t* d_bgr, uint32_t n, uint32_t w, uint32_t hgt, infur_b200_device_out* out,
                                   int32_t sync) {
  if ((int)n > h->cfg.max_batch) return fail(h, INFUR_E_INVALID_ARG, "advance_device: batch larger than max_batch");
  cudaSetDevice(h->cfg.device);
  h->dirty = false;
  Plan* pp = nullptr;
  Status st = get_plan(h, (int)n, (int)w, (int)hgt, &pp);
  if (!st.ok()) return fail(h, st);
  Plan& p = *pp;
  const size_t px = (size_t)n * p.ow * p.oh;
  out->out_w = (uint32_t)p.ow; out->out_h = (uint32_t)p.oh; out->num_classes = p.has_model ? (uint32_t)p.k : 0; out->has_decoded = p.has_model ? 1 : 0;
  out->required[0] = p.has_model ? px : 0; out->required[1] = p.has_model ? px * 4 : 0; out->required[2] = (p.has_model && h->cfg.blend) ? px * 4 : 0;
  if (!out->d_class_map && !out->d_decoded_rgba && !out->d_blended_rgba) return INFUR_OK;   // size query
  if (out->d_blended_rgba && !h->cfg.blend) return fail(h, INFUR_E_INVALID_ARG, "advance_device: blended output requested but cfg.blend == 0");
  if (!p.has_model || px == 0) return INFUR_OK;   // decoded_img = None (app.rs:127-129): buffers untouched
  if ((out->d_class_map && out->class_map_cap < px) || (out->d_decoded_rgba && out->decoded_rgba_cap < px * 4) ||
      (out->d_blended_rgba && out->blended_rgba_cap < px * 4))
    return fail(h, INFUR_E_BUFFER_TOO_SMALL, "advance_device: an output buffer is too small for the current Scale factor (see required[])");
  if (!d_bgr) return fail(h, INFUR_E_INVALID_ARG, "advance_device: d_bgr is NULL");
  OutPtrs o;
  o.class_map = out->d_class_map; o.decoded = reinterpret_cast<uint32_t*>(out->d_decoded_rgba); o.blended = reinterpret_cast<uint32_t*>(out->d_blended_rgba);
  st = run_forward(h, p, d_bgr, o, h->stream, nullptr, nullptr);
  if (!st.ok()) return fail(h, st);
  if (sync) API_CU(h, cudaStreamSynchronize(h->stream));
  return INFUR_OK;
}

extern "C" {

int32_t infur_b200_advance_device(infur_b200_handle* h, const uint8_t* d_bgr, uint32_t n, uint32_t w, uint32_t hgt, infur_b200_device_out* out,
                                  int32_t sync) {
  if (!h || !out || n == 0) return fail(h, INFUR_E_INVALID_ARG, "advance_device: NULL argument or empty batch");
  if (out->struct_size != sizeof(infur_b200_device_out)) return fail(h, INFUR_E_INVALID_ARG, "advance_device: out struct_size mismatch");
  h->dirty = false;
  return run_on(h, 0, [&](infur_b200_handle* c) { return advance_device_impl(c, d_bgr, n, w, hgt, out, sync); });
}

void* infur_b200_compute_stream(const infur_b200_handle* h) { return h ? (void*)(is_group(h) ? h->devs[0]->stream : h->stream) : nullptr; }
uint64_t infur_b200_launch_count(const infur_b200_handle* h) {
  if (!h) return 0;
  if (!is_group(h)) return h->launches.load();
  uint64_t t = 0;
  for (auto* c : h->devs) t += c->launches.load();
  return t;
}

}  // extern "C"

// ---- pinned ring -------------------------------------------------------------------------------

static RingSlot* find_slot(infur_b200_handle* c, uint64_t ticket) {
  for (auto& r : c->ring) if (r.state != 0 && r.ticket == ticket) return &r;
  return nullptr;
}

// Output buffers of a slot for out_px pixels (class map + decoded RGBA, + blended / frame RGBA when configured).
static int32_t slot_reserve_out(infur_b200_handle* h, RingSlot* s, size_t out_px) {
  const bool blend = h->cfg.blend != 0, frame = h->cfg.frame_rgba != 0;
  if (s->out_cap_px >= out_px && s->h_class && s->out_blend == blend && s->out_frame == frame) return INFUR_OK;
  if (s->h_class) cudaFreeHost(s->h_class);
  if (s->h_decoded) cudaFreeHost(s->h_decoded);
  if (s->h_blended) cudaFreeHost(s->h_blended);
  if (s->h_frame) cudaFreeHost(s->h_frame);
  if (s->d_class) cudaFree(s->d_class);
  if (s->d_decoded) cudaFree(s->d_decoded);
  if (s->d_blended) cudaFree(s->d_blended);
  if (s->d_frame) cudaFree(s->d_frame);
  s->h_class = s->h_decoded = s->h_blended = s->h_frame = nullptr; s->d_class = nullptr; s->d_decoded = s->d_blended = s->d_frame = nullptr; s->out_cap_px = 0;
  const size_t px = std::max<size_t>(out_px, 16);
  API_CU(h, cudaHostAlloc(&s->h_class, px, cudaHostAllocPortable));
  API_CU(h, cudaHostAlloc(&s->h_decoded, px * 4, cudaHostAllocPortable));
  API_CU(h, cudaMalloc(&s->d_class, px));
  API_CU(h, cudaMalloc(&s->d_decoded, px * 4));
  if (blend) { API_CU(h, cudaHostAlloc(&s->h_blended, px * 4, cudaHostAllocPortable)); API_CU(h, cudaMalloc(&s->d_blended, px * 4)); }
  if (frame) { API_CU(h, cudaHostAlloc(&s->h_frame, px * 4, cudaHostAllocPortable)); API_CU(h, cudaMalloc(&s->d_frame, px * 4)); }
  s->out_cap_px = out_px; s->out_blend = blend; s->out_frame = frame;
  return INFUR_OK;
}

static int32_t ring_acquire_impl(infur_b200_handle* h, uint32_t n, uint32_t w, uint32_t hgt, uint64_t ticket, infur_b200_slot* slot) {
  cudaSetDevice(h->cfg.device);
  RingSlot* s = nullptr;
  for (auto& r : h->ring) if (r.state == 0) { s = &r; break; }
  if (!s) return fail(h, INFUR_E_TICKET, "ring_acquire: all ring slots are in flight");
  Plan* pp = nullptr;
  Status st = get_plan(h, (int)n, (int)w, (int)hgt, &pp);
  if (!st.ok()) return fail(h, st);
  const size_t in_bytes = (size_t)n * w * hgt * 3, out_px = (size_t)n * pp->ow * pp->oh;
  if (!s->ev_h2d) {
    API_CU(h, cudaEventCreateWithFlags(&s->ev_h2d, cudaEventDisableTiming));
    API_CU(h, cudaEventCreateWithFlags(&s->ev_done, cudaEventDisableTiming));
    API_CU(h, cudaEventCreateWithFlags(&s->ev_out, cudaEventDisableTiming));
  }
  if (s->in_cap < in_bytes || !s->h_in) {
    if (s->h_in) cudaFreeHost(s->h_in);
    if (s->d_in) cudaFree(s->d_in);
    s->h_in = nullptr; s->d_in = nullptr; s->in_cap = 0;
    API_CU(h, cudaHostAlloc(&s->h_in, std::max<size_t>(in_bytes, 16), cudaHostAllocPortable));
    API_CU(h, cudaMalloc(&s->d_in, std::max<size_t>(in_bytes, 16)));
    s->in_cap = in_bytes;
  }
  const int32_t rc = slot_reserve_out(h, s, out_px);
  if (rc != INFUR_OK) return rc;
  s->state = 1; s->ticket = ticket; s->n = n; s->w = w; s->h = hgt; s->ow = (uint32_t)pp->ow; s->oh = (uint32_t)pp->oh; s->sub.reset();
  memset(slot, 0, sizeof(*slot));
  slot->ticket = s->ticket; slot->n = n; slot->w = w; slot->h = hgt; slot->bgr_in = s->h_in; slot->device = h->cfg.device;
  return INFUR_OK;
}

// Enqueue one slot: H2D on the copy-in stream, the path on the compute stream, D2H on the copy-out stream.
static int32_t ring_submit_impl(infur_b200_handle* h, RingSlot* s) {
  cudaSetDevice(h->cfg.device);
  h->dirty = false;
  Plan* pp = nullptr;
  Status st = get_plan(h, (int)s->n, (int)s->w, (int)s->h, &pp);
  if (!st.ok()) return fail(h, st);
  Plan& p = *pp;
  const size_t in_bytes = (size_t)s->n * s->w * s->h * 3, out_px = (size_t)s->n * p.ow * p.oh;
  // the Scale factor may have been raised since ring_acquire sized the slot: grow the output buffers instead of overrunning them
  if (out_px > s->out_cap_px) {
    API_CU(h, cudaStreamSynchronize(h->d2h));
    const int32_t rc = slot_reserve_out(h, s, out_px);
    if (rc != INFUR_OK) return rc;
  }
  s->ow = (uint32_t)p.ow; s->oh = (uint32_t)p.oh; s->has_decoded = p.has_model ? 1 : 0; s->k = p.has_model ? (uint32_t)p.k : 0;
  API_CU(h, cudaMemcpyAsync(s->d_in, s->h_in, in_bytes, cudaMemcpyHostToDevice, h->h2d));
  API_CU(h, cudaEventRecord(s->ev_h2d, h->h2d));
  API_CU(h, cudaStreamWaitEvent(h->stream, s->ev_h2d, 0));
  OutPtrs o;
  o.class_map = s->d_class; o.decoded = s->d_decoded; o.blended = h->cfg.blend ? s->d_blended : nullptr;
  o.frame_rgba = h->cfg.frame_rgba ? s->d_frame : nullptr;
  st = run_forward(h, p, s->d_in, o, h->stream, nullptr, nullptr);
  if (!st.ok()) return fail(h, st);
  API_CU(h, cudaEventRecord(s->ev_done, h->stream));
  API_CU(h, cudaStreamWaitEvent(h->d2h, s->ev_done, 0));
  if (out_px) {
    if (p.has_model) {
      API_CU(h, cudaMemcpyAsync(s->h_class, s->d_class, out_px, cudaMemcpyDeviceToHost, h->d2h));
      API_CU(h, cudaMemcpyAsync(s->h_decoded, s->d_decoded, out_px * 4, cudaMemcpyDeviceToHost, h->d2h));
      if (h->cfg.blend) API_CU(h, cudaMemcpyAsync(s->h_blended, s->d_blended, out_px * 4, cudaMemcpyDeviceToHost, h->d2h));
    }
    if (h->cfg.frame_rgba) API_CU(h, cudaMemcpyAsync(s->h_frame, s->d_frame, out_px * 4, cudaMemcpyDeviceToHost, h->d2h));
  }
  API_CU(h, cudaEventRecord(s->ev_out, h->d2h));
  return INFUR_OK;
}

// ---- routed ring operations (work for single- and multi-device handles; `d` = index into the device list)

static int32_t ring_acquire_on(infur_b200_handle* h, int d, uint32_t n, uint32_t w, uint32_t hgt, infur_b200_slot* slot) {
  const uint64_t ticket = h->next_ticket;
  const int32_t rc = run_on(h, d, [&](infur_b200_handle* c) { return ring_acquire_impl(c, n, w, hgt, ticket, slot); });
  if (rc == INFUR_OK) { h->next_ticket++; h->ticket_dev[ticket] = d; }
  return rc;
}

static int dev_of_ticket(infur_b200_handle* h, uint64_t ticket) {
  auto it = h->ticket_dev.find(ticket);
  return it == h->ticket_dev.end() ? -1 : it->second;
}

static void slot_free(infur_b200_handle* h, int d, RingSlot* s) {
  s->state = 0; s->sub.reset();
  h->ticket_dev.erase(s->ticket);
}

static int32_t ring_submit_on(infur_b200_handle* h, uint64_t ticket) {
  const int d = dev_of_ticket(h, ticket);
  infur_b200_handle* c = d < 0 ? nullptr : ctx_of(h, d);
  RingSlot* s = c ? find_slot(c, ticket) : nullptr;
  if (!s || s->state != 1) return fail(h, INFUR_E_TICKET, "ring_submit: unknown ticket or slot already submitted");
  h->dirty = false;
  if (!is_group(h)) {
    const int32_t rc = ring_submit_impl(c, s);
    if (rc != INFUR_OK) { slot_free(h, d, s); return rc; }   // a failed submit gives the slot back (it could never be waited)
    s->state = 2;
    return INFUR_OK;
  }
  // multi-device: the GPU's worker issues the copies and launches; this thread goes on feeding the other GPUs
  auto sub = std::make_shared<SubmitState>();
  s->sub = sub; s->state = 2;
  h->workers[(size_t)d]->post([c, s, sub] {
    const int32_t rc = ring_submit_impl(c, s);
    std::lock_guard<std::mutex> lk(sub->m);
    sub->rc = rc; if (rc != INFUR_OK) sub->err = c->last_error;
    sub->done = true; sub->cv.notify_all();
  });
  return INFUR_OK;
}

// Wait until a submitted slot's results are in pinned memory.  On failure the slot is freed.
static int32_t slot_sync(infur_b200_handle* h, int d, RingSlot* s) {
  infur_b200_handle* c = ctx_of(h, d);
  if (s->sub) {
    std::unique_lock<std::mutex> lk(s->sub->m);
    s->sub->cv.wait(lk, [&] { return s->sub->done; });
    if (s->sub->rc != INFUR_OK) { const int32_t rc = s->sub->rc; const std::string err = s->sub->err; lk.unlock(); slot_free(h, d, s); return fail(h, rc, err); }
  }
  cudaSetDevice(c->cfg.device);
  const cudaError_t e = cudaEventSynchronize(s->ev_out);
  if (e != cudaSuccess) { slot_free(h, d, s); return fail(h, INFUR_E_RUNTIME, std::string("ring_wait: ") + cudaGetErrorString(e)); }
  return INFUR_OK;
}

static void fill_slot_result(infur_b200_handle* h, int d, const RingSlot* s, infur_b200_slot* slot) {
  infur_b200_handle* c = ctx_of(h, d);
  memset(slot, 0, sizeof(*slot));
  slot->ticket = s->ticket; slot->n = s->n; slot->w = s->w; slot->h = s->h; slot->out_w = s->ow; slot->out_h = s->oh;
  slot->num_classes = s->k; slot->has_decoded = s->has_decoded; slot->bgr_in = s->h_in; slot->device = c->cfg.device;
  if (s->has_decoded) { slot->class_map = s->h_class; slot->decoded_rgba = s->h_decoded; slot->blended_rgba = c->cfg.blend ? s->h_blended : nullptr; }
  slot->frame_rgba = c->cfg.frame_rgba ? s->h_frame : nullptr;
}

extern "C" {

int32_t infur_b200_ring_acquire(infur_b200_handle* h, uint32_t n, uint32_t w, uint32_t hgt, infur_b200_slot* slot) {
  if (!h || !slot || n == 0 || (int)n > h->cfg.max_batch) return fail(h, INFUR_E_INVALID_ARG, "ring_acquire: bad argument");
  return ring_acquire_on(h, (int)((h->next_ticket - 1) % (uint64_t)num_devs(h)), n, w, hgt, slot);
}

int32_t infur_b200_ring_read(infur_b200_handle* h, uint64_t ticket, int32_t fd, uint32_t* frames_read, size_t* partial_bytes) {
  if (!h || !frames_read) return INFUR_E_INVALID_ARG;
  *frames_read = 0;
  if (partial_bytes) *partial_bytes = 0;
  const int d = dev_of_ticket(h, ticket);
  RingSlot* s = d < 0 ? nullptr : find_slot(ctx_of(h, d), ticket);
  if (!s || s->state != 1) return fail(h, INFUR_E_TICKET, "ring_read: unknown ticket or slot already submitted");
  const size_t frame_bytes = (size_t)s->w * s->h * 3;
  int32_t rc = INFUR_OK;
  uint32_t got = 0;
  for (; got < s->n && rc == INFUR_OK; ++got) {
    uint8_t* dst = s->h_in + (size_t)got * frame_bytes;
    size_t have = 0;
    while (have < frame_bytes) {   // read_exact
      const ssize_t r = ::read(fd, dst + have, frame_bytes - have);
      if (r > 0) { have += (size_t)r; continue; }
      if (r < 0 && errno == EINTR) continue;
      if (r == 0) { rc = INFUR_E_STREAM_END; if (partial_bytes) *partial_bytes = have; h->last_error = "failed to fill whole buffer"; }
      else rc = fail(h, INFUR_E_RUNTIME, std::string("ring_read: ") + strerror(errno));
      break;
    }
    if (rc != INFUR_OK) break;
  }
  *frames_read = got;
  if (got == 0) slot_free(h, d, s);   // nothing to process: the slot goes back to the ring
  else s->n = got;
  return rc;
}

int32_t infur_b200_ring_submit(infur_b200_handle* h, uint64_t ticket) {
  if (!h) return INFUR_E_INVALID_ARG;
  return ring_submit_on(h, ticket);
}

int32_t infur_b200_ring_wait(infur_b200_handle* h, uint64_t ticket, infur_b200_slot* slot) {
  if (!h || !slot) return INFUR_E_INVALID_ARG;
  const int d = dev_of_ticket(h, ticket);
  RingSlot* s = d < 0 ? nullptr : find_slot(ctx_of(h, d), ticket);
  if (!s || s->state != 2) return fail(h, INFUR_E_TICKET, "ring_wait: unknown ticket or slot not submitted");
  const int32_t rc = slot_sync(h, d, s);
  if (rc != INFUR_OK) return rc;
  fill_slot_result(h, d, s, slot);
  slot_free(h, d, s);   // results stay valid until the slot is acquired again
  return INFUR_OK;
}

int32_t infur_b200_ring_release(infur_b200_handle* h, uint64_t ticket) {
  if (!h) return INFUR_E_INVALID_ARG;
  const int d = dev_of_ticket(h, ticket);
  RingSlot* s = d < 0 ? nullptr : find_slot(ctx_of(h, d), ticket);
  if (!s || s->state != 1) return fail(h, INFUR_E_TICKET, "ring_release: unknown ticket or slot already submitted");
  slot_free(h, d, s);
  return INFUR_OK;
}

// ---- frame-level API ---------------------------------------------------------------------------

static int32_t submit_open(infur_b200_handle* h, int d) {
  OpenSlot& os = h->open_slots[(size_t)d];
  if (!os.open) return INFUR_OK;
  os.open = false;
  RingSlot* s = find_slot(ctx_of(h, d), os.slot_ticket);
  if (!s) return fail(h, INFUR_E_TICKET, "submit: open slot vanished");
  s->n = os.count;   // a partially filled slot runs with the frames it has
  h->slot_waits[os.slot_ticket] = os.count;
  const int32_t rc = ring_submit_on(h, os.slot_ticket);
  if (rc != INFUR_OK) {   // the frames of this slot can never complete
    for (auto it = h->frames.begin(); it != h->frames.end();) it = it->second.slot_ticket == os.slot_ticket ? h->frames.erase(it) : std::next(it);
    h->slot_waits.erase(os.slot_ticket);
  }
  return rc;
}

int32_t infur_b200_submit(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt, uint64_t id, uint64_t* ticket) {
  if (!h || !ticket || (!bgr && (size_t)w * hgt != 0)) return fail(h, INFUR_E_INVALID_ARG, "submit: NULL argument");
  const uint64_t eff = id ? id : h->submit_counter + 1;
  const int nd = num_devs(h);
  const int d = (int)((eff - 1) % (uint64_t)nd);
  OpenSlot& os = h->open_slots[(size_t)d];
  if (os.open && (os.w != w || os.h != hgt)) { const int32_t rc = submit_open(h, d); if (rc != INFUR_OK) return rc; }
  if (!os.open) {
    infur_b200_slot sl;
    const int32_t rc = ring_acquire_on(h, d, (uint32_t)h->cfg.max_batch, w, hgt, &sl);
    if (rc != INFUR_OK) return rc;   // INFUR_E_TICKET: every slot of that GPU is in flight or lent -- wait for earlier frames first
    os.open = true; os.slot_ticket = sl.ticket; os.count = 0; os.w = w; os.h = hgt; os.bgr_in = sl.bgr_in;
  }
  const size_t frame_bytes = (size_t)w * hgt * 3;
  if (frame_bytes) memcpy(os.bgr_in + (size_t)os.count * frame_bytes, bgr, frame_bytes);
  h->submit_counter++;
  const uint64_t ft = h->next_frame_ticket++;
  FrameRef ref; ref.dev = d; ref.slot_ticket = os.slot_ticket; ref.index = os.count; ref.id = id;
  h->frames[ft] = ref;
  os.count++;
  *ticket = ft;
  if ((int)os.count == h->cfg.max_batch) return submit_open(h, d);
  return INFUR_OK;
}

// slots whose frames had all been waited two calls ago go back to the ring
static void recycle_lent(infur_b200_handle* h) {
  for (auto& pr : h->lent_prev) { RingSlot* s = find_slot(ctx_of(h, pr.first), pr.second); if (s && s->state == 3) slot_free(h, pr.first, s); }
  h->lent_prev.swap(h->lent);
  h->lent.clear();
}

int32_t infur_b200_flush(infur_b200_handle* h) {
  if (!h) return INFUR_E_INVALID_ARG;
  recycle_lent(h);
  int32_t rc = INFUR_OK;
  for (int d = 0; d < num_devs(h); ++d) { const int32_t r = submit_open(h, d); if (rc == INFUR_OK) rc = r; }
  return rc;
}

int32_t infur_b200_wait(infur_b200_handle* h, uint64_t ticket, infur_b200_result* out) {
  if (!h || !out) return INFUR_E_INVALID_ARG;
  recycle_lent(h);
  auto it = h->frames.find(ticket);
  if (it == h->frames.end()) return fail(h, INFUR_E_TICKET, "wait: unknown frame ticket (never submitted, already waited, or its slot failed)");
  const FrameRef ref = it->second;
  OpenSlot& os = h->open_slots[(size_t)ref.dev];
  if (os.open && os.slot_ticket == ref.slot_ticket) { const int32_t rc = submit_open(h, ref.dev); if (rc != INFUR_OK) return rc; }
  RingSlot* s = find_slot(ctx_of(h, ref.dev), ref.slot_ticket);
  if (!s || (s->state != 2 && s->state != 3)) return fail(h, INFUR_E_TICKET, "wait: the frame's slot is not in flight");
  if (s->state == 2) {
    const int32_t rc = slot_sync(h, ref.dev, s);
    if (rc != INFUR_OK) {
      for (auto f = h->frames.begin(); f != h->frames.end();) f = f->second.slot_ticket == ref.slot_ticket ? h->frames.erase(f) : std::next(f);
      h->slot_waits.erase(ref.slot_ticket);
      return rc;
    }
    s->state = 3;
  }
  infur_b200_handle* c = ctx_of(h, ref.dev);
  const size_t px = (size_t)s->ow * s->oh;
  memset(out, 0, sizeof(*out));
  out->ticket = ticket; out->id = ref.id; out->out_w = s->ow; out->out_h = s->oh; out->num_classes = s->k; out->has_decoded = s->has_decoded;
  out->device = c->cfg.device;
  if (s->has_decoded) {
    out->class_map = s->h_class + (size_t)ref.index * px;
    out->decoded_rgba = s->h_decoded + (size_t)ref.index * px * 4;
    out->blended_rgba = c->cfg.blend ? s->h_blended + (size_t)ref.index * px * 4 : nullptr;
  }
  out->frame_rgba = c->cfg.frame_rgba ? s->h_frame + (size_t)ref.index * px * 4 : nullptr;
  h->frames.erase(it);
  auto sw = h->slot_waits.find(ref.slot_ticket);
  if (sw != h->slot_waits.end() && --sw->second == 0) { h->slot_waits.erase(sw); h->lent.emplace_back(ref.dev, ref.slot_ticket); }
  return INFUR_OK;
}

}  // extern "C"

// ---- single-stage entry points and diagnostics (device 0 of a multi-device handle) ---------------


static int32_t scale_advance_impl(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt, uint8_t* out_bgr, size_t out_cap,
                                 uint32_t* out_w, uint32_t* out_h) {
  if (!h) return INFUR_E_INVALID_ARG;
  cudaSetDevice(h->cfg.device);
  h->dirty = false;
  if (!bgr && !out_bgr && !out_w && !out_h) return INFUR_OK;   // advance(&None, ..): clears dirty, nothing else (processing.rs:233-237)
  uint32_t ow, oh;
  if (h->factor == 1.0f) { ow = w; oh = hgt; }
  else {
    if (w == 0 || hgt == 0) return fail(h, INFUR_E_ZERO_SIZE_IN, "scaling from 0-sized input");
    ow = scaled_dim(w, h->factor); oh = scaled_dim(hgt, h->factor);
    if (ow == 0 || oh == 0) return fail(h, INFUR_E_ZERO_SIZE_OUT, "scaling to 0-sized output");
  }
  if (out_w) *out_w = ow;
  if (out_h) *out_h = oh;
  const size_t need = (size_t)ow * oh * 3;
  if (need == 0) return INFUR_OK;
  if (!out_bgr || out_cap < need) return fail(h, INFUR_E_BUFFER_TOO_SMALL, "scale_advance: output buffer too small");
  if (!bgr) return fail(h, INFUR_E_INVALID_ARG, "scale_advance: bgr is NULL");
  if (h->factor == 1.0f) { memcpy(out_bgr, bgr, need); return INFUR_OK; }   // deep copy (processing.rs:238-241)
  if (ow > (1u << 20) || oh > (1u << 20)) return fail(h, INFUR_E_UNSUPPORTED, "scale_advance: output larger than 2^20 per side");
  // the plan of this (size, factor) owns the resampling tables (nearest maps or bilinear taps) and the staging buffers
  Plan* pp = nullptr;
  Status st = get_plan(h, 1, (int)w, (int)hgt, &pp);
  if (!st.ok()) return fail(h, st);
  Plan& p = *pp;
  API_CU(h, cudaMemcpyAsync(p.d_in, bgr, (size_t)w * hgt * 3, cudaMemcpyHostToDevice, h->stream));
  PreArgs pa;
  fill_pre_args(h, p, p.d_in, pa);
  pa.stem_in = nullptr;   // Scale alone: no normalised copy for the network
  cudaError_t e = launch_pre(pa, h->stream);
  h->launches++;
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_bgr, p.scaled, need, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(h, INFUR_E_RUNTIME, std::string("scale_advance: ") + cudaGetErrorString(e));
  return INFUR_OK;
}

static int32_t model_advance_impl(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt, float* logits_f32, size_t logits_cap,
                                 float* aux_logits_f32, size_t aux_cap, uint32_t* num_classes, int32_t* has_model) {
  if (!h) return INFUR_E_INVALID_ARG;
  if (has_model) *has_model = h->model ? 1 : 0;
  if (num_classes) *num_classes = 0;
  if (!h->model) return INFUR_OK;   // no session: Ok(()) and `out` untouched (predict_onnx.rs:321-323)
  const float factor = h->factor;
  const bool dirty = h->dirty;
  h->factor = 1.0f;                 // the Model stage receives the already scaled image
  infur_b200_out o;
  memset(&o, 0, sizeof(o));
  o.struct_size = sizeof(o);
  o.logits_f32 = logits_f32; o.logits_cap = logits_cap; o.aux_logits_f32 = aux_logits_f32; o.aux_logits_cap = aux_cap;
  infur_b200_out* op = &o;
  const int32_t rc = advance_impl(h, &bgr, 1, w, hgt, nullptr, &op);
  h->factor = factor; h->dirty = dirty;
  if (num_classes) *num_classes = o.num_classes;
  return rc;
}

static int32_t model_lowres_impl(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt, float* lowres, size_t cap_floats,
                                uint32_t* k, uint32_t* lw, uint32_t* lh) {
  if (!h) return INFUR_E_INVALID_ARG;
  if (k) *k = 0;
  if (lw) *lw = 0;
  if (lh) *lh = 0;
  if (!h->model) return INFUR_OK;
  const float factor = h->factor;
  const bool dirty = h->dirty;
  h->factor = 1.0f;
  infur_b200_out o;
  memset(&o, 0, sizeof(o));
  o.struct_size = sizeof(o);
  std::vector<uint8_t> cls((size_t)w * hgt);   // one requested output makes advance run the step (no buffer = size query)
  o.class_map = cls.data(); o.class_map_cap = cls.size();
  infur_b200_out* op = &o;
  int32_t rc = advance_impl(h, &bgr, 1, w, hgt, nullptr, &op);
  Plan* pp = nullptr;
  if (rc == INFUR_OK) { Status st = get_plan(h, 1, (int)w, (int)hgt, &pp); if (!st.ok()) rc = fail(h, st); }
  h->factor = factor; h->dirty = dirty;
  if (rc != INFUR_OK) return rc;
  const Plan& p = *pp;
  if (k) *k = (uint32_t)p.k;
  if (lw) *lw = (uint32_t)p.lw;
  if (lh) *lh = (uint32_t)p.lh;
  const size_t px = (size_t)p.lh * p.lw;
  if (!lowres || cap_floats < px * p.k) return fail(h, INFUR_E_BUFFER_TOO_SMALL, "model_lowres: buffer too small");
  std::vector<float> tmp(px * p.ldk);
  API_CU(h, cudaMemcpy(tmp.data(), p.lowres, tmp.size() * 4, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < px; ++i)
    for (int c = 0; c < p.k; ++c) lowres[(size_t)c * px + i] = tmp[i * p.ldk + c];
  return INFUR_OK;
}

static int32_t preprocess_impl(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt, float* out_nchw, size_t out_cap_bytes) {
  if (!h) return INFUR_E_INVALID_ARG;
  cudaSetDevice(h->cfg.device);
  const size_t px = (size_t)w * hgt;
  if (px == 0) return INFUR_OK;
  if (!bgr || !out_nchw || out_cap_bytes < px * 12) return fail(h, INFUR_E_BUFFER_TOO_SMALL, "preprocess: bad buffers");
  uint8_t* d_in = nullptr; float* d_out = nullptr;
  if (cudaMalloc(&d_in, px * 3) != cudaSuccess || cudaMalloc(&d_out, px * 12) != cudaSuccess) { cudaFree(d_in); cudaFree(d_out); return fail(h, INFUR_E_RUNTIME, "preprocess: cudaMalloc failed"); }
  cudaMemcpyAsync(d_in, bgr, px * 3, cudaMemcpyHostToDevice, h->stream);
  cudaError_t e = launch_preprocess_f32(d_in, (int)hgt, (int)w, h->d_lut_f, d_out, h->stream);
  h->launches++;
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_nchw, d_out, px * 12, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_in); cudaFree(d_out);
  if (e != cudaSuccess) return fail(h, INFUR_E_RUNTIME, std::string("preprocess: ") + cudaGetErrorString(e));
  return INFUR_OK;
}

static int32_t color_code_impl(infur_b200_handle* h, const float* hm, uint32_t k, uint32_t w, uint32_t hgt, uint8_t* rgba, uint8_t* class_map) {
  if (!h || k == 0) return fail(h, INFUR_E_INVALID_ARG, "color_code: bad argument");
  cudaSetDevice(h->cfg.device);
  const size_t px = (size_t)w * hgt;
  if (px == 0) return INFUR_OK;
  if (!hm || !rgba) return fail(h, INFUR_E_INVALID_ARG, "color_code: NULL buffer");
  float* d_hm = nullptr; uint32_t* d_rgba = nullptr; uint8_t* d_cls = nullptr;
  if (cudaMalloc(&d_hm, px * k * 4) != cudaSuccess || cudaMalloc(&d_rgba, px * 4) != cudaSuccess || cudaMalloc(&d_cls, px) != cudaSuccess) {
    cudaFree(d_hm); cudaFree(d_rgba); cudaFree(d_cls); return fail(h, INFUR_E_RUNTIME, "color_code: cudaMalloc failed");
  }
  cudaMemcpyAsync(d_hm, hm, px * k * 4, cudaMemcpyHostToDevice, h->stream);
  cudaError_t e = launch_color_code(d_hm, (int)k, (int)hgt, (int)w, h->d_color_lut, d_rgba, d_cls, h->stream);
  h->launches++;
  if (e == cudaSuccess) e = cudaMemcpyAsync(rgba, d_rgba, px * 4, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && class_map) e = cudaMemcpyAsync(class_map, d_cls, px, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_hm); cudaFree(d_rgba); cudaFree(d_cls);
  if (e != cudaSuccess) return fail(h, INFUR_E_RUNTIME, std::string("color_code: ") + cudaGetErrorString(e));
  return INFUR_OK;
}

static int32_t upsample_color_impl(infur_b200_handle* h, const float* lowres, uint32_t k, uint32_t lw, uint32_t lh, uint32_t out_w,
                                  uint32_t out_h, const uint8_t* frame_bgr, uint8_t* class_map, uint8_t* decoded_rgba, uint8_t* blended_rgba,
                                  float* logits_f32) {
  if (!h || !lowres || !decoded_rgba || k == 0 || lw == 0 || lh == 0) return fail(h, INFUR_E_INVALID_ARG, "upsample_color: bad argument");
  if (blended_rgba && !frame_bgr) return fail(h, INFUR_E_INVALID_ARG, "upsample_color: blending needs frame_bgr");
  cudaSetDevice(h->cfg.device);
  const size_t px = (size_t)out_w * out_h;
  if (px == 0) return INFUR_OK;
  Plan tmp;   // owns the scratch allocations
  tmp.n = 1; tmp.ow = (int)out_w; tmp.oh = (int)out_h;
  const int ldk = (int)((k + 3) / 4 * 4);
  // repack [k][lh][lw] -> [lh][lw][ldk]
  std::vector<float> packed((size_t)lh * lw * ldk, 0.f);
  for (uint32_t c = 0; c < k; ++c)
    for (size_t i = 0; i < (size_t)lh * lw; ++i) packed[i * ldk + c] = lowres[(size_t)c * lh * lw + i];
  auto up = [&](auto*& ptr, const void* src, size_t bytes) -> bool {
    void* q = nullptr;
    if (cudaMalloc(&q, std::max<size_t>(bytes, 16)) != cudaSuccess) return false;
    tmp.owned.push_back(q);
    if (src && cudaMemcpy(q, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return false;
    ptr = reinterpret_cast<typename std::remove_reference<decltype(ptr)>::type>(q);
    return true;
  };
  std::vector<int32_t> i0, i1; std::vector<float> l0, l1;
  PostArgs q; memset(&q, 0, sizeof(q));
  float* d_low = nullptr; uint8_t* d_frame = nullptr; uint8_t* d_cls = nullptr; uint32_t *d_dec = nullptr, *d_bl = nullptr; float* d_log = nullptr;
  int32_t *dy0 = nullptr, *dy1 = nullptr, *dx0 = nullptr, *dx1 = nullptr; float *dly0 = nullptr, *dly1 = nullptr, *dlx0 = nullptr, *dlx1 = nullptr;
  bool ok = up(d_low, packed.data(), packed.size() * 4);
  int32_t *d_xs = nullptr, *d_ys = nullptr;
  auto cell_starts = [](const std::vector<int32_t>& t0, int n_in) {
    std::vector<int32_t> cs((size_t)n_in + 1, (int32_t)t0.size());
    size_t x = 0;
    for (int c = 0; c <= n_in; ++c) { while (x < t0.size() && t0[x] < c) ++x; cs[(size_t)c] = (int32_t)x; }
    return cs;
  };
  build_bilinear_table((int)lh, (int)out_h, i0, i1, l0, l1);
  { const std::vector<int32_t> cs = cell_starts(i0, (int)lh); ok = ok && up(d_ys, cs.data(), cs.size() * 4); }
  int max_lr = 1, max_lc = 1;
  for (int Y0 = 0; Y0 < (int)out_h; Y0 += 32) max_lr = std::max(max_lr, i1[std::min<int>(Y0 + 32, out_h) - 1] - i0[Y0] + 1);
  ok = ok && up(dy0, i0.data(), out_h * 4) && up(dy1, i1.data(), out_h * 4) && up(dly0, l0.data(), out_h * 4) && up(dly1, l1.data(), out_h * 4);
  build_bilinear_table((int)lw, (int)out_w, i0, i1, l0, l1);
  { const std::vector<int32_t> cs = cell_starts(i0, (int)lw); ok = ok && up(d_xs, cs.data(), cs.size() * 4); }
  for (int X0 = 0; X0 < (int)out_w; X0 += 32) max_lc = std::max(max_lc, i1[std::min<int>(X0 + 32, out_w) - 1] - i0[X0] + 1);
  ok = ok && up(dx0, i0.data(), out_w * 4) && up(dx1, i1.data(), out_w * 4) && up(dlx0, l0.data(), out_w * 4) && up(dlx1, l1.data(), out_w * 4);
  if (frame_bgr) ok = ok && up(d_frame, frame_bgr, px * 3);
  ok = ok && up(d_cls, nullptr, px) && up(d_dec, nullptr, px * 4);
  if (blended_rgba) ok = ok && up(d_bl, nullptr, px * 4);
  if (logits_f32) ok = ok && up(d_log, nullptr, px * 4 * k);
  if (!ok) return fail(h, INFUR_E_RUNTIME, "upsample_color: device allocation or upload failed");
  q.lowres = d_low; q.n = 1; q.lh = (int)lh; q.lw = (int)lw; q.ldk = ldk; q.k = (int)k; q.oh = (int)out_h; q.ow = (int)out_w;
  q.y0 = dy0; q.y1 = dy1; q.ly0 = dly0; q.ly1 = dly1; q.x0 = dx0; q.x1 = dx1; q.lx0 = dlx0; q.lx1 = dlx1;
  q.color_lut = h->d_color_lut; q.frame_bgr = d_frame; q.class_map = d_cls; q.decoded = d_dec; q.blended = d_bl; q.logits = d_log;
  q.max_lr = max_lr; q.max_lc = max_lc;
  q.softmax = h->cfg.confidence == INFUR_CONF_SOFTMAX ? 1 : 0;
  q.xs = d_xs; q.ys = d_ys;
  if (post_smem_bytes(q) > 200 * 1024) return fail(h, INFUR_E_UNSUPPORTED, "upsample_color: low-res patch per tile does not fit shared memory (upsampling ratio too small / too many classes)");
  cudaError_t e = launch_post(q, h->stream);
  h->launches += (uint64_t)post_launch_count(q);
  if (e == cudaSuccess) e = cudaMemcpyAsync(decoded_rgba, d_dec, px * 4, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && class_map) e = cudaMemcpyAsync(class_map, d_cls, px, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && blended_rgba) e = cudaMemcpyAsync(blended_rgba, d_bl, px * 4, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && logits_f32) e = cudaMemcpyAsync(logits_f32, d_log, px * 4 * k, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(h, INFUR_E_RUNTIME, std::string("upsample_color: ") + cudaGetErrorString(e));
  return INFUR_OK;
}


// ---- diagnostics -------------------------------------------------------------------------------

static int32_t conv_test_api(infur_b200_handle* h, const infur_b200_conv_desc* d, const uint16_t* x, const uint16_t* wgt, const float* bias,
                             const uint16_t* residual, uint16_t* y, float* y_f32, float* elapsed_ms) {
  if (!h || !d || !x || !wgt || !bias || (!y && !y_f32)) return fail(h, INFUR_E_INVALID_ARG, "conv_test: NULL argument");
  cudaSetDevice(h->cfg.device);
  Status st = conv_test_impl(h, d, x, wgt, bias, residual, y, y_f32, elapsed_ms);
  if (!st.ok()) return fail(h, st);
  return INFUR_OK;
}

static int32_t plan_text_impl(infur_b200_handle* h, uint32_t n, uint32_t w, uint32_t hgt, char* buf, size_t cap, size_t* required) {
  if (!h) return INFUR_E_INVALID_ARG;
  cudaSetDevice(h->cfg.device);
  Plan* pp = nullptr;
  Status st = get_plan(h, (int)n, (int)w, (int)hgt, &pp);
  if (!st.ok()) return fail(h, st);
  std::ostringstream os;
  os << "plan n=" << pp->n << " in=" << pp->w << "x" << pp->h << " scaled=" << pp->ow << "x" << pp->oh << " lowres=" << pp->lw << "x" << pp->lh
     << " classes=" << pp->k << " activation_bytes=" << pp->act_bytes << "\n";
  double fl = 0, by = 0;
  for (auto& po : pp->ops) { os << po.text << "\n"; fl += po.flops; by += po.bytes; }
  os << "total GFLOP " << fl * 1e-9 << " layer-wise MB " << by * 1e-6 << "\n";
  return copy_text(os.str(), buf, cap, required);
}

static int32_t profile_ops_impl(infur_b200_handle* h, const uint8_t* d_bgr, uint32_t n, uint32_t w, uint32_t hgt, int32_t iters, float* ms,
                               int32_t cap, int32_t* count) {
  if (!h || !d_bgr || !ms || !count || iters < 1) return fail(h, INFUR_E_INVALID_ARG, "profile_ops: bad argument");
  cudaSetDevice(h->cfg.device);
  Plan* pp = nullptr;
  Status st = get_plan(h, (int)n, (int)w, (int)hgt, &pp);
  if (!st.ok()) return fail(h, st);
  if (!pp->has_model) return fail(h, INFUR_E_INVALID_ARG, "profile_ops: no model loaded");
  const int nops = (int)pp->ops.size();
  *count = nops + 2;
  if (cap < nops + 2) return fail(h, INFUR_E_BUFFER_TOO_SMALL, "profile_ops: ms[] too small");
  // events: [0] start, [1] after the pre-kernel, [2 .. nops+1] after each op, [nops+2] after the post-kernel
  std::vector<cudaEvent_t> evs((size_t)nops + 3);
  for (auto& e : evs) cudaEventCreate(&e);
  std::vector<double> acc((size_t)nops + 2, 0.0);
  OutPtrs o; o.class_map = pp->d_class; o.decoded = pp->d_decoded;
  for (int it = 0; it < iters && st.ok(); ++it) {
    st = run_forward(h, *pp, d_bgr, o, h->stream, nullptr, evs.data());
    if (!st.ok()) break;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) { st = Status::error(INFUR_E_RUNTIME, "profile_ops: stream sync failed"); break; }
    for (int i = 0; i < nops + 2; ++i) { float t = 0; cudaEventElapsedTime(&t, evs[i], evs[i + 1]); acc[i] += t; }
  }
  for (auto& e : evs) cudaEventDestroy(e);
  if (!st.ok()) return fail(h, st);
  // output order: the nops plan ops, then the pre-kernel, then the post-kernel
  for (int i = 0; i < nops; ++i) ms[i] = (float)(acc[i + 1] / iters);
  ms[nops] = (float)(acc[0] / iters);
  ms[nops + 1] = (float)(acc[nops + 1] / iters);
  return INFUR_OK;
}



// In-loop profiling: one step with an event after every kernel, no host synchronisation (infur_b200_profile_step / _collect).
static int32_t profile_step_impl(infur_b200_handle* h, const uint8_t* d_bgr, uint32_t n, uint32_t w, uint32_t hgt) {
  if (!d_bgr) return fail(h, INFUR_E_INVALID_ARG, "profile_step: d_bgr is NULL");
  cudaSetDevice(h->cfg.device);
  Plan* pp = nullptr;
  Status st = get_plan(h, (int)n, (int)w, (int)hgt, &pp);
  if (!st.ok()) return fail(h, st);
  if (!pp->has_model) return fail(h, INFUR_E_INVALID_ARG, "profile_step: no model loaded");
  const size_t nev = pp->ops.size() + 3;
  if (!h->prof_sets.empty() && h->prof_sets[0].size() != nev) return fail(h, INFUR_E_INVALID_ARG, "profile_step: the plan changed since the last profile_collect");
  std::vector<cudaEvent_t> evs(nev);
  for (auto& e : evs) API_CU(h, cudaEventCreate(&e));
  OutPtrs o; o.class_map = pp->d_class; o.decoded = pp->d_decoded;
  st = run_forward(h, *pp, d_bgr, o, h->stream, nullptr, evs.data());
  h->prof_sets.push_back(std::move(evs));
  if (!st.ok()) return fail(h, st);
  return INFUR_OK;
}

static int32_t profile_collect_impl(infur_b200_handle* h, float* ms, int32_t cap, int32_t* count, int32_t* steps) {
  if (!ms || !count) return fail(h, INFUR_E_INVALID_ARG, "profile_collect: NULL argument");
  cudaSetDevice(h->cfg.device);
  API_CU(h, cudaStreamSynchronize(h->stream));
  const int nsets = (int)h->prof_sets.size();
  if (steps) *steps = nsets;
  *count = 0;
  if (nsets == 0) return INFUR_OK;
  const int nops = (int)h->prof_sets[0].size() - 3;
  *count = nops + 2;
  int32_t rc = INFUR_OK;
  if (cap < nops + 2) rc = fail(h, INFUR_E_BUFFER_TOO_SMALL, "profile_collect: ms[] too small");
  else {
    std::vector<double> acc((size_t)nops + 2, 0.0);
    for (auto& evs : h->prof_sets)
      for (int i = 0; i < nops + 2; ++i) { float t = 0; cudaEventElapsedTime(&t, evs[(size_t)i], evs[(size_t)i + 1]); acc[(size_t)i] += t; }
    for (int i = 0; i < nops; ++i) ms[i] = (float)(acc[(size_t)i + 1] / nsets);
    ms[nops] = (float)(acc[0] / nsets);
    ms[nops + 1] = (float)(acc[(size_t)nops + 1] / nsets);
  }
  for (auto& evs : h->prof_sets) for (auto& e : evs) cudaEventDestroy(e);
  h->prof_sets.clear();
  return rc;
}

extern "C" {

int32_t infur_b200_scale_advance(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt, uint8_t* out_bgr, size_t out_cap,
                                 uint32_t* out_w, uint32_t* out_h) {
  if (!h) return INFUR_E_INVALID_ARG;
  h->dirty = false;
  return run_on(h, 0, [&](infur_b200_handle* c) { return scale_advance_impl(c, bgr, w, hgt, out_bgr, out_cap, out_w, out_h); });
}
int32_t infur_b200_model_advance(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt, float* logits_f32, size_t logits_cap,
                                 float* aux_logits_f32, size_t aux_cap, uint32_t* num_classes, int32_t* has_model) {
  if (!h) return INFUR_E_INVALID_ARG;
  return run_on(h, 0, [&](infur_b200_handle* c) { return model_advance_impl(c, bgr, w, hgt, logits_f32, logits_cap, aux_logits_f32, aux_cap, num_classes, has_model); });
}
int32_t infur_b200_model_lowres(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt, float* lowres, size_t cap_floats,
                                uint32_t* k, uint32_t* lw, uint32_t* lh) {
  if (!h) return INFUR_E_INVALID_ARG;
  return run_on(h, 0, [&](infur_b200_handle* c) { return model_lowres_impl(c, bgr, w, hgt, lowres, cap_floats, k, lw, lh); });
}
int32_t infur_b200_preprocess(infur_b200_handle* h, const uint8_t* bgr, uint32_t w, uint32_t hgt, float* out_nchw, size_t out_cap_bytes) {
  if (!h) return INFUR_E_INVALID_ARG;
  return run_on(h, 0, [&](infur_b200_handle* c) { return preprocess_impl(c, bgr, w, hgt, out_nchw, out_cap_bytes); });
}
int32_t infur_b200_color_code(infur_b200_handle* h, const float* hm, uint32_t k, uint32_t w, uint32_t hgt, uint8_t* rgba, uint8_t* class_map) {
  if (!h) return INFUR_E_INVALID_ARG;
  return run_on(h, 0, [&](infur_b200_handle* c) { return color_code_impl(c, hm, k, w, hgt, rgba, class_map); });
}
int32_t infur_b200_upsample_color(infur_b200_handle* h, const float* lowres, uint32_t k, uint32_t lw, uint32_t lh, uint32_t out_w,
                                  uint32_t out_h, const uint8_t* frame_bgr, uint8_t* class_map, uint8_t* decoded_rgba, uint8_t* blended_rgba,
                                  float* logits_f32) {
  if (!h) return INFUR_E_INVALID_ARG;
  return run_on(h, 0, [&](infur_b200_handle* c) { return upsample_color_impl(c, lowres, k, lw, lh, out_w, out_h, frame_bgr, class_map, decoded_rgba, blended_rgba, logits_f32); });
}
int32_t infur_b200_conv_test(infur_b200_handle* h, const infur_b200_conv_desc* d, const uint16_t* x, const uint16_t* wgt, const float* bias,
                             const uint16_t* residual, uint16_t* y, float* y_f32, float* elapsed_ms) {
  if (!h) return INFUR_E_INVALID_ARG;
  return run_on(h, 0, [&](infur_b200_handle* c) { return conv_test_api(c, d, x, wgt, bias, residual, y, y_f32, elapsed_ms); });
}
int32_t infur_b200_plan_text(infur_b200_handle* h, uint32_t n, uint32_t w, uint32_t hgt, char* buf, size_t cap, size_t* required) {
  if (!h) return INFUR_E_INVALID_ARG;
  return run_on(h, 0, [&](infur_b200_handle* c) { return plan_text_impl(c, n, w, hgt, buf, cap, required); });
}
int32_t infur_b200_profile_ops(infur_b200_handle* h, const uint8_t* d_bgr, uint32_t n, uint32_t w, uint32_t hgt, int32_t iters, float* ms,
                               int32_t cap, int32_t* count) {
  if (!h) return INFUR_E_INVALID_ARG;
  return run_on(h, 0, [&](infur_b200_handle* c) { return profile_ops_impl(c, d_bgr, n, w, hgt, iters, ms, cap, count); });
}
int32_t infur_b200_profile_step(infur_b200_handle* h, const uint8_t* d_bgr, uint32_t n, uint32_t w, uint32_t hgt) {
  if (!h) return INFUR_E_INVALID_ARG;
  return run_on(h, 0, [&](infur_b200_handle* c) { return profile_step_impl(c, d_bgr, n, w, hgt); });
}
int32_t infur_b200_profile_collect(infur_b200_handle* h, float* ms, int32_t cap, int32_t* count, int32_t* steps) {
  if (!h) return INFUR_E_INVALID_ARG;
  return run_on(h, 0, [&](infur_b200_handle* c) { return profile_collect_impl(c, ms, cap, count, steps); });
}
int32_t infur_b200_tune_export(infur_b200_handle* h, char* buf, size_t cap, size_t* required) {
  if (!h) return INFUR_E_INVALID_ARG;
  std::string text;
  const int32_t rc = run_on(h, 0, [&](infur_b200_handle* c) -> int32_t {
    std::ostringstream os;
    for (auto& kv : c->tune_cache) {
      const TuneKey& k = kv.first;
      os << k.cin << ' ' << k.cout << ' ' << k.kh << ' ' << k.stride << ' ' << k.dil << ' ' << k.mode << ' ' << k.has_res << ' ' << k.cin2 << ' ' << k.bucket << ' '
         << kv.second.block_n << ' ' << kv.second.variant << '\n';
    }
    text = os.str();
    return INFUR_OK;
  });
  if (rc != INFUR_OK) return rc;
  return copy_text(text, buf, cap, required);
}

int32_t infur_b200_tune_import(infur_b200_handle* h, const char* text) {
  if (!h || !text) return fail(h, INFUR_E_INVALID_ARG, "tune_import: NULL argument");
  std::vector<std::pair<TuneKey, TuneChoice>> items;
  std::istringstream is(text);
  std::string line;
  while (std::getline(is, line)) {
    if (line.find_first_not_of(" \t\r") == std::string::npos) continue;
    std::istringstream ls(line);
    TuneKey k; TuneChoice c;
    if (!(ls >> k.cin >> k.cout >> k.kh >> k.stride >> k.dil >> k.mode >> k.has_res >> k.cin2 >> k.bucket >> c.block_n >> c.variant) ||
        (c.block_n != 32 && c.block_n != 64 && c.block_n != 128 && c.block_n != 256) || c.variant < 0 || c.variant > 3 || k.cout <= 0 || k.cout % c.block_n != 0)
      return fail(h, INFUR_E_INVALID_ARG, "tune_import: malformed line '" + line + "'");
    items.emplace_back(k, c);
  }
  return run_all(h, [&](infur_b200_handle* c, int) -> int32_t {
    for (auto& it : items) c->tune_cache[it.first] = it.second;
    return INFUR_OK;
  });
}

int32_t infur_b200_plan_build_stats(const infur_b200_handle* h, float* ms, int32_t* tuned_convs) {
  if (!h) return INFUR_E_INVALID_ARG;
  float m = 0.f; int t = 0;
  if (is_group(h)) for (auto* c : h->devs) { m = std::max(m, c->last_build_ms); t = std::max(t, c->last_build_tuned); }
  else { m = h->last_build_ms; t = h->last_build_tuned; }
  if (ms) *ms = m;
  if (tuned_convs) *tuned_convs = t;
  return INFUR_OK;
}

int32_t infur_b200_color_lut(const infur_b200_handle* h, uint8_t* lut, size_t cap) {
  if (!h) return INFUR_E_INVALID_ARG;
  const std::vector<uint8_t>& src = is_group(h) ? h->devs[0]->color_lut : h->color_lut;
  if (!h || !lut || cap < src.size()) return INFUR_E_BUFFER_TOO_SMALL;
  memcpy(lut, src.data(), src.size());
  return INFUR_OK;
}

int32_t infur_b200_onnx_describe(const char* utf8_path, char* buf, size_t cap, size_t* required) {
  if (!utf8_path) return INFUR_E_INVALID_ARG;
  std::string text;
  int32_t rc = INFUR_OK;
  try {
    std::vector<uint8_t> bytes;
    read_file(utf8_path, bytes);
    OnnxGraph g;
    parse_onnx(std::move(bytes), g);
    LoweredModel lm;
    lower_model(g, lm);
    text = describe(lm);
  } catch (const ModelError& e) {
    text = e.msg; rc = e.code;
  } catch (const std::exception& e) {
    text = std::string("Failed to load model: ") + e.what(); rc = INFUR_E_MODEL_LOAD;
  }
  int32_t c = copy_text(text, buf, cap, required);
  return rc != INFUR_OK ? rc : c;
}

}  // extern "C"

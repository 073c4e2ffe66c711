// C++17 host-side mirror of infur's plug-in surface over the C ABI (include/infur_b200.h).
//
// The reference's stages all implement `trait Processor` (infur/src/processing.rs:23-60): control / advance /
// is_dirty / generate.  The reference is Rust and no Rust toolchain exists in this environment, so the compiled-
// language host side is written in C++ with the reference's names, argument meaning and error behaviour:
//
//   infur::Processor<...>  <- trait Processor                         processing.rs:23-60
//   infur::Frame           <- struct Frame (equality on id)           processing.rs:9-18
//   infur::Scale           <- Scale                                   processing.rs:179-282
//   infur::Model           <- Model                                   predict_onnx.rs:146-346
//   infur::ColorCode       <- ColorCode                               decode_predict.rs:38-84
//   infur::GpuPipeline     <- the scale -> model -> decoder section of ProcessingApp    app.rs:53-158
//
// Rust's `Result<_, E>` becomes a C++ exception of the matching error type; `Option<T>` becomes std::optional<T>;
// `advance(&In, &mut Out)` keeps its two-argument shape.  All arithmetic runs in libinfur_b200.so on the GPU; this
// header only owns host buffers.  Header-only; link with -linfur_b200.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "infur_b200.h"

namespace infur {

// ---- errors -------------------------------------------------------------------------------------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
struct ValidScaleError : Error { using Error::Error; };   // processing.rs:145-168  "Cannot scale by negative number"
struct ScaleProcError : Error {                            // processing.rs:201-211
  using Error::Error;
  bool zero_size_in() const { return code == INFUR_E_ZERO_SIZE_IN; }
  bool zero_size_out() const { return code == INFUR_E_ZERO_SIZE_OUT; }
};
struct ModelCmdError : Error { using Error::Error; };      // predict_onnx.rs:41-54
struct ModelProcError : Error { using Error::Error; };     // predict_onnx.rs:32-39

// ---- data types ---------------------------------------------------------------------------------
// image-ext/src/image_bgr.rs:7-11: tight row-major [H][W][3], B,G,R, no row padding
struct BgrImage {
  uint32_t width = 0, height = 0;
  std::vector<uint8_t> data;
  BgrImage() = default;
  BgrImage(uint32_t w, uint32_t h) : width(w), height(h), data((size_t)w * h * 3, 0) {}
};

struct Frame {   // processing.rs:9-18
  uint64_t id = 0;
  BgrImage img;
  bool operator==(const Frame& o) const { return id == o.id; }
};

struct Color32 {   // epaint::Color32: premultiplied r,g,b,a
  uint8_t r = 0, g = 0, b = 0, a = 0;
  bool operator==(const Color32& o) const { return r == o.r && g == o.g && b == o.b && a == o.a; }
};
static_assert(sizeof(Color32) == 4, "Color32 must be 4 tightly packed bytes");

struct ColorImage {   // epaint::ColorImage: size = [width, height]
  std::array<size_t, 2> size{0, 0};
  std::vector<Color32> pixels;
  size_t width() const { return size[0]; }
  size_t height() const { return size[1]; }
};

struct ModelInfo {   // predict_onnx.rs:56-62
  std::vector<std::string> input_names;
  std::string input0_dtype;
  std::vector<std::string> output_names;
};

struct GUIFrame {   // app.rs:65-69
  uint64_t id = 0;
  ColorImage buffer;
  std::optional<ColorImage> decoded_buffer;
  std::vector<uint8_t> class_map;   // extra: k_max per pixel
};

// K x H x W confidences, C-contiguous (ndarray Array3<f32>)
struct Array3f {
  size_t k = 0, h = 0, w = 0;
  std::vector<float> data;
  Array3f() = default;
  Array3f(size_t k_, size_t h_, size_t w_) : k(k_), h(h_), w(w_), data(k_ * h_ * w_, 0.f) {}
};

// ---- trait Processor (processing.rs:23-60) --------------------------------------------------------
template <class CommandT, class InputT, class OutputT, class ProcessResultT>
struct Processor {
  using Command = CommandT;
  using Input = InputT;
  using Output = OutputT;
  using ProcessResult = ProcessResultT;
  virtual ~Processor() = default;
  virtual void control(const Command& cmd) = 0;                         // throws the stage's ControlError
  virtual ProcessResult advance(const Input& inp, Output& out) = 0;     // throws the stage's processing error
  virtual bool is_dirty() const = 0;
  ProcessResult generate() {   // processing.rs:53-59
    Input i{};
    Output o{};
    return advance(i, o);
  }
};

// ---- owner of one infur_b200_handle (one GPU, one owner thread) ------------------------------------
class Handle {
 public:
  explicit Handle(int device = 0, int max_batch = 8, bool compute_aux = false, bool blend = false) {
    infur_b200_config cfg;
    infur_b200_default_config(&cfg);
    cfg.device = device; cfg.max_batch = max_batch; cfg.compute_aux = compute_aux; cfg.blend = blend;
    const int rc = infur_b200_create(&cfg, &h_);
    if (rc != INFUR_OK) throw Error(rc, infur_b200_last_error(nullptr));
    compute_aux_ = compute_aux;
  }
  // Several GPUs of the box behind ONE handle (one owner thread, like the reference's "Proc" thread, main.rs:36-40,105-112):
  // weights are ncclBroadcast inside model_load, frame id goes to devices[(id - 1) % n], results come back in submission order.
  explicit Handle(const std::vector<int>& devices, int max_batch = 8, bool compute_aux = false, bool blend = false) {
    infur_b200_config cfg;
    infur_b200_default_config(&cfg);
    if (devices.empty() || devices.size() > INFUR_B200_MAX_DEVICES) throw Error(INFUR_E_INVALID_ARG, "1..8 devices");
    cfg.device = devices[0]; cfg.num_devices = (int32_t)devices.size();
    for (size_t i = 0; i < devices.size(); ++i) cfg.devices[i] = devices[i];
    cfg.max_batch = max_batch; cfg.compute_aux = compute_aux; cfg.blend = blend;
    const int rc = infur_b200_create(&cfg, &h_);
    if (rc != INFUR_OK) throw Error(rc, infur_b200_last_error(nullptr));
    compute_aux_ = compute_aux;
  }
  int num_devices() const { return infur_b200_num_devices(h_); }
  bool compute_aux() const { return compute_aux_; }
  ~Handle() { infur_b200_destroy(h_); }
  Handle(const Handle&) = delete;
  Handle& operator=(const Handle&) = delete;
  infur_b200_handle* get() const { return h_; }
  std::string last_error() const { return infur_b200_last_error(h_); }

 private:
  infur_b200_handle* h_ = nullptr;
  bool compute_aux_ = false;
};

namespace detail {
[[noreturn]] inline void raise(const Handle& h, int rc) {
  const std::string m = h.last_error();
  switch (rc) {
    case INFUR_E_SCALE_NONPOSITIVE: throw ValidScaleError(rc, m);
    case INFUR_E_ZERO_SIZE_IN:
    case INFUR_E_ZERO_SIZE_OUT: throw ScaleProcError(rc, m);
    case INFUR_E_MODEL_LOAD:
    case INFUR_E_MODEL_INPUT_FORMAT: throw ModelCmdError(rc, m);
    case INFUR_E_SHAPE:
    case INFUR_E_RUNTIME: throw ModelProcError(rc, m);
    default: throw Error(rc, m);
  }
}
inline std::vector<std::string> split(const std::string& s, char sep) {
  std::vector<std::string> out;
  size_t a = 0;
  while (a <= s.size()) {
    size_t b = s.find(sep, a);
    if (b == std::string::npos) b = s.size();
    if (b > a) out.push_back(s.substr(a, b - a));
    a = b + 1;
  }
  return out;
}
}  // namespace detail

// ---- Scale (processing.rs:179-282): Command = f32, Input = Output = Option<Frame> -------------------
class Scale : public Processor<float, std::optional<Frame>, std::optional<Frame>, void> {
 public:
  explicit Scale(Handle& h) : h_(h) {}
  void control(const float& factor) override {
    const int rc = infur_b200_scale_control(h_.get(), factor);
    if (rc != INFUR_OK) detail::raise(h_, rc);
  }
  bool is_dirty() const override { return infur_b200_is_dirty(h_.get()) != 0; }
  void advance(const std::optional<Frame>& inp, std::optional<Frame>& out) override {
    if (!inp) {   // processing.rs:233-237: clears dirty, nothing else
      infur_b200_scale_advance(h_.get(), nullptr, 0, 0, nullptr, 0, nullptr, nullptr);
      return;
    }
    uint32_t ow = 0, oh = 0;
    int rc = infur_b200_scale_advance(h_.get(), inp->img.data.data(), inp->img.width, inp->img.height, nullptr, 0, &ow, &oh);
    if (rc != INFUR_OK && rc != INFUR_E_BUFFER_TOO_SMALL) detail::raise(h_, rc);
    if (!out || out->img.width != ow || out->img.height != oh) out = Frame{inp->id, BgrImage(ow, oh)};   // re-created on size change (:260-268)
    out->id = inp->id;
    if ((size_t)ow * oh == 0) return;
    rc = infur_b200_scale_advance(h_.get(), inp->img.data.data(), inp->img.width, inp->img.height, out->img.data.data(), out->img.data.size(), &ow, &oh);
    if (rc != INFUR_OK) detail::raise(h_, rc);
  }

 private:
  Handle& h_;
};

// ---- Model (predict_onnx.rs:146-346): Command = ModelCmd::Load(path), Input = BgrImage, Output = Vec<ArrayD<f32>> ----
struct ModelCmdLoad { std::string path; };
class Model : public Processor<ModelCmdLoad, BgrImage, std::vector<Array3f>, void> {
 public:
  explicit Model(Handle& h) : h_(h) {}
  void control(const ModelCmdLoad& cmd) override {
    const int rc = infur_b200_model_load(h_.get(), cmd.path.c_str());
    if (rc != INFUR_OK) detail::raise(h_, rc);
  }
  bool is_dirty() const override { return false; }   // predict_onnx.rs:336-338
  std::optional<ModelInfo> get_info() const {        // predict_onnx.rs:341-345
    size_t need = 0;
    if (infur_b200_model_info(h_.get(), nullptr, 0, &need) == INFUR_E_INVALID_ARG) return std::nullopt;
    std::string buf(need, '\0');
    if (infur_b200_model_info(h_.get(), buf.data(), need, &need) != INFUR_OK) return std::nullopt;
    buf.resize(need ? need - 1 : 0);
    const auto parts = detail::split(buf, '\t');
    ModelInfo mi;
    if (parts.size() > 0) mi.input_names.push_back(parts[0]);
    if (parts.size() > 1) mi.input0_dtype = parts[1];
    if (parts.size() > 2) mi.output_names = detail::split(parts[2], ',');
    return mi;
  }
  // no session: Ok(()) and `out` untouched (predict_onnx.rs:321-323); else out = the network's outputs, batch dim stripped
  void advance(const BgrImage& img, std::vector<Array3f>& out) override {
    uint32_t k = 0;
    int32_t has = 0;
    int rc = infur_b200_model_advance(h_.get(), img.data.data(), img.width, img.height, nullptr, 0, nullptr, 0, &k, &has);
    if (rc != INFUR_OK) detail::raise(h_, rc);
    if (!has) return;
    const auto info = get_info();
    const bool want_aux = h_.compute_aux() && info && info->output_names.size() > 1;   // `aux` needs cfg.compute_aux
    Array3f lg(k, img.height, img.width), aux;
    if (want_aux) aux = Array3f(k, img.height, img.width);
    rc = infur_b200_model_advance(h_.get(), img.data.data(), img.width, img.height, lg.data.data(), lg.data.size() * 4,
                                  want_aux ? aux.data.data() : nullptr, want_aux ? aux.data.size() * 4 : 0, &k, &has);
    if (rc != INFUR_OK) detail::raise(h_, rc);
    out.clear();
    out.push_back(std::move(lg));
    if (want_aux) out.push_back(std::move(aux));
  }

 private:
  Handle& h_;
};

// ---- ColorCode (decode_predict.rs:38-84): Input = Array3<f32> [K][H][W], Output = Option<ColorImage> ----
class ColorCode : public Processor<std::monostate, Array3f, std::optional<ColorImage>, void> {
 public:
  explicit ColorCode(Handle& h) : h_(h) {}
  void control(const std::monostate&) override {}
  bool is_dirty() const override { return false; }   // decode_predict.rs:81-83
  void advance(const Array3f& inp, std::optional<ColorImage>& out) override {
    if (!out || out->width() != inp.w || out->height() != inp.h) {   // get or re-create (decode_predict.rs:58-65)
      ColorImage img;
      img.size = {inp.w, inp.h};
      img.pixels.assign(inp.w * inp.h, Color32{0, 0, 0, 255});
      out = std::move(img);
    }
    if (inp.w * inp.h == 0) return;
    const int rc = infur_b200_color_code(h_.get(), inp.data.data(), (uint32_t)inp.k, (uint32_t)inp.w, (uint32_t)inp.h,
                                         reinterpret_cast<uint8_t*>(out->pixels.data()), nullptr);
    if (rc != INFUR_OK) detail::raise(h_, rc);
  }

 private:
  Handle& h_;
};

// ---- the scale -> model -> decoder section of ProcessingApp (app.rs:53-158) as ONE fused GPU call ----
struct AppCmdScale { float factor; };
struct AppCmdModel { std::string path; };
using AppCmd = std::variant<AppCmdScale, AppCmdModel>;   // the AppCmd subset of app.rs:39-51 that reaches the hot path

class GpuPipeline : public Processor<AppCmd, std::optional<Frame>, std::optional<GUIFrame>, void> {
 public:
  explicit GpuPipeline(Handle& h) : scale(h), model(h), decoder(h), h_(h) {}
  void control(const AppCmd& cmd) override {   // app.rs:91-105
    if (auto* s = std::get_if<AppCmdScale>(&cmd)) scale.control(s->factor);
    else model.control(ModelCmdLoad{std::get<AppCmdModel>(cmd).path});
  }
  bool is_dirty() const override { return scale.is_dirty(); }
  // app.rs:109-149: None in -> None out; else GUIFrame{id, buffer, decoded_buffer (None without a model)}
  void advance(const std::optional<Frame>& inp, std::optional<GUIFrame>& out) override {
    if (!inp) { out = std::nullopt; return; }
    infur_b200_out o{};
    o.struct_size = sizeof(o);
    int rc = infur_b200_advance(h_.get(), inp->img.data.data(), inp->img.width, inp->img.height, inp->id, &o);   // size query
    if (rc != INFUR_OK) detail::raise(h_, rc);
    GUIFrame g;
    g.id = inp->id;
    const size_t px = (size_t)o.out_w * o.out_h;
    g.buffer.size = {o.out_w, o.out_h};
    g.buffer.pixels.resize(px);
    o.frame_rgba = reinterpret_cast<uint8_t*>(g.buffer.pixels.data()); o.frame_rgba_cap = px * 4;
    ColorImage dec;
    if (o.has_decoded) {
      dec.size = {o.out_w, o.out_h};
      dec.pixels.resize(px);
      g.class_map.resize(px);
      o.decoded_rgba = reinterpret_cast<uint8_t*>(dec.pixels.data()); o.decoded_rgba_cap = px * 4;
      o.class_map = g.class_map.data(); o.class_map_cap = px;
    }
    if (px) {
      rc = infur_b200_advance(h_.get(), inp->img.data.data(), inp->img.width, inp->img.height, inp->id, &o);
      if (rc != INFUR_OK) detail::raise(h_, rc);
    }
    if (o.has_decoded) g.decoded_buffer = std::move(dec);
    out = std::move(g);
  }

  // Streaming use (configs 3-5): submit() copies the frame into the pinned ring slot of GPU (id - 1) % n and returns a ticket;
  // wait() hands back that frame's GUIFrame (copied out of the library's pinned memory).  Tickets complete in submission order.
  uint64_t submit(const Frame& f) {
    uint64_t t = 0;
    const int rc = infur_b200_submit(h_.get(), f.img.data.data(), f.img.width, f.img.height, f.id, &t);
    if (rc != INFUR_OK) detail::raise(h_, rc);
    return t;
  }
  void flush() { const int rc = infur_b200_flush(h_.get()); if (rc != INFUR_OK) detail::raise(h_, rc); }
  GUIFrame wait(uint64_t ticket) {
    infur_b200_result r{};
    const int rc = infur_b200_wait(h_.get(), ticket, &r);
    if (rc != INFUR_OK) detail::raise(h_, rc);
    GUIFrame g;
    g.id = r.id;
    const size_t px = (size_t)r.out_w * r.out_h;
    g.buffer.size = {r.out_w, r.out_h};
    if (r.frame_rgba) { g.buffer.pixels.resize(px); std::memcpy(g.buffer.pixels.data(), r.frame_rgba, px * 4); }
    if (r.has_decoded && r.decoded_rgba) {
      ColorImage dec;
      dec.size = {r.out_w, r.out_h};
      dec.pixels.resize(px);
      std::memcpy(dec.pixels.data(), r.decoded_rgba, px * 4);
      g.decoded_buffer = std::move(dec);
      g.class_map.assign(r.class_map, r.class_map + px);
    }
    return g;
  }

  Scale scale;
  Model model;
  ColorCode decoder;

 private:
  Handle& h_;
};

}  // namespace infur

// The reference's own unit tests for the hot path, ported to the C++ host mirror (include/infur_b200_processors.hpp)
// and run against libinfur_b200.so:
//   processing.rs:288-303     scale_from_size0, scale_to_size0
//   decode_predict.rs:93-116  color_2, decode_0to1
//   predict_onnx.rs:356-381   load_seg_model, infer_seg_model              (needs a model file: argv[1])
//   app.rs:174-252            void, scale, switch_scale, switch_video_then_scale, scaled_frame_after_stopped_video
//                             (frames come from a synthetic source instead of ffmpeg; sizes, ids and dirty flags as asserted there)
// Usage:  host_check <model.onnx>      on a B200: runs everything, prints one line per test, exit 0 iff all pass
//         host_check --no-gpu          anywhere: checks that creating a handle fails loudly without a device
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>

#include "infur_b200_processors.hpp"

using namespace infur;

static int g_failed = 0;
#define CHECK(cond)                                                               \
  do {                                                                            \
    if (!(cond)) { std::printf("    CHECK failed: %s (line %d)\n", #cond, __LINE__); throw std::runtime_error("check"); } \
  } while (0)

static void run(const char* name, const std::function<void()>& f) {
  try { f(); std::printf("ok   %s\n", name); }
  catch (const std::exception& e) { std::printf("FAIL %s: %s\n", name, e.what()); ++g_failed; }
}

static Frame synthetic_frame(uint64_t id, uint32_t w, uint32_t h) {
  Frame f{id, BgrImage(w, h)};
  for (uint32_t y = 0; y < h; ++y)
    for (uint32_t x = 0; x < w; ++x) {
      uint8_t* p = &f.img.data[((size_t)y * w + x) * 3];
      p[0] = (uint8_t)((x + 3 * id) & 255); p[1] = (uint8_t)((y * 2 + id) & 255); p[2] = (uint8_t)(((x / 8 + y / 8) & 1) * 200 + 20);
    }
  return f;
}

int main(int argc, char** argv) {
  if (argc < 2) { std::printf("usage: host_check <model.onnx> | --no-gpu\n"); return 2; }
  if (std::strcmp(argv[1], "--no-gpu") == 0) {
    try { Handle h; (void)h; std::printf("a device is present: nothing to check\n"); return 0; }
    catch (const Error& e) {
      std::printf("create() without a device: code %d, \"%s\"\n", e.code, e.what());
      return e.code == INFUR_E_NO_DEVICE ? 0 : 1;
    }
  }
  const std::string model_path = argv[1];
  Handle h(0, 8, /*compute_aux=*/true, /*blend=*/false);

  run("scale_from_size0", [&] {
    std::optional<Frame> zero = Frame{0, BgrImage(0, 10)}, out;
    Scale scale(h);
    scale.control(0.99f);
    bool got = false;
    try { scale.advance(zero, out); } catch (const ScaleProcError& e) { got = e.zero_size_in(); }
    CHECK(got);
  });
  run("scale_to_size0", [&] {
    std::optional<Frame> img = Frame{0, BgrImage(10, 10)}, out;
    Scale scale(h);
    scale.control(0.00000001f);
    bool got = false;
    try { scale.advance(img, out); } catch (const ScaleProcError& e) { got = e.zero_size_out(); }
    CHECK(got);
  });
  run("valid_scale_rejects_non_positive", [&] {
    Scale scale(h);
    scale.control(1.0f);
    for (float f : {0.0f, -1.0f}) {
      bool got = false;
      try { scale.control(f); } catch (const ValidScaleError& e) { got = std::string(e.what()) == "Cannot scale by negative number"; }
      CHECK(got);
    }
  });
  uint8_t lut[20 * 256 * 4];
  CHECK(infur_b200_color_lut(h.get(), lut, sizeof(lut)) == INFUR_OK);
  auto color_code = [&](size_t klass, float alpha) {   // decode_predict.rs:32-36 through the library's colour table
    float a255 = alpha * 255.0f;
    int a = a255 >= 255.f ? 255 : (a255 > 0.f ? (int)a255 : 0);
    const uint8_t* p = &lut[((klass % 20) * 256 + a) * 4];
    return Color32{p[0], p[1], p[2], p[3]};
  };
  run("color_2", [&] {
    const Color32 c = color_code(2, 0.5f);
    CHECK(c.a == 127);                              // (0.5 * 255) as u8
    CHECK(c.r < 25 && c.g < 225 && c.b < 255);      // premultiplied (25, 225, 255) at alpha 127
    CHECK(c.r > 5 && c.g > 100 && c.b > 120);
  });
  run("decode_0to1", [&] {
    Array3f hm(22, 24, 32);
    const size_t n = hm.data.size();
    for (size_t i = 0; i < n; ++i) hm.data[i] = (float)((double)i / (double)(n - 1));   // linspace(0, 1, n)
    std::optional<ColorImage> img;
    ColorCode decoder(h);
    decoder.advance(hm, img);
    CHECK(img && img->width() == 32 && img->height() == 24);
    int conf = 0;
    for (const Color32& p : img->pixels) {
      CHECK(p == color_code(21, p.a / 255.0f));
      CHECK(conf <= p.a);
      conf = p.a;
    }
    CHECK(conf == 255);
  });
  run("void", [&] {   // app.rs:174-179: nothing to process -> None, twice
    GpuPipeline app(h);
    std::optional<GUIFrame> out;
    app.advance(std::nullopt, out); CHECK(!out);
    app.advance(std::nullopt, out); CHECK(!out);
  });
  run("load_seg_model / infer_seg_model", [&] {
    Model m(h);
    m.control(ModelCmdLoad{model_path});
    const auto info = m.get_info();
    CHECK(info && info->input_names.size() == 1 && info->output_names.size() == 2);
    std::printf("    model %s: %s (%s) -> %s,%s\n", model_path.c_str(), info->input_names[0].c_str(), info->input0_dtype.c_str(),
                info->output_names[0].c_str(), info->output_names[1].c_str());
    BgrImage img(320, 240);
    std::vector<Array3f> tensors;
    m.advance(img, tensors);
    CHECK(tensors.size() == 2);                                                      // "should return two tensors"
    CHECK(tensors[0].k == 21 && tensors[0].h == 240 && tensors[0].w == 320);       // out: 21 classes upscaled
    CHECK(tensors[1].k == 21 && tensors[1].h == 240 && tensors[1].w == 320);       // aux
  });
  run("model_load_failure_keeps_previous_model", [&] {   // predict_onnx.rs:289-308
    Model m(h);
    bool got = false;
    try { m.control(ModelCmdLoad{"/nonexistent/model.onnx"}); } catch (const ModelCmdError&) { got = true; }
    CHECK(got && m.get_info().has_value());
  });
  run("scale / switch_scale / switch_video_then_scale", [&] {   // app.rs:181-216 output sizes
    GpuPipeline app(h);
    std::optional<GUIFrame> f;
    app.control(AppCmdScale{0.5f});
    app.advance(synthetic_frame(1, 1280, 720), f);
    CHECK(f && f->buffer.size == (std::array<size_t, 2>{1280 / 2, 720 / 2}));
    app.control(AppCmdScale{1.0f});
    app.advance(synthetic_frame(2, 640, 480), f);
    CHECK(f && f->buffer.size == (std::array<size_t, 2>{640, 480}));
    app.control(AppCmdScale{0.5f});
    app.advance(synthetic_frame(3, 640, 480), f);
    CHECK(f && f->buffer.size == (std::array<size_t, 2>{640 / 2, 480 / 2}));
    CHECK(f->decoded_buffer && f->decoded_buffer->size == f->buffer.size && f->class_map.size() == 320 * 240);
    app.control(AppCmdScale{2.0f});
    app.advance(synthetic_frame(4, 320, 180), f);
    CHECK(f && f->buffer.size == (std::array<size_t, 2>{320 * 2, 180 * 2}));
    app.control(AppCmdScale{1.0f});
  });
  run("scaled_frame_after_stopped_video", [&] {   // app.rs:219-235: re-scaling the SAME frame when the scale changes
    GpuPipeline app(h);
    std::optional<GUIFrame> f1, f2, f3;
    const auto frame = std::optional<Frame>(synthetic_frame(7, 1280, 720));
    app.advance(frame, f1);
    CHECK(f1 && f1->buffer.size == (std::array<size_t, 2>{1280, 720}));
    app.advance(frame, f2);
    CHECK(f1->id == f2->id && !app.is_dirty());
    app.control(AppCmdScale{0.5f});
    CHECK(app.is_dirty());
    app.advance(frame, f3);
    CHECK(f2->id == f3->id && f3->buffer.size == (std::array<size_t, 2>{1280 / 2, 720 / 2}) && !app.is_dirty());
    app.control(AppCmdScale{1.0f});
  });
  run("unload_model_gives_no_decoded_buffer", [&] {   // app.rs:127-129
    GpuPipeline app(h);
    app.control(AppCmdModel{""});
    std::optional<GUIFrame> f;
    app.advance(synthetic_frame(9, 64, 48), f);
    CHECK(f && !f->decoded_buffer && f->buffer.pixels.size() == 64 * 48);
    const Frame src = synthetic_frame(9, 64, 48);
    CHECK(f->buffer.pixels[5].r == src.img.data[5 * 3 + 2] && f->buffer.pixels[5].b == src.img.data[5 * 3] && f->buffer.pixels[5].a == 255);
  });
  run("stream_submit_wait_in_order", [&] {   // new surface (configs 3-5): per-frame submit / wait, results in submission order
    GpuPipeline app(h);
    app.control(AppCmdModel{model_path});
    app.control(AppCmdScale{1.0f});
    std::vector<uint64_t> tickets;
    for (uint64_t id = 1; id <= 5; ++id) tickets.push_back(app.submit(synthetic_frame(id, 96, 64)));
    for (uint64_t id = 1; id <= 5; ++id) {
      GUIFrame g = app.wait(tickets[id - 1]);
      std::optional<GUIFrame> ref;
      app.advance(synthetic_frame(id, 96, 64), ref);
      CHECK(g.id == id && g.buffer.size == (std::array<size_t, 2>{96, 64}) && g.decoded_buffer && ref && ref->decoded_buffer);
      CHECK(std::memcmp(g.decoded_buffer->pixels.data(), ref->decoded_buffer->pixels.data(), 96 * 64 * 4) == 0);
      CHECK(std::memcmp(g.buffer.pixels.data(), ref->buffer.pixels.data(), 96 * 64 * 4) == 0 && g.class_map == ref->class_map);
    }
    app.flush(); app.flush();
  });
  std::printf("%s (%d failed)\n", g_failed ? "FAILED" : "all reference tests passed", g_failed);
  return g_failed ? 1 : 0;
}

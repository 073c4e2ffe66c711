#include "tables.h"

#include <cmath>

namespace infur {

void build_norm_lut(float out[3 * 256]) {
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int c = 0; c < 3; ++c) {
    volatile float std1 = 1.0f / stdv[c];
    for (int v = 0; v < 256; ++v) {
      volatile float x = ((float)v * 1.0f) / 255.0f;
      volatile float d = x - mean[c];
      volatile float r = d * std1;
      out[c * 256 + v] = r;
    }
  }
}

namespace {

// decode_predict.rs:9-30
const uint8_t kPalette[20][3] = {
    {75, 180, 60},  {75, 25, 230},   {25, 225, 255}, {200, 130, 0},   {48, 130, 245}, {240, 240, 70}, {230, 50, 240},
    {60, 245, 210}, {180, 30, 145},  {190, 190, 250}, {128, 128, 0},  {255, 190, 230}, {40, 110, 170}, {200, 250, 255},
    {0, 0, 128},    {195, 255, 170}, {0, 128, 128},  {180, 215, 255}, {128, 0, 0},    {128, 128, 128},
};

// powf evaluated in f64 on the f32-rounded operands, rounded once to f32 (a correctly rounded powf)
float powf_cr(float x, float e) { return (float)std::pow((double)x, (double)e); }

float linear_from_gamma_u8(uint8_t s) {
  volatile float sf = (float)s;
  if (s <= 10) { volatile float r = sf / 3294.6f; return r; }
  volatile float t = sf + 14.025f;
  volatile float u = t / 269.025f;
  return powf_cr(u, 2.4f);
}

uint8_t gamma_u8_from_linear(float l) {
  if (l <= 0.0f) return 0;
  volatile float r;
  if (l <= 0.0031308f) {
    volatile float m = 3294.6f * l;
    r = std::floor(m + 0.5f);
  } else if (l <= 1.0f) {
    volatile float inv = 1.0f / 2.4f;
    volatile float p = powf_cr(l, inv);
    volatile float m = 269.025f * p;
    volatile float d = m - 14.025f;
    r = std::floor(d + 0.5f);
  } else {
    return 255;
  }
  if (r <= 0.0f) return 0;
  if (r >= 255.0f) return 255;
  return (uint8_t)r;
}

}  // namespace

void build_color_lut(uint8_t out[20 * 256 * 4]) {
  for (int k = 0; k < 20; ++k) {
    for (int a = 0; a < 256; ++a) {
      uint8_t* o = out + ((size_t)k * 256 + a) * 4;
      if (a == 255) { o[0] = kPalette[k][0]; o[1] = kPalette[k][1]; o[2] = kPalette[k][2]; o[3] = 255; continue; }
      if (a == 0) { o[0] = o[1] = o[2] = o[3] = 0; continue; }
      volatile float a_lin = (float)a / 255.0f;
      for (int c = 0; c < 3; ++c) {
        volatile float pm = linear_from_gamma_u8(kPalette[k][c]) * a_lin;
        o[c] = gamma_u8_from_linear(pm);
      }
      o[3] = (uint8_t)a;
    }
  }
}

uint32_t scaled_dim(uint32_t v, float factor) {
  volatile float f = (float)v * factor;
  if (f != f) return 0;
  if (f <= 0.0f) return 0;
  if (f >= 4294967296.0f) return 4294967295u;
  return (uint32_t)f;
}

void build_nearest_map(int src, int dst, std::vector<int32_t>& idx) {
  idx.resize((size_t)dst);
  const double s = (double)src / (double)dst;
  volatile double half_s = 0.5 * s;
  for (int x = 0; x < dst; ++x) {
    volatile double m = s * (double)x;
    volatile double v = half_s + m;
    long long i = (long long)v;
    if (i > src - 1) i = src - 1;
    idx[(size_t)x] = (int32_t)i;
  }
}

void build_bilinear_table(int n_in, int n_out, std::vector<int32_t>& i0, std::vector<int32_t>& i1, std::vector<float>& l0,
                          std::vector<float>& l1) {
  i0.resize((size_t)n_out); i1.resize((size_t)n_out); l0.resize((size_t)n_out); l1.resize((size_t)n_out);
  volatile float scale = (float)n_in / (float)n_out;
  for (int d = 0; d < n_out; ++d) {
    volatile float c = (float)d + 0.5f;
    volatile float m = scale * c;
    volatile float src = m - 0.5f;
    if (src < 0.0f) src = 0.0f;
    int a = (int)src;
    if (a > n_in - 1) a = n_in - 1;
    int b = a + 1 > n_in - 1 ? n_in - 1 : a + 1;
    volatile float w1 = src - (float)a;
    volatile float w0 = 1.0f - w1;
    i0[(size_t)d] = a; i1[(size_t)d] = b; l0[(size_t)d] = w0; l1[(size_t)d] = w1;
  }
}

}  // namespace infur

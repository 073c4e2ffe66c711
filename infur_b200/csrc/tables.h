// Host-built lookup tables that make the integer / LUT stages bit-exact with the reference semantics.
// Compiled with -ffp-contract=off: every f32 operation below is individually rounded, in the
// reference's order.
#pragma once
#include <cstdint>
#include <vector>

namespace infur {

// ColorNorm::new_torchvision_rgb + the float path of ImageSession::forward
// (infur/src/predict_onnx.rs:126-137,175-180): lut[c][v] = ((v * 1.0f) / 255.0f - mean[c]) * (1.0f / std[c]),
// c in R,G,B order.
void build_norm_lut(float out[3 * 256]);

// color_code (infur/src/decode_predict.rs:9-36) for every (class % 20, alpha byte) through epaint 0.19's
// Color32::from_rgba_unmultiplied: [20][256] x (r,g,b,a) premultiplied.
void build_color_lut(uint8_t out[20 * 256 * 4]);

// Rust `(v as f32 * factor) as u32` (infur/src/processing.rs:253-254): truncate, saturate, NaN -> 0.
uint32_t scaled_dim(uint32_t v, float factor);

// fast_image_resize 1.x nearest: src = trunc(0.5*s + s*x), s = src/dst in f64.
void build_nearest_map(int src, int dst, std::vector<int32_t>& idx);

// ONNX Resize(linear, half_pixel) / F.interpolate(bilinear, align_corners=False) taps in f32.
void build_bilinear_table(int n_in, int n_out, std::vector<int32_t>& i0, std::vector<int32_t>& i1, std::vector<float>& l0,
                          std::vector<float>& l1);

}  // namespace infur

// HBM-bound kernels on either side of the network, plus the slow validation convolution.
// Launchers only; see kernels.cu for the device code.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace infur {

// Stem input layout: NHWC with C padded 3 -> 4 (R,G,B,0 fp16), a zero border of kStemPadTop rows above /
// below and kStemPadLeft pixels left (and at least 16 zero pixels right), so that the 7x7/s2 stem reads 8-pixel
// windows without a bounds test.  The row pitch is a multiple of 16 pixels (128 B): the stem kernel fetches
// whole 128-byte groups of a row with TMA.
constexpr int kStemPadTop = 3;
constexpr int kStemPadLeft = 4;
inline int stem_pitch_px(int w) { return ((w + kStemPadLeft + 16 + 15) / 16) * 16; }
inline int stem_rows(int h) { return h + 2 * kStemPadTop + 1; }

struct PreArgs {
  const uint8_t* src;    // [n][h][w][3] BGR u8
  int n, h, w;           // input size
  int oh, ow;            // size after Scale
  const int32_t* xmap;   // [ow] nearest source column (nullptr = identity)
  const int32_t* ymap;   // [oh]
  // bilinear mode (all eight non-null): taps and weights per output column / row
  const int32_t* bx0; const int32_t* bx1; const float* blx0; const float* blx1;
  const int32_t* by0; const int32_t* by1; const float* bly0; const float* bly1;
  const __half* lut_h;   // [3][256] fp16, one table per network input channel
  int bgr_order;         // 0: network channels = (R, G, B) of the pixel (Float models); 1: (B, G, R) as stored (Uint8 models)
  __half* stem_in;       // [n][stem_rows(oh)][stem_pitch_px(ow)][4] fp16, or nullptr
  uint8_t* scaled_bgr;   // [n][oh][ow][3], or nullptr
};
cudaError_t launch_pre(const PreArgs& a, cudaStream_t s);

// ImageSession::forward pre-processing alone: [h][w][3] u8 BGR -> [3][h][w] f32 (lut_f: [3][256] f32, R,G,B)
cudaError_t launch_preprocess_f32(const uint8_t* bgr, int h, int w, const float* lut_f, float* out, cudaStream_t s);

// MaxPool k x k / stride / pad over NHWC fp16, c % 8 == 0
// 3x3 / stride 2 / pad 1 pooling of a u8 NHWC tensor (int8 plans), c % 16 == 0
cudaError_t launch_maxpool3s2_u8(const uint8_t* in, uint8_t* out, int n, int h, int w, int c, int oh, int ow, cudaStream_t s);
cudaError_t launch_maxpool(const __half* in, __half* out, int n, int h, int w, int c, int oh, int ow, int k, int stride, int pad,
                           cudaStream_t s);

struct PostArgs {
  const float* lowres;   // [n][lh][lw][ldk] f32
  int n, lh, lw, ldk, k;
  int oh, ow;
  const int32_t* y0; const int32_t* y1; const float* ly0; const float* ly1;  // [oh]
  const int32_t* x0; const int32_t* x1; const float* lx0; const float* lx1;  // [ow]
  const uint32_t* color_lut;  // [20][256] premultiplied RGBA (little-endian r,g,b,a)
  const uint8_t* frame_bgr;   // [n][oh][ow][3] (for frame_rgba / blended), may be nullptr
  uint8_t* class_map;         // [n][oh][ow], may be nullptr
  uint32_t* decoded;          // [n][oh][ow] RGBA
  uint32_t* blended;          // may be nullptr
  uint32_t* frame_rgba;       // may be nullptr
  float* logits;              // [n][k][oh][ow] f32, may be nullptr (debug)
  int max_lr, max_lc;         // largest low-res patch (rows, cols) any 32x32 output tile touches
  // cell tables of the cell-per-thread kernel (K = 21 path): low-res cell c covers output columns [xs[c], xs[c + 1]) -- the run of
  // columns whose left tap x0[] is c -- and likewise rows; nullptr: strip kernel
  const int32_t* xs; const int32_t* ys;
  int softmax;                // INFUR_CONF_SOFTMAX: confidence = softmax probability of the winning class (README.md:76), else the raw logit
  int32_t* top_code;          // scratch [n][lh][lw] (K = 21 path): winning class of a low-res pixel if it wins by a safe margin, else -1; may be nullptr
};
cudaError_t launch_post(const PostArgs& a, cudaStream_t s);
int post_launch_count(const PostArgs& a);   // kernels launch_post enqueues for these arguments
size_t post_smem_bytes(const PostArgs& a);

// ColorCode::advance alone on a planar [k][h][w] f32 map
cudaError_t launch_color_code(const float* hm, int k, int h, int w, const uint32_t* color_lut, uint32_t* rgba, uint8_t* class_map,
                              cudaStream_t s);

// BGR u8 -> RGBA (r,g,b,255)
cudaError_t launch_frame_rgba(const uint8_t* bgr, size_t npix, uint32_t* rgba, cudaStream_t s);

// Validation-only direct convolution on CUDA cores (fp16 in, f32 accumulate); generic shapes; slow.
struct DirectConvArgs {
  const __half* x; const __half* w; const float* bias; const __half* residual;
  __half* y; float* y_f32;
  int n, h, wd, cin, cout, kh, kw, stride, pad, dil, oh, ow, relu;
  int x_pitch_px, x_rows, x_c;   // physical layout of x: pixels per row, rows per image, channels per pixel
  int x_off_y, x_off_x;          // logical (0,0) sits at physical (x_off_y, x_off_x)
  int out_ld;
};
cudaError_t launch_direct_conv(const DirectConvArgs& a, cudaStream_t s);

}  // namespace infur

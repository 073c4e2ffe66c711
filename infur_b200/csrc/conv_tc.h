// tcgen05 implicit-GEMM convolution for NHWC fp16 activations (the arithmetic the reference hands to
// ONNX Runtime's Conv/Relu/Add nodes inside session.run, infur/src/predict_onnx.rs:138).
//
//   out[n][oy][ox][co] = act( bias[co] + sum_{tap,ci} in[n][oy*s+dy(tap)][ox*s+dx(tap)][ci] * w[co][tap][ci] (+ residual) )
//
// GEMM view: M = output pixels (tiles of 128 = bw x bh rectangle of one image), N = cout, K = taps*cin.
// A tiles are fetched by TMA straight from the activation tensor (4-D tiled maps, zero fill outside
// the image = the conv padding; strided convs read through per-parity strided views), B tiles from
// the [cout][taps*cin] weight matrix; both land in 128B-swizzled smem and feed tcgen05.mma with the
// accumulator in TMEM.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace infur {

constexpr int kMaxTaps = 50;   // 7x7 + one fused shortcut tap
constexpr int kStemWBytes = 7 * 4 * 64 * 8 * 2;   // 28 KB
constexpr int kStemRowGroups = 17;                // 17 x 128 B = 272 input pixels feed 128 output pixels
constexpr int kMaxViews = 4;

struct ConvTcGeom {
  int32_t n_img, oh, ow;
  int32_t bw_log2;            // tile is (1 << bw_log2) wide, 128 >> bw_log2 high
  int32_t tiles_x, tiles_y, tiles_n, num_tiles;
  int32_t num_taps, num_kb;   // K blocks = sum over taps of tap_cc[tap], 64 input channels each
  int32_t main_taps, cchunks; // the kh*kw filter taps all have `cchunks` 64-channel chunks; taps beyond them are fused shortcut taps
  int32_t out_ld;             // elements between consecutive output pixels
  int32_t relu;
  int32_t store_mode;         // 0: per-thread vector stores (f32 head); 1: smem-staged TMA store; 2: + TMA residual prefetch
  int32_t stages, epi_bufs;   // smem pipeline depth and epilogue chunk buffers (0 / 2 / 4), see conv_tc_stages
  int32_t pair, num_work;     // 1: conv_tc_pair_kernel (cta_group::2); work items = ceil(M tiles / 2) * tiles_n
  int32_t halo, halo_dil;     // 1: conv_halo_kernel (3x3 / stride 1 / pad = dilation = halo_dil): maps.a[0] has a halo-patch box
  // 1: conv_b2b_kernel -- a 3x3 / stride 1 convolution (cmid -> cmid, ReLU) fused with the 1x1 convolution that follows it
  // (cmid -> tiles_n * 128 channels, + residual, ReLU): the cmid-channel intermediate never leaves shared memory.  The fields
  // above describe the 3x3 (maps.a / maps.b, bias, taps); maps.b2 / bias2 / residual / out / maps.c / maps.r the 1x1.
  int32_t b2b, b2b_cmid;
  const float* bias2;
  // 1: stem_pool_kernel -- the 7x7/s2 stem fused with the 3x3/s2/p1 max-pool that follows it: `out` is the POOLED tensor
  // [n][sp_oh][sp_ow][64]; oh / ow stay the convolution's output size.  Work unit = (image, sp_rc pooled rows, 60 pooled columns).
  int32_t sp_fused, sp_oh, sp_ow, sp_rc, sp_chunks, sp_strips;
  int32_t stem;               // 1: 7x7/s2 RGB stem through stem_tc_kernel (maps.a[0] = row-group view of the padded NHWC4 input)
  const __half* stem_w;       // stem weights in smem order [7 ky][4 k-cores][64 cout][8], kStemWBytes
  const float* bias;          // [tiles_n * BLOCK_N]
  const __half* residual;     // NHWC like out, or nullptr
  __half* out;                // fp16 NHWC, or nullptr when out_f32 is used
  float* out_f32;             // f32 NHWC (logit head)
  // Quantised layer (QLinearConv [+ QLinearAdd], onnx_reader.h ConvOp): operands are integers carried in fp16, the epilogue
  // requantises  r = clamp(rne((acc + bias) * qmul[c]), q_lo, q_hi);  with a residual  r = clamp(rne(r * q_ra + res * q_rb),
  // q_lo2, q_hi2);  an f32 head stores r * q_deq.
  int32_t quant;
  const float* qmul;          // [tiles_n * BLOCK_N]
  float q_lo, q_hi, q_ra, q_rb, q_lo2, q_hi2, q_deq;
  // int8 plans: mode 2 = fp16-carried operands, u8 output (the stem); mode 3 = u8 activations x s8 weights, s32 accumulators,
  // u8 output / residual (or f32 head).  mode 0 / 1 = !quant / quant with fp16 tensors.  u8 tensors hold the raw q:
  // q_zres = zero point of the residual tensor, q_zmagic = zero point of the output tensor + 1.5 * 2^23.
  int32_t mode;
  const int32_t* bias_i32;    // mode 3: [tiles_n * BLOCK_N]
  float q_zres, q_zmagic;
  // u8 outputs (mode >= 2): which form of the requantisation tail the epilogue runs (requant_u8_pass in conv_tc.cu) -- decided
  // once per layer on the host from bounds that hold whatever the input:
  //   0  generic;
  //   1  integer tail: the value that reaches the final rounding add is below 2^22 in magnitude (with a residual:
  //      max|q_lo, q_hi| * |q_ra| + 255 * |q_rb|), and the final bounds are [q_floor - zero_point, 255 - zero_point];
  //   2  integer tail + small accumulators (mode 3): |accumulator + bias| < 2^22 for every output channel (255 * sum|w| + |b|),
  //      and without a residual also max qmul <= 1 -- int -> float becomes an integer add and a float subtract instead of I2F.
  // q_floor = lower bound of the stored byte (0 unless a ReLU follows a tensor with a non-zero zero point).
  int32_t q_tail, q_floor;
  int8_t tap_view[kMaxTaps + 3];
  uint8_t tap_cc[kMaxTaps + 3];  // 64-channel chunks of each tap (a fused shortcut tap may differ from the main taps)
  int16_t tap_dx[kMaxTaps + 1];
  int16_t tap_dy[kMaxTaps + 1];
};

struct alignas(64) ConvTcMaps {
  CUtensorMap a[kMaxViews];  // activation views
  CUtensorMap b;             // weights [cout][K]
  CUtensorMap c;             // output (store_mode 1, 2): 4-D NHWC, box 64 channels x tile
  CUtensorMap r;             // residual (store_mode 2), same geometry as c
  CUtensorMap b2;            // conv_b2b_kernel: weights of the fused 1x1 convolution [cout][cmid], box 64 x 128
};

// block_n in {32, 64, 128, 256}.  Returns cudaSuccess or the launch error.
cudaError_t conv_tc_launch(int block_n, const ConvTcMaps& maps, const ConvTcGeom& g, int num_sms, cudaStream_t stream);
// One-time: opt in to the dynamic shared memory each instantiation needs.
cudaError_t conv_tc_init();
int conv_tc_stages(int block_n, int epi_bufs, bool i8 = false);
int conv_tc_pair_stages(int epi_bufs, bool i8 = false);
bool conv_b2b_supported(int cmid);   // 64 or 128

}  // namespace infur

// Pre-kernel (Scale + ImageSession::forward pre-processing), max-pool, post-kernel (final bilinear
// Resize fused with ColorCode) and the slow validation convolution.  All integer/LUT paths are
// bit-exact with the reference semantics; float paths use explicitly un-fused f32 operations in the
// reference's order.
#include "kernels.h"

#include <cstdlib>

namespace infur {

namespace {

// ------------------------------------------------------------------------------------------------
// K0: Scale::advance (nearest gather, infur/src/processing.rs:232-281) fused with the u8 -> normalised
// conversion of ImageSession::forward (infur/src/predict_onnx.rs:103-137).  The normalisation has only
// 256 inputs per channel, so a [3][256] table built on the host with the reference's three separately
// rounded f32 operations reproduces it exactly; the table is pre-rounded to fp16 (the network's
// activation type).
__global__ void __launch_bounds__(256) pre_generic_kernel(PreArgs a, int stem_pitch, int stem_rows_) {
  __shared__ __half lut[3 * 256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = a.lut_h[i];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int img = blockIdx.z;
  if (x >= a.ow) return;
  uint8_t b, g, r;
  if (a.bx0) {
    // opt-in bilinear Scale: ly0*(lx0*a + lx1*b) + ly1*(lx0*c + lx1*d), un-fused f32 in this order, u8 = floor(v + 0.5)
    const int x0 = a.bx0[x], x1 = a.bx1[x], y0 = a.by0[y], y1 = a.by1[y];
    const float lx0 = a.blx0[x], lx1 = a.blx1[x], ly0 = a.bly0[y], ly1 = a.bly1[y];
    const uint8_t* base = a.src + (size_t)img * a.h * a.w * 3;
    const uint8_t* p00 = base + ((size_t)y0 * a.w + x0) * 3; const uint8_t* p01 = base + ((size_t)y0 * a.w + x1) * 3;
    const uint8_t* p10 = base + ((size_t)y1 * a.w + x0) * 3; const uint8_t* p11 = base + ((size_t)y1 * a.w + x1) * 3;
    uint8_t o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float top = __fadd_rn(__fmul_rn(lx0, (float)p00[c]), __fmul_rn(lx1, (float)p01[c]));
      const float bot = __fadd_rn(__fmul_rn(lx0, (float)p10[c]), __fmul_rn(lx1, (float)p11[c]));
      const float v = floorf(__fadd_rn(__fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot)), 0.5f));
      o[c] = (uint8_t)fminf(fmaxf(v, 0.f), 255.f);
    }
    b = o[0]; g = o[1]; r = o[2];
  } else {
    const int sx = a.xmap ? a.xmap[x] : x;
    const int sy = a.ymap ? a.ymap[y] : y;
    const uint8_t* p = a.src + (((size_t)img * a.h + sy) * a.w + sx) * 3;
    b = p[0]; g = p[1]; r = p[2];
  }
  if (a.scaled_bgr) {
    uint8_t* q = a.scaled_bgr + (((size_t)img * a.oh + y) * a.ow + x) * 3;
    q[0] = b; q[1] = g; q[2] = r;
  }
  if (a.stem_in) {
    const uint8_t c0 = a.bgr_order ? b : r, c2 = a.bgr_order ? r : b;
    __half2 rg = __halves2half2(lut[c0], lut[256 + g]);
    __half2 b0 = __halves2half2(lut[512 + c2], __ushort_as_half((unsigned short)0));
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&rg);
    v.y = *reinterpret_cast<uint32_t*>(&b0);
    uint2* dst = reinterpret_cast<uint2*>(a.stem_in) + ((size_t)img * stem_rows_ + (y + kStemPadTop)) * stem_pitch + (x + kStemPadLeft);
    *dst = v;
  }
}

// Unit-scale fast path: 4 pixels per thread, 3 x 32-bit loads -> 2 x 128-bit stores.  Needs w % 4 == 0.
__global__ void __launch_bounds__(256) pre_unit_vec4_kernel(PreArgs a, int stem_pitch, int stem_rows_) {
  __shared__ __half lut[3 * 256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = a.lut_h[i];
  __syncthreads();
  const int x4 = blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 pixels
  const int y = blockIdx.y;
  const int img = blockIdx.z;
  if (x4 * 4 >= a.w) return;
  const uint32_t* p = reinterpret_cast<const uint32_t*>(a.src + (((size_t)img * a.h + y) * a.w + (size_t)x4 * 4) * 3);
  const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
  // bytes: b0 g0 r0 b1 | g1 r1 b2 g2 | r2 b3 g3 r3
  const uint8_t bb[4] = {(uint8_t)(w0), (uint8_t)(w0 >> 24), (uint8_t)(w1 >> 16), (uint8_t)(w2 >> 8)};
  const uint8_t gg[4] = {(uint8_t)(w0 >> 8), (uint8_t)(w1), (uint8_t)(w1 >> 24), (uint8_t)(w2 >> 16)};
  const uint8_t rr[4] = {(uint8_t)(w0 >> 16), (uint8_t)(w1 >> 8), (uint8_t)(w2), (uint8_t)(w2 >> 24)};
  uint32_t o[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint8_t c0 = a.bgr_order ? bb[i] : rr[i], c2 = a.bgr_order ? rr[i] : bb[i];
    __half2 rg = __halves2half2(lut[c0], lut[256 + gg[i]]);
    __half2 b0 = __halves2half2(lut[512 + c2], __ushort_as_half((unsigned short)0));
    o[2 * i] = *reinterpret_cast<uint32_t*>(&rg);
    o[2 * i + 1] = *reinterpret_cast<uint32_t*>(&b0);
  }
  // kStemPadLeft = 4 pixels = 32 bytes, x4*4 pixels = 32*x4 bytes: 16-byte aligned stores
  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint2*>(a.stem_in) +
                                        ((size_t)img * stem_rows_ + (y + kStemPadTop)) * stem_pitch + ((size_t)x4 * 4 + kStemPadLeft));
  dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
  dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// Unit-scale path for w % 128 == 0: a warp converts 128 pixels per step.  Lanes 0..23 fetch the 384 input bytes with one 128-bit
// load each (fully coalesced), the bytes are staged in shared memory, and every lane then emits pixel pairs as 128-bit stores to
// consecutive addresses (512 contiguous bytes per store instruction) -- both directions move whole 128-byte lines.
__global__ void __launch_bounds__(128) pre_unit_warp128_kernel(PreArgs a, int stem_pitch, int stem_rows_) {
  __shared__ __half lut[3 * 256];
  __shared__ __align__(16) uint8_t stage[4][384];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = a.lut_h[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y = blockIdx.y, img = blockIdx.z;
  const int chunks = a.w >> 7;
  const uint8_t* row = a.src + ((size_t)img * a.h + y) * (size_t)a.w * 3;
  uint4* drow = reinterpret_cast<uint4*>(reinterpret_cast<uint2*>(a.stem_in) + ((size_t)img * stem_rows_ + (y + kStemPadTop)) * stem_pitch + kStemPadLeft);
  for (int ch = blockIdx.x * 4 + warp; ch < chunks; ch += gridDim.x * 4) {
    if (lane < 24) *reinterpret_cast<uint4*>(&stage[warp][lane * 16]) = __ldg(reinterpret_cast<const uint4*>(row + (size_t)ch * 384) + lane);
    __syncwarp();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const uint16_t* sp = reinterpret_cast<const uint16_t*>(&stage[warp][hh * 192 + lane * 6]);   // b0 g0 | r0 b1 | g1 r1
      const uint32_t h0 = sp[0], h1 = sp[1], h2 = sp[2];
      const uint8_t b0 = (uint8_t)h0, g0 = (uint8_t)(h0 >> 8), r0 = (uint8_t)h1, b1 = (uint8_t)(h1 >> 8), g1 = (uint8_t)h2, r1 = (uint8_t)(h2 >> 8);
      const uint8_t p0c0 = a.bgr_order ? b0 : r0, p0c2 = a.bgr_order ? r0 : b0, p1c0 = a.bgr_order ? b1 : r1, p1c2 = a.bgr_order ? r1 : b1;
      const __half zero = __ushort_as_half((unsigned short)0);
      __half2 o0 = __halves2half2(lut[p0c0], lut[256 + g0]), o1 = __halves2half2(lut[512 + p0c2], zero);
      __half2 o2 = __halves2half2(lut[p1c0], lut[256 + g1]), o3 = __halves2half2(lut[512 + p1c2], zero);
      drow[(size_t)ch * 64 + hh * 32 + lane] = make_uint4(*reinterpret_cast<uint32_t*>(&o0), *reinterpret_cast<uint32_t*>(&o1),
                                                          *reinterpret_cast<uint32_t*>(&o2), *reinterpret_cast<uint32_t*>(&o3));
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) preprocess_f32_kernel(const uint8_t* __restrict__ bgr, int h, int w, const float* __restrict__ lut_f,
                                                              float* __restrict__ out) {
  __shared__ float lut[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = lut_f[i];
  __syncthreads();
  const size_t npix = (size_t)h * w;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const uint8_t* p = bgr + i * 3;
  out[i] = lut[p[2]];
  out[npix + i] = lut[256 + p[1]];
  out[2 * npix + i] = lut[512 + p[0]];
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_kernel(const __half* __restrict__ in, __half* __restrict__ out, int n, int h, int w, int c8,
                                                       int oh, int ow, int k, int stride, int pad) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)n * oh * ow * c8;
  if (idx >= total) return;
  const int cg = (int)(idx % c8);
  size_t t = idx / c8;
  const int ox = (int)(t % ow); t /= ow;
  const int oy = (int)(t % oh);
  const int img = (int)(t / oh);
  const __half2 ninf = __float2half2_rn(-65504.f);
  __half2 m[4] = {ninf, ninf, ninf, ninf};
  bool any = false;
  for (int ky = 0; ky < k; ++ky) {
    const int iy = oy * stride + ky - pad;
    if (iy < 0 || iy >= h) continue;
    for (int kx = 0; kx < k; ++kx) {
      const int ix = ox * stride + kx - pad;
      if (ix < 0 || ix >= w) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)img * h + iy) * w + ix) * (size_t)(c8 * 8)) + cg);
      const __half2* hv = reinterpret_cast<const __half2*>(&v);
      if (!any) { m[0] = hv[0]; m[1] = hv[1]; m[2] = hv[2]; m[3] = hv[3]; any = true; }
      else {
#pragma unroll
        for (int q = 0; q < 4; ++q) m[q] = __hmax2(m[q], hv[q]);
      }
    }
  }
  uint4 o;
  __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int q = 0; q < 4; ++q) ho[q] = m[q];
  reinterpret_cast<uint4*>(out)[idx] = o;
}

// The FCN stem pool (3x3 / stride 2 / pad 1), row-streaming: thread = (8 channels, output column) walks down a strip of
// output rows; the horizontal maxima of input row 2*oy+1 serve output rows oy and oy+1 from registers, and the column
// overlap between neighbouring threads hits L1, so every input element crosses the L2->SM port about once instead of
// 2.25 times (the gather kernel above is bound by exactly that port).  Same __hmax2 reductions: identical results.
constexpr int kPoolStrip = 16;
__global__ void __launch_bounds__(256) maxpool3s2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int h, int w, int c8, int oh, int ow) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ow * c8) return;
  const int cg = idx % c8, ox = idx / c8;
  const int img = blockIdx.z;
  const int oy0 = blockIdx.y * kPoolStrip, oy1 = min(oy0 + kPoolStrip, oh);
  const uint4* src = reinterpret_cast<const uint4*>(in) + (size_t)img * h * w * c8 + cg;
  uint4* dst = reinterpret_cast<uint4*>(out) + (size_t)img * oh * ow * c8 + (size_t)ox * c8 + cg;
  const __half2 ninf = __halves2half2(__ushort_as_half((unsigned short)0xFC00), __ushort_as_half((unsigned short)0xFC00));
  const int ix = 2 * ox;
  auto hrow = [&](int iy, __half2 (&r)[4]) {
    const uint4* row = src + (size_t)iy * w * c8;
    const uint4 v1 = __ldg(row + (size_t)ix * c8);
    const __half2* h1 = reinterpret_cast<const __half2*>(&v1);
#pragma unroll
    for (int q = 0; q < 4; ++q) r[q] = h1[q];
    if (ix > 0) {
      const uint4 v0 = __ldg(row + (size_t)(ix - 1) * c8);
      const __half2* h0 = reinterpret_cast<const __half2*>(&v0);
#pragma unroll
      for (int q = 0; q < 4; ++q) r[q] = __hmax2(r[q], h0[q]);
    }
    if (ix + 1 < w) {
      const uint4 v2 = __ldg(row + (size_t)(ix + 1) * c8);
      const __half2* h2 = reinterpret_cast<const __half2*>(&v2);
#pragma unroll
      for (int q = 0; q < 4; ++q) r[q] = __hmax2(r[q], h2[q]);
    }
  };
  __half2 prev[4] = {ninf, ninf, ninf, ninf};
  if (oy0 > 0) hrow(2 * oy0 - 1, prev);
  for (int oy = oy0; oy < oy1; ++oy) {
    __half2 a[4], b[4] = {ninf, ninf, ninf, ninf};
    hrow(2 * oy, a);
    if (2 * oy + 1 < h) hrow(2 * oy + 1, b);
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) { ho[q] = __hmax2(__hmax2(prev[q], a[q]), b[q]); prev[q] = b[q]; }
    dst[(size_t)oy * ow * c8] = o;
  }
}

// The same pool on u8 tensors (int8 plans: the raw quantised values; max commutes with the affine quantisation map):
// thread = (16 channels, output column), __vmaxu4 on packed bytes.
__global__ void __launch_bounds__(256) maxpool3s2_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int h, int w, int c16, int oh, int ow) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ow * c16) return;
  const int cg = idx % c16, ox = idx / c16;
  const int img = blockIdx.z;
  const int oy0 = blockIdx.y * kPoolStrip, oy1 = min(oy0 + kPoolStrip, oh);
  const uint4* src = reinterpret_cast<const uint4*>(in) + (size_t)img * h * w * c16 + cg;
  uint4* dst = reinterpret_cast<uint4*>(out) + (size_t)img * oh * ow * c16 + (size_t)ox * c16 + cg;
  const int ix = 2 * ox;
  auto vmax = [](uint4 a, uint4 b) { return make_uint4(__vmaxu4(a.x, b.x), __vmaxu4(a.y, b.y), __vmaxu4(a.z, b.z), __vmaxu4(a.w, b.w)); };
  auto hrow = [&](int iy) {
    const uint4* row = src + (size_t)iy * w * c16;
    uint4 r = __ldg(row + (size_t)ix * c16);
    if (ix > 0) r = vmax(r, __ldg(row + (size_t)(ix - 1) * c16));
    if (ix + 1 < w) r = vmax(r, __ldg(row + (size_t)(ix + 1) * c16));
    return r;
  };
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);   // identity of max on u8
  uint4 prev = oy0 > 0 ? hrow(2 * oy0 - 1) : zero;
  for (int oy = oy0; oy < oy1; ++oy) {
    const uint4 a = hrow(2 * oy);
    const uint4 b = 2 * oy + 1 < h ? hrow(2 * oy + 1) : zero;
    dst[(size_t)oy * ow * c16] = vmax(vmax(prev, a), b);
    prev = b;
  }
}

// ------------------------------------------------------------------------------------------------
// K5: the network's final Resize(linear, half_pixel) fused with ColorCode::advance
// (infur/src/decode_predict.rs:53-79) and color_code (:32-36).  One CTA = one 32 x 32 output tile:
//   phase 0  copy the low-res patch the tile touches into smem (coalesced 128-bit loads)
//   phase 1  horizontal interpolation  t[k][lr][x] = lx0*L[lr][x0][k] + lx1*L[lr][x1][k]   (once per low-res row)
//   phase 2  per pixel: v = ly0*t[k][y0][x] + ly1*t[k][y1][x]; strict '>' scan from (0, 0.0); alpha = trunc(sat(c*255));
//            colour from the 20x256 premultiplied table; optional "over" blend and BGR->RGBA of the frame.
// The full-resolution logits (174 MB per 1080p frame in the reference) are never materialised.
constexpr int kPostTile = 32;
// smem pixel stride of the low-res patch: odd, so that neighbouring low-res pixels fall into different banks
__host__ __device__ inline int post_pad(int k) { return (((k + 3) >> 2) << 2) | 1; }

__global__ void __launch_bounds__(128) post_kernel(PostArgs a) {
  extern __shared__ float sm[];
  const int X0 = blockIdx.x * kPostTile, Y0 = blockIdx.y * kPostTile, img = blockIdx.z;
  const int X1 = min(X0 + kPostTile, a.ow) - 1, Y1 = min(Y0 + kPostTile, a.oh) - 1;
  const int lc0 = a.x0[X0], lc1 = a.x1[X1], lr0 = a.y0[Y0], lr1 = a.y1[Y1];
  const int ncol = lc1 - lc0 + 1, nrow = lr1 - lr0 + 1;
  const int kPostPad = post_pad(a.k);
  float* patch = sm;                                   // [nrow][ncol][kPostPad]
  float* hor = sm + (size_t)a.max_lr * a.max_lc * kPostPad;  // [k][nrow][32]
  const int tid = threadIdx.x;
  const int kq = (a.k + 3) >> 2;  // float4 per low-res pixel that carry classes
  for (int i = tid; i < nrow * ncol * kq; i += blockDim.x) {
    const int q = i % kq;
    const int pc = (i / kq) % ncol;
    const int pr = i / (kq * ncol);
    const float4 v = __ldg(reinterpret_cast<const float4*>(a.lowres + (((size_t)img * a.lh + (lr0 + pr)) * a.lw + (lc0 + pc)) * a.ldk) + q);
    float* d = patch + (pr * ncol + pc) * kPostPad + q * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  const int x = X0 + lane;
  const bool xin = x < a.ow;
  const int xi = xin ? x : a.ow - 1;
  {
    const int c0 = a.x0[xi] - lc0, c1 = a.x1[xi] - lc0;
    const float w0 = a.lx0[xi], w1 = a.lx1[xi];
    for (int i = warp; i < a.k * nrow; i += 4) {
      const int k = i % a.k, pr = i / a.k;
      const float* rowp = patch + (size_t)pr * ncol * kPostPad + k;
      const float t = __fadd_rn(__fmul_rn(w0, rowp[c0 * kPostPad]), __fmul_rn(w1, rowp[c1 * kPostPad]));
      hor[((size_t)k * nrow + pr) * kPostTile + lane] = t;
    }
  }
  __syncthreads();
  if (!xin) return;
  const size_t plane = (size_t)a.oh * a.ow;
  for (int ry = warp * 8; ry < warp * 8 + 8; ++ry) {
    const int y = Y0 + ry;
    if (y >= a.oh) break;
    const int r0 = a.y0[y] - lr0, r1 = a.y1[y] - lr0;
    const float w0 = a.ly0[y], w1 = a.ly1[y];
    int k_max = 0;
    float c_max = a.softmax ? -INFINITY : 0.f;
    const size_t pix = (size_t)y * a.ow + x;
    for (int k = 0; k < a.k; ++k) {
      const float* hk = hor + (size_t)k * nrow * kPostTile + lane;
      const float v = __fadd_rn(__fmul_rn(w0, hk[r0 * kPostTile]), __fmul_rn(w1, hk[r1 * kPostTile]));
      if (a.logits) a.logits[((size_t)img * a.k + k) * plane + pix] = v;
      if (v > c_max) { k_max = k; c_max = v; }
    }
    if (a.softmax) {
      // confidence = softmax probability of the winning class: p = 1 / sum_k exp(v_k - max) (the winner's own term is exp(0) = 1);
      // every probability is > 0, so ColorCode's strict '>' scan from (0, 0.0) picks the first maximum of the logits
      float sum = 0.f;
      for (int k = 0; k < a.k; ++k) {
        const float* hk = hor + (size_t)k * nrow * kPostTile + lane;
        const float v = __fadd_rn(__fmul_rn(w0, hk[r0 * kPostTile]), __fmul_rn(w1, hk[r1 * kPostTile]));
        sum = __fadd_rn(sum, expf(__fsub_rn(v, c_max)));
      }
      c_max = c_max == c_max && sum > 0.f ? __fdiv_rn(1.0f, sum) : 0.f;
      if (!(c_max > 0.f)) k_max = 0;   // NaN logits: nothing beats (0, 0.0)
    }
    const float av = __fmul_rn(c_max, 255.0f);
    const int alpha = av >= 255.0f ? 255 : (int)av;  // c_max >= 0 always; trunc toward zero, saturate
    const uint32_t col = __ldg(a.color_lut + (k_max % 20) * 256 + alpha);
    const size_t gp = (size_t)img * plane + pix;
    if (a.class_map) a.class_map[gp] = (uint8_t)k_max;
    a.decoded[gp] = col;
    if (a.frame_bgr && (a.blended || a.frame_rgba)) {
      const uint8_t* f = a.frame_bgr + gp * 3;
      const uint32_t fb = f[0], fg = f[1], fr = f[2];
      if (a.frame_rgba) a.frame_rgba[gp] = fr | (fg << 8) | (fb << 16) | 0xff000000u;
      if (a.blended) {
        const uint32_t ia = 255u - (col >> 24);
        const uint32_t r = min(255u, (col & 0xff) + (fr * ia + 127u) / 255u);
        const uint32_t g = min(255u, ((col >> 8) & 0xff) + (fg * ia + 127u) / 255u);
        const uint32_t b = min(255u, ((col >> 16) & 0xff) + (fb * ia + 127u) / 255u);
        a.blended[gp] = r | (g << 8) | (b << 16) | 0xff000000u;
      }
    }
  }
}

// Strip variant of K5 for a compile-time class count K (the network path, K = 21): one warp = 32 output columns x
// kPostStripRows rows, lane = column.  The horizontally interpolated logits of the two low-res rows the current
// output row blends (top[k], bot[k]) live in registers and are refreshed only when the low-res row pair changes
// (every 8 output rows at the network's x8 upsample); when the pair slides down by one, bot becomes top without a
// reload.  Same f32 operations in the same order as post_kernel / the oracle, hence the same bits.
constexpr int kPostStripRows = 64;

// alpha = trunc(sat(c * 255)), colour from the premultiplied table, class byte, optional blend / frame RGBA
__device__ __forceinline__ void post_emit(const PostArgs& a, int img, size_t plane, size_t pix, int k_max, float c_max) {
  const float av = __fmul_rn(c_max, 255.0f);
  const int alpha = av >= 255.0f ? 255 : (int)av;  // c_max >= 0 always; trunc toward zero, saturate
  const uint32_t col = __ldg(a.color_lut + (k_max % 20) * 256 + alpha);
  const size_t gp = (size_t)img * plane + pix;
  if (a.class_map) a.class_map[gp] = (uint8_t)k_max;
  a.decoded[gp] = col;
  if (a.frame_bgr && (a.blended || a.frame_rgba)) {
    const uint8_t* f = a.frame_bgr + gp * 3;
    const uint32_t fb = f[0], fg = f[1], fr = f[2];
    if (a.frame_rgba) a.frame_rgba[gp] = fr | (fg << 8) | (fb << 16) | 0xff000000u;
    if (a.blended) {
      const uint32_t ia = 255u - (col >> 24);
      const uint32_t r = min(255u, (col & 0xff) + (fr * ia + 127u) / 255u);
      const uint32_t g = min(255u, ((col >> 8) & 0xff) + (fg * ia + 127u) / 255u);
      const uint32_t b = min(255u, ((col >> 16) & 0xff) + (fb * ia + 127u) / 255u);
      a.blended[gp] = r | (g << 8) | (b << 16) | 0xff000000u;
    }
  }
}

// Pre-pass of the fast path below: per low-res pixel the strict first-maximum class k* of its K logits, kept only
// when it beats every other class by a margin far above the rounding error of the interpolation
// (1e-4 * (1 + |v1| + |v2|) vs. a few f32 ulp); -1 otherwise.
template <int K>
__global__ void __launch_bounds__(256) lowres_top_kernel(const float* __restrict__ lowres, int ldk, size_t npix, int32_t* __restrict__ code) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const float* p = lowres + i * ldk;
  int k1 = 0;
  float v1 = p[0], v2 = -INFINITY;
#pragma unroll
  for (int k = 1; k < K; ++k) {
    const float v = p[k];
    if (v > v1) { v2 = v1; v1 = v; k1 = k; }
    else if (v > v2 || !(v2 == v2)) v2 = v;
  }
  const bool safe = (v1 - v2) > 1e-4f * (1.f + fabsf(v1) + fabsf(v2));   // false for NaN / inf oddities: those pixels take the full path
  code[i] = safe ? k1 : -1;
}

template <int K, bool SOFTMAX = false>
__global__ void __launch_bounds__(128, 3) post_strip_kernel(PostArgs a) {
  constexpr int KQ = (K + 3) / 4;     // float4 per low-res pixel that carry classes
  constexpr int KP = KQ * 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x = blockIdx.x * 32 + lane;
  const int Y0 = (blockIdx.y * 4 + warp) * kPostStripRows;
  const int img = blockIdx.z;
  if (Y0 >= a.oh) return;
  const int Y1 = min(Y0 + kPostStripRows, a.oh);
  const bool xin = x < a.ow;
  const int xi = xin ? x : a.ow - 1;
  const int c0 = __ldg(a.x0 + xi), c1 = __ldg(a.x1 + xi);
  const float wx0 = __ldg(a.lx0 + xi), wx1 = __ldg(a.lx1 + xi);
  float top[KP], bot[KP];
  auto load_row = [&](int r, float* dst) {
    const float4* p0 = reinterpret_cast<const float4*>(a.lowres + (((size_t)img * a.lh + r) * a.lw + c0) * a.ldk);
    const float4* p1 = reinterpret_cast<const float4*>(a.lowres + (((size_t)img * a.lh + r) * a.lw + c1) * a.ldk);
#pragma unroll
    for (int q = 0; q < KQ; ++q) {
      const float4 u = __ldg(p0 + q), v = __ldg(p1 + q);
      dst[4 * q + 0] = __fadd_rn(__fmul_rn(wx0, u.x), __fmul_rn(wx1, v.x));
      dst[4 * q + 1] = __fadd_rn(__fmul_rn(wx0, u.y), __fmul_rn(wx1, v.y));
      dst[4 * q + 2] = __fadd_rn(__fmul_rn(wx0, u.z), __fmul_rn(wx1, v.z));
      dst[4 * q + 3] = __fadd_rn(__fmul_rn(wx0, u.w), __fmul_rn(wx1, v.w));
    }
  };
  int cur0 = -1, cur1 = -1;
  const size_t plane = (size_t)a.oh * a.ow;
  // fast path state (per lane, refreshed when the low-res row pair changes)
  int f0 = -2, f1 = -2, fk = -1;
  float ft = 0.f, fb = 0.f;
  for (int y = Y0; y < Y1; ++y) {
    const int r0 = __ldg(a.y0 + y), r1 = __ldg(a.y1 + y);   // warp-uniform
    if (!SOFTMAX && a.top_code && !a.logits) {
      // Fast path: when ONE class wins all four low-res pixels a lane's output pixel blends, by a safe margin, it wins
      // the blend too (each interpolation step is monotone in its inputs, and the margin dwarfs the rounding error),
      // so only that class needs interpolating -- with the very same operations, hence the same bits.
      if (r0 != f0 || r1 != f1) {
        f0 = r0; f1 = r1;
        const size_t b0 = ((size_t)img * a.lh + r0) * a.lw, b1 = ((size_t)img * a.lh + r1) * a.lw;
        const int k00 = __ldg(a.top_code + b0 + c0), k01 = __ldg(a.top_code + b0 + c1);
        const int k10 = __ldg(a.top_code + b1 + c0), k11 = __ldg(a.top_code + b1 + c1);
        fk = (k00 >= 0 && k00 == k01 && k00 == k10 && k00 == k11) ? k00 : -1;
        if (fk >= 0) {
          ft = __fadd_rn(__fmul_rn(wx0, __ldg(a.lowres + (b0 + c0) * a.ldk + fk)), __fmul_rn(wx1, __ldg(a.lowres + (b0 + c1) * a.ldk + fk)));
          fb = __fadd_rn(__fmul_rn(wx0, __ldg(a.lowres + (b1 + c0) * a.ldk + fk)), __fmul_rn(wx1, __ldg(a.lowres + (b1 + c1) * a.ldk + fk)));
        }
      }
      if (__all_sync(0xffffffffu, fk >= 0)) {
        const float wy0 = __ldg(a.ly0 + y), wy1 = __ldg(a.ly1 + y);
        const float v = __fadd_rn(__fmul_rn(wy0, ft), __fmul_rn(wy1, fb));
        const bool pos = v > 0.f;                       // the scan starts from (class 0, 0.0): nothing <= 0 replaces it
        const int k_max = pos ? fk : 0;
        const float c_max = pos ? v : 0.f;
        if (xin) post_emit(a, img, plane, (size_t)y * a.ow + x, k_max, c_max);
        continue;
      }
    }
    if (r0 != cur0) {
      if (r0 == cur1) {
#pragma unroll
        for (int k = 0; k < KP; ++k) top[k] = bot[k];
      } else {
        load_row(r0, top);
      }
      cur0 = r0;
    }
    if (r1 != cur1) {
      if (r1 == cur0) {
#pragma unroll
        for (int k = 0; k < KP; ++k) bot[k] = top[k];
      } else {
        load_row(r1, bot);
      }
      cur1 = r1;
    }
    const float wy0 = __ldg(a.ly0 + y), wy1 = __ldg(a.ly1 + y);
    // ColorCode's scan (strict '>', first maximum wins, start (0, 0.0)) split into 3 contiguous class blocks that
    // run as independent dependency chains and are merged left to right with the same strict '>': identical result
    // (blocks 1, 2 start from -inf, which no value beats or ties differently than in one chain; a NaN never wins
    // either way).  The running maximum is fmaxf (keeps the old value on NaN and on ties, like a failed '>').
    constexpr int KB = (K + 2) / 3;
    int bk[3] = {0, KB, 2 * KB};
    float bv[3] = {SOFTMAX ? -INFINITY : 0.f, -INFINITY, -INFINITY};
    const size_t pix = (size_t)y * a.ow + x;
#pragma unroll
    for (int i = 0; i < KB; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int k = j * KB + i;
        if (k < K) {
          const float v = __fadd_rn(__fmul_rn(wy0, top[k]), __fmul_rn(wy1, bot[k]));
          if (a.logits && xin) a.logits[((size_t)img * K + k) * plane + pix] = v;
          const bool gt = v > bv[j];
          bk[j] = gt ? k : bk[j];
          bv[j] = fmaxf(bv[j], v);
        }
      }
    }
    int k_max = bk[0];
    float c_max = bv[0];
#pragma unroll
    for (int j = 1; j < 3; ++j) {
      const bool gt = bv[j] > c_max;
      k_max = gt ? bk[j] : k_max;
      c_max = fmaxf(c_max, bv[j]);
    }
    if (SOFTMAX) {
      // see post_kernel: p(winner) = 1 / sum_k exp(v_k - max), classes summed in index order
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float v = __fadd_rn(__fmul_rn(wy0, top[k]), __fmul_rn(wy1, bot[k]));
        sum = __fadd_rn(sum, expf(__fsub_rn(v, c_max)));
      }
      c_max = c_max == c_max && sum > 0.f ? __fdiv_rn(1.0f, sum) : 0.f;
      if (!(c_max > 0.f)) k_max = 0;
    }
    if (xin) post_emit(a, img, plane, pix, k_max, c_max);
  }
}


// Cell-per-thread variant of K5 for K = 21 (the network path).  A "cell" is the block of output pixels that blend the same four
// low-res logit vectors (8 x 8 pixels at the network's x8 upsample; 12 wide / tall along the top-left border, 4 along the
// bottom-right one).  One thread owns one cell:
//   1. candidate pruning: class k can win somewhere in the cell only if its largest corner value reaches L = max_k min_corner(k)
//      (class argmax-of-min is >= L everywhere in the cell, every blend is a convex combination of the corners) and is not
//      clearly negative (the scan starts from (0, 0.0)); the margin 1e-5 * (1 + |.|) is ~50x the rounding error of the three
//      un-fused f32 operations of a blend.  Typically 1-4 of the 21 classes survive.  Non-finite corners: all classes stay.
//   2. for each surviving class, in class order: the horizontal blends of the cell's columns once, then per pixel the vertical
//      blend and ColorCode's strict '>' update -- the very same f32 operations in the same order as the strip / generic kernels
//      and the oracle, so the same bits; pruned classes are strictly below the winner and cannot change the scan's result.
// ~35 instructions per pixel instead of ~130.  Only class map + decoded RGBA are written here; frame RGBA / blend are a
// separate streaming pass (frame_blend_kernel).
template <int K, int ROWS>
__global__ void __launch_bounds__(128, 4) post_cell_kernel(PostArgs a) {
  constexpr int KQ = (K + 3) / 4;
  const int cells = a.lh * a.lw;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int img = blockIdx.y;
  if (idx >= cells) return;
  const int r = idx / a.lw, c = idx - r * a.lw;
  const int X0 = __ldg(a.xs + c), X1 = __ldg(a.xs + c + 1), Y0 = __ldg(a.ys + r), Y1 = __ldg(a.ys + r + 1);
  if (X0 >= X1 || Y0 >= Y1) return;
  const int c1 = __ldg(a.x1 + X0), r1 = __ldg(a.y1 + Y0);
  const float* p00 = a.lowres + (((size_t)img * a.lh + r) * a.lw + c) * a.ldk;
  const float* p01 = a.lowres + (((size_t)img * a.lh + r) * a.lw + c1) * a.ldk;
  const float* p10 = a.lowres + (((size_t)img * a.lh + r1) * a.lw + c) * a.ldk;
  const float* p11 = a.lowres + (((size_t)img * a.lh + r1) * a.lw + c1) * a.ldk;
  // ---- 1. candidate classes
  float mx[KQ * 4];
  float L = -INFINITY;
  bool odd = false;
#pragma unroll
  for (int q = 0; q < KQ; ++q) {
    const float4 u0 = __ldg(reinterpret_cast<const float4*>(p00) + q), u1 = __ldg(reinterpret_cast<const float4*>(p01) + q);
    const float4 u2 = __ldg(reinterpret_cast<const float4*>(p10) + q), u3 = __ldg(reinterpret_cast<const float4*>(p11) + q);
    const float v0[4] = {u0.x, u0.y, u0.z, u0.w}, v1[4] = {u1.x, u1.y, u1.z, u1.w}, v2[4] = {u2.x, u2.y, u2.z, u2.w}, v3[4] = {u3.x, u3.y, u3.z, u3.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = 4 * q + j;
      if (k < K) {
        const float hi = fmaxf(fmaxf(v0[j], v1[j]), fmaxf(v2[j], v3[j])), lo = fminf(fminf(v0[j], v1[j]), fminf(v2[j], v3[j]));
        odd |= !(fabsf(v0[j]) <= 3.0e38f) || !(fabsf(v1[j]) <= 3.0e38f) || !(fabsf(v2[j]) <= 3.0e38f) || !(fabsf(v3[j]) <= 3.0e38f);
        mx[k] = hi;
        L = fmaxf(L, lo);
      } else {
        mx[k] = -INFINITY;
      }
    }
  }
  uint32_t mask = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float eps = 1e-5f * (1.f + fabsf(L) + fabsf(mx[k]));
    if (mx[k] >= L - eps && mx[k] >= -eps) mask |= 1u << k;
  }
  if (odd) mask = (1u << K) - 1u;
  const size_t plane = (size_t)a.oh * a.ow;
  // ---- 2. chunks of ROWS rows x 8 columns (the per-pixel scan state lives in registers: fewer rows = more resident warps)
  for (int ya = Y0; ya < Y1; ya += ROWS) {
    float wy0[ROWS], wy1[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) { const int y = min(ya + i, Y1 - 1); wy0[i] = __ldg(a.ly0 + y); wy1[i] = __ldg(a.ly1 + y); }
    for (int xa = X0; xa < X1; xa += 8) {
      float wx0[8], wx1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { const int x = min(xa + j, X1 - 1); wx0[j] = __ldg(a.lx0 + x); wx1[j] = __ldg(a.lx1 + x); }
      float cm[ROWS][8];
      int km[ROWS][8];
#pragma unroll
      for (int i = 0; i < ROWS; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { cm[i][j] = 0.f; km[i][j] = 0; }
      for (uint32_t m = mask; m != 0; m &= m - 1) {
        const int k = __ffs((int)m) - 1;
        const float a00 = __ldg(p00 + k), a01 = __ldg(p01 + k), a10 = __ldg(p10 + k), a11 = __ldg(p11 + k);
        float top[8], bot[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          top[j] = __fadd_rn(__fmul_rn(wx0[j], a00), __fmul_rn(wx1[j], a01));
          bot[j] = __fadd_rn(__fmul_rn(wx0[j], a10), __fmul_rn(wx1[j], a11));
        }
#pragma unroll
        for (int i = 0; i < ROWS; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float v = __fadd_rn(__fmul_rn(wy0[i], top[j]), __fmul_rn(wy1[i], bot[j]));
            const bool gt = v > cm[i][j];
            km[i][j] = gt ? k : km[i][j];
            cm[i][j] = gt ? v : cm[i][j];
          }
      }
      // emit: alpha = trunc(sat(c * 255)), colour from the premultiplied table
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        const int y = ya + i;
        if (y >= Y1) break;
        const size_t rowp = (size_t)img * plane + (size_t)y * a.ow;
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
          const int xg = xa + 4 * gq;
          if (xg >= X1) break;
          uint32_t col[4];
          uint32_t cls = 0;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int j = 4 * gq + t;
            const float av = __fmul_rn(cm[i][j], 255.0f);
            const int alpha = av >= 255.0f ? 255 : (int)av;
            const int kk = km[i][j];
            col[t] = __ldg(a.color_lut + (kk >= 20 ? kk - 20 : kk) * 256 + alpha);
            cls |= (uint32_t)kk << (8 * t);
          }
          const size_t gp = rowp + (size_t)xg;
          if (xg + 4 <= X1 && ((gp & 3) == 0)) {
            *reinterpret_cast<uint4*>(a.decoded + gp) = make_uint4(col[0], col[1], col[2], col[3]);
            if (a.class_map) *reinterpret_cast<uint32_t*>(a.class_map + gp) = cls;
          } else {
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (xg + t < X1) {
                a.decoded[gp + t] = col[t];
                if (a.class_map) a.class_map[gp + t] = (uint8_t)(cls >> (8 * t));
              }
          }
        }
      }
    }
  }
}

// Display buffer (app.rs:132-144) and the optional "over" blend as one streaming pass: 4 pixels per thread.
__global__ void __launch_bounds__(256) frame_blend_kernel(const uint8_t* __restrict__ bgr, const uint32_t* __restrict__ decoded, size_t npix,
                                                           uint32_t* __restrict__ frame_rgba, uint32_t* __restrict__ blended) {
  const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= npix) return;
  auto one = [&](size_t i, uint32_t fb, uint32_t fg, uint32_t fr) {
    if (frame_rgba) frame_rgba[i] = fr | (fg << 8) | (fb << 16) | 0xff000000u;
    if (blended) {
      const uint32_t col = decoded[i];
      const uint32_t ia = 255u - (col >> 24);
      const uint32_t r = min(255u, (col & 0xff) + (fr * ia + 127u) / 255u);
      const uint32_t g = min(255u, ((col >> 8) & 0xff) + (fg * ia + 127u) / 255u);
      const uint32_t b = min(255u, ((col >> 16) & 0xff) + (fb * ia + 127u) / 255u);
      blended[i] = r | (g << 8) | (b << 16) | 0xff000000u;
    }
  };
  if (i4 + 4 <= npix && !blended && (reinterpret_cast<uintptr_t>(bgr + i4 * 3) & 3) == 0) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(bgr + i4 * 3);   // 12 bytes, 4-byte aligned (i4 % 4 == 0)
    const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
    // bytes: b0 g0 r0 b1 | g1 r1 b2 g2 | r2 b3 g3 r3
    const uint32_t o0 = ((w0 >> 16) & 0xff) | (w0 & 0xff00) | ((w0 & 0xff) << 16) | 0xff000000u;
    const uint32_t o1 = ((w1 >> 8) & 0xff) | ((w1 & 0xff) << 8) | ((w0 >> 24) << 16) | 0xff000000u;
    const uint32_t o2 = (w2 & 0xff) | ((w1 >> 24) << 8) | (((w1 >> 16) & 0xff) << 16) | 0xff000000u;
    const uint32_t o3 = (w2 >> 24) | (((w2 >> 16) & 0xff) << 8) | (((w2 >> 8) & 0xff) << 16) | 0xff000000u;
    *reinterpret_cast<uint4*>(frame_rgba + i4) = make_uint4(o0, o1, o2, o3);
    return;
  }
  for (size_t i = i4; i < npix && i < i4 + 4; ++i) { const uint8_t* f = bgr + i * 3; one(i, f[0], f[1], f[2]); }
}

__global__ void __launch_bounds__(256) color_code_kernel(const float* __restrict__ hm, int k, size_t npix, const uint32_t* __restrict__ lut,
                                                          uint32_t* __restrict__ rgba, uint8_t* __restrict__ class_map) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  int k_max = 0;
  float c_max = 0.f;
  for (int c = 0; c < k; ++c) {
    const float v = __ldg(hm + (size_t)c * npix + i);
    if (v > c_max) { k_max = c; c_max = v; }
  }
  const float av = __fmul_rn(c_max, 255.0f);
  const int alpha = av >= 255.0f ? 255 : (int)av;
  rgba[i] = __ldg(lut + (k_max % 20) * 256 + alpha);
  if (class_map) class_map[i] = (uint8_t)k_max;
}

__global__ void __launch_bounds__(256) frame_rgba_kernel(const uint8_t* __restrict__ bgr, size_t npix, uint32_t* __restrict__ rgba) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const uint8_t* f = bgr + i * 3;
  rgba[i] = (uint32_t)f[2] | ((uint32_t)f[1] << 8) | ((uint32_t)f[0] << 16) | 0xff000000u;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) direct_conv_kernel(DirectConvArgs a) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)a.n * a.oh * a.ow * a.cout;
  if (idx >= total) return;
  const int co = (int)(idx % a.cout);
  size_t t = idx / a.cout;
  const int ox = (int)(t % a.ow); t /= a.ow;
  const int oy = (int)(t % a.oh);
  const int img = (int)(t / a.oh);
  float acc = 0.f;
  for (int ky = 0; ky < a.kh; ++ky) {
    const int iy = oy * a.stride + ky * a.dil - a.pad;
    if (iy < 0 || iy >= a.h) continue;
    for (int kx = 0; kx < a.kw; ++kx) {
      const int ix = ox * a.stride + kx * a.dil - a.pad;
      if (ix < 0 || ix >= a.wd) continue;
      const __half* xp = a.x + (((size_t)img * a.x_rows + (iy + a.x_off_y)) * a.x_pitch_px + (ix + a.x_off_x)) * a.x_c;
      const __half* wp = a.w + (((size_t)co * a.kh + ky) * a.kw + kx) * a.cin;
      for (int ci = 0; ci < a.cin; ++ci) acc = fmaf(__half2float(xp[ci]), __half2float(wp[ci]), acc);
    }
  }
  acc += a.bias[co];
  const size_t o = (((size_t)img * a.oh + oy) * a.ow + ox) * a.out_ld + co;
  if (a.residual) acc += __half2float(a.residual[o]);
  if (a.relu) acc = fmaxf(acc, 0.f);
  if (a.y_f32) a.y_f32[o] = acc;
  else a.y[o] = __float2half_rn(acc);
}

}  // namespace

cudaError_t launch_pre(const PreArgs& a, cudaStream_t s) {
  const int pitch = stem_pitch_px(a.ow), rows = stem_rows(a.oh);
  const bool unit = a.xmap == nullptr && a.ymap == nullptr && a.bx0 == nullptr && a.scaled_bgr == nullptr && a.stem_in != nullptr && (a.w % 4) == 0 &&
                    a.oh == a.h && a.ow == a.w;
  if (unit && (a.w % 128) == 0 && (reinterpret_cast<uintptr_t>(a.src) & 15) == 0) {
    dim3 grid(((a.w >> 7) + 3) / 4, a.h, a.n);
    pre_unit_warp128_kernel<<<grid, 128, 0, s>>>(a, pitch, rows);
  } else if (unit) {
    dim3 grid((a.w / 4 + 255) / 256, a.h, a.n);
    pre_unit_vec4_kernel<<<grid, 256, 0, s>>>(a, pitch, rows);
  } else {
    dim3 grid((a.ow + 255) / 256, a.oh, a.n);
    pre_generic_kernel<<<grid, 256, 0, s>>>(a, pitch, rows);
  }
  return cudaGetLastError();
}

cudaError_t launch_preprocess_f32(const uint8_t* bgr, int h, int w, const float* lut_f, float* out, cudaStream_t s) {
  const size_t npix = (size_t)h * w;
  preprocess_f32_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(bgr, h, w, lut_f, out);
  return cudaGetLastError();
}

cudaError_t launch_maxpool(const __half* in, __half* out, int n, int h, int w, int c, int oh, int ow, int k, int stride, int pad,
                           cudaStream_t s) {
  if (k == 3 && stride == 2 && pad == 1 && oh == (h - 1) / 2 + 1 && ow == (w - 1) / 2 + 1 && n <= 65535) {
    const int c8 = c / 8;
    dim3 grid((unsigned)((ow * c8 + 255) / 256), (unsigned)((oh + kPoolStrip - 1) / kPoolStrip), (unsigned)n);
    maxpool3s2_kernel<<<grid, 256, 0, s>>>(in, out, h, w, c8, oh, ow);
    return cudaGetLastError();
  }
  const size_t total = (size_t)n * oh * ow * (c / 8);
  maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, out, n, h, w, c / 8, oh, ow, k, stride, pad);
  return cudaGetLastError();
}

cudaError_t launch_maxpool3s2_u8(const uint8_t* in, uint8_t* out, int n, int h, int w, int c, int oh, int ow, cudaStream_t s) {
  const int c16 = c / 16;
  dim3 grid((unsigned)((ow * c16 + 255) / 256), (unsigned)((oh + kPoolStrip - 1) / kPoolStrip), (unsigned)n);
  maxpool3s2_u8_kernel<<<grid, 256, 0, s>>>(in, out, h, w, c16, oh, ow);
  return cudaGetLastError();
}

// which kernel(s) launch_post uses for these arguments
static bool post_uses_cells(const PostArgs& a) { return a.k == 21 && a.ldk % 4 == 0 && a.ldk >= 24 && a.xs && a.ys && !a.softmax && !a.logits && a.n <= 65535; }
int post_launch_count(const PostArgs& a) {
  if (post_uses_cells(a)) return (a.frame_bgr && (a.frame_rgba || a.blended)) ? 2 : 1;
  if (a.k == 21 && a.ldk % 4 == 0 && a.ldk >= 24) return (a.top_code && !a.softmax) ? 2 : 1;
  return 1;
}

size_t post_smem_bytes(const PostArgs& a) {
  return ((size_t)a.max_lr * a.max_lc * post_pad(a.k) + (size_t)a.k * a.max_lr * kPostTile) * sizeof(float);
}

cudaError_t launch_post(const PostArgs& a, cudaStream_t s) {
  if (a.k == 21 && a.ldk % 4 == 0 && a.ldk >= 24) {   // the 21 VOC classes of fcn-resnet50; other K: generic kernel below
    if (post_uses_cells(a)) {
      dim3 cgrid((unsigned)((a.lh * a.lw + 127) / 128), (unsigned)a.n);
      post_cell_kernel<21, 4><<<cgrid, 128, 0, s>>>(a);   // 2-row chunks at 6 CTAs / SM (80 registers, spills) measured slower: 0.21 vs 0.156 ms
      if (a.frame_bgr && (a.frame_rgba || a.blended)) {
        const size_t npix = (size_t)a.n * a.oh * a.ow;
        frame_blend_kernel<<<(unsigned)((npix / 4 + 256) / 256), 256, 0, s>>>(a.frame_bgr, a.decoded, npix, a.frame_rgba, a.blended);
      }
      return cudaGetLastError();
    }
    dim3 grid((a.ow + 31) / 32, (a.oh + 4 * kPostStripRows - 1) / (4 * kPostStripRows), a.n);
    if (a.softmax) { post_strip_kernel<21, true><<<grid, 128, 0, s>>>(a); return cudaGetLastError(); }
    if (a.top_code) {
      const size_t npix = (size_t)a.n * a.lh * a.lw;
      lowres_top_kernel<21><<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(a.lowres, a.ldk, npix, a.top_code);
    }
    post_strip_kernel<21><<<grid, 128, 0, s>>>(a);
    return cudaGetLastError();
  }
  const size_t smem = post_smem_bytes(a);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {   // function attributes are per device: set on every launch that needs it (cheap)
    cudaError_t e = cudaFuncSetAttribute(post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((a.ow + kPostTile - 1) / kPostTile, (a.oh + kPostTile - 1) / kPostTile, a.n);
  post_kernel<<<grid, 128, smem, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_color_code(const float* hm, int k, int h, int w, const uint32_t* color_lut, uint32_t* rgba, uint8_t* class_map,
                              cudaStream_t s) {
  const size_t npix = (size_t)h * w;
  if (npix == 0) return cudaSuccess;
  color_code_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(hm, k, npix, color_lut, rgba, class_map);
  return cudaGetLastError();
}

cudaError_t launch_frame_rgba(const uint8_t* bgr, size_t npix, uint32_t* rgba, cudaStream_t s) {
  if (npix == 0) return cudaSuccess;
  frame_rgba_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(bgr, npix, rgba);
  return cudaGetLastError();
}

cudaError_t launch_direct_conv(const DirectConvArgs& a, cudaStream_t s) {
  const size_t total = (size_t)a.n * a.oh * a.ow * a.cout;
  direct_conv_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace infur

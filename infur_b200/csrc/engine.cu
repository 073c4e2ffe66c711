// Device model, execution plans and the forward pass behind the C ABI.
#include "engine.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>

#include "tables.h"

namespace infur {

#define CU_TRY(expr)                                                                                        \
  do {                                                                                                      \
    cudaError_t e__ = (expr);                                                                               \
    if (e__ != cudaSuccess)                                                                                 \
      return Status::error(INFUR_E_RUNTIME, std::string(#expr) + ": " + cudaGetErrorString(e__));           \
  } while (0)

// ------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point: libcuda is not linked, so the
// library still loads (and exports its symbols) on a machine without a driver.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// esz = 2: fp16 elements, 128B swizzle (or none); esz = 1: u8 elements, 64B swizzle (a 64-channel K block is a 64-byte row) or none
static Status make_tmap(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box, int swizzle_bytes, int esz) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return Status::error(INFUR_E_RUNTIME, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
  CUresult r = enc(m, esz == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx,
                   es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::ostringstream os;
    os << "cuTensorMapEncodeTiled failed (CUresult " << (int)r << ") rank " << rank << " dims";
    for (int i = 0; i < rank; ++i) os << " " << dims[i];
    os << " strides";
    for (int i = 0; i + 1 < rank; ++i) os << " " << strides_bytes[i];
    os << " box";
    for (int i = 0; i < rank; ++i) os << " " << box[i];
    return Status::error(INFUR_E_RUNTIME, os.str());
  }
  return Status();
}

static Status make_tmap_f16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                            const uint32_t* box, bool swizzle128 = true) {
  return make_tmap(m, base, rank, dims, strides_bytes, box, swizzle128 ? 128 : 0, 2);
}

DeviceModel::~DeviceModel() { if (arena) cudaFree(arena); }
Plan::~Plan() {
  for (auto& g : graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
  for (void* p : owned) cudaFree(p);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
// Stem weights in the order stem_tc_kernel keeps them in smem: [ky][k / 8][cout][k % 8] with k = 4 * (kx + 1) + ci
// (an un-swizzled K-major UMMA operand: 8-row core matrices 128 B apart along cout, 1024 B apart along k).
static inline size_t stem_w_index(int co, int ky, int kx, int ci) {
  const int k = 4 * (kx + 1) + ci;
  return (((size_t)ky * 4 + (k >> 3)) * 64 + co) * 8 + (k & 7);
}
static inline int floordiv(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; }

// ------------------------------------------------------------------------------------------------
// Model: classify each conv, pack weights into one device arena.

static bool skip_fusion_env() { const char* e = getenv("INFUR_B200_NO_SHORTCUT_FUSION"); return e && e[0] == '1'; }

static bool halo_disabled_env() { const char* e = getenv("INFUR_B200_NO_HALO"); return e && e[0] == '1'; }
static bool i8_enabled_env() { const char* e = getenv("INFUR_B200_I8"); return !(e && e[0] == '0'); }   // INFUR_B200_I8=0: keep quantised models on fp16-carried tensors
static bool pair_disabled_env() { const char* e = getenv("INFUR_B200_NO_CTA_PAIR"); return e && e[0] == '1'; }

// 255 * sum|w| + |b| < 2^22 for every output channel: the s32 accumulator + bias of a u8 x s8 convolution then converts to f32 with
// the magic-number add (ConvTcGeom::q_tail == 2) whatever the input
static bool small_acc_bound(const float* w, size_t k_per_out, int cout, const float* bias) {
  if (!w || k_per_out == 0) return false;
  for (int co = 0; co < cout; ++co) {
    double s = 0.0;
    for (size_t j = 0; j < k_per_out; ++j) s += fabs((double)w[(size_t)co * k_per_out + j]);
    if (255.0 * s + fabs((double)bias[co]) >= 4194304.0) return false;
  }
  return true;
}

static void classify_conv(const ConvOp& c, bool reads_input, bool is_head, DevConv& d) {
  d.cin = c.cin; d.cout = c.cout; d.kh = c.kh; d.kw = c.kw; d.stride = c.stride; d.pad = c.pad; d.dil = c.dil; d.relu = c.relu;
  d.quant = c.quant; d.q_lo = c.q_lo; d.q_hi = c.q_hi; d.q_ra = c.q_ra; d.q_rb = c.q_rb; d.q_lo2 = c.q_lo2; d.q_hi2 = c.q_hi2; d.q_deq = c.deq_scale;
  d.stem = false; d.tc_ok = false;
  static const int cands[4] = {256, 128, 64, 32};
  d.block_n = 32;
  for (int bn : cands) {
    const int padded = (int)align_up((size_t)c.cout, (size_t)bn);
    if (padded - c.cout < bn / 2 || bn == 32) { d.block_n = bn; d.cout_pad = padded; break; }
  }
  if (!is_head && d.cout_pad != c.cout) { d.why_not = "cout is not a multiple of 64 for an fp16 NHWC output"; }
  if (reads_input && c.cin == 3 && c.kh == 7 && c.kw == 7 && c.stride == 2 && c.pad == 3 && c.dil == 1) {
    d.stem = true; d.taps = 7; d.cchunks = 1; d.kdim = 7 * 64;
    d.tc_ok = d.why_not.empty();
    return;
  }
  d.taps = c.kh * c.kw; d.cchunks = c.cin / 64; d.kdim = d.taps * c.cin + c.cin2;
  d.cin2 = c.in2 >= 0 ? c.cin2 : 0; d.stride2 = c.stride2;
  if (c.cin % 64 != 0) d.why_not = "cin is not a multiple of 64";
  else if (c.stride != 1 && c.stride != 2) d.why_not = "stride must be 1 or 2";
  else if (d.taps + (d.cin2 ? 1 : 0) > kMaxTaps) d.why_not = "more than 49 filter taps";
  d.tc_ok = d.why_not.empty();
}

Status build_device_model(LoweredModel&& lm, const infur_b200_config& cfg, bool skip_weights, std::unique_ptr<DeviceModel>& out) {
  auto dm = std::unique_ptr<DeviceModel>(new DeviceModel());
  dm->lm = std::move(lm);
  LoweredModel& m = dm->lm;
  if (m.heads.empty()) return Status::error(INFUR_E_MODEL_LOAD, "model has no output head");
  if (cfg.conv_impl == INFUR_CONV_TCGEN05 && !skip_fusion_env()) fuse_projection_shortcuts(m);
  dm->out_head = 0;                                   // the reference consumes out[0] only (app.rs:116)
  dm->aux_head = m.heads.size() > 1 ? 1 : -1;
  // ops needed for the computed heads
  std::vector<int> producer(m.num_tensors, -1);
  for (size_t i = 0; i < m.ops.size(); ++i) producer[m.ops[i].out] = (int)i;
  dm->needed.assign(m.ops.size(), 0);
  std::vector<int> stack;
  stack.push_back(m.heads[dm->out_head].tensor);
  if (cfg.compute_aux && dm->aux_head >= 0) stack.push_back(m.heads[dm->aux_head].tensor);
  while (!stack.empty()) {
    int t = stack.back(); stack.pop_back();
    int p = t >= 0 ? producer[t] : -1;
    if (p < 0 || dm->needed[p]) continue;
    dm->needed[p] = 1;
    stack.push_back(m.ops[p].in);
    if (m.ops[p].kind == OpKind::Conv && m.ops[p].conv.residual >= 0) stack.push_back(m.ops[p].conv.residual);
    if (m.ops[p].kind == OpKind::Conv && m.ops[p].conv.in2 >= 0) stack.push_back(m.ops[p].conv.in2);
  }
  std::vector<char> is_head_tensor(m.num_tensors, 0);
  for (auto& h : m.heads) is_head_tensor[h.tensor] = 1;

  dm->convs.resize(m.ops.size());
  // Int8 plan (onnx_reader.h int8_plan_eligible): u8 activation tensors, native int8 convolutions for all but the RGB stem
  const bool i8 = cfg.conv_impl == INFUR_CONV_TCGEN05 && i8_enabled_env() && int8_plan_eligible(m, &dm->needed, nullptr);
  dm->i8 = i8;
  size_t off = 0;
  for (size_t i = 0; i < m.ops.size(); ++i) {
    if (m.ops[i].kind != OpKind::Conv || !dm->needed[i]) continue;
    const ConvOp& c = m.ops[i].conv;
    DevConv& d = dm->convs[i];
    classify_conv(c, m.ops[i].in == m.input_tensor, is_head_tensor[m.ops[i].out] != 0, d);
    if (i8) { d.mode = d.stem ? 2 : 3; d.res_zp = c.res_zp; d.out_zp = c.out_zp; }
    if (d.mode == 3) d.small_acc = small_acc_bound(c.weight.data(), c.weight.size() / (size_t)std::max(c.cout, 1), c.cout, c.bias.data());
    for (float q : c.qmul) d.qmul_max = std::max(d.qmul_max, std::fabs(q));
    if (!d.tc_ok && cfg.conv_impl == INFUR_CONV_TCGEN05)
      return Status::error(INFUR_E_MODEL_LOAD, "convolution '" + m.ops[i].name + "' cannot run on the tcgen05 path: " + d.why_not);
    d.w_off = off; off = align_up(off + (size_t)d.cout_pad * d.kdim * 2, 256);
    if (d.stem) { d.wv_off = off; off = align_up(off + (size_t)c.cout * c.kh * c.kw * c.cin * 2, 256); }
    else d.wv_off = d.w_off;
    d.b_off = off; off = align_up(off + (size_t)d.cout_pad * 4, 256);
    if (d.quant) {
      if (cfg.conv_impl != INFUR_CONV_TCGEN05) return Status::error(INFUR_E_UNSUPPORTED, "quantised models run on the tcgen05 path only (cfg.conv_impl)");
      d.q_off = off; off = align_up(off + (size_t)d.cout_pad * 4, 256);
      if (d.mode == 3) { d.bi_off = off; off = align_up(off + (size_t)d.cout_pad * 4, 256); }
    }
  }
  if (m.quant) { dm->lut_q_off = off; off = align_up(off + 768 * 2, 256); }
  dm->arena_bytes = off;
  CU_TRY(cudaMalloc(&dm->arena, off ? off : 256));
  if (skip_weights) {
    CU_TRY(cudaMemset(dm->arena, 0, off));
  } else {
    std::vector<uint8_t> host(off, 0);
    for (size_t i = 0; i < m.ops.size(); ++i) {
      if (m.ops[i].kind != OpKind::Conv || !dm->needed[i]) continue;
      const ConvOp& c = m.ops[i].conv;
      const DevConv& d = dm->convs[i];
      __half* w16 = reinterpret_cast<__half*>(host.data() + d.w_off);
      if (d.stem) {
        for (int co = 0; co < c.cout; ++co)
          for (int ky = 0; ky < 7; ++ky)
            for (int kx = 0; kx < 7; ++kx)
              for (int ci = 0; ci < 3; ++ci)
                w16[stem_w_index(co, ky, kx, ci)] = __float2half_rn(c.weight[(((size_t)co * 7 + ky) * 7 + kx) * 3 + ci]);
        __half* wv = reinterpret_cast<__half*>(host.data() + d.wv_off);
        for (size_t j = 0; j < c.weight.size(); ++j) wv[j] = __float2half_rn(c.weight[j]);
      } else if (d.mode == 3) {
        // native int8: s8 [cout][kdim] (one byte per weight, zero point already subtracted) and the int32 bias
        int8_t* w8 = reinterpret_cast<int8_t*>(host.data() + d.w_off);
        for (size_t j = 0; j < c.weight.size(); ++j) w8[j] = (int8_t)c.weight[j];
        int32_t* bi = reinterpret_cast<int32_t*>(host.data() + d.bi_off);
        for (int co = 0; co < c.cout; ++co) bi[co] = (int32_t)c.bias[co];
      } else {
        // [cout][kh][kw][cin] (+ [cout][cin2] of a fused shortcut) == [cout][kdim]
        const size_t k1 = (size_t)c.kh * c.kw * c.cin;
        for (int co = 0; co < c.cout; ++co) {
          for (size_t j = 0; j < k1; ++j) w16[(size_t)co * d.kdim + j] = __float2half_rn(c.weight[(size_t)co * k1 + j]);
          for (int j = 0; j < d.cin2; ++j) w16[(size_t)co * d.kdim + k1 + j] = __float2half_rn(c.weight2[(size_t)co * d.cin2 + j]);
        }
      }
      float* b = reinterpret_cast<float*>(host.data() + d.b_off);
      for (int co = 0; co < c.cout; ++co) b[co] = c.bias[co];
      if (d.quant) {
        float* q = reinterpret_cast<float*>(host.data() + d.q_off);
        for (int co = 0; co < c.cout; ++co) q[co] = c.qmul[co];
      }
    }
    if (m.quant) {
      // QuantizeLinear on the network input folded into the pre-kernel table: q = sat(rne(x / scale) + zp), stored as q - zp.
      // x is the value ImageSession::forward would feed (predict_onnx.rs:117-137): normalised f32, or the raw byte.
      float lut_f[768];
      if (m.io.float_input) build_norm_lut(lut_f);
      else for (int i = 0; i < 768; ++i) lut_f[i] = (float)(i & 255);
      __half* lq = reinterpret_cast<__half*>(host.data() + dm->lut_q_off);
      for (int i = 0; i < 768; ++i) {
        volatile float t = lut_f[i] / m.in_scale;
        float r = nearbyintf(t) + (float)m.in_zp;
        r = fminf(fmaxf(r, (float)m.in_qmin), (float)m.in_qmax);
        lq[i] = __float2half_rn(r - (float)m.in_zp);
      }
    }
    CU_TRY(cudaMemcpy(dm->arena, host.data(), off, cudaMemcpyHostToDevice));
  }
  CU_TRY(cudaDeviceSynchronize());   // legacy-stream uploads do not order with the handle's non-blocking streams
  for (auto& op : m.ops) { std::vector<float>().swap(op.conv.weight); std::vector<float>().swap(op.conv.weight2); }
  out = std::move(dm);
  return Status();
}

// ------------------------------------------------------------------------------------------------
// One tcgen05 conv launch description: tensor maps + geometry.

struct ConvIO {
  const __half* x = nullptr;   // NHWC fp16 [n][h][w][cin], or the padded stem buffer
  int n = 0, h = 0, w = 0;     // logical input size
  int oh = 0, ow = 0;
  const __half* wgt = nullptr; // [cout_pad][kdim]
  const float* bias = nullptr;
  const __half* residual = nullptr;
  const __half* x2 = nullptr;  // fused shortcut source [n][h2][w2][cin2]
  int h2 = 0, w2 = 0;
  __half* y = nullptr;
  float* y_f32 = nullptr;
  int out_ld = 0;
  const float* qmul = nullptr;   // quantised layer (DevConv::quant)
  const int32_t* bias_i32 = nullptr;   // mode 3
};

enum { kVarPlain = 0, kVarPair = 1, kVarHalo = 2, kVarPairDeep = 3 /* CTA pair with 8 epilogue chunk buffers (layers with a residual) */ };
static bool halo_ok(const DevConv& d) { return !d.stem && d.kh == 3 && d.kw == 3 && d.stride == 1 && d.pad == d.dil && (d.dil == 1 || d.dil == 2 || d.dil == 4) && d.cin2 == 0; }

static Status setup_conv_tc(const DevConv& d, const ConvIO& io, PlanOp& po, int block_n, int variant = kVarPlain) {
  const bool pair = variant == kVarPair || variant == kVarPairDeep, halo = variant == kVarHalo;
  po.block_n = block_n; po.pair = pair; po.variant = variant;
  ConvTcGeom& g = po.geom;
  memset(&g, 0, sizeof(g));
  memset(&po.maps, 0, sizeof(po.maps));
  g.n_img = io.n; g.oh = io.oh; g.ow = io.ow;
  // tile shape: bw x bh = 128 output pixels; least overhang, ties to 16 x 8
  int best = -1; long best_cost = 0;
  static const int order[6] = {4, 3, 5, 6, 7, 2};
  for (int l2 : order) {
    const int bw = 1 << l2, bh = 128 >> l2;
    const long cost = (long)((io.ow + bw - 1) / bw) * ((io.oh + bh - 1) / bh);
    if (best < 0 || cost < best_cost) { best = l2; best_cost = cost; }
  }
  if (halo) best = 3;   // 8 wide x 16 high: an 8-row core-matrix group = one patch row (conv_halo_kernel)
  g.bw_log2 = best;
  const int bw = 1 << best, bh = 128 >> best;
  g.tiles_x = (io.ow + bw - 1) / bw; g.tiles_y = (io.oh + bh - 1) / bh;
  g.tiles_n = d.cout_pad / block_n;
  g.num_tiles = io.n * g.tiles_x * g.tiles_y * g.tiles_n;
  g.num_taps = d.taps + (d.cin2 ? 1 : 0); g.num_kb = d.taps * d.cchunks + d.cin2 / 64;
  g.main_taps = d.taps; g.cchunks = d.cchunks;
  g.out_ld = io.out_ld; g.relu = d.relu ? 1 : 0;
  g.store_mode = io.y_f32 ? 0 : (io.residual ? 2 : 1);
  // chunk buffers of the epilogue: 4 whenever the smem pipeline keeps >= 4 stages beside them, or the K loop is so short
  // (<= 8 blocks) that the epilogue, not the operand pipeline, sets the pace; else 2
  g.epi_bufs = g.store_mode == 0 ? 0 : 4;
  if (g.store_mode == 1 && !pair && !halo && !d.stem && conv_tc_stages(block_n, 4) < 4 && g.num_kb > 8) g.epi_bufs = 2;
  const int esz = d.mode == 3 ? 1 : 2;      // operand element size
  const int osz = d.mode >= 2 ? 1 : 2;      // output / residual element size
  if ((d.mode >= 2 && halo) || (d.mode == 2 && pair)) return Status::error(INFUR_E_UNSUPPORTED, "int8 plans use the plain, CTA-pair and stem kernels only");
  if (d.mode == 3) g.epi_bufs = g.store_mode == 0 ? 0 : 4;
  if (variant == kVarPairDeep) {
    // four stores in flight + four residual chunks prefetched instead of two + two: the +residual 1x1 layers stream 2 x 64 KB per
    // tile through these buffers, and with two stores in flight the chunk rate is bound by the TMA store's ~2 us read-out latency
    if (g.store_mode != 2) return Status::error(INFUR_E_UNSUPPORTED, "the deep-epilogue pair variant is for layers with a residual");
    g.epi_bufs = 8;
  }
  g.stages = pair ? conv_tc_pair_stages(g.epi_bufs, d.mode == 3) : conv_tc_stages(block_n, g.epi_bufs, d.mode == 3);
  g.pair = pair ? 1 : 0;
  g.halo = halo ? 1 : 0; g.halo_dil = d.dil;
  if (halo) { if (!halo_ok(d) || io.y_f32 || io.residual) return Status::error(INFUR_E_UNSUPPORTED, "halo variant: needs a 3x3 / stride 1 / pad = dilation convolution without residual"); g.stages = 0; }
  g.num_work = ((io.n * g.tiles_x * g.tiles_y + 1) / 2) * g.tiles_n;
  g.bias = io.bias; g.residual = io.residual; g.out = io.y; g.out_f32 = io.y_f32;
  if (d.quant) {
    g.quant = 1; g.qmul = io.qmul;
    g.q_lo = d.q_lo; g.q_hi = d.q_hi; g.q_ra = d.q_ra; g.q_rb = d.q_rb; g.q_lo2 = d.q_lo2; g.q_hi2 = d.q_hi2; g.q_deq = d.q_deq;
    g.mode = d.mode ? d.mode : 1;
    g.bias_i32 = io.bias_i32; g.q_zres = (float)d.res_zp; g.q_zmagic = (float)d.out_zp + 12582912.f;
    g.q_tail = 0; g.q_floor = 0;
    if (g.mode >= 2 && !io.y_f32) {
      // final bounds of the stored byte, and the bound on what reaches the last rounding add (ConvTcGeom::q_tail)
      const bool has_res = io.residual != nullptr;
      const float lo_f = has_res ? (d.relu ? std::max(d.q_lo2, 0.f) : d.q_lo2) : (d.relu ? std::max(d.q_lo, 0.f) : d.q_lo);
      const float hi_f = has_res ? d.q_hi2 : d.q_hi;
      const double lo_b = (double)lo_f + d.out_zp, hi_b = (double)hi_f + d.out_zp;
      const bool range_ok = hi_b == 255.0 && lo_b >= 0.0 && lo_b <= 255.0 && lo_b == std::floor(lo_b);
      const bool small = g.mode == 3 && d.small_acc;
      bool bounded;
      if (has_res) bounded = std::max(std::fabs((double)d.q_lo), std::fabs((double)d.q_hi)) * std::fabs((double)d.q_ra) + 255.0 * std::fabs((double)d.q_rb) + 1.0 < 4194304.0;
      else bounded = small && d.qmul_max <= 1.f;          // |(acc + bias) * qmul| < 2^22
      if (range_ok && bounded) { g.q_tail = small ? 2 : 1; g.q_floor = (int)lo_b; }
    }
  }
  Status st;
  const uint32_t box[4] = {64, (uint32_t)(d.stem ? 128 : bw), (uint32_t)(d.stem ? 1 : bh), 1};
  if (d.stem) {
    // stem_tc_kernel: tile = 128 pixels of one output row; A = 7 raw input row segments (conv_tc.cu)
    g.stem = 1; g.stem_w = io.wgt;
    g.bw_log2 = 7;
    g.tiles_x = (io.ow + 127) / 128; g.tiles_y = io.oh; g.tiles_n = 1;
    g.num_tiles = io.n * g.tiles_x * g.tiles_y;
    const int pitch = stem_pitch_px(io.w), rows = stem_rows(io.h);
    const uint64_t dims[4] = {64, (uint64_t)(pitch / 16), (uint64_t)rows, (uint64_t)io.n};
    const uint64_t strides[3] = {128, (uint64_t)pitch * 8, (uint64_t)rows * pitch * 8};
    const uint32_t abox[4] = {64, (uint32_t)kStemRowGroups, 7, 1};
    st = make_tmap_f16(&po.maps.a[0], io.x, 4, dims, strides, abox, false);
    if (!st.ok()) return st;
    po.maps.a[1] = po.maps.a[0]; po.maps.a[2] = po.maps.a[0]; po.maps.a[3] = po.maps.a[0];
  } else if (halo) {
    const uint64_t dims[4] = {(uint64_t)d.cin, (uint64_t)io.w, (uint64_t)io.h, (uint64_t)io.n};
    const uint64_t strides[3] = {(uint64_t)d.cin * 2, (uint64_t)io.w * d.cin * 2, (uint64_t)io.h * io.w * d.cin * 2};
    const uint32_t pbox[4] = {64, 16, (uint32_t)(16 + 2 * d.dil), 1};   // halo patch: 16 px pitch x (16 + 2d) rows
    st = make_tmap_f16(&po.maps.a[0], io.x, 4, dims, strides, pbox);
    if (!st.ok()) return st;
    po.maps.a[1] = po.maps.a[0]; po.maps.a[2] = po.maps.a[0]; po.maps.a[3] = po.maps.a[0];
    for (int t = 0; t < 9; ++t) g.tap_cc[t] = (uint8_t)d.cchunks;
  } else {
    const int s = d.stride;
    bool have[4] = {false, false, false, false};
    for (int py = 0; py < s; ++py)
      for (int px = 0; px < s; ++px) {
        const int vw = (io.w - px + s - 1) / s, vh = (io.h - py + s - 1) / s;
        if (vw <= 0 || vh <= 0) continue;
        const uint64_t dims[4] = {(uint64_t)d.cin, (uint64_t)vw, (uint64_t)vh, (uint64_t)io.n};
        const uint64_t strides[3] = {(uint64_t)s * d.cin * esz, (uint64_t)s * io.w * d.cin * esz, (uint64_t)io.h * io.w * d.cin * esz};
        st = make_tmap(&po.maps.a[py * s + px], reinterpret_cast<const uint8_t*>(io.x) + ((size_t)py * io.w + px) * d.cin * esz, 4, dims, strides, box, esz == 1 ? 64 : 128, esz);
        if (!st.ok()) return st;
        have[py * s + px] = true;
      }
    int first = -1;
    for (int v = 0; v < 4; ++v) if (have[v]) { first = v; break; }
    if (first < 0) return Status::error(INFUR_E_SHAPE, "convolution input is empty");
    for (int v = 0; v < 4; ++v) if (!have[v]) po.maps.a[v] = po.maps.a[first];
    for (int ky = 0; ky < d.kh; ++ky)
      for (int kx = 0; kx < d.kw; ++kx) {
        const int t = ky * d.kw + kx;
        const int offy = ky * d.dil - d.pad, offx = kx * d.dil - d.pad;
        const int qy = floordiv(offy, s), qx = floordiv(offx, s);
        const int py = offy - qy * s, px = offx - qx * s;
        // a tap whose parity view does not exist reads only padding: point it far outside any view
        g.tap_view[t] = (int8_t)(have[py * s + px] ? py * s + px : first);
        g.tap_dy[t] = (int16_t)(have[py * s + px] ? qy : 30000);
        g.tap_dx[t] = (int16_t)qx;
        g.tap_cc[t] = (uint8_t)d.cchunks;
      }
    if (d.cin2) {
      // fused projection shortcut: one more tap reading the block input through its own (strided) view.  The main
      // convolution is 1x1 / stride 1 (fuse_projection_shortcuts), so view 1 is free.
      const int s2 = d.stride2;
      const uint64_t dims[4] = {(uint64_t)d.cin2, (uint64_t)((io.w2 + s2 - 1) / s2), (uint64_t)((io.h2 + s2 - 1) / s2), (uint64_t)io.n};
      const uint64_t strides[3] = {(uint64_t)s2 * d.cin2 * 2, (uint64_t)s2 * io.w2 * d.cin2 * 2, (uint64_t)io.h2 * io.w2 * d.cin2 * 2};
      st = make_tmap_f16(&po.maps.a[1], io.x2, 4, dims, strides, box);
      if (!st.ok()) return st;
      const int t = d.taps;
      g.tap_view[t] = 1; g.tap_dx[t] = 0; g.tap_dy[t] = 0; g.tap_cc[t] = (uint8_t)(d.cin2 / 64);
    }
  }
  if (!d.stem) {
    const uint64_t dims[2] = {(uint64_t)d.kdim, (uint64_t)d.cout_pad};
    const uint64_t strides[1] = {(uint64_t)d.kdim * esz};
    const uint32_t bbox[2] = {64, (uint32_t)(pair ? 128 : block_n)};   // a CTA pair splits the weight tile between its CTAs
    st = make_tmap(&po.maps.b, io.wgt, 2, dims, strides, bbox, esz == 1 ? 64 : 128, esz);
    if (!st.ok()) return st;
  } else {
    po.maps.b = po.maps.a[0];
  }
  po.maps.c = po.maps.b; po.maps.r = po.maps.b;
  if (g.store_mode != 0) {
    // output / residual: NHWC fp16 [n][oh][ow][out_ld], stored (loaded) in 64-channel x tile boxes
    const uint64_t dims[4] = {(uint64_t)d.cout, (uint64_t)io.ow, (uint64_t)io.oh, (uint64_t)io.n};
    const uint64_t strides[3] = {(uint64_t)io.out_ld * osz, (uint64_t)io.ow * io.out_ld * osz, (uint64_t)io.oh * io.ow * io.out_ld * osz};
    // output chunks (conv_tc.cu epilogue_tma): fp16: 64 channels = 128-byte rows, 128B-swizzled; u8 (int8 plans): 128 channels =
    // 128-byte rows, 128B-swizzled, when the N tile has them, else 64 channels = 64-byte rows, unswizzled
    const uint32_t cw = (osz == 1 && block_n >= 128) ? 128u : 64u;
    const uint32_t cbox[4] = {cw, box[1], box[2], box[3]};
    const int csw = (osz == 2 || cw == 128) ? 128 : 0;
    st = make_tmap(&po.maps.c, io.y, 4, dims, strides, cbox, csw, osz);
    if (!st.ok()) return st;
    if (io.residual) {
      st = make_tmap(&po.maps.r, io.residual, 4, dims, strides, cbox, csw, osz);
      if (!st.ok()) return st;
    }
  }
  return Status();
}

static void setup_direct(const DevConv& d, const ConvIO& io, const __half* wv, DirectConvArgs& a) {
  memset(&a, 0, sizeof(a));
  a.x = io.x; a.w = wv; a.bias = io.bias; a.residual = io.residual; a.y = io.y; a.y_f32 = io.y_f32;
  a.n = io.n; a.h = io.h; a.wd = io.w; a.cin = d.cin; a.cout = d.cout; a.kh = d.kh; a.kw = d.kw; a.stride = d.stride; a.pad = d.pad;
  a.dil = d.dil; a.oh = io.oh; a.ow = io.ow; a.relu = d.relu ? 1 : 0; a.out_ld = io.out_ld;
  if (d.stem) { a.x_pitch_px = stem_pitch_px(io.w); a.x_rows = stem_rows(io.h); a.x_c = 4; a.x_off_y = kStemPadTop; a.x_off_x = kStemPadLeft; }
  else { a.x_pitch_px = io.w; a.x_rows = io.h; a.x_c = d.cin; a.x_off_y = 0; a.x_off_x = 0; }
}

// ------------------------------------------------------------------------------------------------
// Per-op choice of the N tile (256 / 128 / 64 output channels) by measurement at plan-build time.  The
// accumulation order of every output element is the same for all choices (K blocks in sequence), so the
// result is bit-identical whichever wins; only time differs: a narrower tile leaves room for more pipeline
// stages (more HBM bytes in flight for the memory-bound 1x1 convs), a wider one halves the activation re-reads.
static Status tune_block_n(infur_b200_handle* H, const DevConv& d, const ConvIO& io, PlanOp& po, double flops) {
  struct Cand { int bn; int var; };
  static const Cand cands[8] = {{256, kVarPair}, {256, kVarPlain}, {128, kVarPlain}, {64, kVarPlain}, {256, kVarHalo}, {128, kVarHalo}, {64, kVarHalo}, {256, kVarPairDeep}};
  if (d.stem) return Status();
  // Decisions are kept per layer-shape class: layer parameters + the number of 128-pixel M tiles of the whole batch in
  // half-octave buckets (what decides waves per SM and hence which variant wins).  The 91 positions of the reference's scale
  // slider (gui.rs:278-285) span 100x in area = 14 buckets, so most ticks -- and every re-visited factor -- tune nothing.
  const double mtiles = (double)io.n * ((io.ow + 15) / 16) * ((io.oh + 7) / 8);
  const TuneKey key{d.cin, d.cout, d.kh, d.stride, d.dil, d.mode, io.residual ? 1 : 0, d.cin2, (int)lround(2.0 * log2(std::max(1.0, mtiles)))};
  {
    auto it = H->tune_cache.find(key);
    // a measured decision of the same layer within one octave of work (+-2 buckets) is reused without measuring: moving the
    // scale slider by a tick must not cost an autotune (the choice is about waves per SM, which changes slowly with size)
    for (int delta = 1; delta <= 2 && it == H->tune_cache.end(); ++delta)
      for (int sgn = -1; sgn <= 1 && it == H->tune_cache.end(); sgn += 2) {
        TuneKey k2 = key; k2.bucket += sgn * delta;
        it = H->tune_cache.find(k2);
      }
    if (it != H->tune_cache.end()) {
      if (it->second.block_n != po.block_n || it->second.variant != po.variant) return setup_conv_tc(d, io, po, it->second.block_n, it->second.variant);
      return Status();
    }
  }
  // a layer of less than 2 GFLOP (~2 us of tensor-core time: small frames, the first 1x1 of layer1 at batch 1) is bound by launch
  // latency whatever the tile shape: keep the default instead of spending ~80 launches on measuring it
  if (flops < 2e9) return Status();
  H->last_build_tuned++;
  if (const char* dbg = getenv("INFUR_B200_DEBUG_TUNE"))
    if (dbg[0] == '1' || dbg[0] == '2') fprintf(stderr, "[infur_b200] autotune: %dx%d conv %d->%d s%d d%d mode %d res %d cin2 %d, n %d out %dx%d, bucket %d\n", d.kh, d.kw, d.cin, d.cout,
                               d.stride, d.dil, d.mode, io.residual ? 1 : 0, d.cin2, io.n, io.ow, io.oh, key.bucket);
  const bool allow_pair = !pair_disabled_env();
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0));
  CU_TRY(cudaEventCreate(&e1));
  const bool allow_halo = halo_ok(d) && !io.residual && !io.y_f32 && !halo_disabled_env();
  Cand best = {po.block_n, po.variant};
  float best_ms = 1e30f;
  Status st;
  // three interleaved rounds, minimum per candidate: a single short measurement is at the mercy of clock ramps
  for (int round = 0; round < 3 && st.ok(); ++round) {
    for (const Cand& c : cands) {
      if (c.bn > d.block_n || d.cout_pad % c.bn != 0 || (c.var == kVarPair && !allow_pair) || (c.var == kVarHalo && !allow_halo)) continue;
      if (c.var == kVarPairDeep && (!allow_pair || !io.residual || d.mode == 2)) continue;
      if (d.mode >= 2 && c.var == kVarHalo) continue;   // int8 plans: plain and CTA-pair kernels
      PlanOp trial;
      if (!(st = setup_conv_tc(d, io, trial, c.bn, c.var)).ok()) break;
      cudaError_t e = conv_tc_launch(c.bn, trial.maps, trial.geom, H->num_sms, H->stream);   // warm-up
      cudaEventRecord(e0, H->stream);
      for (int r = 0; r < 3 && e == cudaSuccess; ++r) e = conv_tc_launch(c.bn, trial.maps, trial.geom, H->num_sms, H->stream);
      cudaEventRecord(e1, H->stream);
      if (e == cudaSuccess) e = cudaEventSynchronize(e1);
      H->launches += 4;
      if (e != cudaSuccess) { st = Status::error(INFUR_E_RUNTIME, std::string("autotune: ") + cudaGetErrorString(e)); break; }
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      const bool same = c.bn == best.bn && c.var == best.var;
      if (const char* dbg = getenv("INFUR_B200_DEBUG_TUNE"))
        if (dbg[0] == '2') fprintf(stderr, "[infur_b200]   round %d: N%d variant %d: %.4f ms\n", round, c.bn, c.var, ms / 3);
      if (ms < best_ms * 0.97f || (same && ms < best_ms)) { best_ms = ms < best_ms ? ms : best_ms; best = c; }
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (!st.ok()) return st;
  H->tune_cache[key] = TuneChoice{best.bn, best.var};
  if (best.bn != po.block_n || best.var != po.variant) st = setup_conv_tc(d, io, po, best.bn, best.var);
  return st;
}


// ------------------------------------------------------------------------------------------------
// Bottleneck-tail fusion (conv_b2b_kernel): a 3x3 / stride 1 convolution (cmid -> cmid, ReLU) whose only consumer is a 1x1
// convolution with a residual (cmid -> 4 cmid, + identity, ReLU) becomes ONE launch; the cmid-channel tensor between them never
// reaches HBM.  Results are bit-identical to the two launches (same K order, same rounding points), so -- like the tile shapes --
// the choice is made by measurement per layer-shape class and cached.
static bool b2b_disabled_env() { const char* e = getenv("INFUR_B200_NO_B2B"); return e && e[0] == '1'; }
static bool b2b_forced_env() { const char* e = getenv("INFUR_B200_B2B"); return e && e[0] == 'f'; }   // INFUR_B200_B2B=force: skip the measurement

static Status setup_b2b(const DevConv& da, const ConvIO& ioa, const DevConv& db, const ConvIO& iob, PlanOp& pf) {
  Status st = setup_conv_tc(da, ioa, pf, da.cout, kVarPlain);
  if (!st.ok()) return st;
  ConvTcGeom& g = pf.geom;
  const int m_tiles = ioa.n * g.tiles_x * g.tiles_y;
  g.b2b = 1; g.b2b_cmid = da.cout;
  g.tiles_n = db.cout / 128;
  g.num_tiles = m_tiles * g.tiles_n;
  g.bias2 = iob.bias; g.residual = iob.residual; g.out = iob.y; g.out_f32 = nullptr; g.out_ld = iob.out_ld; g.relu = db.relu ? 1 : 0;
  g.store_mode = 2; g.epi_bufs = 4; g.pair = 0; g.halo = 0;
  pf.b2b = true; pf.pair = false; pf.variant = kVarPlain; pf.block_n = da.cout;
  const int bw = 1 << g.bw_log2, bh = 128 >> g.bw_log2;
  {
    const uint64_t dims[2] = {(uint64_t)db.kdim, (uint64_t)db.cout_pad};
    const uint64_t strides[1] = {(uint64_t)db.kdim * 2};
    const uint32_t bbox[2] = {64, 128};
    if (!(st = make_tmap(&pf.maps.b2, iob.wgt, 2, dims, strides, bbox, 128, 2)).ok()) return st;
  }
  const uint64_t dims[4] = {(uint64_t)db.cout, (uint64_t)iob.ow, (uint64_t)iob.oh, (uint64_t)iob.n};
  const uint64_t strides[3] = {(uint64_t)iob.out_ld * 2, (uint64_t)iob.ow * iob.out_ld * 2, (uint64_t)iob.oh * iob.ow * iob.out_ld * 2};
  const uint32_t cbox[4] = {64, (uint32_t)bw, (uint32_t)bh, 1};
  if (!(st = make_tmap(&pf.maps.c, iob.y, 4, dims, strides, cbox, 128, 2)).ok()) return st;
  return make_tmap(&pf.maps.r, iob.residual, 4, dims, strides, cbox, 128, 2);
}

static Status fuse_b2b_pairs(infur_b200_handle* H, const DeviceModel& M, Plan& p, const std::vector<ConvIO>& ios, const std::vector<int>& last_use) {
  const LoweredModel& m = M.lm;
  std::vector<int> uses(m.num_tensors, 0);
  for (size_t i = 0; i < m.ops.size(); ++i) {
    if (!M.needed[i]) continue;
    const LoweredOp& op = m.ops[i];
    if (op.in >= 0) uses[op.in]++;
    if (op.kind == OpKind::Conv) { if (op.conv.residual >= 0) uses[op.conv.residual]++; if (op.conv.in2 >= 0) uses[op.conv.in2]++; }
  }
  for (auto& hd : m.heads) uses[hd.tensor]++;
  (void)last_use;
  for (size_t k = 0; k + 1 < p.ops.size(); ++k) {
    PlanOp &a = p.ops[k], &b = p.ops[k + 1];
    if (!a.is_conv || !b.is_conv || a.skip || b.skip) continue;
    const LoweredOp &opa = m.ops[a.op], &opb = m.ops[b.op];
    const DevConv &da = M.convs[a.op], &db = M.convs[b.op];
    const ConvIO &ioa = ios[k], &iob = ios[k + 1];
    if (!(da.tc_ok && db.tc_ok && !da.stem && da.kh == 3 && da.kw == 3 && da.stride == 1 && da.pad == da.dil && da.relu && da.cin2 == 0 && !da.quant && da.mode == 0 &&
          da.cin == da.cout && conv_b2b_supported(da.cout) && opa.conv.residual < 0 && !ioa.y_f32))
      continue;
    if (!(db.kh == 1 && db.kw == 1 && db.stride == 1 && db.pad == 0 && db.cin == da.cout && db.cin2 == 0 && !db.quant && db.mode == 0 && db.cout % 128 == 0 &&
          db.cout_pad == db.cout && opb.conv.residual >= 0 && iob.residual && !iob.y_f32 && opb.in == opa.out && uses[opa.out] == 1))
      continue;
    PlanOp pf = a;
    Status st = setup_b2b(da, ioa, db, iob, pf);
    if (!st.ok()) return st;
    const double mtiles = (double)ioa.n * ((ioa.ow + 15) / 16) * ((ioa.oh + 7) / 8);
    const TuneKey key{da.cin, db.cout, 3, 1, da.dil, 100, 1, 0, (int)lround(2.0 * log2(std::max(1.0, mtiles)))};
    bool fuse = true;
    if (H->cfg.autotune && !b2b_forced_env()) {
      auto it = H->tune_cache.find(key);
      for (int delta = 1; delta <= 2 && it == H->tune_cache.end(); ++delta)
        for (int sgn = -1; sgn <= 1 && it == H->tune_cache.end(); sgn += 2) { TuneKey k2 = key; k2.bucket += sgn * delta; it = H->tune_cache.find(k2); }
      if (it != H->tune_cache.end()) fuse = it->second.variant == 1;
      else if (a.flops + b.flops < 2e9) fuse = true;   // too small to be worth measuring (see tune_block_n): one launch instead of two
      else {
        H->last_build_tuned++;
        cudaEvent_t e0, e1, e2;
        CU_TRY(cudaEventCreate(&e0)); CU_TRY(cudaEventCreate(&e1)); CU_TRY(cudaEventCreate(&e2));
        float best_f = 1e30f, best_s = 1e30f;
        cudaError_t e = cudaSuccess;
        for (int round = 0; round < 3 && e == cudaSuccess; ++round) {
          e = conv_tc_launch(pf.block_n, pf.maps, pf.geom, H->num_sms, H->stream);   // warm-up
          cudaEventRecord(e0, H->stream);
          for (int r = 0; r < 3 && e == cudaSuccess; ++r) e = conv_tc_launch(pf.block_n, pf.maps, pf.geom, H->num_sms, H->stream);
          cudaEventRecord(e1, H->stream);
          for (int r = 0; r < 3 && e == cudaSuccess; ++r) {
            e = conv_tc_launch(a.block_n, a.maps, a.geom, H->num_sms, H->stream);
            if (e == cudaSuccess) e = conv_tc_launch(b.block_n, b.maps, b.geom, H->num_sms, H->stream);
          }
          cudaEventRecord(e2, H->stream);
          if (e == cudaSuccess) e = cudaEventSynchronize(e2);
          H->launches += 10;
          float tf = 0.f, ts = 0.f;
          if (e == cudaSuccess) { cudaEventElapsedTime(&tf, e0, e1); cudaEventElapsedTime(&ts, e1, e2); best_f = std::min(best_f, tf); best_s = std::min(best_s, ts); }
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
        if (e != cudaSuccess) return Status::error(INFUR_E_RUNTIME, std::string("autotune (b2b): ") + cudaGetErrorString(e));
        fuse = best_f < best_s * 0.98f;
        H->tune_cache[key] = TuneChoice{da.cout, fuse ? 1 : 0};
        if (const char* dbg = getenv("INFUR_B200_DEBUG_TUNE"))
          if (dbg[0] == '1') fprintf(stderr, "[infur_b200] b2b %d -> %d d%d, n %d out %dx%d: fused %.3f ms vs separate %.3f ms -> %s\n", da.cin, db.cout, da.dil, ioa.n,
                                     ioa.ow, ioa.oh, best_f / 3, best_s / 3, fuse ? "fused" : "separate");
      }
    }
    if (!fuse) continue;
    const double inter = (double)ioa.n * ioa.oh * ioa.ow * da.cout * 2.0;   // the intermediate: written once and read once when unfused
    pf.flops = a.flops + b.flops;
    pf.bytes = a.bytes + b.bytes - 2.0 * inter;
    pf.text = a.text.substr(0, a.text.find(" | ")) + " ++ " + opb.name + " 1x1 -> " + std::to_string(db.cout) + " +res relu | tcgen05 fused 3x3 -> 1x1 (conv_b2b_kernel, cmid " +
              std::to_string(da.cout) + ") tiles " + std::to_string(pf.geom.num_tiles / pf.geom.tiles_n) + " | GFLOP " + std::to_string(pf.flops * 1e-9) + " MB " + std::to_string(pf.bytes * 1e-6);
    b.skip = true;
    b.text = b.text.substr(0, b.text.find(" | ")) + " | (fused into previous) | GFLOP 0 MB 0";
    b.flops = 0; b.bytes = 0;
    p.ops[k] = pf;
  }
  return Status();
}

// ------------------------------------------------------------------------------------------------
// Stem + max-pool fusion (stem_pool_kernel): the 7x7/s2 stem followed by the 3x3/s2/p1 pool as one launch; the stem's output
// (531 MB per 8 x 1080p frames) never reaches HBM.  Same bits as the two kernels; chosen by measurement like the other variants.
static bool stem_pool_disabled_env() { const char* e = getenv("INFUR_B200_NO_STEM_POOL"); return e && e[0] == '1'; }
static bool stem_pool_forced_env() { const char* e = getenv("INFUR_B200_STEM_POOL"); return e && e[0] == 'f'; }

static Status fuse_stem_pool(infur_b200_handle* H, const DeviceModel& M, Plan& p, const std::vector<ConvIO>& ios) {
  const LoweredModel& m = M.lm;
  for (size_t k = 0; k + 1 < p.ops.size(); ++k) {
    PlanOp &a = p.ops[k], &b = p.ops[k + 1];
    if (!a.is_conv || b.is_conv || a.skip || b.skip || a.b2b) continue;
    const LoweredOp &opa = m.ops[a.op], &opb = m.ops[b.op];
    const DevConv& da = M.convs[a.op];
    // the fp16 model's stem (ReLU in the epilogue) or the stem of an int8 plan (mode 2: requantised u8 output, pooled as bytes)
    if (!(da.stem && da.tc_ok && ((da.mode == 0 && !da.quant && da.relu) || da.mode == 2) && da.cout == 64 && opa.conv.residual < 0)) continue;
    if (!(opb.kind == OpKind::MaxPool && opb.pool_k == 3 && opb.pool_s == 2 && opb.pool_p == 1 && opb.in == opa.out)) continue;
    int uses = 0;
    for (size_t i = 0; i < m.ops.size(); ++i) {
      if (!M.needed[i]) continue;
      if (m.ops[i].in == opa.out) ++uses;
      if (m.ops[i].kind == OpKind::Conv && (m.ops[i].conv.residual == opa.out || m.ops[i].conv.in2 == opa.out)) ++uses;
    }
    for (auto& hd : m.heads) if (hd.tensor == opa.out) ++uses;
    if (uses != 1) continue;
    const TensorInfo& ti = p.tensors[opb.in];
    const TensorInfo& to = p.tensors[opb.out];
    if (to.h != (ti.h - 1) / 2 + 1 || to.w != (ti.w - 1) / 2 + 1 || to.c != 64) continue;
    PlanOp pf = a;
    ConvTcGeom& g = pf.geom;
    const int n = ios[k].n;
    g.sp_fused = 1; g.sp_oh = to.h; g.sp_ow = to.w;
    g.sp_strips = (to.w + 59) / 60;
    // pooled rows per unit: a unit costs 2 rc + 1 convolution rows; pick the rc whose (waves over the SMs) x (rows per unit) is smallest
    long best_cost = -1;
    for (int rc = 4; rc <= to.h; ++rc) {
      const int chunks = (to.h + rc - 1) / rc;
      const long units = (long)n * g.sp_strips * chunks;
      const long waves = (units + H->num_sms - 1) / H->num_sms;
      const long cost = waves * (2 * rc + 1);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; g.sp_rc = rc; g.sp_chunks = chunks; }
    }
    if (best_cost < 0) { g.sp_rc = to.h; g.sp_chunks = 1; }
    g.num_tiles = n * g.sp_strips * g.sp_chunks;
    g.out = reinterpret_cast<__half*>(to.ptr);
    const TuneKey key{3, 64, 7, 2, 1, 101 + da.mode, 0, 0, (int)lround(2.0 * log2(std::max(1.0, (double)n * to.h * to.w / 128.0)))};
    bool fuse = true;
    if (H->cfg.autotune && !stem_pool_forced_env()) {
      auto it = H->tune_cache.find(key);
      for (int delta = 1; delta <= 2 && it == H->tune_cache.end(); ++delta)
        for (int sgn = -1; sgn <= 1 && it == H->tune_cache.end(); sgn += 2) { TuneKey k2 = key; k2.bucket += sgn * delta; it = H->tune_cache.find(k2); }
      if (it != H->tune_cache.end()) fuse = it->second.variant == 1;
      else if (a.flops < 2e9) fuse = true;             // small frames: one launch instead of two, unmeasured
      else {
        H->last_build_tuned++;
        cudaEvent_t e0, e1, e2;
        CU_TRY(cudaEventCreate(&e0)); CU_TRY(cudaEventCreate(&e1)); CU_TRY(cudaEventCreate(&e2));
        float best_f = 1e30f, best_s = 1e30f;
        cudaError_t e = cudaSuccess;
        for (int round = 0; round < 3 && e == cudaSuccess; ++round) {
          e = conv_tc_launch(pf.block_n, pf.maps, pf.geom, H->num_sms, H->stream);
          cudaEventRecord(e0, H->stream);
          for (int r = 0; r < 3 && e == cudaSuccess; ++r) e = conv_tc_launch(pf.block_n, pf.maps, pf.geom, H->num_sms, H->stream);
          cudaEventRecord(e1, H->stream);
          for (int r = 0; r < 3 && e == cudaSuccess; ++r) {
            e = conv_tc_launch(a.block_n, a.maps, a.geom, H->num_sms, H->stream);
            if (e == cudaSuccess)
              e = M.i8 ? launch_maxpool3s2_u8(reinterpret_cast<const uint8_t*>(ti.ptr), reinterpret_cast<uint8_t*>(to.ptr), n, ti.h, ti.w, ti.c, to.h, to.w, H->stream)
                       : launch_maxpool(reinterpret_cast<const __half*>(ti.ptr), reinterpret_cast<__half*>(to.ptr), n, ti.h, ti.w, ti.c, to.h, to.w, 3, 2, 1, H->stream);
          }
          cudaEventRecord(e2, H->stream);
          if (e == cudaSuccess) e = cudaEventSynchronize(e2);
          H->launches += 10;
          float tf = 0.f, ts = 0.f;
          if (e == cudaSuccess) { cudaEventElapsedTime(&tf, e0, e1); cudaEventElapsedTime(&ts, e1, e2); best_f = std::min(best_f, tf); best_s = std::min(best_s, ts); }
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
        if (e != cudaSuccess) return Status::error(INFUR_E_RUNTIME, std::string("autotune (stem + pool): ") + cudaGetErrorString(e));
        fuse = best_f < best_s * 0.98f;
        H->tune_cache[key] = TuneChoice{64, fuse ? 1 : 0};
        if (const char* dbg = getenv("INFUR_B200_DEBUG_TUNE"))
          if (dbg[0] == '1' || dbg[0] == '2') fprintf(stderr, "[infur_b200] stem + pool, n %d conv out %dx%d: fused %.3f ms vs separate %.3f ms -> %s (rc %d, %d units)\n", n, ti.w, ti.h,
                                     best_f / 3, best_s / 3, fuse ? "fused" : "separate", g.sp_rc, g.num_tiles);
      }
    }
    if (!fuse) continue;
    pf.flops = a.flops;
    pf.bytes = a.bytes - (double)ti.bytes + (double)to.bytes;   // the stem's own output is never written; the pooled tensor is
    pf.text = a.text.substr(0, a.text.find(" | ")) + " ++ " + opb.name + " 3x3/s2 -> [" + std::to_string(to.h) + "x" + std::to_string(to.w) + "] | tcgen05 fused stem + max-pool (stem_pool_kernel) units " +
              std::to_string(g.num_tiles) + " rows/unit " + std::to_string(g.sp_rc) + " | GFLOP " + std::to_string(pf.flops * 1e-9) + " MB " + std::to_string(pf.bytes * 1e-6);
    b.skip = true;
    b.text = b.text.substr(0, b.text.find(" | ")) + " | (fused into previous) | MB 0";
    b.flops = 0; b.bytes = 0;
    p.ops[k] = pf;
    break;
  }
  return Status();
}

// ------------------------------------------------------------------------------------------------
// Plan
constexpr size_t kMaxPlans = 6;

template <typename T>
static Status dev_alloc(Plan& p, T** ptr, size_t count) {
  void* q = nullptr;
  CU_TRY(cudaMalloc(&q, std::max<size_t>(count * sizeof(T), 256)));
  p.owned.push_back(q);
  *ptr = reinterpret_cast<T*>(q);
  return Status();
}
template <typename T>
static Status dev_upload(Plan& p, T** ptr, const std::vector<T>& v) {
  Status st = dev_alloc(p, ptr, v.size());
  if (!st.ok()) return st;
  if (!v.empty()) CU_TRY(cudaMemcpy(*ptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return Status();
}

// first output index of every low-res cell: cell c = the run of outputs whose left tap is c (post_cell_kernel)
static std::vector<int32_t> cell_starts(const std::vector<int32_t>& i0, int n_in) {
  std::vector<int32_t> s((size_t)n_in + 1, (int32_t)i0.size());
  size_t x = 0;
  for (int c = 0; c <= n_in; ++c) {
    while (x < i0.size() && i0[x] < c) ++x;
    s[(size_t)c] = (int32_t)x;
  }
  return s;
}

static Status build_bilinear(Plan& p, int lh, int lw) {
  std::vector<int32_t> i0, i1; std::vector<float> l0, l1;
  Status st;
  build_bilinear_table(lh, p.oh, i0, i1, l0, l1);
  if (!(st = dev_upload(p, &p.cell_ys, cell_starts(i0, lh))).ok()) return st;
  p.max_lr = 1;
  for (int Y0 = 0; Y0 < p.oh; Y0 += 32) p.max_lr = std::max(p.max_lr, i1[std::min(Y0 + 32, p.oh) - 1] - i0[Y0] + 1);
  if (!(st = dev_upload(p, &p.y0, i0)).ok() || !(st = dev_upload(p, &p.y1, i1)).ok() || !(st = dev_upload(p, &p.ly0, l0)).ok() ||
      !(st = dev_upload(p, &p.ly1, l1)).ok())
    return st;
  build_bilinear_table(lw, p.ow, i0, i1, l0, l1);
  if (!(st = dev_upload(p, &p.cell_xs, cell_starts(i0, lw))).ok()) return st;
  p.max_lc = 1;
  for (int X0 = 0; X0 < p.ow; X0 += 32) p.max_lc = std::max(p.max_lc, i1[std::min(X0 + 32, p.ow) - 1] - i0[X0] + 1);
  if (!(st = dev_upload(p, &p.x0, i0)).ok() || !(st = dev_upload(p, &p.x1, i1)).ok() || !(st = dev_upload(p, &p.lx0, l0)).ok() ||
      !(st = dev_upload(p, &p.lx1, l1)).ok())
    return st;
  return Status();
}

Status build_plan(infur_b200_handle* H, int n, int w, int h, std::unique_ptr<Plan>& out) {
  auto pl = std::unique_ptr<Plan>(new Plan());
  Plan& p = *pl;
  p.n = n; p.w = w; p.h = h; p.factor = H->factor;
  const bool unit = H->factor == 1.0f;
  if (unit) { p.ow = w; p.oh = h; }
  else {
    if (w == 0 || h == 0) return Status::error(INFUR_E_ZERO_SIZE_IN, "scaling from 0-sized input");
    p.ow = (int)std::min<uint32_t>(scaled_dim((uint32_t)w, H->factor), 1u << 20);
    p.oh = (int)std::min<uint32_t>(scaled_dim((uint32_t)h, H->factor), 1u << 20);
    if (p.ow == 0 || p.oh == 0) return Status::error(INFUR_E_ZERO_SIZE_OUT, "scaling to 0-sized output");
  }
  Status st;
  const size_t in_bytes = (size_t)n * w * h * 3, out_px = (size_t)n * p.ow * p.oh;
  if (!(st = dev_alloc(p, &p.d_in, in_bytes)).ok()) return st;
  if (!unit) {
    if (H->cfg.resize_mode == INFUR_RESIZE_BILINEAR) {
      std::vector<int32_t> i0, i1; std::vector<float> l0, l1;
      build_bilinear_table(w, p.ow, i0, i1, l0, l1);
      if (!(st = dev_upload(p, &p.sbx0, i0)).ok() || !(st = dev_upload(p, &p.sbx1, i1)).ok() || !(st = dev_upload(p, &p.sblx0, l0)).ok() ||
          !(st = dev_upload(p, &p.sblx1, l1)).ok())
        return st;
      build_bilinear_table(h, p.oh, i0, i1, l0, l1);
      if (!(st = dev_upload(p, &p.sby0, i0)).ok() || !(st = dev_upload(p, &p.sby1, i1)).ok() || !(st = dev_upload(p, &p.sbly0, l0)).ok() ||
          !(st = dev_upload(p, &p.sbly1, l1)).ok())
        return st;
    } else {
      std::vector<int32_t> xm, ym;
      build_nearest_map(w, p.ow, xm); build_nearest_map(h, p.oh, ym);
      if (!(st = dev_upload(p, &p.xmap, xm)).ok() || !(st = dev_upload(p, &p.ymap, ym)).ok()) return st;
    }
    if (!(st = dev_alloc(p, &p.scaled, out_px * 3)).ok()) return st;
  }
  if (!(st = dev_alloc(p, &p.d_frame_rgba, out_px)).ok()) return st;
  DeviceModel* M = H->model.get();
  p.has_model = M != nullptr;
  if (!M || out_px == 0) { p.has_model = M != nullptr && out_px != 0; out = std::move(pl); return Status(); }

  if (!(st = dev_alloc(p, &p.d_class, out_px)).ok() || !(st = dev_alloc(p, &p.d_decoded, out_px)).ok()) return st;
  if (H->cfg.blend && !(st = dev_alloc(p, &p.d_blended, out_px)).ok()) return st;

  const LoweredModel& m = M->lm;
  // ---- shapes
  p.tensors.assign(m.num_tensors, TensorInfo());
  p.tensors[m.input_tensor].h = p.oh; p.tensors[m.input_tensor].w = p.ow; p.tensors[m.input_tensor].c = 3;
  std::vector<char> is_head_tensor(m.num_tensors, 0);
  for (auto& hd : m.heads) is_head_tensor[hd.tensor] = 1;
  std::vector<int> last_use(m.num_tensors, -1);
  for (size_t i = 0; i < m.ops.size(); ++i) {
    if (!M->needed[i]) continue;
    const LoweredOp& op = m.ops[i];
    const TensorInfo& ti = p.tensors[op.in];
    TensorInfo& to = p.tensors[op.out];
    int k, s, pad, dil;
    if (op.kind == OpKind::Conv) { k = op.conv.kh; s = op.conv.stride; pad = op.conv.pad; dil = op.conv.dil; to.c = op.conv.cout; }
    else { k = op.pool_k; s = op.pool_s; pad = op.pool_p; dil = 1; to.c = ti.c; }
    const int eh = ti.h + 2 * pad - dil * (k - 1) - 1, ew = ti.w + 2 * pad - dil * (k - 1) - 1;
    if (eh < 0 || ew < 0) return Status::error(INFUR_E_SHAPE, "Invalid input shape: image too small for '" + op.name + "'");
    to.h = eh / s + 1; to.w = ew / s + 1;
    if (op.kind == OpKind::Conv && is_head_tensor[op.out]) { to.f32 = true; to.ld = M->convs[i].cout_pad; to.bytes = (size_t)n * to.h * to.w * to.ld * 4; }
    else { to.ld = to.c; to.bytes = (size_t)n * to.h * to.w * to.c * (M->i8 ? 1 : 2); }
    last_use[op.in] = (int)i;
    if (op.kind == OpKind::Conv && op.conv.residual >= 0) {
      const TensorInfo& tr = p.tensors[op.conv.residual];
      if (tr.h != to.h || tr.w != to.w || tr.c != to.c) return Status::error(INFUR_E_SHAPE, "Invalid input shape: residual size mismatch at '" + op.name + "'");
      last_use[op.conv.residual] = (int)i;
    }
    if (op.kind == OpKind::Conv && op.conv.in2 >= 0) {
      const TensorInfo& t2 = p.tensors[op.conv.in2];
      const int s2 = op.conv.stride2;
      if ((t2.h - 1) / s2 + 1 != to.h || (t2.w - 1) / s2 + 1 != to.w || t2.c != op.conv.cin2)
        return Status::error(INFUR_E_SHAPE, "Invalid input shape: shortcut size mismatch at '" + op.name + "'");
      last_use[op.conv.in2] = (int)i;
    }
    if (op.kind == OpKind::MaxPool && ti.c % (M->i8 ? 16 : 8) != 0) return Status::error(INFUR_E_UNSUPPORTED, "MaxPool needs channels % 8 == 0 (16 in an int8 plan)");
  }
  for (auto& hd : m.heads) last_use[hd.tensor] = 1 << 30;

  // ---- stem input (zero border written once; the pre-kernel only touches the interior)
  {
    const size_t bytes = (size_t)n * stem_rows(p.oh) * stem_pitch_px(p.ow) * 8;
    if (!(st = dev_alloc(p, reinterpret_cast<uint8_t**>(&p.stem_in), bytes)).ok()) return st;
    CU_TRY(cudaMemset(p.stem_in, 0, bytes));
    p.tensors[m.input_tensor].ptr = p.stem_in;
  }
  // ---- activation buffers by liveness
  struct Buf { void* ptr; size_t bytes; bool free_; };
  std::vector<Buf> pool;
  std::vector<int> tensor_buf(m.num_tensors, -1);
  for (size_t i = 0; i < m.ops.size(); ++i) {
    if (!M->needed[i]) continue;
    const LoweredOp& op = m.ops[i];
    TensorInfo& to = p.tensors[op.out];
    int pick = -1;
    for (size_t b = 0; b < pool.size(); ++b)
      if (pool[b].free_ && pool[b].bytes >= to.bytes && (pick < 0 || pool[b].bytes < pool[pick].bytes)) pick = (int)b;
    if (pick < 0) {
      uint8_t* q = nullptr;
      if (!(st = dev_alloc(p, &q, to.bytes)).ok()) return st;
      pool.push_back({q, to.bytes, false});
      pick = (int)pool.size() - 1;
      p.act_bytes += to.bytes;
    }
    pool[pick].free_ = false;
    tensor_buf[op.out] = pick;
    to.ptr = pool[pick].ptr;
    auto release = [&](int t) { if (t >= 0 && tensor_buf[t] >= 0 && last_use[t] == (int)i) pool[tensor_buf[t]].free_ = true; };
    release(op.in);
    if (op.kind == OpKind::Conv) { release(op.conv.residual); release(op.conv.in2); }
  }
  // ---- ops
  std::vector<ConvIO> ios;   // parallel to p.ops (default-constructed for pools)
  for (size_t i = 0; i < m.ops.size(); ++i) {
    if (!M->needed[i]) continue;
    const LoweredOp& op = m.ops[i];
    const TensorInfo& ti = p.tensors[op.in];
    const TensorInfo& to = p.tensors[op.out];
    PlanOp po;
    po.op = (int)i;
    ios.emplace_back();
    std::ostringstream os;
    if (op.kind == OpKind::Conv) {
      const DevConv& d = M->convs[i];
      po.is_conv = true;
      ConvIO io;
      io.x = reinterpret_cast<const __half*>(ti.ptr); io.n = n; io.h = ti.h; io.w = ti.w; io.oh = to.h; io.ow = to.w;
      io.wgt = reinterpret_cast<const __half*>(M->arena + d.w_off);
      io.bias = reinterpret_cast<const float*>(M->arena + d.b_off);
      if (d.quant) io.qmul = reinterpret_cast<const float*>(M->arena + d.q_off);
      if (d.mode == 3) io.bias_i32 = reinterpret_cast<const int32_t*>(M->arena + d.bi_off);
      io.residual = op.conv.residual >= 0 ? reinterpret_cast<const __half*>(p.tensors[op.conv.residual].ptr) : nullptr;
      if (op.conv.in2 >= 0) {
        const TensorInfo& t2 = p.tensors[op.conv.in2];
        io.x2 = reinterpret_cast<const __half*>(t2.ptr); io.h2 = t2.h; io.w2 = t2.w;
      }
      if (to.f32) io.y_f32 = reinterpret_cast<float*>(to.ptr); else io.y = reinterpret_cast<__half*>(to.ptr);
      io.out_ld = to.ld;
      ios.back() = io;
      if (d.tc_ok) { if (!(st = setup_conv_tc(d, io, po, d.block_n)).ok()) return st; }
      po.flops = 2.0 * n * to.h * to.w * (double)d.cout * (d.kh * d.kw * d.cin + d.cin2);
      if (d.tc_ok && !to.f32 && H->cfg.autotune && H->cfg.conv_impl == INFUR_CONV_TCGEN05) {
        if (!(st = tune_block_n(H, d, io, po, po.flops)).ok()) return st;
      }
      setup_direct(d, io, reinterpret_cast<const __half*>(M->arena + d.wv_off), po.direct);
      const double xsz = d.mode == 3 ? 1.0 : 2.0, rsz = d.mode >= 2 ? 1.0 : 2.0;   // element sizes of the operands / of the residual
      po.bytes = (double)n * ti.h * ti.w * d.cin * xsz + (double)to.bytes + (io.residual ? (double)n * to.h * to.w * to.c * rsz : 0.0) +
                 (double)d.cout * (d.kh * d.kw * d.cin + d.cin2) * xsz + (io.x2 ? (double)n * io.h2 * io.w2 * d.cin2 * 2 : 0.0);
      os << "conv " << op.name << " [" << n << "x" << ti.h << "x" << ti.w << "x" << d.cin << "] -> [" << to.h << "x" << to.w << "x" << d.cout
         << "] k" << d.kh << " s" << d.stride << " p" << d.pad << " d" << d.dil << (io.residual ? " +res" : "") << (d.cin2 ? " +shortcut1x1" : "") << (d.relu ? " relu" : "");
      if (d.mode == 3) os << (d.small_acc ? " int8 (acc < 2^22)" : " int8");
      if (d.tc_ok)
        os << " | tcgen05 tile " << (1 << po.geom.bw_log2) << "x" << (128 >> po.geom.bw_log2) << "px x N" << po.block_n << (po.variant == kVarPair ? " pair" : (po.variant == kVarHalo ? " halo" : (po.variant == kVarPairDeep ? " pair deep-epilogue" : ""))) << " tiles "
           << po.geom.num_tiles << " kblocks " << po.geom.num_kb;
      os << " | GFLOP " << po.flops * 1e-9 << " MB " << po.bytes * 1e-6;
    } else {
      po.flops = 0;
      po.bytes = (double)n * ti.h * ti.w * ti.c * (M->i8 ? 1 : 2) + (double)to.bytes;
      os << "maxpool " << op.name << " [" << n << "x" << ti.h << "x" << ti.w << "x" << ti.c << "] -> [" << to.h << "x" << to.w << "] k" << op.pool_k
         << " s" << op.pool_s << " | MB " << po.bytes * 1e-6;
    }
    po.text = os.str();
    p.ops.push_back(po);
  }
  if (H->cfg.conv_impl == INFUR_CONV_TCGEN05 && !M->i8 && !m.quant && !b2b_disabled_env()) {
    if (!(st = fuse_b2b_pairs(H, *M, p, ios, last_use)).ok()) return st;
  }
  if (H->cfg.conv_impl == INFUR_CONV_TCGEN05 && (M->i8 || !m.quant) && !stem_pool_disabled_env()) {
    if (!(st = fuse_stem_pool(H, *M, p, ios)).ok()) return st;
  }
  // ---- head / post
  const LoweredHead& hd = m.heads[M->out_head];
  const TensorInfo& th = p.tensors[hd.tensor];
  if (!th.f32) return Status::error(INFUR_E_UNSUPPORTED, "output head is not produced by a convolution");
  p.lowres = reinterpret_cast<float*>(th.ptr); p.lh = th.h; p.lw = th.w; p.k = hd.num_classes; p.ldk = th.ld;
  if (H->cfg.compute_aux && M->aux_head >= 0) p.aux_lowres = reinterpret_cast<float*>(p.tensors[m.heads[M->aux_head].tensor].ptr);
  if (!(st = build_bilinear(p, p.lh, p.lw)).ok()) return st;
  if (p.k == 21 && !(st = dev_alloc(p, &p.top_code, (size_t)n * p.lh * p.lw)).ok()) return st;
  out = std::move(pl);
  return Status();
}

Status get_plan(infur_b200_handle* H, int n, int w, int h, Plan** out) {
  uint32_t fbits; memcpy(&fbits, &H->factor, 4);
  auto key = std::make_tuple(n, w, h, fbits, H->model_gen);
  auto it = H->plans.find(key);
  if (it == H->plans.end()) {
    H->last_build_ms = 0.f; H->last_build_tuned = 0;
    // bounded cache, least recently used plan retired first (a plan owns its activation buffers: ~0.3 GB per 1080p frame of batch)
    while (H->plans.size() >= kMaxPlans) {
      auto victim = H->plans.begin();
      for (auto p = H->plans.begin(); p != H->plans.end(); ++p)
        if (H->plan_used[p->first] < H->plan_used[victim->first]) victim = p;
      cudaStreamSynchronize(H->stream); cudaStreamSynchronize(H->d2h);
      H->plan_used.erase(victim->first);
      H->plans.erase(victim);
    }
    const auto t0 = std::chrono::steady_clock::now();
    std::unique_ptr<Plan> p;
    Status st = build_plan(H, n, w, h, p);
    if (!st.ok()) return st;
    // build_plan uploads tables and clears borders on the legacy stream, which does not order with the handle's
    // non-blocking streams: make all of it visible before the first forward
    if (cudaDeviceSynchronize() != cudaSuccess) return Status::error(INFUR_E_RUNTIME, "plan build: device synchronisation failed");
    H->last_build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    it = H->plans.emplace(key, std::move(p)).first;
  }
  H->plan_used[key] = ++H->plan_clock;
  *out = it->second.get();
  return Status();
}

// ------------------------------------------------------------------------------------------------
// Forward: Scale -> Model -> ColorCode on device buffers.

// Scale (+ normalise) arguments of a plan: nearest maps or bilinear taps, the model's input convention.
void fill_pre_args(infur_b200_handle* H, Plan& p, const uint8_t* d_bgr, PreArgs& pa) {
  memset(&pa, 0, sizeof(pa));
  const bool unit = p.factor == 1.0f;
  pa.src = d_bgr; pa.n = p.n; pa.h = p.h; pa.w = p.w; pa.oh = p.oh; pa.ow = p.ow;
  pa.xmap = p.xmap; pa.ymap = p.ymap;
  pa.bx0 = p.sbx0; pa.bx1 = p.sbx1; pa.blx0 = p.sblx0; pa.blx1 = p.sblx1; pa.by0 = p.sby0; pa.by1 = p.sby1; pa.bly0 = p.sbly0; pa.bly1 = p.sbly1;
  // Float models: RGB + torchvision normalisation; Uint8 models: the raw bytes in B,G,R order (predict_onnx.rs:103-137,296-306)
  const bool u8_model = p.has_model && !H->model->lm.io.float_input;
  pa.lut_h = u8_model ? H->d_lut_u8 : H->d_lut_h; pa.bgr_order = u8_model ? 1 : 0;
  if (p.has_model && H->model->lm.quant) pa.lut_h = reinterpret_cast<const __half*>(H->model->arena + H->model->lut_q_off);
  pa.stem_in = p.has_model ? p.stem_in : nullptr;
  pa.scaled_bgr = unit ? nullptr : p.scaled;
}

// Enqueue the kernels of one step on `s` (plain launches; `launched` counts them).
static Status issue_forward(infur_b200_handle* H, Plan& p, const uint8_t* d_bgr, const OutPtrs& o, cudaStream_t s, cudaEvent_t* evs, uint64_t& launched);

static bool graphs_disabled_env() { const char* e = getenv("INFUR_B200_NO_GRAPH"); return e && e[0] == '1'; }

// One step = ~56 kernels.  With cfg.use_cuda_graph the sequence is captured ONCE per (plan, buffer set) into a CUDA graph
// (the conv kernels' programmatic-dependent-launch edges included) and replayed with a single cudaGraphLaunch: the host cost of
// a step drops from ~56 launches to one, which is what a single-frame call (configs[1]) waits on.
Status run_forward(infur_b200_handle* H, Plan& p, const uint8_t* d_bgr, const OutPtrs& o, cudaStream_t s, float* op_ms = nullptr,
                   cudaEvent_t* evs = nullptr) {
  (void)op_ms;
  const size_t out_px = (size_t)p.n * p.ow * p.oh;
  if (out_px == 0) return Status();
  uint64_t launched = 0;
  if (!H->cfg.use_cuda_graph || evs || !p.has_model || p.graphs_off || o.logits || o.aux_logits || H->cfg.conv_impl != INFUR_CONV_TCGEN05 || graphs_disabled_env()) {
    Status st = issue_forward(H, p, d_bgr, o, s, evs, launched);
    H->launches += launched;
    return st;
  }
  const GraphKey key{d_bgr, o.class_map, o.decoded, o.blended, o.frame_rgba, o.logits, o.aux_logits};
  auto it = p.graphs.find(key);
  if (it == p.graphs.end()) {
    if (p.graphs.size() >= 24) {   // a caller that hands in new buffers every step: graphs would never be re-used
      Status st = issue_forward(H, p, d_bgr, o, s, evs, launched);
      H->launches += launched;
      return st;
    }
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
    Status st;
    if (e == cudaSuccess) {
      st = issue_forward(H, p, d_bgr, o, s, nullptr, launched);
      e = cudaStreamEndCapture(s, &graph);
      if (st.ok() && e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
    }
    if (!st.ok() || e != cudaSuccess || !exec) {
      // capture is an optimisation of HOW the same kernels are launched: without it the step runs as plain launches
      cudaGetLastError();
      p.graphs_off = true;
      if (!st.ok()) return st;
      launched = 0;
      st = issue_forward(H, p, d_bgr, o, s, evs, launched);
      H->launches += launched;
      return st;
    }
    it = p.graphs.emplace(key, PlanGraph{exec, launched}).first;
  }
  CU_TRY(cudaGraphLaunch(it->second.exec, s));
  H->launches += it->second.kernels;
  return Status();
}

static Status issue_forward(infur_b200_handle* H, Plan& p, const uint8_t* d_bgr, const OutPtrs& o, cudaStream_t s, cudaEvent_t* evs, uint64_t& launched) {
  const size_t out_px = (size_t)p.n * p.ow * p.oh;
  const bool unit = p.factor == 1.0f;
  PreArgs pa;
  fill_pre_args(H, p, d_bgr, pa);
  const uint8_t* frame = unit ? d_bgr : p.scaled;
  int ei = 0;
  if (evs) CU_TRY(cudaEventRecord(evs[ei++], s));
  if (pa.stem_in || pa.scaled_bgr) { CU_TRY(launch_pre(pa, s)); launched++; }
  if (!p.has_model) {
    if (o.frame_rgba) { CU_TRY(launch_frame_rgba(frame, out_px, o.frame_rgba, s)); launched++; }
    return Status();
  }
  const DeviceModel& M = *H->model;
  if (evs) CU_TRY(cudaEventRecord(evs[ei++], s));
  for (PlanOp& po : p.ops) {
    const LoweredOp& op = M.lm.ops[po.op];
    if (po.skip) {   // runs inside the previous op's fused kernel; keep the event list aligned with the op list
      if (evs) CU_TRY(cudaEventRecord(evs[ei++], s));
      continue;
    }
    if (po.is_conv) {
      const DevConv& d = M.convs[po.op];
      if (H->cfg.conv_impl == INFUR_CONV_TCGEN05) CU_TRY(conv_tc_launch(po.block_n, po.maps, po.geom, H->num_sms, s));
      else CU_TRY(launch_direct_conv(po.direct, s));
    } else {
      const TensorInfo& ti = p.tensors[op.in];
      const TensorInfo& to = p.tensors[op.out];
      if (H->model->i8)
        CU_TRY(launch_maxpool3s2_u8(reinterpret_cast<const uint8_t*>(ti.ptr), reinterpret_cast<uint8_t*>(to.ptr), p.n, ti.h, ti.w, ti.c, to.h, to.w, s));
      else
        CU_TRY(launch_maxpool(reinterpret_cast<const __half*>(ti.ptr), reinterpret_cast<__half*>(to.ptr), p.n, ti.h, ti.w, ti.c, to.h, to.w,
                              op.pool_k, op.pool_s, op.pool_p, s));
    }
    launched++;
    if (evs) CU_TRY(cudaEventRecord(evs[ei++], s));
  }
  PostArgs q;
  memset(&q, 0, sizeof(q));
  q.lowres = p.lowres; q.n = p.n; q.lh = p.lh; q.lw = p.lw; q.ldk = p.ldk; q.k = p.k; q.oh = p.oh; q.ow = p.ow;
  q.y0 = p.y0; q.y1 = p.y1; q.ly0 = p.ly0; q.ly1 = p.ly1; q.x0 = p.x0; q.x1 = p.x1; q.lx0 = p.lx0; q.lx1 = p.lx1;
  q.color_lut = H->d_color_lut; q.frame_bgr = frame;
  q.class_map = o.class_map; q.decoded = o.decoded; q.blended = o.blended; q.frame_rgba = o.frame_rgba; q.logits = o.logits;
  q.max_lr = p.max_lr; q.max_lc = p.max_lc; q.top_code = p.top_code;
  q.softmax = H->cfg.confidence == INFUR_CONF_SOFTMAX ? 1 : 0;
  q.xs = p.cell_xs; q.ys = p.cell_ys;
  if (!q.decoded) q.decoded = p.d_decoded;
  if (o.aux_logits && p.aux_lowres) {
    // debug path: the aux head through the same post kernel first; only its logits are kept
    PostArgs qa = q;
    qa.lowres = p.aux_lowres; qa.logits = o.aux_logits; qa.class_map = nullptr; qa.blended = nullptr; qa.frame_rgba = nullptr;
    qa.decoded = p.d_decoded;
    CU_TRY(launch_post(qa, s));
    launched += post_launch_count(qa);
  }
  CU_TRY(launch_post(q, s));
  launched += post_launch_count(q);
  if (evs) CU_TRY(cudaEventRecord(evs[ei++], s));
  return Status();
}

// ------------------------------------------------------------------------------------------------
// Diagnostics: one convolution through either implementation, host tensors in / out.

Status conv_test_impl(infur_b200_handle* H, const infur_b200_conv_desc* cd, const uint16_t* x, const uint16_t* wgt, const float* bias,
                      const uint16_t* residual, uint16_t* y, float* y_f32, float* elapsed_ms) {
  ConvOp c;
  c.cin = (int)cd->cin; c.cout = (int)cd->cout; c.kh = (int)cd->kh; c.kw = (int)cd->kw; c.stride = (int)cd->stride; c.pad = (int)cd->pad;
  c.dil = (int)cd->dil; c.relu = cd->relu != 0;
  if (cd->qmul) {
    c.quant = true; c.q_lo = cd->q_lo; c.q_hi = cd->q_hi; c.q_ra = cd->q_ra; c.q_rb = cd->q_rb; c.q_lo2 = cd->q_lo2; c.q_hi2 = cd->q_hi2;
    c.deq_scale = cd->q_deq;
  }
  if (cd->n == 0 || cd->h == 0 || cd->w == 0 || c.cin <= 0 || c.cout <= 0 || c.kh <= 0 || c.kw <= 0 || c.stride <= 0 || c.dil <= 0 || c.pad < 0)
    return Status::error(INFUR_E_INVALID_ARG, "conv_test: bad descriptor");
  const int n = (int)cd->n, h = (int)cd->h, w = (int)cd->w;
  const int eh = h + 2 * c.pad - c.dil * (c.kh - 1) - 1, ew = w + 2 * c.pad - c.dil * (c.kw - 1) - 1;
  if (eh < 0 || ew < 0) return Status::error(INFUR_E_SHAPE, "conv_test: input smaller than the filter");
  const int oh = eh / c.stride + 1, ow = ew / c.stride + 1;
  const bool f32out = y_f32 != nullptr;
  DevConv d;
  classify_conv(c, c.cin == 3, f32out, d);
  const bool deep = cd->impl == INFUR_CONV_TCGEN05_PAIR_DEEP || cd->impl == INFUR_CONV_TCGEN05_I8_PAIR_DEEP;
  const bool pair = cd->impl == INFUR_CONV_TCGEN05_PAIR || cd->impl == INFUR_CONV_TCGEN05_I8_PAIR || deep;
  const bool halo = cd->impl == INFUR_CONV_TCGEN05_HALO;
  const bool i8 = cd->impl == INFUR_CONV_TCGEN05_I8 || cd->impl == INFUR_CONV_TCGEN05_I8_PAIR || cd->impl == INFUR_CONV_TCGEN05_I8_PAIR_DEEP;
  const bool tc = cd->impl == INFUR_CONV_TCGEN05 || pair || halo || i8;
  if (i8) {
    // int8 plan form of the layer: the RGB stem keeps fp16-carried operands and writes u8 (mode 2), everything else is native int8
    if (!c.quant) return Status::error(INFUR_E_INVALID_ARG, "conv_test: INFUR_CONV_TCGEN05_I8 needs the quantisation parameters (qmul)");
    d.mode = d.stem ? 2 : 3; d.res_zp = cd->q_zres; d.out_zp = cd->q_zout;
    if (d.mode == 3) {
      std::vector<float> wf((size_t)c.cout * c.kh * c.kw * c.cin);
      for (size_t j = 0; j < wf.size(); ++j) wf[j] = __half2float(reinterpret_cast<const __half*>(wgt)[j]);
      d.small_acc = small_acc_bound(wf.data(), wf.size() / (size_t)c.cout, c.cout, bias);
      for (int co = 0; co < c.cout; ++co) d.qmul_max = std::max(d.qmul_max, std::fabs(cd->qmul[co]));
    }
  }
  if (c.quant && !tc) return Status::error(INFUR_E_UNSUPPORTED, "conv_test: quantised layers run on the tcgen05 implementations only");
  if (tc && !d.tc_ok) return Status::error(INFUR_E_UNSUPPORTED, "conv_test: shape not supported by the tcgen05 kernel: " + d.why_not);
  Plan tmp;
  Status st;
  const size_t wcount = (size_t)c.cout * c.kh * c.kw * c.cin;
  // weights
  std::vector<__half> wp((size_t)d.cout_pad * d.kdim, __float2half_rn(0.f));
  const __half* wh = reinterpret_cast<const __half*>(wgt);
  if (d.stem) {
    for (int co = 0; co < c.cout; ++co)
      for (int ky = 0; ky < 7; ++ky)
        for (int kx = 0; kx < 7; ++kx)
          for (int ci = 0; ci < 3; ++ci) wp[stem_w_index(co, ky, kx, ci)] = wh[(((size_t)co * 7 + ky) * 7 + kx) * 3 + ci];
  } else if (d.tc_ok) {
    for (size_t j = 0; j < wcount; ++j) wp[j] = wh[j];
  }
  __half *d_w = nullptr, *d_wv = nullptr, *d_x = nullptr, *d_res = nullptr, *d_y = nullptr;
  float *d_b = nullptr, *d_yf = nullptr, *d_q = nullptr;
  std::vector<float> bp((size_t)d.cout_pad, 0.f);
  for (int co = 0; co < c.cout; ++co) bp[co] = bias[co];
  if (!(st = dev_upload(tmp, &d_w, wp)).ok() || !(st = dev_upload(tmp, &d_b, bp)).ok()) return st;
  int32_t* d_bi = nullptr;
  if (d.mode == 3) {   // native int8: s8 weights [cout_pad][kdim], int32 bias
    std::vector<int8_t> w8((size_t)d.cout_pad * d.kdim, 0);
    for (size_t j = 0; j < wcount; ++j) {
      const float f = __half2float(wh[j]);
      if (f < -128.f || f > 127.f || f != nearbyintf(f)) return Status::error(INFUR_E_INVALID_ARG, "conv_test: int8 weights must be integers in [-128, 127]");
      w8[j] = (int8_t)f;
    }
    int8_t* d_w8 = nullptr;
    if (!(st = dev_upload(tmp, &d_w8, w8)).ok()) return st;
    d_w = reinterpret_cast<__half*>(d_w8);
    std::vector<int32_t> bi((size_t)d.cout_pad, 0);
    for (int co = 0; co < c.cout; ++co) bi[co] = (int32_t)bias[co];
    if (!(st = dev_upload(tmp, &d_bi, bi)).ok()) return st;
  }
  if (c.quant) {
    std::vector<float> qp((size_t)d.cout_pad, 0.f);
    for (int co = 0; co < c.cout; ++co) qp[co] = cd->qmul[co];
    if (!(st = dev_upload(tmp, &d_q, qp)).ok()) return st;
  }
  {
    std::vector<__half> wv(wh, wh + wcount);
    if (!(st = dev_upload(tmp, &d_wv, wv)).ok()) return st;
  }
  // input
  if (d.stem) {
    const int pitch = stem_pitch_px(w), rows = stem_rows(h);
    std::vector<__half> xp((size_t)n * rows * pitch * 4, __float2half_rn(0.f));
    const __half* xh = reinterpret_cast<const __half*>(x);
    for (int i = 0; i < n; ++i)
      for (int yy = 0; yy < h; ++yy)
        for (int xx = 0; xx < w; ++xx)
          for (int ci = 0; ci < 3; ++ci)
            xp[(((size_t)i * rows + yy + kStemPadTop) * pitch + xx + kStemPadLeft) * 4 + ci] = xh[(((size_t)i * h + yy) * w + xx) * 3 + ci];
    if (!(st = dev_upload(tmp, &d_x, xp)).ok()) return st;
  } else if (d.mode == 3) {   // u8 activations (zero point 0: the caller's centred values are the raw q)
    const __half* xh = reinterpret_cast<const __half*>(x);
    std::vector<uint8_t> x8((size_t)n * h * w * c.cin);
    for (size_t j = 0; j < x8.size(); ++j) {
      const float f = __half2float(xh[j]);
      if (f < 0.f || f > 255.f) return Status::error(INFUR_E_INVALID_ARG, "conv_test: int8 activations must be in [0, 255] (input zero point 0)");
      x8[j] = (uint8_t)f;
    }
    uint8_t* d_x8 = nullptr;
    if (!(st = dev_upload(tmp, &d_x8, x8)).ok()) return st;
    d_x = reinterpret_cast<__half*>(d_x8);
  } else {
    std::vector<__half> xv(reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(x) + (size_t)n * h * w * c.cin);
    if (!(st = dev_upload(tmp, &d_x, xv)).ok()) return st;
  }
  const int out_ld = f32out ? d.cout_pad : c.cout;
  const size_t ocount = (size_t)n * oh * ow * out_ld;
  if (residual) {
    std::vector<__half> rv(reinterpret_cast<const __half*>(residual), reinterpret_cast<const __half*>(residual) + (size_t)n * oh * ow * c.cout);
    if (f32out) return Status::error(INFUR_E_UNSUPPORTED, "conv_test: residual with f32 output is not supported");
    if (d.mode >= 2) {   // u8 residual tensor: raw q = centred value + its zero point
      std::vector<uint8_t> r8(rv.size());
      for (size_t j = 0; j < rv.size(); ++j) r8[j] = (uint8_t)(__half2float(rv[j]) + (float)d.res_zp);
      uint8_t* d_r8 = nullptr;
      if (!(st = dev_upload(tmp, &d_r8, r8)).ok()) return st;
      d_res = reinterpret_cast<__half*>(d_r8);
    } else if (!(st = dev_upload(tmp, &d_res, rv)).ok()) return st;
  }
  if (f32out) { if (!(st = dev_alloc(tmp, &d_yf, ocount)).ok()) return st; CU_TRY(cudaMemset(d_yf, 0, ocount * 4)); }
  else { if (!(st = dev_alloc(tmp, &d_y, ocount)).ok()) return st; CU_TRY(cudaMemset(d_y, 0, ocount * 2)); }
  ConvIO io;
  io.x = d_x; io.n = n; io.h = h; io.w = w; io.oh = oh; io.ow = ow; io.wgt = d_w; io.bias = d_b; io.residual = d_res; io.y = d_y; io.y_f32 = d_yf;
  io.out_ld = out_ld; io.qmul = d_q; io.bias_i32 = d_bi;
  PlanOp po;
  if (pair && (d.block_n != 256 || d.stem || f32out)) return Status::error(INFUR_E_UNSUPPORTED, "conv_test: the CTA-pair variant needs cout % 256 == 0 and an fp16 output");
  if (halo && !halo_ok(d)) return Status::error(INFUR_E_UNSUPPORTED, "conv_test: the halo variant needs a 3x3 / stride 1 / pad = dilation convolution");
  if (deep && !residual) return Status::error(INFUR_E_UNSUPPORTED, "conv_test: the deep-epilogue pair variant needs a residual");
  if (tc) { if (!(st = setup_conv_tc(d, io, po, d.block_n, deep ? kVarPairDeep : (pair ? kVarPair : (halo ? kVarHalo : kVarPlain)))).ok()) return st; }
  else setup_direct(d, io, d_wv, po.direct);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = elapsed_ms ? 3 : 1;
  cudaError_t err = cudaSuccess;
  for (int r = 0; r < reps && err == cudaSuccess; ++r) {
    if (r == reps - 1) cudaEventRecord(e0, H->stream);
    err = tc ? conv_tc_launch(po.block_n, po.maps, po.geom, H->num_sms, H->stream) : launch_direct_conv(po.direct, H->stream);
    H->launches++;
  }
  cudaEventRecord(e1, H->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(H->stream);
  float ms = 0.f;
  if (err == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (err != cudaSuccess) return Status::error(INFUR_E_RUNTIME, std::string("conv_test: ") + cudaGetErrorString(err));
  if (elapsed_ms) *elapsed_ms = ms;
  if (f32out) {
    std::vector<float> tmpo(ocount);
    CU_TRY(cudaMemcpy(tmpo.data(), d_yf, ocount * 4, cudaMemcpyDeviceToHost));
    for (size_t px = 0; px < (size_t)n * oh * ow; ++px)
      for (int co = 0; co < c.cout; ++co) y_f32[px * c.cout + co] = tmpo[px * out_ld + co];
  } else {
    if (d.mode >= 2) {   // u8 output: hand back the centred values (q - zero point) like the fp16-carried form does
      std::vector<uint8_t> y8(ocount);
      CU_TRY(cudaMemcpy(y8.data(), d_y, ocount, cudaMemcpyDeviceToHost));
      __half* yh = reinterpret_cast<__half*>(y);
      for (size_t j = 0; j < ocount; ++j) yh[j] = __float2half_rn((float)((int)y8[j] - d.out_zp));
    } else
    CU_TRY(cudaMemcpy(y, d_y, ocount * 2, cudaMemcpyDeviceToHost));
  }
  return Status();
}

}  // namespace infur

// tcgen05 implicit-GEMM convolution kernels (see conv_tc.h for the math and the data layout).
//
// Four persistent, warp-specialised kernels (one CTA per SM, 384 threads) that share one epilogue:
//   conv_tc_kernel<N>     one CTA per 128-pixel x N tile; K blocks = (64-channel chunk, filter tap), tap-wise TMA loads
//   conv_tc_pair_kernel   CTA pair (cta_group::2): 256 x 256 tile, each CTA loads its A half and half of the weights
//   conv_halo_kernel<N>   3x3 / stride 1: one halo patch per chunk, nine taps = nine shifted UMMA descriptors
//   stem_tc_kernel        7x7 / stride 2 on RGB: raw input rows read as Toeplitz operands, weights resident in smem
// Warp roles in all of them:
//   warp 0    TMA producer   - one elected lane issues the operand loads into a ring of smem stages
//   warp 1    MMA issuer     - one elected lane issues 4 x tcgen05.mma (M128 x N x K16) per K block into one of the
//                              2 (N = 256) or 4 (N <= 128) TMEM accumulator stages; tcgen05.commit releases the smem
//                              stage and signals the epilogue
//   warp 2    TMEM allocator
//   warp 3    epilogue DMA   - one elected lane issues the TMA stores of finished output chunks and the TMA
//                              prefetch of residual chunks (all bulk-group accounting in one thread)
//   warps 4-11 epilogue      - tcgen05.ld the finished accumulator (thread = output pixel, 32 channels), + bias
//                              (+ residual) -> fp16 -> ReLU, overlapping the next tiles' MMAs through the other
//                              accumulator stages.  Outputs are staged per 64-channel chunk in 128B-swizzled smem
//                              and written with TMA stores (which also clip tiles that overhang the image); the
//                              residual chunk is TMA-prefetched into the same buffer up to 4 chunks ahead, so the
//                              HBM traffic of the epilogue is fully asynchronous.  The f32 logit head (21 channels)
//                              uses direct stores.
// Every variant accumulates the K blocks of an output element in the same (chunk-major) order: results are
// bit-identical whichever variant the plan-time autotuner picks.
// Each kernel is instantiated per MODE: 0 the fp16 form; 1 quantised layers carried in fp16 (the epilogue requantises exact
// integer accumulators like QLinearConv / QLinearAdd, ConvTcGeom::quant, onnx_reader.h ConvOp); 2 the same with u8 outputs
// (the stem of an int8 plan); 3 native int8: u8 activations x s8 weights through tcgen05.mma.kind::i8 into s32 accumulators.
// All are launched with programmatic stream serialization: the prologue (barrier init, TMEM allocation, descriptor
// prefetch) overlaps the previous kernel's tail, griddepcontrol.wait precedes the first global access.
#include "conv_tc.h"

#include <cstdlib>

#include <type_traits>

#include "ptx.cuh"

namespace infur {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                         // fp16 elements = one 128-byte swizzle span
constexpr int kABytes = kBlockM * kBlockK * 2;      // 16 KB
constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr int kDmaWarp = 3;

constexpr int kMaxStages = 8;
constexpr int kEpiBufBytes = kBlockM * 64 * 2;      // one 64-channel output chunk: 128 rows x 128 B
constexpr int kMaxEpiBufs = 8;                       // int8 plans split the same 64 KB into eight 8 KB chunk buffers
constexpr int kMaxAcc = 4;
constexpr int kSmemLimit = 232448;                  // 227 KB per CTA
constexpr int kBarBytes = 512;

// I8: u8 activations / s8 weights (1 byte per element): a K block of 64 channels is a 64-byte row, 64B swizzle
template <int BLOCK_N, bool I8 = false>
struct Cfg {
  static constexpr int kEsz = I8 ? 1 : 2;
  static constexpr int kA = kBlockM * kBlockK * kEsz;
  static constexpr int kBBytes = BLOCK_N * kBlockK * kEsz;
  static constexpr int kStageBytes = kA + kBBytes;
  static constexpr int kAcc = BLOCK_N >= 256 ? 2 : 4;     // TMEM accumulator stages: as many 128 x BLOCK_N f32 tiles as fit in 512 columns (max 4)
  static constexpr int kTmemCols = kAcc * BLOCK_N;        // power of two, 128 .. 512
  // epilogue buffers: 0 (direct stores), 2 (TMA store), 4 (TMA store + TMA residual prefetch)
  static constexpr int stages(int epi_bufs) {
    int s = (kSmemLimit - 1024 - kBarBytes - epi_bufs * kEpiBufBytes) / kStageBytes;
    return s > kMaxStages ? kMaxStages : s;
  }
  static constexpr int smem_bytes(int epi_bufs) { return stages(epi_bufs) * kStageBytes + epi_bufs * kEpiBufBytes + 1024 + kBarBytes; }
};

struct TileCoord { int nt, ox0, oy0, img; };
__device__ __forceinline__ TileCoord decode_tile(const ConvTcGeom& g, int tile) {
  TileCoord t;
  t.nt = tile % g.tiles_n;
  int m = tile / g.tiles_n;
  const int tx = m % g.tiles_x; m /= g.tiles_x;
  const int ty = m % g.tiles_y;
  t.img = m / g.tiles_y;
  t.ox0 = tx << g.bw_log2;
  t.oy0 = ty * (kBlockM >> g.bw_log2);
  return t;
}

// Static persistent schedules: which linear tile (m * tiles_n + nt) a CTA works on in its it-th round, -1 = done.
struct Sched1 {   // one CTA per tile
  int first, step, total;
  __device__ __forceinline__ int tile(int it) const { const int t = first + it * step; return t < total ? t : -1; }
  __device__ __forceinline__ int count() const { return first < total ? (total - first + step - 1) / step : 0; }
};
struct Sched2 {   // one CTA pair per (two consecutive M tiles, one N tile); an odd last M tile pairs with a tile outside the batch
  int cluster, nclusters, rank, tiles_n, num_work;
  __device__ __forceinline__ int tile(int it) const {
    const int w = cluster + it * nclusters;
    if (w >= num_work) return -1;
    return (2 * (w / tiles_n) + rank) * tiles_n + w % tiles_n;
  }
  __device__ __forceinline__ int count() const { return cluster < num_work ? (num_work - cluster + nclusters - 1) / nclusters : 0; }
};

// Round-half-even of a value already clamped to |v| <= 2^22, as two full-rate adds (the conversion instruction behind
// rintf runs at a quarter of that rate and paced the epilogue of the HBM-bound quantised layers).  Clamping first is
// equivalent: the clamp bounds are integers, so clamp(rne(v)) == rne(clamp(v)).
constexpr float kRneMagic = 12582912.f;   // 1.5 * 2^23
__device__ __forceinline__ float rne_small(float v) { return __fadd_rn(__fadd_rn(v, kRneMagic), -kRneMagic); }
__device__ __forceinline__ float requant(float v, float m, float lo, float hi) { return rne_small(fminf(fmaxf(__fmul_rn(v, m), lo), hi)); }
__device__ __forceinline__ float requant_add(float a, float ra, float b, float rb, float lo, float hi) {
  return rne_small(fminf(fmaxf(__fadd_rn(__fmul_rn(a, ra), __fmul_rn(b, rb)), lo), hi));
}

// Direct-store epilogue (f32 logit head): thread = output pixel, 32 channels at a time.
template <int BLOCK_N, int ACC, int MODE>
__device__ __forceinline__ void epilogue_direct(const ConvTcGeom& g, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0, int quad, int lane,
                                                int row) {
  const int bw_log2 = g.bw_log2;
  const int px = row & ((1 << bw_log2) - 1), py = row >> bw_log2;
  int it = 0;
  for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
    const TileCoord tc = decode_tile(g, tile);
    const int ox = tc.ox0 + px, oy = tc.oy0 + py;
    const bool valid = ox < g.ow && oy < g.oh;
    const size_t pix = ((size_t)tc.img * g.oh + oy) * g.ow + ox;
    const int n0 = tc.nt * BLOCK_N;
    const int as = it % ACC;
    const uint32_t aphase = (uint32_t)(it / ACC) & 1u;
    ptx::mbar_wait(tfull0 + 8u * as, aphase);
    ptx::tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BLOCK_N);
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t acc[32];
      ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)c0, acc);
      ptx::tmem_ld_wait();
      if (valid) {
        float v[32];
        if (MODE == 3) {   // s32 accumulators + int32 bias, converted like QLinearConv does: f32(acc + bias)
          const int4* b4 = reinterpret_cast<const int4*>(g.bias_i32 + n0 + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int4 b = __ldg(b4 + j);
            v[4 * j + 0] = __int2float_rn((int)acc[4 * j + 0] + b.x);
            v[4 * j + 1] = __int2float_rn((int)acc[4 * j + 1] + b.y);
            v[4 * j + 2] = __int2float_rn((int)acc[4 * j + 2] + b.z);
            v[4 * j + 3] = __int2float_rn((int)acc[4 * j + 3] + b.w);
          }
        } else {
          const float4* b4 = reinterpret_cast<const float4*>(g.bias + n0 + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            v[4 * j + 0] = __uint_as_float(acc[4 * j + 0]) + b.x;
            v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + b.y;
            v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + b.z;
            v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + b.w;
          }
        }
        const size_t off = pix * (size_t)g.out_ld + (size_t)(n0 + c0);
        if (MODE >= 1) {   // QLinearConv requantisation (ConvTcGeom::quant)
          const float4* m4 = reinterpret_cast<const float4*>(g.qmul + n0 + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 m = __ldg(m4 + j);
            v[4 * j + 0] = requant(v[4 * j + 0], m.x, g.q_lo, g.q_hi);
            v[4 * j + 1] = requant(v[4 * j + 1], m.y, g.q_lo, g.q_hi);
            v[4 * j + 2] = requant(v[4 * j + 2], m.z, g.q_lo, g.q_hi);
            v[4 * j + 3] = requant(v[4 * j + 3], m.w, g.q_lo, g.q_hi);
          }
        }
        if (g.residual != nullptr) {
          const uint4* r4 = reinterpret_cast<const uint4*>(g.residual + off);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 r = __ldg(r4 + j);
            const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 f = __half22float2(h[q]);
              if (MODE >= 1) {   // QLinearAdd
                v[8 * j + 2 * q] = requant_add(v[8 * j + 2 * q], g.q_ra, f.x, g.q_rb, g.q_lo2, g.q_hi2);
                v[8 * j + 2 * q + 1] = requant_add(v[8 * j + 2 * q + 1], g.q_ra, f.y, g.q_rb, g.q_lo2, g.q_hi2);
              } else {
                v[8 * j + 2 * q] += f.x;
                v[8 * j + 2 * q + 1] += f.y;
              }
            }
          }
        }
        if (g.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (g.out_f32 != nullptr) {
          if (MODE >= 1 && g.q_deq != 0.f) {   // DequantizeLinear of the logits
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __fmul_rn(v[j], g.q_deq);
          }
          float4* o4 = reinterpret_cast<float4*>(g.out_f32 + off);
#pragma unroll
          for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
          uint4* o4 = reinterpret_cast<uint4*>(g.out + off);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int q = 0; q < 4; ++q) h[q] = __floats2half2_rn(v[8 * j + 2 * q], v[8 * j + 2 * q + 1]);
            o4[j] = o;
          }
        }
      }
    }
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(tempty0 + 8u * as);
  }
}

// Requantisation of 32 channels of one pixel to u8 (int8 plans), in f32 exactly as the fp16-carried form does it (requant /
// requant_add above); what differs is getting integers in and out cheaply:
//  * a residual byte b becomes the float 2^23 + b by placing it in the low mantissa byte of 0x4B000000 (PRMT), so b - zero_point
//    is ONE add of -(2^23 + zero_point)  (exact);
//  * rne(v) for |v| < 2^22 is v + 1.5 * 2^23 (one add): the low mantissa bits of the result hold rne(v) as a two's-complement
//    integer.  (Adding the zero point before rounding would break ties differently whenever it is odd.)
// VAR 0: generic -- the clamp happens in f32 before the rounding add (so that add is valid whatever the accumulator), the zero
//        point is added to the mantissa bits and PRMT gathers the low bytes.
// VAR 1: integer tail (ConvTcGeom::q_tail: the value that reaches the LAST rounding add is below 2^22 in magnitude whatever the
//        input) -- no f32 clamp there: the integer rne(v) + zero_point saturates to [0, 255] inside the pack instruction
//        (I2IP.SAT, two per four channels), then one byte-wise max with the lower bound when that is above 0.  Same result:
//        clamp and rne commute (integer bounds), and q_hi + zero_point == 255 is checked where q_tail is set.
// VAR 2: VAR 1 + small accumulators (MODE 3, |acc + bias| < 2^22 proven from the weights): int -> float is an integer add of the
//        bit pattern of 1.5 * 2^23 (folded into the bias add, IADD3) and one packed float subtract instead of two I2F.
// Multiplies that feed an add use mul_f32x2_sep: ptxas would otherwise fuse them into FFMA2 and drop a rounding (ptx.cuh).
template <int MODE, bool HAS_RES, int VAR>
__device__ __forceinline__ void requant_u8_pass(const ConvTcGeom& g, const uint32_t (&acc)[32], const uint32_t (&rw)[8], uint32_t (&ow)[8], int cofs,
                                                float lo_out, float hi1, float lo2, float hi2, float res_bias, uint32_t zout, uint32_t lfloor) {
  constexpr bool ITAIL = VAR >= 1, SMALL = VAR == 2;
  constexpr int kMagicBits = 0x4B400000;          // bits of 1.5 * 2^23
  const uint32_t zadj = ITAIL ? zout - (uint32_t)kMagicBits : zout;
#pragma unroll
  for (int j = 0; j < 8; ++j) {            // 4 channels -> one 32-bit word of u8
    const float4 m = __ldg(reinterpret_cast<const float4*>(g.qmul + cofs) + j);
    float t[4];
    if (MODE == 3) {
      const int4 bi = __ldg(reinterpret_cast<const int4*>(g.bias_i32 + cofs) + j);
      if (SMALL) {
        t[0] = __int_as_float((int)acc[4 * j + 0] + bi.x + kMagicBits); t[1] = __int_as_float((int)acc[4 * j + 1] + bi.y + kMagicBits);
        t[2] = __int_as_float((int)acc[4 * j + 2] + bi.z + kMagicBits); t[3] = __int_as_float((int)acc[4 * j + 3] + bi.w + kMagicBits);
        ptx::add_f32x2(t[0], t[1], -kRneMagic, -kRneMagic); ptx::add_f32x2(t[2], t[3], -kRneMagic, -kRneMagic);
      } else {
        t[0] = __int2float_rn((int)acc[4 * j + 0] + bi.x); t[1] = __int2float_rn((int)acc[4 * j + 1] + bi.y);
        t[2] = __int2float_rn((int)acc[4 * j + 2] + bi.z); t[3] = __int2float_rn((int)acc[4 * j + 3] + bi.w);
      }
    } else {
      const float4 bf = __ldg(reinterpret_cast<const float4*>(g.bias + cofs) + j);
      t[0] = __uint_as_float(acc[4 * j + 0]); t[1] = __uint_as_float(acc[4 * j + 1]);
      t[2] = __uint_as_float(acc[4 * j + 2]); t[3] = __uint_as_float(acc[4 * j + 3]);
      ptx::add_f32x2(t[0], t[1], bf.x, bf.y); ptx::add_f32x2(t[2], t[3], bf.z, bf.w);
    }
    ptx::mul_f32x2_sep(t[0], t[1], m.x, m.y); ptx::mul_f32x2_sep(t[2], t[3], m.z, m.w);
    uint32_t bits[4];
#pragma unroll
    for (int x = 0; x < 4; x += 2) {
      float a0 = t[x], a1 = t[x + 1];
      if (HAS_RES || !ITAIL) { a0 = fminf(fmaxf(a0, lo_out), hi1); a1 = fminf(fmaxf(a1, lo_out), hi1); }
      ptx::add_f32x2(a0, a1, kRneMagic, kRneMagic);                       // rne(v) + 1.5 * 2^23
      if (HAS_RES) {
        ptx::add_f32x2(a0, a1, -kRneMagic, -kRneMagic);                   // rne(v) as a float
        float b0 = __uint_as_float(__byte_perm(rw[j], 0x4B000000u, 0x7650 + x));        // 2^23 + residual byte
        float b1 = __uint_as_float(__byte_perm(rw[j], 0x4B000000u, 0x7650 + x + 1));
        ptx::add_f32x2(b0, b1, res_bias, res_bias);                         // residual - its zero point
        ptx::mul_f32x2_sep(a0, a1, g.q_ra, g.q_ra);
        ptx::mul_f32x2_sep(b0, b1, g.q_rb, g.q_rb);
        ptx::add_f32x2(a0, a1, b0, b1);
        if (!ITAIL) { a0 = fminf(fmaxf(a0, lo2), hi2); a1 = fminf(fmaxf(a1, lo2), hi2); }
        ptx::add_f32x2(a0, a1, kRneMagic, kRneMagic);
      }
      bits[x] = __float_as_uint(a0) + zadj; bits[x + 1] = __float_as_uint(a1) + zadj;
    }
    if (ITAIL) {
      ow[j] = ptx::pack_sat_u8x4((int)bits[0], (int)bits[1], (int)bits[2], (int)bits[3]);
      if (lfloor) ow[j] = __vmaxu4(ow[j], lfloor);
    } else {
      ow[j] = __byte_perm(__byte_perm(bits[0], bits[1], 0x0040), __byte_perm(bits[2], bits[3], 0x0040), 0x5410);
    }
  }
}

// TMA-staged epilogue for NHWC outputs (fp16, or u8 in int8 plans: MODE >= 2).  The tile's output is produced in chunks
// of CW channels (64 for fp16; 128 for u8 where the N tile has them); chunk q of this CTA (running count over all its
// tiles) lives in smem buffer q % EB as 128 rows (pixels) x 128 B with the 128B swizzle the tensor maps expect
// (u8 chunks of 64-channel tiles: 64-byte rows, unswizzled).  Eight warps: warp ew owns TMEM lane quadrant ew % 4
// (32 pixels) and channel half ew / 4 of the chunk.  HAS_RES: the residual chunk was TMA-loaded into the buffer
// ahead of time by the DMA warp; the sum is written back in place.  No CTA-wide barrier: a warp announces its part
// through the chunk_ready mbarrier and moves on.
// MODE: 0 float model; 1 quantised, fp16-carried integers in and out; 2 the same accumulators, u8 out (stem of an int8
// plan); 3 s32 accumulators of tcgen05.mma.kind::i8 + int32 bias, u8 residual and output.
struct EpiBars { uint32_t res, ready, free_; };

template <int BLOCK_N, int ACC, bool HAS_RES, int MODE, class Sched, bool PAIR = false, int EB = 4>
__device__ __forceinline__ void epilogue_tma(const ConvTcGeom& g, const Sched sched, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                             const EpiBars eb, uint32_t epi_base, int ew, int lane) {
  // Output chunk: 64 channels (128-byte fp16 rows); u8 outputs use 128-channel chunks where the tile has them, so that a chunk
  // row is again a full 128-byte line (64-byte rows are partial-line writes: the TMA stores of an int8 plan's HBM-bound layers
  // ran at half the byte rate of the fp16 ones).
  constexpr int CW = (MODE >= 2 && BLOCK_N >= 128) ? 128 : 64;
  constexpr int CH = BLOCK_N / CW;            // chunks per tile
  static_assert(!HAS_RES || EB == 4 || EB == 8, "the residual prefetch cycles through all chunk buffers");
  const int quad = ew & 3, half = ew >> 2;
  const int row = quad * 32 + lane;
  const uint32_t sw = (uint32_t)(row & 7);
  int q = 0, it = 0;
  for (int tile = sched.tile(0); tile >= 0; tile = sched.tile(++it)) {
    const int n0 = (tile % g.tiles_n) * BLOCK_N;
    const int as = it % ACC;
    const uint32_t aphase = (uint32_t)(it / ACC) & 1u;
    ptx::mbar_wait(tfull0 + 8u * as, aphase);
    ptx::tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BLOCK_N + half * 32);
    if constexpr (MODE >= 2 && CH >= 1) {
      // ---- u8 outputs (int8 plan).  CW = 128: rows of 128 bytes, 128B-swizzled like the fp16 chunks, a thread owns 64 bytes of
      // its row (two passes of 32 channels); CW = 64 (64-channel tiles): rows of 64 bytes, unswizzled, 32 bytes per thread.
      constexpr int PASSES = CW / 64;
      constexpr int NP = CH * PASSES;             // passes per tile: 1, 2 or 4
      constexpr uint32_t kChunk = CW == 128 ? kEpiBufBytes : kEpiBufBytes / 2;
      const float lo1 = g.q_lo, hi1 = g.q_hi;
      const float lo2 = g.relu ? fmaxf(g.q_lo2, 0.f) : g.q_lo2, hi2 = g.q_hi2;
      const float lo_out = (!HAS_RES && g.relu) ? fmaxf(lo1, 0.f) : lo1;
      const float res_bias = -(8388608.f + g.q_zres);
      const uint32_t zout = (uint32_t)(int)(g.q_zmagic - kRneMagic);
      const uint32_t lfloor = (uint32_t)g.q_floor * 0x01010101u;
      const uint32_t t_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BLOCK_N + half * (CW / 2));
      // ONE pass body in a rolled loop: four unrolled copies (with the next pass's TMEM load in flight) were tried and measured no
      // faster -- tools/epi_tput.cu: TMEM delivers 800 B/clk/SM, the load is 1 % of a pass -- while quadrupling the loop's code.
      uint32_t acc[32];
#pragma unroll 1
      for (int p = 0; p < NP; ++p) {
        const int c = p / PASSES, hh = p % PASSES;
        ptx::tmem_ld_32x32b_x32(t_acc + (uint32_t)(c * CW + hh * 32), acc);
        ptx::tmem_ld_wait();
        if (p == NP - 1) {                       // accumulator stage fully read
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) ptx::mbar_arrive_cluster_relaxed(ptx::mapa(tempty0 + 8u * as, 0));
            else ptx::mbar_arrive(tempty0 + 8u * as);
          }
        }
        const int qq = q + c;
        const int b = qq % EB;
        const uint32_t use = (uint32_t)(qq / EB);
        const uint32_t rowb = epi_base + b * kChunk + (uint32_t)row * (uint32_t)CW;
        const int cofs = n0 + c * CW + half * (CW / 2) + hh * 32;
        uint32_t ga[2];                          // 16-byte group addresses of this pass
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
          ga[jj] = CW == 128 ? rowb + (((uint32_t)(half * 4 + hh * 2 + jj) ^ sw) << 4) : rowb + (uint32_t)(half * 32 + jj * 16);
        uint32_t rw[8];
        if (hh == 0) {
          if (HAS_RES) ptx::mbar_wait(eb.res + 8u * b, use & 1u);
          else if (use >= 1) ptx::mbar_wait(eb.free_ + 8u * b, (use - 1u) & 1u);
        }
        if (HAS_RES) {
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rw[0]), "=r"(rw[1]), "=r"(rw[2]), "=r"(rw[3]) : "r"(ga[0]));
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rw[4]), "=r"(rw[5]), "=r"(rw[6]), "=r"(rw[7]) : "r"(ga[1]));
        }
        uint32_t ow[8];
        // uniform branches, not predicates: each variant is straight-line code of its own (a predicated-off instruction still
        // takes its issue slot)
        if (MODE == 3 && g.q_tail == 2) requant_u8_pass<MODE, HAS_RES, 2>(g, acc, rw, ow, cofs, lo_out, hi1, lo2, hi2, res_bias, zout, lfloor);
        else if (g.q_tail == 1) requant_u8_pass<MODE, HAS_RES, 1>(g, acc, rw, ow, cofs, lo_out, hi1, lo2, hi2, res_bias, zout, lfloor);
        else requant_u8_pass<MODE, HAS_RES, 0>(g, acc, rw, ow, cofs, lo_out, hi1, lo2, hi2, res_bias, zout, lfloor);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ga[0]), "r"(ow[0]), "r"(ow[1]), "r"(ow[2]), "r"(ow[3]) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ga[1]), "r"(ow[4]), "r"(ow[5]), "r"(ow[6]), "r"(ow[7]) : "memory");
        if (hh == PASSES - 1) {
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(eb.ready + 8u * b);
        }
      }
      static_assert(NP == 1 || NP == 2 || NP == 4, "passes per tile");
      q += CH;
      continue;
    }
    // fp16 outputs: one 64-channel chunk after the other, software-pipelined like the u8 passes above -- the TMEM load of chunk
    // c + 1 is in flight during the arithmetic of chunk c, and the per-channel constants are requested before the wait.
    auto chunk = [&](auto cc, uint32_t (&acc)[32], uint32_t (&nxt)[32]) {
      constexpr int c = decltype(cc)::value;
      const int b = (q + c) % EB;
      const uint32_t use = (uint32_t)((q + c) / EB);
      const uint32_t rowp = epi_base + b * kEpiBufBytes + (uint32_t)row * 128u;
      float4 bias[8], qm[8];
      {
        const float4* b4 = reinterpret_cast<const float4*>(g.bias + n0 + c * 64 + half * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) bias[j] = __ldg(b4 + j);
      }
      if (MODE >= 1) {
        const float4* m4 = reinterpret_cast<const float4*>(g.qmul + n0 + c * 64 + half * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) qm[j] = __ldg(m4 + j);
      }
      ptx::tmem_ld_wait(acc);
      if constexpr (c + 1 < CH) {
        ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)((c + 1) * 64), nxt);
      } else {             // accumulator stage fully read: hand it back to the MMA warp (of the leader CTA in a pair)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) ptx::mbar_arrive_cluster_relaxed(ptx::mapa(tempty0 + 8u * as, 0));
          else ptx::mbar_arrive(tempty0 + 8u * as);
        }
      }
      uint4 res[4];
      if (HAS_RES) {
        ptx::mbar_wait(eb.res + 8u * b, use & 1u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t addr = rowp + (((uint32_t)(half * 4 + j) ^ sw) << 4);
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(res[j].x), "=r"(res[j].y), "=r"(res[j].z), "=r"(res[j].w) : "r"(addr));
        }
      } else if (use >= 1) {
        ptx::mbar_wait(eb.free_ + 8u * b, (use - 1u) & 1u);   // the previous store out of this buffer has been read
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {            // 16-byte group = 8 channels of this pixel
        const uint32_t addr = rowp + (((uint32_t)(half * 4 + j) ^ sw) << 4);
        const float4 bl = bias[2 * j], bh = bias[2 * j + 1];
        float v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(acc[8 * j + t]);
        // + bias: packed f32x2 adds (FADD2); + residual: f32 + f16 mixed adds (FHADD), no unpacking; both round like
        // the scalar f32 operations they replace
        ptx::add_f32x2(v[0], v[1], bl.x, bl.y); ptx::add_f32x2(v[2], v[3], bl.z, bl.w);
        ptx::add_f32x2(v[4], v[5], bh.x, bh.y); ptx::add_f32x2(v[6], v[7], bh.z, bh.w);
        if (MODE >= 1) {
          // quantised layer: v is the exact integer accumulator + bias; requantise like QLinearConv, then (HAS_RES) add the
          // residual like QLinearAdd.  Separate f32 multiplies and adds (no FMA contraction), round half to even.
          const float4 ml = qm[2 * j], mh = qm[2 * j + 1];
          const float mm[8] = {ml.x, ml.y, ml.z, ml.w, mh.x, mh.y, mh.z, mh.w};
#pragma unroll
          for (int t = 0; t < 8; t += 2) {
            ptx::mul_f32x2(v[t], v[t + 1], mm[t], mm[t + 1]);
            v[t] = fminf(fmaxf(v[t], g.q_lo), g.q_hi); v[t + 1] = fminf(fmaxf(v[t + 1], g.q_lo), g.q_hi);
            ptx::add_f32x2(v[t], v[t + 1], kRneMagic, kRneMagic);
            ptx::add_f32x2(v[t], v[t + 1], -kRneMagic, -kRneMagic);
          }
          if (HAS_RES) {
            const uint32_t rw[4] = {res[j].x, res[j].y, res[j].z, res[j].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 r = __half22float2(*reinterpret_cast<const __half2*>(&rw[t]));
              float a0 = v[2 * t], a1 = v[2 * t + 1], b0 = r.x, b1 = r.y;
              ptx::mul_f32x2_sep(a0, a1, g.q_ra, g.q_ra);
              ptx::mul_f32x2_sep(b0, b1, g.q_rb, g.q_rb);
              ptx::add_f32x2(a0, a1, b0, b1);
              a0 = fminf(fmaxf(a0, g.q_lo2), g.q_hi2); a1 = fminf(fmaxf(a1, g.q_lo2), g.q_hi2);
              ptx::add_f32x2(a0, a1, kRneMagic, kRneMagic);
              ptx::add_f32x2(a0, a1, -kRneMagic, -kRneMagic);
              v[2 * t] = a0; v[2 * t + 1] = a1;
            }
          }
        } else if (HAS_RES) {
          const uint32_t rw[4] = {res[j].x, res[j].y, res[j].z, res[j].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            v[2 * t] = ptx::add_f32_f16(v[2 * t], (unsigned short)(rw[t] & 0xffffu));
            v[2 * t + 1] = ptx::add_f32_f16(v[2 * t + 1], (unsigned short)(rw[t] >> 16));
          }
        }
        uint4 o;
        __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int t = 0; t < 4; ++t) h[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
        if (g.relu) {   // rounding is monotonic and keeps 0, so ReLU after the fp16 rounding equals ReLU before it
          const __half2 z = __float2half2_rn(0.f);
#pragma unroll
          for (int t = 0; t < 4; ++t) h[t] = __hmax2(h[t], z);
        }
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
      }
      ptx::fence_proxy_async_smem();           // generic-proxy writes -> visible to the TMA store
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(eb.ready + 8u * b);
    };
    {
      uint32_t accA[32], accB[32];
      ptx::tmem_ld_32x32b_x32(t_row, accA);
      chunk(std::integral_constant<int, 0>{}, accA, accB);
      if constexpr (CH >= 2) chunk(std::integral_constant<int, 1>{}, accB, accA);
      if constexpr (CH >= 4) {
        chunk(std::integral_constant<int, 2>{}, accA, accB);
        chunk(std::integral_constant<int, 3>{}, accB, accA);
      }
      static_assert(CH == 1 || CH == 2 || CH == 4, "chunks per tile");
      q += CH;
    }
  }
}

// The single thread that owns every bulk copy of the epilogue: stores chunk q when all eight warps have
// written it, then (one store later, so it never waits on the store it just issued) recycles the previous
// buffer: HAS_RES -> prefetch the residual of chunk q-1+EB into it, else -> mark it free.
template <int BLOCK_N, bool HAS_RES, class Sched, int EB = 4, int CHUNK_BYTES = kEpiBufBytes, int CW = 64>
__device__ __forceinline__ void epilogue_dma(const ConvTcMaps& maps, const ConvTcGeom& g, const Sched sched, const EpiBars eb, uint32_t epi_base) {
  constexpr int CH = BLOCK_N / CW;   // CW = channels per chunk (epilogue_tma)
  // stores kept in flight: their read-out of smem takes ~2 us to observe, so the chunk rate is (stores in flight) / that.  With
  // a residual the buffers are shared between stores in flight and prefetched residual chunks: half each.
  constexpr int D = HAS_RES ? EB / 2 : EB - 1;
  const int total = sched.count() * CH;
  auto issue_res = [&](int qq) {
    const TileCoord tc = decode_tile(g, sched.tile(qq / CH));
    const int b = qq % EB;
    ptx::mbar_expect_tx(eb.res + 8u * b, (uint32_t)CHUNK_BYTES);
    ptx::tma_load_4d(epi_base + b * CHUNK_BYTES, &maps.r, eb.res + 8u * b, tc.nt * BLOCK_N + (qq % CH) * CW, tc.ox0, tc.oy0, tc.img);
  };
  if (HAS_RES) {
    for (int p = 0; p < EB && p < total; ++p) issue_res(p);
  }
  int q = 0, it = 0;
  for (int tile = sched.tile(0); tile >= 0; tile = sched.tile(++it)) {
    const TileCoord tc = decode_tile(g, tile);
    for (int c = 0; c < CH; ++c, ++q) {
      const int b = q % EB;
      ptx::mbar_wait(eb.ready + 8u * b, (uint32_t)(q / EB) & 1u);
      ptx::tma_store_4d(&maps.c, epi_base + b * CHUNK_BYTES, tc.nt * BLOCK_N + c * CW, tc.ox0, tc.oy0, tc.img);
      ptx::tma_store_commit();
      if (HAS_RES) {
        if (q >= D - 1) {
          ptx::tma_store_wait_read<D - 1>();      // the store of chunk q-(D-1) has left its buffer: prefetch into it
          if (q - (D - 1) + EB < total) issue_res(q - (D - 1) + EB);
        }
      } else if (q >= D) {
        ptx::tma_store_wait_read<D>();
        ptx::mbar_arrive(eb.free_ + 8u * ((q - D) % EB));
      }
    }
  }
  ptx::tma_store_wait<0>();
}

template <int BLOCK_N, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ ConvTcMaps maps, const __grid_constant__ ConvTcGeom g) {
  constexpr bool I8 = MODE == 3;
  constexpr int kCW = (MODE >= 2 && BLOCK_N >= 128) ? 128 : 64;                           // channels per output chunk (epilogue_tma)
  constexpr int kChunkBytes = (MODE >= 2 && kCW == 64) ? kEpiBufBytes / 2 : kEpiBufBytes;
  constexpr int kEB4 = kChunkBytes == kEpiBufBytes ? 4 : 8, kEB2 = kEB4 / 2;               // chunk buffers in 64 KB / 32 KB of smem
  using C = Cfg<BLOCK_N, I8>;
  extern __shared__ uint8_t smem_raw[];
  const int num_stages = g.stages;
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + num_stages * C::kStageBytes;          // 1024-aligned: stage sizes are multiples of 4 KB
  const uint32_t bar_base = epi_base + g.epi_bufs * kEpiBufBytes;
  // barrier block: full[8], empty[8], tmem_full[2], tmem_empty[2], res_full[4], chunk_ready[4], buf_free[4], tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + kMaxAcc + s); };
  EpiBars eb;
  eb.res = bar_base + 8u * (2 * kMaxStages + 2 * kMaxAcc);
  eb.ready = eb.res + 8u * kMaxEpiBufs;
  eb.free_ = eb.ready + 8u * kMaxEpiBufs;
  const uint32_t tmem_ptr_addr = eb.free_ + 8u * kMaxEpiBufs;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = g.num_kb;

  if (warp == 0 && ptx::elect_one()) {
    for (int v = 0; v < kMaxViews; ++v) ptx::prefetch_tmap(&maps.a[v]);
    ptx::prefetch_tmap(&maps.b);
    if (g.store_mode != 0) ptx::prefetch_tmap(&maps.c);
    if (g.store_mode == 2) ptx::prefetch_tmap(&maps.r);
  }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < num_stages; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < C::kAcc; ++s) { ptx::mbar_init(tfull_bar(s), 1); ptx::mbar_init(tempty_bar(s), g.store_mode == 0 ? 4 : kEpiWarps); }
    for (int s = 0; s < kMaxEpiBufs; ++s) {
      ptx::mbar_init(eb.res + 8u * s, 1);
      ptx::mbar_init(eb.ready + 8u * s, kEpiWarps);
      ptx::mbar_init(eb.free_ + 8u * s, 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr_addr, C::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // everything above touched only this CTA's smem / TMEM: with a programmatic launch it overlaps the previous kernel's tail
  ptx::grid_dep_launch();
  ptx::grid_dep_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  const int bw_log2 = g.bw_log2;
  const int bh = kBlockM >> bw_log2;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        const int nt = tile % g.tiles_n;
        int m = tile / g.tiles_n;
        const int tx = m % g.tiles_x; m /= g.tiles_x;
        const int ty = m % g.tiles_y;
        const int img = m / g.tiles_y;
        const int ox0 = tx << bw_log2, oy0 = ty * bh;
        // K order: for every 64-channel chunk all filter taps (chunk-major: the same order in every kernel variant, so
        // each output element accumulates identically whichever variant the autotuner picks), then a fused shortcut tap
        auto load_kblock = [&](int tap, int cc, int kcoord) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = smem_base + stage * C::kStageBytes;
          ptx::mbar_expect_tx(full_bar(stage), (uint32_t)C::kStageBytes);
          ptx::tma_load_4d(a_dst, &maps.a[g.tap_view[tap]], full_bar(stage), cc * kBlockK, ox0 + g.tap_dx[tap], oy0 + g.tap_dy[tap], img);
          ptx::tma_load_2d(a_dst + C::kA, &maps.b, full_bar(stage), kcoord, nt * BLOCK_N);
          if (++stage == num_stages) { stage = 0; phase ^= 1u; }
        };
        for (int cc = 0; cc < g.cchunks; ++cc)
          for (int tap = 0; tap < g.main_taps; ++tap) load_kblock(tap, cc, (tap * g.cchunks + cc) * kBlockK);
        for (int tap = g.main_taps; tap < g.num_taps; ++tap)
          for (int cc = 0; cc < g.tap_cc[tap]; ++cc) load_kblock(tap, cc, (g.main_taps * g.cchunks + cc) * kBlockK);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = I8 ? ptx::make_idesc_i8(kBlockM, BLOCK_N) : ptx::make_idesc_f16(kBlockM, BLOCK_N);
      constexpr uint32_t kRowBytes = I8 ? 64 : 128;   // one K block of 64 channels per row
      constexpr int ACC = C::kAcc;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
        const int as = it % ACC;
        const uint32_t aphase = (uint32_t)(it / ACC) & 1u;
        ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);  // epilogue drained this accumulator stage
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = smem_base + stage * C::kStageBytes;
          const uint64_t a_desc = ptx::make_smem_desc(a_addr, kRowBytes);
          const uint64_t b_desc = ptx::make_smem_desc(a_addr + C::kA, kRowBytes);
#pragma unroll
          for (int k = 0; k < (I8 ? kBlockK / 32 : kBlockK / 16); ++k) {
            // advancing one instruction's K (16 fp16 or 32 int8) inside the swizzle span = +32 bytes = +2 in the (addr >> 4) field
            if (I8) ptx::umma_i8(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            else ptx::umma_f16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          ptx::umma_commit(empty_bar(stage));
          if (++stage == num_stages) { stage = 0; phase ^= 1u; }
        }
        ptx::umma_commit(tfull_bar(as));
      }
    }
  } else if (warp == kDmaWarp) {
    // ===================== epilogue DMA =====================
    if constexpr (BLOCK_N >= 64) {
      if (ptx::elect_one()) {
        const Sched1 sched{(int)blockIdx.x, (int)gridDim.x, g.num_tiles};
        if (g.store_mode == 1) {
          if (g.epi_bufs == 4) epilogue_dma<BLOCK_N, false, Sched1, kEB4, kChunkBytes, kCW>(maps, g, sched, eb, epi_base);
          else epilogue_dma<BLOCK_N, false, Sched1, kEB2, kChunkBytes, kCW>(maps, g, sched, eb, epi_base);
        } else if (g.store_mode == 2) epilogue_dma<BLOCK_N, true, Sched1, kEB4, kChunkBytes, kCW>(maps, g, sched, eb, epi_base);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    const int ew = warp - kEpiWarp0;         // ew % 4 == warp % 4: the TMEM lane quadrant this warp may read
    if (g.store_mode == 0) {
      if (ew < 4) epilogue_direct<BLOCK_N, C::kAcc, MODE>(g, tmem_base, tfull_bar(0), tempty_bar(0), ew, lane, ew * 32 + lane);
    } else if constexpr (BLOCK_N >= 64) {
      const Sched1 sched{(int)blockIdx.x, (int)gridDim.x, g.num_tiles};
      if (g.store_mode == 1) {
        if (g.epi_bufs == 4) epilogue_tma<BLOCK_N, C::kAcc, false, MODE, Sched1, false, kEB4>(g, sched, tmem_base, tfull_bar(0), tempty_bar(0), eb, epi_base, ew, lane);
        else epilogue_tma<BLOCK_N, C::kAcc, false, MODE, Sched1, false, kEB2>(g, sched, tmem_base, tfull_bar(0), tempty_bar(0), eb, epi_base, ew, lane);
      } else epilogue_tma<BLOCK_N, C::kAcc, true, MODE, Sched1, false, kEB4>(g, sched, tmem_base, tfull_bar(0), tempty_bar(0), eb, epi_base, ew, lane);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) for 256-wide N tiles: the two CTAs of a cluster compute a 256 (M) x 256 (N)
// tile.  Each CTA loads its own 128 pixels of A and HALF of the weight tile (128 of the 256 output channels), the
// leader's single thread issues tcgen05.mma.cta_group::2 (M = 256), which reads A from each CTA's smem and B from
// both, and each CTA keeps the accumulator of its own 128 pixels in its own TMEM.  Per SM and K block that is
// 32 KB of operands instead of 48 KB: less L2 -> SM traffic per FLOP and room for 5-6 pipeline stages instead of
// 3-4.  Everything after the accumulator (epilogue, TMA stores, residual prefetch) is per CTA and unchanged.
//   full[s]   lives in the leader: its producer arms 64 KB; both CTAs' TMA loads complete_tx on it
//   empty[s], tmem_full[a]   one per CTA, signalled by the leader's multicast tcgen05.commit
//   tmem_empty[a]   in the leader, counts the epilogue warps of BOTH CTAs (remote mbarrier arrive)
constexpr int kPairStageBytes = kABytes + 128 * kBlockK * 2;   // 32 KB; int8 plans: half of it

__host__ __device__ constexpr int pair_stages(int epi_bufs, bool i8 = false) {
  const int s = (kSmemLimit - 1024 - kBarBytes - epi_bufs * kEpiBufBytes) / (i8 ? kPairStageBytes / 2 : kPairStageBytes);
  return s > kMaxStages ? kMaxStages : s;
}
__host__ __device__ constexpr int pair_smem_bytes(int epi_bufs, bool i8 = false) {
  return pair_stages(epi_bufs, i8) * (i8 ? kPairStageBytes / 2 : kPairStageBytes) + epi_bufs * kEpiBufBytes + 1024 + kBarBytes;
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_tc_pair_kernel(const __grid_constant__ ConvTcMaps maps, const __grid_constant__ ConvTcGeom g) {
  constexpr int BLOCK_N = 256;
  constexpr bool I8 = MODE == 3;
  constexpr int kStage = I8 ? kPairStageBytes / 2 : kPairStageBytes;   // A (128 px) + half of B (128 channels) per CTA
  constexpr int kAB = I8 ? kABytes / 2 : kABytes;
  constexpr uint32_t kRowBytes = I8 ? 64 : 128;
  constexpr int ACC = 2;
  extern __shared__ uint8_t smem_raw[];
  const int num_stages = g.stages;
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + num_stages * kStage;
  const uint32_t bar_base = epi_base + g.epi_bufs * kEpiBufBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + kMaxAcc + s); };
  EpiBars eb;
  eb.res = bar_base + 8u * (2 * kMaxStages + 2 * kMaxAcc);
  eb.ready = eb.res + 8u * kMaxEpiBufs;
  eb.free_ = eb.ready + 8u * kMaxEpiBufs;
  const uint32_t tmem_ptr_addr = eb.free_ + 8u * kMaxEpiBufs;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();
  const Sched2 sched{(int)blockIdx.x >> 1, (int)gridDim.x >> 1, rank, g.tiles_n, g.num_work};

  if (warp == 0 && ptx::elect_one()) {
    for (int v = 0; v < kMaxViews; ++v) ptx::prefetch_tmap(&maps.a[v]);
    ptx::prefetch_tmap(&maps.b);
    ptx::prefetch_tmap(&maps.c);
    if (g.store_mode == 2) ptx::prefetch_tmap(&maps.r);
  }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < num_stages; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < ACC; ++s) { ptx::mbar_init(tfull_bar(s), 1); ptx::mbar_init(tempty_bar(s), 2 * kEpiWarps); }
    for (int s = 0; s < kMaxEpiBufs; ++s) {
      ptx::mbar_init(eb.res + 8u * s, 1);
      ptx::mbar_init(eb.ready + 8u * s, kEpiWarps);
      ptx::mbar_init(eb.free_ + 8u * s, 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2sm(tmem_ptr_addr, 2 * BLOCK_N);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();      // barrier inits and TMEM of both CTAs are in place before anything crosses the pair
  ptx::tc_fence_after();
  ptx::grid_dep_launch();
  ptx::grid_dep_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = sched.tile(0); tile >= 0; tile = sched.tile(++it)) {
        const TileCoord tc = decode_tile(g, tile);
        auto load_kblock = [&](int tap, int cc, int kcoord) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = smem_base + stage * kStage;
          const uint32_t lead_full = ptx::mapa(full_bar(stage), 0);
          if (rank == 0) ptx::mbar_expect_tx(full_bar(stage), 2u * (uint32_t)kStage);
          ptx::tma_load_4d_2sm(a_dst, &maps.a[g.tap_view[tap]], lead_full, cc * kBlockK, tc.ox0 + g.tap_dx[tap], tc.oy0 + g.tap_dy[tap], tc.img);
          ptx::tma_load_2d_2sm(a_dst + kAB, &maps.b, lead_full, kcoord, tc.nt * BLOCK_N + rank * 128);
          if (++stage == num_stages) { stage = 0; phase ^= 1u; }
        };
        for (int cc = 0; cc < g.cchunks; ++cc)      // chunk-major K order, see conv_tc_kernel
          for (int tap = 0; tap < g.main_taps; ++tap) load_kblock(tap, cc, (tap * g.cchunks + cc) * kBlockK);
        for (int tap = g.main_taps; tap < g.num_taps; ++tap)
          for (int cc = 0; cc < g.tap_cc[tap]; ++cc) load_kblock(tap, cc, (g.main_taps * g.cchunks + cc) * kBlockK);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t idesc = I8 ? ptx::make_idesc_i8(2 * kBlockM, BLOCK_N) : ptx::make_idesc_f16(2 * kBlockM, BLOCK_N);
      const int num_kb = g.num_kb;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = sched.tile(0); tile >= 0; tile = sched.tile(++it)) {
        const int as = it % ACC;
        const uint32_t aphase = (uint32_t)(it / ACC) & 1u;
        ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);  // the epilogues of BOTH CTAs drained this accumulator stage
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = smem_base + stage * kStage;
          const uint64_t a_desc = ptx::make_smem_desc(a_addr, kRowBytes);
          const uint64_t b_desc = ptx::make_smem_desc(a_addr + kAB, kRowBytes);
#pragma unroll
          for (int k = 0; k < (I8 ? kBlockK / 32 : kBlockK / 16); ++k) {
            if (I8) ptx::umma_i8_2sm(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            else ptx::umma_f16_2sm(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          ptx::umma_commit_2sm(empty_bar(stage), 3);
          if (++stage == num_stages) { stage = 0; phase ^= 1u; }
        }
        ptx::umma_commit_2sm(tfull_bar(as), 3);
      }
    }
  } else if (warp == kDmaWarp) {
    if (ptx::elect_one()) {
      if (g.store_mode == 1) epilogue_dma<BLOCK_N, false, Sched2, 4, kEpiBufBytes, (MODE >= 2 ? 128 : 64)>(maps, g, sched, eb, epi_base);
      else if (g.epi_bufs == 8) epilogue_dma<BLOCK_N, true, Sched2, 8, kEpiBufBytes, (MODE >= 2 ? 128 : 64)>(maps, g, sched, eb, epi_base);
      else epilogue_dma<BLOCK_N, true, Sched2, 4, kEpiBufBytes, (MODE >= 2 ? 128 : 64)>(maps, g, sched, eb, epi_base);
    }
  } else if (warp >= kEpiWarp0) {
    const int ew = warp - kEpiWarp0;
    if (g.store_mode == 1) epilogue_tma<BLOCK_N, ACC, false, MODE, Sched2, true, 4>(g, sched, tmem_base, tfull_bar(0), tempty_bar(0), eb, epi_base, ew, lane);
    else if (g.epi_bufs == 8) epilogue_tma<BLOCK_N, ACC, true, MODE, Sched2, true, 8>(g, sched, tmem_base, tfull_bar(0), tempty_bar(0), eb, epi_base, ew, lane);
    else epilogue_tma<BLOCK_N, ACC, true, MODE, Sched2, true, 4>(g, sched, tmem_base, tfull_bar(0), tempty_bar(0), eb, epi_base, ew, lane);
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();      // nobody leaves while the peer may still touch this CTA's smem, barriers or TMEM
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 2 * BLOCK_N);
  }
}

// ------------------------------------------------------------------------------------------------
// Halo variant for 3x3 / stride 1 convolutions (pad = dilation d in {1, 2, 4}).  The tap-wise kernels above fetch the
// activation tile nine times, once per filter tap, shifted by the tap offset: 9 x 16 KB through the L2 -> SM port
// per 64-channel chunk.  Here the tile's HALO PATCH ((16 + 2d) rows x 16 pixels x 64 channels) is fetched ONCE per
// chunk, and the nine taps are nine UMMA descriptors into it: with an 8-pixel-wide x 16-row output tile, accumulator
// row m = 8 py + px, an 8-row core-matrix group is one patch row, so start address = patch + ((ky d) * 16 + kx d) * 128 B
// and stride-between-groups = one patch row (2 KB) select exactly the shifted window.  The 128B swizzle is a function of
// the shared-memory ADDRESS bits (that is what lets TMA and UMMA agree on it), so a window that starts at any 128-byte
// row reads back what TMA wrote -- verified on hardware with the descriptor's base-offset field left 0 (setting it to
// (start >> 7) & 7 gives wrong results).  Weights stream through their own ring, one (tap, chunk) tile per stage.
constexpr int kHaloPatchPx = 16;                    // patch pitch in pixels: 8 + 2d <= 16, and a multiple of the 8-row swizzle atom
constexpr int kHaloMaxBStages = 12;
constexpr int kHaloMaxPatches = 4;

// smem plan of the halo kernel for dilation d and N tile bn: up to 3 patch buffers (the patch of the NEXT tiles must be
// in flight while this one is consumed: a 64 -> 64 layer spends only ~0.6 us of MMAs per patch), then the weight ring,
// then the epilogue chunk buffers (4, or 2 when that leaves fewer than 4 weight stages).
struct HaloPlan { int patch_bytes, patches, b_stages, epi_bufs, smem_bytes; };
__host__ __device__ constexpr HaloPlan halo_plan(int d, int bn) {
  HaloPlan p{};
  p.patch_bytes = kHaloPatchPx * (16 + 2 * d) * 128;
  p.patches = 3;
  const int b_bytes = bn * kBlockK * 2;
  p.epi_bufs = 4;
  int rest = kSmemLimit - 1024 - kBarBytes - p.patches * p.patch_bytes - p.epi_bufs * kEpiBufBytes;
  if (rest / b_bytes < 4) { p.epi_bufs = 2; rest += 2 * kEpiBufBytes; }
  if (rest / b_bytes < 3) { p.patches = 2; rest += p.patch_bytes; }
  p.b_stages = rest / b_bytes > kHaloMaxBStages ? kHaloMaxBStages : rest / b_bytes;
  p.smem_bytes = p.patches * p.patch_bytes + p.b_stages * b_bytes + p.epi_bufs * kEpiBufBytes + 1024 + kBarBytes;
  return p;
}

template <int BLOCK_N>
struct HaloCfg {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kAcc = BLOCK_N >= 256 ? 2 : 4;
  static constexpr int kTmemCols = kAcc * BLOCK_N;
};

template <int BLOCK_N, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ ConvTcMaps maps, const __grid_constant__ ConvTcGeom g) {
  using C = HaloCfg<BLOCK_N>;
  constexpr int ACC = C::kAcc;
  const int d = g.halo_dil;
  const HaloPlan hp = halo_plan(d, BLOCK_N);
  const int BS = hp.b_stages, NP = hp.patches;
  const uint32_t patch_bytes = (uint32_t)hp.patch_bytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base + NP * patch_bytes;             // patch sizes are multiples of 2 KB: 1024-aligned
  const uint32_t epi_base = b_base + BS * C::kBBytes;
  const uint32_t bar_base = epi_base + hp.epi_bufs * kEpiBufBytes;
  // barriers: b_full[12], b_empty[12], a_full[4], a_empty[4], tmem_full[4], tmem_empty[4], epilogue (res, ready, free) x 4
  auto bfull_bar = [&](int s) { return bar_base + 8u * s; };
  auto bempty_bar = [&](int s) { return bar_base + 8u * (kHaloMaxBStages + s); };
  auto afull_bar = [&](int s) { return bar_base + 8u * (2 * kHaloMaxBStages + s); };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (2 * kHaloMaxBStages + kHaloMaxPatches + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kHaloMaxBStages + 2 * kHaloMaxPatches + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kHaloMaxBStages + 2 * kHaloMaxPatches + kMaxAcc + s); };
  EpiBars eb;
  eb.res = bar_base + 8u * (2 * kHaloMaxBStages + 2 * kHaloMaxPatches + 2 * kMaxAcc);
  eb.ready = eb.res + 8u * kMaxEpiBufs;
  eb.free_ = eb.ready + 8u * kMaxEpiBufs;
  const uint32_t tmem_ptr_addr = eb.free_ + 8u * kMaxEpiBufs;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const Sched1 sched{(int)blockIdx.x, (int)gridDim.x, g.num_tiles};

  if (warp == 0 && ptx::elect_one()) { ptx::prefetch_tmap(&maps.a[0]); ptx::prefetch_tmap(&maps.b); ptx::prefetch_tmap(&maps.c); }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < BS; ++s) { ptx::mbar_init(bfull_bar(s), 1); ptx::mbar_init(bempty_bar(s), 1); }
    for (int s = 0; s < NP; ++s) { ptx::mbar_init(afull_bar(s), 1); ptx::mbar_init(aempty_bar(s), 1); }
    for (int s = 0; s < ACC; ++s) { ptx::mbar_init(tfull_bar(s), 1); ptx::mbar_init(tempty_bar(s), kEpiWarps); }
    for (int s = 0; s < kMaxEpiBufs; ++s) {
      ptx::mbar_init(eb.res + 8u * s, 1);
      ptx::mbar_init(eb.ready + 8u * s, kEpiWarps);
      ptx::mbar_init(eb.free_ + 8u * s, 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr_addr, C::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // everything above touched only this CTA's smem / TMEM: with a programmatic launch it overlaps the previous kernel's tail
  ptx::grid_dep_launch();
  ptx::grid_dep_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      int bs = 0, au = 0;          // B ring position, patch use counter
      uint32_t bphase = 0;
      int it = 0;
      for (int tile = sched.tile(0); tile >= 0; tile = sched.tile(++it)) {
        const TileCoord tc = decode_tile(g, tile);
        for (int cc = 0; cc < g.cchunks; ++cc, ++au) {
          const int ab = au % NP;
          ptx::mbar_wait(aempty_bar(ab), ((uint32_t)(au / NP) & 1u) ^ 1u);
          ptx::mbar_expect_tx(afull_bar(ab), patch_bytes);
          ptx::tma_load_4d(smem_base + ab * patch_bytes, &maps.a[0], afull_bar(ab), cc * kBlockK, tc.ox0 - d, tc.oy0 - d, tc.img);
          for (int tap = 0; tap < 9; ++tap) {
            ptx::mbar_wait(bempty_bar(bs), bphase ^ 1u);
            ptx::mbar_expect_tx(bfull_bar(bs), (uint32_t)C::kBBytes);
            ptx::tma_load_2d(b_base + bs * C::kBBytes, &maps.b, bfull_bar(bs), (tap * g.cchunks + cc) * kBlockK, tc.nt * BLOCK_N);
            if (++bs == BS) { bs = 0; bphase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(kBlockM, BLOCK_N);
      int bs = 0, au = 0;
      uint32_t bphase = 0;
      int it = 0;
      for (int tile = sched.tile(0); tile >= 0; tile = sched.tile(++it)) {
        const int as = it % ACC;
        const uint32_t aphase = (uint32_t)(it / ACC) & 1u;
        ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int cc = 0; cc < g.cchunks; ++cc, ++au) {
          const int ab = au % NP;
          ptx::mbar_wait(afull_bar(ab), (uint32_t)(au / NP) & 1u);
          const uint32_t patch = smem_base + ab * patch_bytes;
          for (int tap = 0; tap < 9; ++tap) {
            ptx::mbar_wait(bfull_bar(bs), bphase);
            ptx::tc_fence_after();
            const int ky = tap / 3, kx = tap - 3 * ky;
            const uint32_t a_addr = patch + (uint32_t)((ky * d * kHaloPatchPx + kx * d) * 128);
            const uint64_t a_desc = ptx::make_smem_desc_sw128(a_addr, kHaloPatchPx * 128);
            const uint64_t b_desc = ptx::make_smem_desc(b_base + bs * C::kBBytes, 128);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              ptx::umma_f16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (cc | tap | k) != 0);
            ptx::umma_commit(bempty_bar(bs));
            if (++bs == BS) { bs = 0; bphase ^= 1u; }
          }
          ptx::umma_commit(aempty_bar(ab));
        }
        ptx::umma_commit(tfull_bar(as));
      }
    }
  } else if (warp == kDmaWarp) {
    if (ptx::elect_one()) {
      if (hp.epi_bufs == 4) epilogue_dma<BLOCK_N, false, Sched1, 4>(maps, g, sched, eb, epi_base);
      else epilogue_dma<BLOCK_N, false, Sched1, 2>(maps, g, sched, eb, epi_base);
    }
  } else if (warp >= kEpiWarp0) {
    if (hp.epi_bufs == 2) epilogue_tma<BLOCK_N, ACC, false, MODE, Sched1, false, 2>(g, sched, tmem_base, tfull_bar(0), tempty_bar(0), eb, epi_base, warp - kEpiWarp0, lane);
    else epilogue_tma<BLOCK_N, ACC, false, MODE, Sched1, false, 4>(g, sched, tmem_base, tfull_bar(0), tempty_bar(0), eb, epi_base, warp - kEpiWarp0, lane);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}


// ------------------------------------------------------------------------------------------------
// Back-to-back fusion of a bottleneck's tail:  y = relu(W3 . relu(W2 (*) x + b2) + b3 + residual)
// with W2 a 3x3 / stride 1 convolution over CMID channels (64 or 128) and W3 the 1x1 expansion that follows it.  Run as two
// kernels, the CMID-channel intermediate makes a round trip through HBM and the 1x1 layer is purely HBM-bound (it streams
// the residual and the output, 4 x CMID channels each) while the 3x3 layer is operand-load-bound -- fused, the 3x3's MMAs
// run in the shadow of the 1x1's HBM streams, and the intermediate lives only in shared memory:
//   GEMM1  acc1[128 px][CMID]  = sum over (64-channel chunk, tap) of A1(TMA, tap-shifted box) . B1          (as conv_tc_kernel)
//   epi1   acc1 + b2 -> ReLU -> fp16 -> A2 in smem, written in the 128B-swizzled K-major layout TMA would have produced
//          (the same bits the unfused 3x3 layer stores to HBM)
//   GEMM2  acc2[128 px][128]   = A2 . B2(n-tile)                        for each 128-channel tile of the 4 x CMID outputs
//   epi2   acc2 + b3 + residual (TMA-prefetched) -> fp16 -> ReLU -> TMA store                                (as epilogue_tma)
// K orders equal the unfused kernels', so the result is bit-identical to running the two layers separately.
// Warps: 0 TMA producer of the GEMM1 ring, 2 (after allocating TMEM) TMA producer of the B2 ring, 1 MMA issuer, 3 epilogue
// DMA (residual prefetch + stores), 4-11 epilogue (epi1 and epi2 of a tile in turn).  TMEM: acc1 double-buffered at columns
// [0, 2 CMID), acc2 double-buffered at [2 CMID, 2 CMID + 256).
template <int CMID>
struct B2BCfg {
  static constexpr int kCC = CMID / 64;                         // 64-channel chunks of the intermediate = K blocks of GEMM2
  static constexpr int kStage1 = kABytes + CMID * kBlockK * 2;  // A1 tile + B1 tile
  static constexpr int kN2 = 128;
  static constexpr int kStage2 = kN2 * kBlockK * 2;             // one B2 tile (128 output channels x 64)
  static constexpr int kS2 = 2;
  static constexpr int kA2 = kCC * kABytes;
  static constexpr int kEB = 4;
  static constexpr int kS1 = (kSmemLimit - 1024 - kBarBytes - kEB * kEpiBufBytes - kA2 - kS2 * kStage2) / kStage1;
  static constexpr int kSmem = kS1 * kStage1 + kS2 * kStage2 + kA2 + kEB * kEpiBufBytes + 1024 + kBarBytes;
  static constexpr int kAcc2Col0 = 2 * CMID;
  static_assert(kS1 >= 3 && kS1 <= kMaxStages, "GEMM1 ring depth");
};

struct SchedB {   // virtual tiles of the 1x1 stage: (M tile of this CTA, 128-channel N tile), N fastest
  int first, step, m_tiles, nt2;
  __device__ __forceinline__ int tile(int v) const { const int m = first + (v / nt2) * step; return m < m_tiles ? m * nt2 + v % nt2 : -1; }
  __device__ __forceinline__ int count() const { return first < m_tiles ? ((m_tiles - first + step - 1) / step) * nt2 : 0; }
};

template <int CMID>
__global__ void __launch_bounds__(kThreads, 1)
conv_b2b_kernel(const __grid_constant__ ConvTcMaps maps, const __grid_constant__ ConvTcGeom g) {
  using C = B2BCfg<CMID>;
  constexpr int S1 = C::kS1, S2 = C::kS2, CC = C::kCC, N2 = C::kN2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ring2_base = smem_base + S1 * C::kStage1;
  const uint32_t a2_base = ring2_base + S2 * C::kStage2;
  const uint32_t epi_base = a2_base + C::kA2;
  const uint32_t bar_base = epi_base + C::kEB * kEpiBufBytes;
  auto full1 = [&](int s) { return bar_base + 8u * s; };
  auto empty1 = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto full2 = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto empty2 = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 2 + s); };
  auto tfull1 = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 4 + s); };
  auto tempty1 = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 6 + s); };
  auto tfull2 = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 8 + s); };
  auto tempty2 = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 10 + s); };
  const uint32_t a2_ready = bar_base + 8u * (2 * kMaxStages + 12);
  const uint32_t a2_free = bar_base + 8u * (2 * kMaxStages + 13);
  EpiBars eb;
  eb.res = bar_base + 8u * (2 * kMaxStages + 14);
  eb.ready = eb.res + 8u * 4;
  eb.free_ = eb.ready + 8u * 4;
  const uint32_t tmem_ptr_addr = eb.free_ + 8u * 4;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nt2 = g.tiles_n;                       // 128-channel tiles of the 1x1 stage
  const int m_tiles = g.num_tiles / nt2;
  const int my_first = (int)blockIdx.x, my_step = (int)gridDim.x;
  const SchedB sched{my_first, my_step, m_tiles, nt2};

  if (warp == 0 && ptx::elect_one()) {
    for (int v = 0; v < kMaxViews; ++v) ptx::prefetch_tmap(&maps.a[v]);
    ptx::prefetch_tmap(&maps.b); ptx::prefetch_tmap(&maps.b2); ptx::prefetch_tmap(&maps.c); ptx::prefetch_tmap(&maps.r);
  }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < S1; ++s) { ptx::mbar_init(full1(s), 1); ptx::mbar_init(empty1(s), 1); }
    for (int s = 0; s < S2; ++s) { ptx::mbar_init(full2(s), 1); ptx::mbar_init(empty2(s), 1); }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tfull1(s), 1); ptx::mbar_init(tempty1(s), kEpiWarps);
      ptx::mbar_init(tfull2(s), 1); ptx::mbar_init(tempty2(s), kEpiWarps);
    }
    ptx::mbar_init(a2_ready, kEpiWarps);
    ptx::mbar_init(a2_free, 1);
    for (int s = 0; s < 4; ++s) {
      ptx::mbar_init(eb.res + 8u * s, 1);
      ptx::mbar_init(eb.ready + 8u * s, kEpiWarps);
      ptx::mbar_init(eb.free_ + 8u * s, 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr_addr, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  ptx::grid_dep_launch();
  ptx::grid_dep_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp == 0) {
    // ===================== TMA producer: GEMM1 operands =====================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int m = my_first; m < m_tiles; m += my_step) {
        const TileCoord tc = decode_tile(g, m * nt2);
        for (int cc = 0; cc < g.cchunks; ++cc)          // chunk-major K order, as in conv_tc_kernel
          for (int tap = 0; tap < g.main_taps; ++tap) {
            ptx::mbar_wait(empty1(stage), phase ^ 1u);
            const uint32_t a_dst = smem_base + stage * C::kStage1;
            ptx::mbar_expect_tx(full1(stage), (uint32_t)C::kStage1);
            ptx::tma_load_4d(a_dst, &maps.a[g.tap_view[tap]], full1(stage), cc * kBlockK, tc.ox0 + g.tap_dx[tap], tc.oy0 + g.tap_dy[tap], tc.img);
            ptx::tma_load_2d(a_dst + kABytes, &maps.b, full1(stage), (tap * g.cchunks + cc) * kBlockK, 0);
            if (++stage == S1) { stage = 0; phase ^= 1u; }
          }
      }
    }
  } else if (warp == 2) {
    // ===================== TMA producer: weights of the 1x1 stage =====================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int m = my_first; m < m_tiles; m += my_step)
        for (int j = 0; j < nt2; ++j)
          for (int k2 = 0; k2 < CC; ++k2) {
            ptx::mbar_wait(empty2(stage), phase ^ 1u);
            ptx::mbar_expect_tx(full2(stage), (uint32_t)C::kStage2);
            ptx::tma_load_2d(ring2_base + stage * C::kStage2, &maps.b2, full2(stage), k2 * kBlockK, j * N2);
            if (++stage == S2) { stage = 0; phase ^= 1u; }
          }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc1 = ptx::make_idesc_f16(kBlockM, CMID);
      constexpr uint32_t idesc2 = ptx::make_idesc_f16(kBlockM, N2);
      const int num_kb = g.num_kb;
      const int n_my = sched.count() / nt2;          // M tiles of this CTA
      // Two work streams share the tensor pipe: GEMM2 of tile t2 (short MMAs that unblock the HBM-bound epilogue) and GEMM1 of
      // tile t1 in {t2, t2 + 1} (the long K loop).  The issuer never blocks on one while the other could run: it polls the
      // barriers (try_wait) and issues whatever is ready, GEMM2 first -- so the 3x3's MMAs of the NEXT tile fill the time the
      // epilogue warps spend streaming this tile's residual and output.
      int stage1 = 0, stage2 = 0;
      uint32_t phase1 = 0, phase2 = 0;
      int t1 = 0, kb = 0;             // GEMM1 cursor
      int t2 = 0, jn = 0, k2 = 0;     // GEMM2 cursor
      int v = 0;                      // virtual tile counter of the 1x1 stage (accumulator stage / phase)
      bool g1_acc_ok = false, g2_ready = false, g2_acc_ok = false;
      long long t_idle = 0;
      while (t2 < n_my) {
        bool progressed = false;
        // ---- GEMM2(t2, jn, k2)
        if (t1 > t2) {
          if (!g2_ready && ptx::mbar_try_wait(a2_ready, (uint32_t)t2 & 1u)) g2_ready = true;
          if (g2_ready) {
            const int as2 = v & 1;
            if (!g2_acc_ok && ptx::mbar_try_wait(tempty2(as2), (((uint32_t)(v >> 1)) & 1u) ^ 1u)) g2_acc_ok = true;
            if (g2_acc_ok && ptx::mbar_try_wait(full2(stage2), phase2)) {
              ptx::tc_fence_after();
              const uint32_t d2 = tmem_base + (uint32_t)(C::kAcc2Col0 + as2 * N2);
              const uint64_t a_desc = ptx::make_smem_desc(a2_base + k2 * kABytes, 128);
              const uint64_t b_desc = ptx::make_smem_desc(ring2_base + stage2 * C::kStage2, 128);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) ptx::umma_f16(d2, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc2, (k2 | k) != 0);
              ptx::umma_commit(empty2(stage2));
              if (++stage2 == S2) { stage2 = 0; phase2 ^= 1u; }
              if (++k2 == CC) {
                k2 = 0;
                ptx::umma_commit(tfull2(as2));
                ++v; g2_acc_ok = false;
                if (++jn == nt2) {
                  jn = 0;
                  ptx::umma_commit(a2_free);      // every read of the intermediate has retired: epi1 of the next tile may overwrite it
                  ++t2; g2_ready = false;
                }
              }
              progressed = true;
            }
          }
        }
        // ---- GEMM1(t1, kb)
        if (!progressed && t1 < n_my && t1 <= t2 + 1) {
          const int as1 = t1 & 1;
          if (!g1_acc_ok && ptx::mbar_try_wait(tempty1(as1), (((uint32_t)(t1 >> 1)) & 1u) ^ 1u)) g1_acc_ok = true;
          if (g1_acc_ok && ptx::mbar_try_wait(full1(stage1), phase1)) {
            ptx::tc_fence_after();
            const uint32_t d1 = tmem_base + (uint32_t)(as1 * CMID);
            const uint32_t a_addr = smem_base + stage1 * C::kStage1;
            const uint64_t a_desc = ptx::make_smem_desc(a_addr, 128);
            const uint64_t b_desc = ptx::make_smem_desc(a_addr + kABytes, 128);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) ptx::umma_f16(d1, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc1, (kb | k) != 0);
            ptx::umma_commit(empty1(stage1));
            if (++stage1 == S1) { stage1 = 0; phase1 ^= 1u; }
            if (++kb == num_kb) {
              kb = 0;
              ptx::umma_commit(tfull1(as1));
              ++t1; g1_acc_ok = false;
            }
            progressed = true;
          }
        }
        if (progressed) t_idle = 0;
        else {   // a protocol bug must trap, never hang the device
          if (t_idle == 0) t_idle = clock64();
          else if (clock64() - t_idle > 4000000000LL) __trap();
        }
      }
    }
  } else if (warp == kDmaWarp) {
    if (ptx::elect_one()) epilogue_dma<N2, true, SchedB, 4, kEpiBufBytes, 64>(maps, g, sched, eb, epi_base);
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue warps: epi1 then epi2 of every tile =====================
    const int ew = warp - kEpiWarp0;
    const int quad = ew & 3, half = ew >> 2;
    const int row = quad * 32 + lane;
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    int it = 0, v = 0, q = 0;
    for (int m = my_first; m < m_tiles; m += my_step, ++it) {
      // ---- epi1: acc1 + b2 -> ReLU -> fp16 -> A2 (128B-swizzled K-major: 16-byte group gidx of row r sits at group gidx ^ (r & 7))
      const int as1 = it & 1;
      ptx::mbar_wait(tfull1(as1), ((uint32_t)(it >> 1)) & 1u);
      ptx::tc_fence_after();
      if (it >= 1) ptx::mbar_wait(a2_free, (uint32_t)(it - 1) & 1u);     // GEMM2 of the previous tile has read A2
#pragma unroll
      for (int c = 0; c < CC; ++c) {
        float4 bias[8];
        const float4* b4 = reinterpret_cast<const float4*>(g.bias + c * 64 + half * 32);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) bias[jj] = __ldg(b4 + jj);
        uint32_t acc[32];
        ptx::tmem_ld_32x32b_x32(lane_base + (uint32_t)(as1 * CMID + c * 64 + half * 32), acc);
        ptx::tmem_ld_wait();
        if (c == CC - 1) {   // accumulator stage fully read
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(tempty1(as1));
        }
        const uint32_t rowp = a2_base + (uint32_t)c * kABytes + (uint32_t)row * 128u;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint32_t addr = rowp + (((uint32_t)(half * 4 + jj) ^ sw) << 4);
          const float4 bl = bias[2 * jj], bh = bias[2 * jj + 1];
          float x[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) x[t] = __uint_as_float(acc[8 * jj + t]);
          ptx::add_f32x2(x[0], x[1], bl.x, bl.y); ptx::add_f32x2(x[2], x[3], bl.z, bl.w);
          ptx::add_f32x2(x[4], x[5], bh.x, bh.y); ptx::add_f32x2(x[6], x[7], bh.z, bh.w);
          uint4 o;
          __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int t = 0; t < 4; ++t) h[t] = __floats2half2_rn(x[2 * t], x[2 * t + 1]);
          const __half2 z = __float2half2_rn(0.f);
#pragma unroll
          for (int t = 0; t < 4; ++t) h[t] = __hmax2(h[t], z);
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
        }
      }
      ptx::fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(a2_ready);
      // ---- epi2: every 128-channel tile of the 1x1 stage, two 64-channel chunks each (as epilogue_tma with a residual)
      for (int j = 0; j < nt2; ++j, ++v) {
        const int as2 = v & 1;
        ptx::mbar_wait(tfull2(as2), ((uint32_t)(v >> 1)) & 1u);
        ptx::tc_fence_after();
        // the two 64-channel chunks of the tile, software-pipelined (see epilogue_tma): chunk 1's TMEM load flies during chunk 0's arithmetic
        auto chunk = [&](auto cc, uint32_t (&acc)[32], uint32_t (&nxt)[32]) {
          constexpr int c = decltype(cc)::value;
          const int b = (q + c) & 3;
          const uint32_t use = (uint32_t)((q + c) >> 2);
          const uint32_t rowp = epi_base + b * kEpiBufBytes + (uint32_t)row * 128u;
          float4 bias[8];
          const float4* b4 = reinterpret_cast<const float4*>(g.bias2 + j * N2 + c * 64 + half * 32);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) bias[jj] = __ldg(b4 + jj);
          ptx::tmem_ld_wait(acc);
          if constexpr (c == 0) {
            ptx::tmem_ld_32x32b_x32(lane_base + (uint32_t)(C::kAcc2Col0 + as2 * N2 + 64 + half * 32), nxt);
          } else {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty2(as2));
          }
          uint4 res[4];
          ptx::mbar_wait(eb.res + 8u * b, use & 1u);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint32_t addr = rowp + (((uint32_t)(half * 4 + jj) ^ sw) << 4);
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(res[jj].x), "=r"(res[jj].y), "=r"(res[jj].z), "=r"(res[jj].w) : "r"(addr));
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint32_t addr = rowp + (((uint32_t)(half * 4 + jj) ^ sw) << 4);
            const float4 bl = bias[2 * jj], bh = bias[2 * jj + 1];
            float x[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) x[t] = __uint_as_float(acc[8 * jj + t]);
            ptx::add_f32x2(x[0], x[1], bl.x, bl.y); ptx::add_f32x2(x[2], x[3], bl.z, bl.w);
            ptx::add_f32x2(x[4], x[5], bh.x, bh.y); ptx::add_f32x2(x[6], x[7], bh.z, bh.w);
            const uint32_t rw[4] = {res[jj].x, res[jj].y, res[jj].z, res[jj].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              x[2 * t] = ptx::add_f32_f16(x[2 * t], (unsigned short)(rw[t] & 0xffffu));
              x[2 * t + 1] = ptx::add_f32_f16(x[2 * t + 1], (unsigned short)(rw[t] >> 16));
            }
            uint4 o;
            __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int t = 0; t < 4; ++t) h[t] = __floats2half2_rn(x[2 * t], x[2 * t + 1]);
            if (g.relu) {
              const __half2 z = __float2half2_rn(0.f);
#pragma unroll
              for (int t = 0; t < 4; ++t) h[t] = __hmax2(h[t], z);
            }
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(eb.ready + 8u * b);
        };
        uint32_t accA[32], accB[32];
        ptx::tmem_ld_32x32b_x32(lane_base + (uint32_t)(C::kAcc2Col0 + as2 * N2 + half * 32), accA);
        chunk(std::integral_constant<int, 0>{}, accA, accB);
        chunk(std::integral_constant<int, 1>{}, accB, accA);
        q += 2;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// 7x7 / stride 2 / pad 3 stem on 3 input channels (ORT Conv of `conv1`, inside session.run predict_onnx.rs:138).
// Input: padded NHWC4 fp16 (kernels.h).  Output tile = 128 consecutive pixels of one output row.  For filter row
// ky the implicit-GEMM operand A[m][k] (m = output pixel, k = 4 * (kx + 1) + channel, 32 wide) is the fp16 at
// byte 16 m + 2 k of input row 2 oy + ky starting at pixel 2 ox0: consecutive output pixels read windows that
// start two pixels = 16 bytes apart, which is exactly the row pitch of an un-swizzled UMMA core matrix.  So the
// 7 raw input row segments (2176 B each, one TMA box) ARE the A operands: 14 tcgen05.mma (M128 N64 K16) per tile
// read them through overlapping core-matrix descriptors, no im2col copy anywhere.  Weights (28 KB) stay in smem
// for the whole persistent CTA.  Epilogue = the TMA-store epilogue above (one 64-channel chunk per tile).
constexpr int kStemRowBytes = kStemRowGroups * 128;          // 2176
constexpr int kStemStageBytes = 16384;                       // 7 rows x 2176 = 15232 B used
constexpr int kStemTxBytes = 7 * kStemRowBytes;
constexpr int kStemStages = 8;
constexpr int kStemSmemBytes = kStemStages * kStemStageBytes + 4 * kEpiBufBytes + kStemWBytes + 1024 + kBarBytes;

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
stem_tc_kernel(const __grid_constant__ ConvTcMaps maps, const __grid_constant__ ConvTcGeom g) {
  constexpr int BLOCK_N = 64;
  constexpr int ACC = 4;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + kStemStages * kStemStageBytes;
  const uint32_t w_base = epi_base + 4 * kEpiBufBytes;
  const uint32_t bar_base = w_base + kStemWBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + kMaxAcc + s); };
  EpiBars eb;
  eb.res = bar_base + 8u * (2 * kMaxStages + 2 * kMaxAcc);
  eb.ready = eb.res + 8u * kMaxEpiBufs;
  eb.free_ = eb.ready + 8u * kMaxEpiBufs;
  const uint32_t tmem_ptr_addr = eb.free_ + 8u * kMaxEpiBufs;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && ptx::elect_one()) { ptx::prefetch_tmap(&maps.a[0]); ptx::prefetch_tmap(&maps.c); }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < kStemStages; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < ACC; ++s) { ptx::mbar_init(tfull_bar(s), 1); ptx::mbar_init(tempty_bar(s), kEpiWarps); }
    for (int s = 0; s < kMaxEpiBufs; ++s) {
      ptx::mbar_init(eb.res + 8u * s, 1);
      ptx::mbar_init(eb.ready + 8u * s, kEpiWarps);
      ptx::mbar_init(eb.free_ + 8u * s, 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr_addr, ACC * BLOCK_N);
    ptx::tmem_relinquish();
  }
  {  // weights: global (already in smem order) -> smem, once per CTA
    const uint4* src = reinterpret_cast<const uint4*>(g.stem_w);
    for (int i = threadIdx.x; i < kStemWBytes / 16; i += kThreads) {
      const uint4 v = __ldg(src + i);
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(w_base + 16u * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
    ptx::fence_proxy_async_smem();   // generic writes -> visible to the tensor core's async-proxy reads
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // everything above touched only this CTA's smem / TMEM: with a programmatic launch it overlaps the previous kernel's tail
  ptx::grid_dep_launch();
  ptx::grid_dep_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(g, tile);
        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
        ptx::mbar_expect_tx(full_bar(stage), (uint32_t)kStemTxBytes);
        // row groups of 64 fp16 = 16 pixels: output pixel ox0 starts at input pixel 2 ox0 = group ox0 / 8
        ptx::tma_load_4d(smem_base + stage * kStemStageBytes, &maps.a[0], full_bar(stage), 0, tc.ox0 >> 3, 2 * tc.oy0, tc.img);
        if (++stage == kStemStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(kBlockM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
        const int as = it % ACC;
        const uint32_t aphase = (uint32_t)(it / ACC) & 1u;
        ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);
        ptx::mbar_wait(full_bar(stage), phase);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
        const uint32_t a_addr = smem_base + stage * kStemStageBytes;
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint64_t a_desc = ptx::make_smem_desc_noswz(a_addr + ky * kStemRowBytes + kk * 32, 16, 128);
            const uint64_t b_desc = ptx::make_smem_desc_noswz(w_base + ky * 4096 + kk * 2048, 1024, 128);
            ptx::umma_f16(d_tmem, a_desc, b_desc, idesc, (ky | kk) != 0);
          }
        }
        ptx::umma_commit(empty_bar(stage));
        ptx::umma_commit(tfull_bar(as));
        if (++stage == kStemStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == kDmaWarp) {
    if (ptx::elect_one()) epilogue_dma<BLOCK_N, false, Sched1, (MODE >= 2 ? 8 : 4), (MODE >= 2 ? kEpiBufBytes / 2 : kEpiBufBytes)>(maps, g, Sched1{(int)blockIdx.x, (int)gridDim.x, g.num_tiles}, eb, epi_base);
  } else if (warp >= kEpiWarp0) {
    epilogue_tma<BLOCK_N, ACC, false, MODE, Sched1, false, (MODE >= 2 ? 8 : 4)>(g, Sched1{(int)blockIdx.x, (int)gridDim.x, g.num_tiles}, tmem_base, tfull_bar(0), tempty_bar(0), eb, epi_base,
                                 warp - kEpiWarp0, lane);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, ACC * BLOCK_N);
  }
}


// ------------------------------------------------------------------------------------------------
// Stem + max-pool in one kernel (ORT Conv + Relu + MaxPool of the network's first three nodes).  Run separately the stem writes
// 531 MB per 8 x 1080p frames (it is HBM-write-bound) only for the pool to read them back; fused, the convolution's output never
// leaves the SM.  A CTA walks DOWN a column strip: per convolution row the same 14 Toeplitz MMAs as stem_tc_kernel (one TMA box of
// 7 input rows, weights resident in smem) fill a TMEM stage; the epilogue warps turn the row into fp16 (+ bias, ReLU), exchange
// it through shared memory to take the horizontal 3-max, and combine rows vertically in registers -- pooled row p is
// max(H[2p-1], H[2p], H[2p+1]) and H[2p+1] is carried into row p+1, so every convolution row is computed once.
// Strip geometry: TMA boxes start on 16-input-pixel (= 8 convolution pixel) boundaries, so a strip's 128 convolution columns start at
// c0 = 120 t - 8 and yield the 60 pooled columns q = 60 t + u, u < 60, from local columns 2u + 7 .. 2u + 9.  Columns / rows outside the
// convolution's output count as 0: the pool pads with -inf, and every real value is >= 0 after the ReLU, so 0 never wins wrongly.
// fp16 values are maxed exactly as maxpool3s2_kernel does: same bits as the two separate kernels.
constexpr int kSpStrip = 60;
constexpr int kSpXchBytes = kBlockM * 128;          // one convolution row: 128 pixels x 64 channels fp16
constexpr int kSpSmemBytes = kStemStages * kStemStageBytes + 2 * kSpXchBytes + kStemWBytes + 1024 + kBarBytes;

template <int MODE>   // 0: fp16 model; 2: the stem of an int8 plan (fp16-carried integer operands, requantised u8 output, pooled as bytes)
__global__ void __launch_bounds__(kThreads, 1)
stem_pool_kernel(const __grid_constant__ ConvTcMaps maps, const __grid_constant__ ConvTcGeom g) {
  constexpr int BLOCK_N = 64;
  constexpr int ACC = 4;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t xch_base = smem_base + kStemStages * kStemStageBytes;
  const uint32_t w_base = xch_base + 2 * kSpXchBytes;
  const uint32_t bar_base = w_base + kStemWBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + kMaxAcc + s); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * kMaxStages + 2 * kMaxAcc);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && ptx::elect_one()) ptx::prefetch_tmap(&maps.a[0]);
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < kStemStages; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < ACC; ++s) { ptx::mbar_init(tfull_bar(s), 1); ptx::mbar_init(tempty_bar(s), kEpiWarps); }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr_addr, ACC * BLOCK_N);
    ptx::tmem_relinquish();
  }
  {  // weights: global (already in smem order) -> smem, once per CTA
    const uint4* src = reinterpret_cast<const uint4*>(g.stem_w);
    for (int i = threadIdx.x; i < kStemWBytes / 16; i += kThreads) {
      const uint4 v = __ldg(src + i);
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(w_base + 16u * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  ptx::grid_dep_launch();
  ptx::grid_dep_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  // unit -> (image, chunk of pooled rows, strip of pooled columns); strips fastest: neighbours share input rows in L2
  auto unit_geom = [&](int unit, int& img, int& pr0, int& pr1, int& strip) {
    strip = unit % g.sp_strips;
    const int rest = unit / g.sp_strips;
    const int chunk = rest % g.sp_chunks;
    img = rest / g.sp_chunks;
    pr0 = chunk * g.sp_rc;
    pr1 = min(pr0 + g.sp_rc, g.sp_oh);
  };

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < g.num_tiles; unit += gridDim.x) {
        int img, pr0, pr1, strip;
        unit_geom(unit, img, pr0, pr1, strip);
        for (int r = 2 * pr0 - 1; r <= 2 * pr1 - 1; ++r) {          // convolution rows of this unit, top to bottom
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          ptx::mbar_expect_tx(full_bar(stage), (uint32_t)kStemTxBytes);
          // column c0 = 120 strip - 8 starts at input pixel 2 c0 = 16-pixel group 15 strip - 1; row r reads input rows 2r .. 2r + 6
          // of the padded buffer; groups / rows outside it are zero-filled by TMA and their outputs masked in the epilogue
          ptx::tma_load_4d(smem_base + stage * kStemStageBytes, &maps.a[0], full_bar(stage), 0, 15 * strip - 1, 2 * r, img);
          if (++stage == kStemStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(kBlockM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int unit = blockIdx.x; unit < g.num_tiles; unit += gridDim.x) {
        int img, pr0, pr1, strip;
        unit_geom(unit, img, pr0, pr1, strip);
        for (int r = 2 * pr0 - 1; r <= 2 * pr1 - 1; ++r, ++it) {
          const int as = it % ACC;
          ptx::mbar_wait(tempty_bar(as), (((uint32_t)(it / ACC)) & 1u) ^ 1u);
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
          const uint32_t a_addr = smem_base + stage * kStemStageBytes;
#pragma unroll
          for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t a_desc = ptx::make_smem_desc_noswz(a_addr + ky * kStemRowBytes + kk * 32, 16, 128);
              const uint64_t b_desc = ptx::make_smem_desc_noswz(w_base + ky * 4096 + kk * 2048, 1024, 128);
              ptx::umma_f16(d_tmem, a_desc, b_desc, idesc, (ky | kk) != 0);
            }
          }
          ptx::umma_commit(empty_bar(stage));
          ptx::umma_commit(tfull_bar(as));
          if (++stage == kStemStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    const int ew = warp - kEpiWarp0;
    const int quad = ew & 3, half = ew >> 2;
    const int m = quad * 32 + lane;                       // convolution column of this thread inside the strip (TMEM lane)
    const uint32_t msw = (uint32_t)(m & 7);
    const int tid2 = ew * 32 + lane;                      // reader role: pooled column u, 16-channel group cg
    const int u = tid2 >> 2, cg = tid2 & 3;
    float4 bias[8];
    {
      const float4* b4 = reinterpret_cast<const float4*>(g.bias + half * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) bias[j] = __ldg(b4 + j);
    }
    float4 qm[8];
    if (MODE == 2) {
      const float4* m4 = reinterpret_cast<const float4*>(g.qmul + half * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) qm[j] = __ldg(m4 + j);
    }
    const __half2 z2 = __float2half2_rn(0.f);
    int it = 0, n_it = 0;                                 // convolution rows of this CTA: done / in all
    for (int unit = blockIdx.x; unit < g.num_tiles; unit += gridDim.x) {
      int img, pr0, pr1, strip;
      unit_geom(unit, img, pr0, pr1, strip);
      n_it += 2 * (pr1 - pr0) + 1;
    }
    // The accumulator row of iteration it + 1 is requested from TMEM as soon as row it has been converted (v is dead then), so the
    // load's latency hides behind the exchange barrier and the pooling reads instead of heading every row.
    uint32_t v[32];
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 32);
    if (n_it > 0) {
      ptx::mbar_wait(tfull_bar(0), 0u);
      ptx::tc_fence_after();
      ptx::tmem_ld_32x32b_x32(t_lane, v);
    }
    for (int unit = blockIdx.x; unit < g.num_tiles; unit += gridDim.x) {
      int img, pr0, pr1, strip;
      unit_geom(unit, img, pr0, pr1, strip);
      const int c = 120 * strip - 8 + m;                  // global convolution column
      const bool col_ok = c >= 0 && c < g.ow;
      const int q = kSpStrip * strip + u;                 // global pooled column of the reader role
      const bool q_ok = u < kSpStrip && q < g.sp_ow;
      __half2 acc_out[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc_out[j] = z2;
      uint4 acc8 = make_uint4(0u, 0u, 0u, 0u);
      for (int r = 2 * pr0 - 1; r <= 2 * pr1 - 1; ++r, ++it) {
        const int as = it % ACC;
        ptx::tmem_ld_wait(v);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty_bar(as));
        // ---- this thread's 32 channels of convolution pixel (r, c): + bias, fp16 + ReLU (or requantised u8); 0 outside the
        // convolution's output
        const bool ok = col_ok && r >= 0 && r < g.oh;
        if (MODE == 0) {
          const uint32_t xb = xch_base + (uint32_t)(it & 1) * kSpXchBytes + (uint32_t)m * 128u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 bl = bias[2 * j], bh = bias[2 * j + 1];
            float x[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) x[t] = __uint_as_float(v[8 * j + t]);
            ptx::add_f32x2(x[0], x[1], bl.x, bl.y); ptx::add_f32x2(x[2], x[3], bl.z, bl.w);
            ptx::add_f32x2(x[4], x[5], bh.x, bh.y); ptx::add_f32x2(x[6], x[7], bh.z, bh.w);
            uint4 o;
            __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int t = 0; t < 4; ++t) h[t] = ok ? __hmax2(__floats2half2_rn(x[2 * t], x[2 * t + 1]), z2) : z2;
            const uint32_t addr = xb + (((uint32_t)(half * 4 + j) ^ msw) << 4);      // 16-byte groups XOR-swizzled by the row: conflict-free both ways
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
          }
        } else {
          // QLinearConv requantisation exactly as epilogue_tma's u8 path: r = clamp(rne((acc + b) * m), lo, hi), byte = r + zero point
          // 64-byte rows: odd row pairs are stored swapped (row m at slot m ^ ((m >> 1) & 1)) so that rows two apart -- what the lanes of
          // a quarter-warp read in the pooling pass -- sit in different halves of the bank space
          const uint32_t xb = xch_base + (uint32_t)(it & 1) * kSpXchBytes + (uint32_t)(m ^ ((m >> 1) & 1)) * 64u;
          const float lo_out = g.relu ? fmaxf(g.q_lo, 0.f) : g.q_lo, hi1 = g.q_hi;
          const uint32_t zout = (uint32_t)(int)(g.q_zmagic - kRneMagic);
#pragma unroll
          for (int j = 0; j < 2; ++j) {            // 16 channels -> one 16-byte group
            uint32_t ow[4];
#pragma unroll
            for (int wq = 0; wq < 4; ++wq) {
              const int c4 = 4 * j + wq;           // float4 index of these 4 channels
              const float4 bf = bias[c4], mq = qm[c4];
              float t0 = __uint_as_float(v[4 * c4 + 0]), t1 = __uint_as_float(v[4 * c4 + 1]), t2 = __uint_as_float(v[4 * c4 + 2]), t3 = __uint_as_float(v[4 * c4 + 3]);
              ptx::add_f32x2(t0, t1, bf.x, bf.y); ptx::add_f32x2(t2, t3, bf.z, bf.w);
              ptx::mul_f32x2(t0, t1, mq.x, mq.y); ptx::mul_f32x2(t2, t3, mq.z, mq.w);
              t0 = fminf(fmaxf(t0, lo_out), hi1); t1 = fminf(fmaxf(t1, lo_out), hi1); t2 = fminf(fmaxf(t2, lo_out), hi1); t3 = fminf(fmaxf(t3, lo_out), hi1);
              ptx::add_f32x2(t0, t1, kRneMagic, kRneMagic); ptx::add_f32x2(t2, t3, kRneMagic, kRneMagic);
              const uint32_t b0 = __float_as_uint(t0) + zout, b1 = __float_as_uint(t1) + zout, b2 = __float_as_uint(t2) + zout, b3 = __float_as_uint(t3) + zout;
              ow[wq] = ok ? __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410) : 0u;
            }
            const uint32_t addr = xb + (((uint32_t)(half * 2 + j) ^ (uint32_t)((m >> 1) & 3)) << 4);   // 64-byte rows: swizzle by the 128-byte line
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(ow[0]), "r"(ow[1]), "r"(ow[2]), "r"(ow[3]) : "memory");
          }
        }
        if (it + 1 < n_it) {                               // v has been consumed: start the next row's load
          const int as1 = (it + 1) % ACC;
          ptx::mbar_wait(tfull_bar(as1), ((uint32_t)((it + 1) / ACC)) & 1u);
          ptx::tc_fence_after();
          ptx::tmem_ld_32x32b_x32(t_lane + (uint32_t)(as1 * BLOCK_N), v);
        }
        ptx::named_bar_sync(1, kEpiWarps * 32);            // the whole row is in the exchange buffer (double-buffered: one barrier per row)
        // ---- reader role: horizontal 3-max for pooled column u, channels [16 cg, 16 cg + 16)
        if (MODE == 0) {
          __half2 hmx[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) hmx[j] = z2;
          if (q_ok) {
            const uint32_t rb = xch_base + (uint32_t)(it & 1) * kSpXchBytes;
#pragma unroll
            for (int dm = 0; dm < 3; ++dm) {
              const int mm = 2 * u + 7 + dm;
              const uint32_t rowp = rb + (uint32_t)mm * 128u;
              const uint32_t sw = (uint32_t)(mm & 7);
#pragma unroll
              for (int gq = 0; gq < 2; ++gq) {
                uint4 w4;
                // which of its two 16-byte groups a thread reads first alternates with the pooled column: the rows of neighbouring
                // columns are two apart (same swizzle parity), so reading the same group in both would put the eight lanes of a
                // quarter-warp on four bank groups -- a 2-way conflict on every load of a kernel that is shared-memory-bound
                // (ncu: LSU + tensor-core wavefronts = 96 % of the pipe, 9.9 M of 16.2 M load wavefronts were conflicts)
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w4.x), "=r"(w4.y), "=r"(w4.z), "=r"(w4.w) : "r"(rowp + (((uint32_t)(2 * cg + (gq ^ (u & 1))) ^ sw) << 4)));
                const __half2* hv = reinterpret_cast<const __half2*>(&w4);
#pragma unroll
                for (int t = 0; t < 4; ++t) hmx[4 * gq + t] = __hmax2(hmx[4 * gq + t], hv[t]);
              }
            }
          }
          // ---- vertical: pooled row p = max(H[2p - 1], H[2p], H[2p + 1]); H[2p + 1] is carried into row p + 1
#pragma unroll
          for (int j = 0; j < 8; ++j) acc_out[j] = __hmax2(acc_out[j], hmx[j]);
          if (r & 1) {
            const int p = (r - 1) >> 1;
            if (r > 2 * pr0 - 1 && q_ok && p < g.sp_oh) {
              uint4* dst = reinterpret_cast<uint4*>(g.out + ((((size_t)img * g.sp_oh + p) * g.sp_ow + q) * 64 + cg * 16));
              uint4 o0, o1;
              __half2* h0 = reinterpret_cast<__half2*>(&o0);
              __half2* h1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
              for (int t = 0; t < 4; ++t) { h0[t] = acc_out[t]; h1[t] = acc_out[4 + t]; }
              dst[u & 1] = o0; dst[(u & 1) ^ 1] = o1;      // acc_out[0..3] holds the group that was read first (see the loads above)
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc_out[j] = hmx[j];
          }
        } else {
          uint4 hm8 = make_uint4(0u, 0u, 0u, 0u);              // 0 is the identity of max on u8
          if (q_ok) {
            const uint32_t rb = xch_base + (uint32_t)(it & 1) * kSpXchBytes;
#pragma unroll
            for (int dm = 0; dm < 3; ++dm) {
              const int mm = 2 * u + 7 + dm;
              uint4 w4;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w4.x), "=r"(w4.y), "=r"(w4.z), "=r"(w4.w)
                           : "r"(rb + (uint32_t)(mm ^ ((mm >> 1) & 1)) * 64u + (((uint32_t)cg ^ (uint32_t)((mm >> 1) & 3)) << 4)));
              hm8 = make_uint4(__vmaxu4(hm8.x, w4.x), __vmaxu4(hm8.y, w4.y), __vmaxu4(hm8.z, w4.z), __vmaxu4(hm8.w, w4.w));
            }
          }
          acc8 = make_uint4(__vmaxu4(acc8.x, hm8.x), __vmaxu4(acc8.y, hm8.y), __vmaxu4(acc8.z, hm8.z), __vmaxu4(acc8.w, hm8.w));
          if (r & 1) {
            const int p = (r - 1) >> 1;
            if (r > 2 * pr0 - 1 && q_ok && p < g.sp_oh)
              *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(g.out) + ((((size_t)img * g.sp_oh + p) * g.sp_ow + q) * 64 + cg * 16)) = acc8;
            acc8 = hm8;
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, ACC * BLOCK_N);
  }
}

// Launch with programmatic stream serialization: the kernel may begin (prologue only, see grid_dep_wait) while the
// previous kernel of the stream drains.  INFUR_B200_NO_PDL=1 turns it off.
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("INFUR_B200_NO_PDL"); return !(e && e[0] == '1'); }();
  return on;
}
template <class Kernel>
cudaError_t launch_conv(Kernel kernel, int grid, int smem, cudaStream_t stream, const ConvTcMaps& maps, const ConvTcGeom& g) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, maps, g);
}

template <int BLOCK_N>
cudaError_t launch_one(const ConvTcMaps& maps, const ConvTcGeom& g, int num_sms, cudaStream_t stream) {
  const int grid = g.num_tiles < num_sms ? g.num_tiles : num_sms;
  if (grid <= 0) return cudaSuccess;
  if (g.store_mode != 0 && BLOCK_N < 64) return cudaErrorInvalidValue;
  if (g.mode == 3) {
    if (g.stages != Cfg<BLOCK_N, true>::stages(g.epi_bufs)) return cudaErrorInvalidValue;
    return launch_conv(conv_tc_kernel<BLOCK_N, 3>, grid, Cfg<BLOCK_N, true>::smem_bytes(g.epi_bufs), stream, maps, g);
  }
  if (g.mode == 2 || g.stages != Cfg<BLOCK_N>::stages(g.epi_bufs)) return cudaErrorInvalidValue;   // mode 2 is the stem's
  if (g.mode == 1) return launch_conv(conv_tc_kernel<BLOCK_N, 1>, grid, Cfg<BLOCK_N>::smem_bytes(g.epi_bufs), stream, maps, g);
  return launch_conv(conv_tc_kernel<BLOCK_N, 0>, grid, Cfg<BLOCK_N>::smem_bytes(g.epi_bufs), stream, maps, g);
}

}  // namespace

int conv_tc_pair_stages(int epi_bufs, bool i8) { return pair_stages(epi_bufs, i8); }

int conv_tc_stages(int block_n, int epi_bufs, bool i8) {
  switch (block_n) {
    case 32: return i8 ? Cfg<32, true>::stages(epi_bufs) : Cfg<32>::stages(epi_bufs);
    case 64: return i8 ? Cfg<64, true>::stages(epi_bufs) : Cfg<64>::stages(epi_bufs);
    case 128: return i8 ? Cfg<128, true>::stages(epi_bufs) : Cfg<128>::stages(epi_bufs);
    default: return i8 ? Cfg<256, true>::stages(epi_bufs) : Cfg<256>::stages(epi_bufs);
  }
}

cudaError_t conv_tc_init() {
  cudaError_t e = cudaSuccess;
  auto opt_in = [&](auto kernel, int bytes) { if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); };
  opt_in(conv_tc_kernel<32, 0>, kSmemLimit); opt_in(conv_tc_kernel<32, 1>, kSmemLimit); opt_in(conv_tc_kernel<32, 3>, kSmemLimit);
  opt_in(conv_tc_kernel<64, 0>, kSmemLimit); opt_in(conv_tc_kernel<64, 1>, kSmemLimit); opt_in(conv_tc_kernel<64, 3>, kSmemLimit);
  opt_in(conv_tc_kernel<128, 0>, kSmemLimit); opt_in(conv_tc_kernel<128, 1>, kSmemLimit); opt_in(conv_tc_kernel<128, 3>, kSmemLimit);
  opt_in(conv_tc_kernel<256, 0>, kSmemLimit); opt_in(conv_tc_kernel<256, 1>, kSmemLimit); opt_in(conv_tc_kernel<256, 3>, kSmemLimit);
  opt_in(stem_tc_kernel<0>, kStemSmemBytes); opt_in(stem_tc_kernel<1>, kStemSmemBytes); opt_in(stem_tc_kernel<2>, kStemSmemBytes);
  opt_in(conv_halo_kernel<64, 0>, kSmemLimit); opt_in(conv_halo_kernel<64, 1>, kSmemLimit);
  opt_in(conv_halo_kernel<128, 0>, kSmemLimit); opt_in(conv_halo_kernel<128, 1>, kSmemLimit);
  opt_in(conv_halo_kernel<256, 0>, kSmemLimit); opt_in(conv_halo_kernel<256, 1>, kSmemLimit);
  opt_in(conv_tc_pair_kernel<0>, kSmemLimit); opt_in(conv_tc_pair_kernel<1>, kSmemLimit); opt_in(conv_tc_pair_kernel<3>, kSmemLimit);
  opt_in(conv_b2b_kernel<64>, B2BCfg<64>::kSmem); opt_in(conv_b2b_kernel<128>, B2BCfg<128>::kSmem);
  opt_in(stem_pool_kernel<0>, kSpSmemBytes); opt_in(stem_pool_kernel<2>, kSpSmemBytes);
  return e;
}

bool conv_b2b_supported(int cmid) { return cmid == 64 || cmid == 128; }

cudaError_t conv_tc_launch(int block_n, const ConvTcMaps& maps, const ConvTcGeom& g, int num_sms, cudaStream_t stream) {
  if (g.b2b) {
    const int m_tiles = g.tiles_n > 0 ? g.num_tiles / g.tiles_n : 0;
    const int grid = m_tiles < num_sms ? m_tiles : num_sms;
    if (grid <= 0) return cudaSuccess;
    if (g.store_mode != 2 || g.mode != 0 || g.main_taps != g.num_taps || g.num_kb != g.main_taps * g.cchunks) return cudaErrorInvalidValue;
    if (g.b2b_cmid == 64 && g.cchunks == 1) return launch_conv(conv_b2b_kernel<64>, grid, B2BCfg<64>::kSmem, stream, maps, g);
    if (g.b2b_cmid == 128 && g.cchunks == 2) return launch_conv(conv_b2b_kernel<128>, grid, B2BCfg<128>::kSmem, stream, maps, g);
    return cudaErrorInvalidValue;
  }
  if (g.stem && g.sp_fused) {
    const int grid = g.num_tiles < num_sms ? g.num_tiles : num_sms;
    if (grid <= 0) return cudaSuccess;
    if (block_n != 64 || (g.mode != 0 && g.mode != 2) || g.sp_rc < 1 || g.sp_strips < 1 || g.sp_chunks < 1) return cudaErrorInvalidValue;
    if (g.mode == 2) return launch_conv(stem_pool_kernel<2>, grid, kSpSmemBytes, stream, maps, g);
    return launch_conv(stem_pool_kernel<0>, grid, kSpSmemBytes, stream, maps, g);
  }
  if (g.stem) {
    const int grid = g.num_tiles < num_sms ? g.num_tiles : num_sms;
    if (grid <= 0) return cudaSuccess;
    if (block_n != 64 || g.store_mode != 1 || g.bw_log2 != 7) return cudaErrorInvalidValue;
    if (g.mode == 2) return launch_conv(stem_tc_kernel<2>, grid, kStemSmemBytes, stream, maps, g);
    if (g.mode == 1) return launch_conv(stem_tc_kernel<1>, grid, kStemSmemBytes, stream, maps, g);
    if (g.mode != 0) return cudaErrorInvalidValue;
    return launch_conv(stem_tc_kernel<0>, grid, kStemSmemBytes, stream, maps, g);
  }
  if (g.halo) {
    const int grid = g.num_tiles < num_sms ? g.num_tiles : num_sms;
    if (grid <= 0) return cudaSuccess;
    if (g.store_mode != 1 || g.bw_log2 != 3 || g.main_taps != 9 || g.num_taps != 9 || g.mode > 1) return cudaErrorInvalidValue;
    switch (block_n) {
      case 64:
        return g.mode ? launch_conv(conv_halo_kernel<64, 1>, grid, halo_plan(g.halo_dil, 64).smem_bytes, stream, maps, g)
                       : launch_conv(conv_halo_kernel<64, 0>, grid, halo_plan(g.halo_dil, 64).smem_bytes, stream, maps, g);
      case 128:
        return g.mode ? launch_conv(conv_halo_kernel<128, 1>, grid, halo_plan(g.halo_dil, 128).smem_bytes, stream, maps, g)
                       : launch_conv(conv_halo_kernel<128, 0>, grid, halo_plan(g.halo_dil, 128).smem_bytes, stream, maps, g);
      case 256:
        return g.mode ? launch_conv(conv_halo_kernel<256, 1>, grid, halo_plan(g.halo_dil, 256).smem_bytes, stream, maps, g)
                       : launch_conv(conv_halo_kernel<256, 0>, grid, halo_plan(g.halo_dil, 256).smem_bytes, stream, maps, g);
      default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
  }
  if (g.pair) {
    const int clusters = g.num_work < num_sms / 2 ? g.num_work : num_sms / 2;
    if (clusters <= 0) return cudaSuccess;
    if (block_n != 256 || g.store_mode == 0 || g.stages != pair_stages(g.epi_bufs, g.mode == 3) || g.mode == 2) return cudaErrorInvalidValue;
    if (g.mode == 3) return launch_conv(conv_tc_pair_kernel<3>, 2 * clusters, pair_smem_bytes(g.epi_bufs, true), stream, maps, g);
    if (g.mode == 1) return launch_conv(conv_tc_pair_kernel<1>, 2 * clusters, pair_smem_bytes(g.epi_bufs), stream, maps, g);
    return launch_conv(conv_tc_pair_kernel<0>, 2 * clusters, pair_smem_bytes(g.epi_bufs), stream, maps, g);
  }
  switch (block_n) {
    case 32: return launch_one<32>(maps, g, num_sms, stream);
    case 64: return launch_one<64>(maps, g, num_sms, stream);
    case 128: return launch_one<128>(maps, g, num_sms, stream);
    case 256: return launch_one<256>(maps, g, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace infur

// How fast can the arithmetic of the int8 +residual epilogue (requant_u8_pass<3, true, 2> in csrc/conv_tc.cu) issue, with nothing
// else in the way (no TMEM, no shared memory, no barriers)?  One CTA per SM with W warps (8 = the kernels' epilogue, 16, 32);
// every warp runs PASSES passes of 32 channels on register-resident data.  Prints cycles per pass per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I infur_b200/csrc -o build/epi_tput tools/epi_tput.cu && build/epi_tput
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace infur;

constexpr float kRneMagic = 12582912.f;

template <bool LDG>
__device__ __forceinline__ void pass(const uint32_t (&acc)[32], const uint32_t (&rw)[8], uint32_t (&ow)[8], const float* qmul, const int* bias_i32, int cofs,
                                     float lo_out, float hi1, float res_bias, float ra, float rb, uint32_t zadj) {
  constexpr int kMagicBits = 0x4B400000;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 m;
    int4 bi;
    if (LDG) {
      m = __ldg(reinterpret_cast<const float4*>(qmul + cofs) + j);
      bi = __ldg(reinterpret_cast<const int4*>(bias_i32 + cofs) + j);
    } else {
      m = make_float4(0.5f + lo_out, 0.25f, 0.125f, 0.75f);
      bi = make_int4(cofs, cofs + 1, cofs + 2, cofs + 3);
    }
    float t[4];
    t[0] = __int_as_float((int)acc[4 * j + 0] + bi.x + kMagicBits); t[1] = __int_as_float((int)acc[4 * j + 1] + bi.y + kMagicBits);
    t[2] = __int_as_float((int)acc[4 * j + 2] + bi.z + kMagicBits); t[3] = __int_as_float((int)acc[4 * j + 3] + bi.w + kMagicBits);
    ptx::add_f32x2(t[0], t[1], -kRneMagic, -kRneMagic); ptx::add_f32x2(t[2], t[3], -kRneMagic, -kRneMagic);
    ptx::mul_f32x2_sep(t[0], t[1], m.x, m.y); ptx::mul_f32x2_sep(t[2], t[3], m.z, m.w);
    uint32_t bits[4];
#pragma unroll
    for (int x = 0; x < 4; x += 2) {
      float a0 = fminf(fmaxf(t[x], lo_out), hi1), a1 = fminf(fmaxf(t[x + 1], lo_out), hi1);
      ptx::add_f32x2(a0, a1, kRneMagic, kRneMagic);
      ptx::add_f32x2(a0, a1, -kRneMagic, -kRneMagic);
      float b0 = __uint_as_float(__byte_perm(rw[j], 0x4B000000u, 0x7650 + x));
      float b1 = __uint_as_float(__byte_perm(rw[j], 0x4B000000u, 0x7650 + x + 1));
      ptx::add_f32x2(b0, b1, res_bias, res_bias);
      ptx::mul_f32x2_sep(a0, a1, ra, ra);
      ptx::mul_f32x2_sep(b0, b1, rb, rb);
      ptx::add_f32x2(a0, a1, b0, b1);
      ptx::add_f32x2(a0, a1, kRneMagic, kRneMagic);
      bits[x] = __float_as_uint(a0) + zadj; bits[x + 1] = __float_as_uint(a1) + zadj;
    }
    ow[j] = ptx::pack_sat_u8x4((int)bits[0], (int)bits[1], (int)bits[2], (int)bits[3]);
  }
}

// TMEM variants (8 or 16 warps: warp w reads lane quadrant w % 4, 32 columns at a time, of 512 allocated columns -- their content is
// whatever the last kernel left, which is all the arithmetic needs): T = 1 loads only, 2 load -> wait -> arithmetic, 3 the load of
// pass p + 1 in flight during the arithmetic of pass p
template <int T, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) kt(unsigned* out, const float* qmul, const int* bias, int passes, float lo, float hi, float ra, float rb, unsigned seed,
                                                    long long* cycles) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tmem_ptr), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t accA[32], accB[32], rw[8], ow[8], x = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { rw[i] = seed ^ (threadIdx.x * 2654435761u + i); ow[i] = 0; }
  const int col0 = (warp >> 2) * 32;
  const long long t0 = clock64();
  if (T == 3) ptx::tmem_ld_32x32b_x32(base + (uint32_t)col0, accA);
  for (int p = 0; p < passes; p += 2) {
    if (T == 1) {
      ptx::tmem_ld_32x32b_x32(base + (uint32_t)((col0 + p * 32) & 511 & ~31), accA);
      ptx::tmem_ld_32x32b_x32(base + (uint32_t)((col0 + p * 32 + 32) & 511 & ~31), accB);
      ptx::tmem_ld_wait(accA); ptx::tmem_ld_wait(accB);
      x ^= accA[0] ^ accA[31] ^ accB[5] ^ accB[17];
    } else if (T == 2) {
      ptx::tmem_ld_32x32b_x32(base + (uint32_t)((col0 + p * 32) & 511 & ~31), accA);
      ptx::tmem_ld_wait(accA);
      pass<true>(accA, rw, ow, qmul, bias, (p & 7) * 32, lo, hi, -(8388608.f + 37.f), ra, rb, 3u - 0x4B400000u);
#pragma unroll
      for (int i = 0; i < 8; ++i) { x ^= ow[i]; rw[i] += ow[i]; }
      ptx::tmem_ld_32x32b_x32(base + (uint32_t)((col0 + p * 32 + 32) & 511 & ~31), accB);
      ptx::tmem_ld_wait(accB);
      pass<true>(accB, rw, ow, qmul, bias, ((p + 1) & 7) * 32, lo, hi, -(8388608.f + 37.f), ra, rb, 3u - 0x4B400000u);
#pragma unroll
      for (int i = 0; i < 8; ++i) { x ^= ow[i]; rw[i] += ow[i]; }
    } else {
      ptx::tmem_ld_wait(accA);
      ptx::tmem_ld_32x32b_x32(base + (uint32_t)((col0 + p * 32 + 32) & 511 & ~31), accB);
      pass<true>(accA, rw, ow, qmul, bias, (p & 7) * 32, lo, hi, -(8388608.f + 37.f), ra, rb, 3u - 0x4B400000u);
#pragma unroll
      for (int i = 0; i < 8; ++i) { x ^= ow[i]; rw[i] += ow[i]; }
      ptx::tmem_ld_wait(accB);
      ptx::tmem_ld_32x32b_x32(base + (uint32_t)((col0 + p * 32 + 64) & 511 & ~31), accA);
      pass<true>(accB, rw, ow, qmul, bias, ((p + 1) & 7) * 32, lo, hi, -(8388608.f + 37.f), ra, rb, 3u - 0x4B400000u);
#pragma unroll
      for (int i = 0; i < 8; ++i) { x ^= ow[i]; rw[i] += ow[i]; }
    }
  }
  if (T == 3) ptx::tmem_ld_wait(accA);
  const long long t1 = clock64();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_ptr, 512); }
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (x == 0x12345678u) out[0] = x + accA[0];
}

// Chunk hand-over variants (8 warps): every pass = TMEM load + arithmetic + two 16-byte shared-memory stores; every second pass ends a
// chunk: F = 0 nothing more, 1 fence.proxy.async + __syncwarp + mbarrier arrive right after the stores (what the kernels do), 2 the same
// but one pass later (the stores have long drained when the fence executes)
template <int F>
__global__ void __launch_bounds__(256, 1) kf(unsigned* out, const float* qmul, const int* bias, int passes, float lo, float hi, float ra, float rb, unsigned seed,
                                            long long* cycles) {
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(16) uint8_t buf[2][128 * 128];
  __shared__ __align__(8) unsigned long long bar[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tmem_ptr), 512); ptx::tmem_relinquish(); }
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar[0]), 1u << 19); ptx::mbar_init(ptx::smem_u32(&bar[1]), 1u << 19); ptx::fence_mbar_init(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc[32], rw[8], ow[8], x = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { rw[i] = seed ^ (threadIdx.x * 2654435761u + i); ow[i] = 0; }
  const int half = warp >> 2, row = (warp & 3) * 32 + lane;
  const uint32_t sw = (uint32_t)(row & 7);
  const long long t0 = clock64();
  int pending = -1;
  for (int p = 0; p < passes; ++p) {
    const int b = (p >> 1) & 1, hh = p & 1;
    ptx::tmem_ld_32x32b_x32(base + (uint32_t)(((half * 64 + hh * 32) + (p & 3) * 128) & 511 & ~31), acc);
    ptx::tmem_ld_wait();
    pass<true>(acc, rw, ow, qmul, bias, (p & 7) * 32, lo, hi, -(8388608.f + 37.f), ra, rb, 3u - 0x4B400000u);
    const uint32_t rowb = ptx::smem_u32(&buf[b][0]) + (uint32_t)row * 128u;
    const uint32_t g0 = rowb + (((uint32_t)(half * 4 + hh * 2) ^ sw) << 4), g1 = rowb + (((uint32_t)(half * 4 + hh * 2 + 1) ^ sw) << 4);
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(g0), "r"(ow[0]), "r"(ow[1]), "r"(ow[2]), "r"(ow[3]) : "memory");
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(g1), "r"(ow[4]), "r"(ow[5]), "r"(ow[6]), "r"(ow[7]) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) { x ^= ow[i]; rw[i] += ow[i]; }
    if (F == 2 && pending >= 0 && hh == 0) {     // hand over the PREVIOUS chunk now
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar[pending]));
      pending = -1;
    }
    if (hh == 1) {
      if (F == 1) {
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar[b]));
      } else if (F == 2) pending = b;
    }
  }
  const long long t1 = clock64();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_ptr, 512); }
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (x == 0x12345678u) out[0] = x + buf[0][threadIdx.x];
}

template <int F>
void runf(int sms, unsigned* d, const float* qmul, const int* bias, long long* dcyc) {
  const int passes = 2048;
  for (int rep = 0; rep < 2; ++rep) kf<F><<<sms, 256>>>(d, qmul, bias, passes, -128.f, 127.f, 1.37f, 0.55f, 12345u, dcyc);
  cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
  const char* names[3] = {"stores only", "fence + arrive right after the stores", "fence + arrive one pass later"};
  printf("chunk hand-over, 8 warps: %9lld cycles, %6.1f scheduler-cycles per warp-pass  [%s; %s]\n", cyc, (double)cyc / passes / 2.0, names[F],
         cudaGetErrorString(cudaGetLastError()));
}

template <int T, int WARPS>
void runt(int sms, unsigned* d, const float* qmul, const int* bias, long long* dcyc) {
  const int passes = 2048, warps = WARPS;
  for (int rep = 0; rep < 2; ++rep) kt<T, WARPS><<<sms, warps * 32>>>(d, qmul, bias, passes, -128.f, 127.f, 1.37f, 0.55f, 12345u, dcyc);
  cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
  const char* names[4] = {"", "tmem loads only          ", "load -> wait -> arithmetic", "pipelined load + arithmetic"};
  printf("%s warps/CTA %2d: %9lld cycles, %6.1f scheduler-cycles per warp-pass; %6.2f outputs (= 4-byte TMEM words)/clk/SM = %6.1f B/clk/SM of TMEM reads  [%s]\n",
         names[T], warps, cyc, (double)cyc / passes / (warps / 4.0), 32.0 * 32.0 * warps * passes / (double)cyc, 4.0 * 32.0 * 32.0 * warps * passes / (double)cyc,
         cudaGetErrorString(cudaGetLastError()));
}

template <bool LDG, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k(unsigned* out, const float* qmul, const int* bias, int passes, float lo, float hi, float ra, float rb, unsigned seed,
                                             long long* cycles) {
  uint32_t acc[32], rw[8], ow[8], x = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = seed * (threadIdx.x + 1) + i * 977u;
#pragma unroll
  for (int i = 0; i < 8; ++i) rw[i] = seed ^ (threadIdx.x * 2654435761u + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int p = 0; p < passes; ++p) {
    pass<LDG>(acc, rw, ow, qmul, bias, (p & 7) * 32, lo, hi, -(8388608.f + 37.f), ra, rb, 3u - 0x4B400000u);
#pragma unroll
    for (int i = 0; i < 8; ++i) { x ^= ow[i]; rw[i] += ow[i]; }
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] += ow[i & 7] & 0xff;
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (x == 0x12345678u) out[0] = x;
}

template <bool LDG, int WARPS>
void run(int sms, unsigned* d, const float* qmul, const int* bias, long long* dcyc) {
  const int passes = 2048, warps = WARPS;
  k<LDG, WARPS><<<sms, warps * 32>>>(d, qmul, bias, passes, -128.f, 127.f, 1.37f, 0.55f, 12345u, dcyc);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<LDG, WARPS><<<sms, warps * 32>>>(d, qmul, bias, passes, -128.f, 127.f, 1.37f, 0.55f, 12345u, dcyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long cyc = 0;
  cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
  const double per_pass_smsp = (double)cyc / passes / (warps / 4.0);      // scheduler cycles per (warp, pass)
  printf("%s warps/CTA %2d: %8.3f ms, %9lld cycles, %6.1f scheduler-cycles per warp-pass of 32 channels = %5.2f per channel pair; %6.2f outputs/clk/SM\n",
         LDG ? "ldg " : "regs", warps, ms, cyc, per_pass_smsp, per_pass_smsp / 16.0, 32.0 * 32.0 * warps * passes / (double)cyc);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* d; float* qmul; int* bias; long long* dcyc;
  cudaMalloc(&d, 4); cudaMalloc(&qmul, 4096); cudaMalloc(&bias, 4096); cudaMalloc(&dcyc, 8);
  cudaMemset(qmul, 0, 4096); cudaMemset(bias, 0, 4096);
  run<false, 4>(sms, d, qmul, bias, dcyc); run<false, 8>(sms, d, qmul, bias, dcyc); run<false, 12>(sms, d, qmul, bias, dcyc);
  run<false, 16>(sms, d, qmul, bias, dcyc); run<false, 32>(sms, d, qmul, bias, dcyc);
  run<true, 4>(sms, d, qmul, bias, dcyc); run<true, 8>(sms, d, qmul, bias, dcyc); run<true, 12>(sms, d, qmul, bias, dcyc);
  run<true, 16>(sms, d, qmul, bias, dcyc); run<true, 32>(sms, d, qmul, bias, dcyc);
  runt<1, 4>(sms, d, qmul, bias, dcyc); runt<1, 8>(sms, d, qmul, bias, dcyc); runt<1, 16>(sms, d, qmul, bias, dcyc);
  runt<2, 8>(sms, d, qmul, bias, dcyc); runt<3, 8>(sms, d, qmul, bias, dcyc);
  runt<2, 16>(sms, d, qmul, bias, dcyc); runt<3, 16>(sms, d, qmul, bias, dcyc);
  runf<0>(sms, d, qmul, bias, dcyc); runf<1>(sms, d, qmul, bias, dcyc); runf<2>(sms, d, qmul, bias, dcyc);
  return 0;
}

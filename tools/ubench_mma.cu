// Micro-benchmark: issue rate of tcgen05.mma M128 x N x K-per-instruction for kind::f16 (K = 16) and kind::i8 (K = 32),
// operands in (zeroed) 128B-swizzled shared memory, accumulator in TMEM.  One CTA per SM, one issuing thread.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I infur_b200/csrc tools/ubench_mma.cu -o build/ubench_mma && build/ubench_mma
//
// Prints cycles per MMA and the implied dense rate per SM.  Planning input for a native int8 path.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include "ptx.cuh"

using namespace infur;

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::i8: D = S32 (c_format 2), A = UINT8 (0), B = INT8 (1), both K-major
__host__ __device__ constexpr uint32_t make_idesc_i8(uint32_t m, uint32_t n) {
  return (2u << 4) | (0u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

template <int N, bool I8>
__global__ void __launch_bounds__(128, 1) ubench(int iters, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_ptr;
  const uint32_t base = (ptx::smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + (base - ptx::smem_u32(smem)))[i] = 0u;
  const uint32_t bar_addr = ptx::smem_u32(&bar);
  if (threadIdx.x == 0) { ptx::mbar_init(bar_addr, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(ptx::smem_u32(&tmem_ptr), N); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t d = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint64_t a_desc = ptx::make_smem_desc(base, 128), b_desc = ptx::make_smem_desc(base + 16384, 128);
    const uint32_t idesc = I8 ? make_idesc_i8(128, N) : ptx::make_idesc_f16(128, N);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {   // four K slices of one 128-byte swizzle span (+32 B each)
        if (I8) umma_i8(d, a_desc + 2 * k, b_desc + 2 * k, idesc, (i | k) != 0);
        else ptx::umma_f16(d, a_desc + 2 * k, b_desc + 2 * k, idesc, (i | k) != 0);
      }
    }
    ptx::umma_commit(bar_addr);
    ptx::mbar_wait(bar_addr, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(d, N); }
}

template <int N, bool I8>
void run(const char* name, int sms) {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 8);
  const int smem = 1024 + 16384 + N * 128;
  cudaFuncSetAttribute(ubench<N, I8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  ubench<N, I8><<<sms, 128, smem>>>(64, d_out);   // warm-up
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  ubench<N, I8><<<sms, 128, smem>>>(iters, d_out);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  unsigned long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double per = (double)cyc / (4.0 * iters);
  const double k = I8 ? 32 : 16;
  const double ops = 2.0 * 128 * N * k * 4.0 * iters * sms;
  printf("%-28s %7.1f cycles per MMA (M128 x N%d x K%d) -> %6.0f MAC/clk/SM; all %d SMs: %7.1f T%s/s over %.3f ms\n", name, per, N, (int)k,
         128.0 * N * k / per, sms, ops / (ms * 1e-3) * 1e-12, I8 ? "OP" : "FLOP", ms);
  cudaFree(d_out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  const int sms = p.multiProcessorCount;
  run<256, false>("kind::f16 N256", sms);
  run<256, true>("kind::i8  N256", sms);
  run<128, false>("kind::f16 N128", sms);
  run<128, true>("kind::i8  N128", sms);
  run<64, false>("kind::f16 N64", sms);
  run<64, true>("kind::i8  N64", sms);
  return 0;
}

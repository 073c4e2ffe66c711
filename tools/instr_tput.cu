// Issue-rate microbenchmark for the instructions of the int8 requantisation epilogue (csrc/conv_tc.cu, MODE 3):
// warp-instructions per clock per SM for FMNMX, FADD2, FMUL2, I2FP, IADD3, PRMT, I2IP.SAT, VIADDMNMX.RELU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/instr_tput tools/instr_tput.cu && build/instr_tput
// Each kernel runs 8 independent dependency chains per thread so that latency never limits; 1 CTA of 1024 threads per SM.
#include <cstdio>
#include <cuda_runtime.h>

#define REP 64
#define ITERS 256

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(unsigned* out, unsigned seed, float f0, float f1) {
  unsigned r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = seed + threadIdx.x * 8 + i;
  unsigned long long p[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = ((unsigned long long)r[2 * i] << 32) | r[2 * i + 1];
  const unsigned long long fc = ((unsigned long long)__float_as_uint(f1) << 32) | __float_as_uint(f0);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int j = 0; j < REP / 8; ++j) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == 0) asm volatile("max.f32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 3) & 7]));
        if (OP == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i & 3]) : "l"(fc));
        if (OP == 2) asm volatile("mul.rn.ftz.f32x2 %0, %0, %1;" : "+l"(p[i & 3]) : "l"(fc));
        if (OP == 3) asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(r[i]));
        if (OP == 4) asm volatile("add.s32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 3) & 7]));
        if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, 0x7650;" : "+r"(r[i]) : "r"(seed));
        if (OP == 6) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(seed), "r"(r[(i + 1) & 7]));
        if (OP == 7) r[i] = (unsigned)__viaddmin_s32_relu((int)r[i], (int)seed, 255 + it);
        if (OP == 8) asm volatile("min.s32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 3) & 7]));
        if (OP == 9) asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(r[i]) : "r"(__float_as_uint(f0)));
      }
    }
  }
  unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= r[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc ^= (unsigned)p[i] ^ (unsigned)(p[i] >> 32);
  if (acc == 0x12345678u) out[0] = acc;
}

template <int OP>
void run(const char* name, int sms, unsigned* d) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<sms, 1024>>>(d, 1u, 1.0001f, 0.9999f);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) k<OP><<<sms, 1024>>>(d, 1u, 1.0001f, 0.9999f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double winstr = 10.0 * 32 /*warps*/ * (double)ITERS * REP;          // per SM
  const double per_us = winstr / (ms * 1e3);
  printf("%-18s %8.3f ms  %7.1f warp-instr/us/SM  = %5.2f per clk at %d MHz nominal (%5.2f at 1.9 GHz)\n", name, ms, per_us, per_us / (clk_khz / 1e3), clk_khz / 1000,
         per_us / 1900.0);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* d;
  cudaMalloc(&d, 4);
  run<0>("FMNMX", sms, d);
  run<9>("FADD", sms, d);
  run<1>("FADD2", sms, d);
  run<2>("FMUL2.FTZ", sms, d);
  run<3>("I2FP.F32.S32", sms, d);
  run<4>("IADD", sms, d);
  run<5>("PRMT", sms, d);
  run<6>("I2IP.U8.S32.SAT", sms, d);
  run<7>("VIADDMNMX.RELU", sms, d);
  run<8>("VIMNMX", sms, d);
  return 0;
}

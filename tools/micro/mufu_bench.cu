// Micro-benchmark: throughput of ex2.approx.f32 vs ex2.approx.f16x2 / bf16x2 (elements per clock per SM).
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

template <int MODE>
__global__ void k(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  uint32_t h0 = threadIdx.x, h1 = threadIdx.x + 77, h2 = threadIdx.x + 99, h3 = threadIdx.x + 1234;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
    } else {
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h3));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + (float)(h0 ^ h1 ^ h2 ^ h3);
}

template <int MODE>
void run(const char* name, int elems_per_instr) {
  float* d;
  cudaMalloc(&d, 148 * 8 * 1024 * 4);
  const int iters = 20000;
  k<MODE><<<148 * 2, 1024>>>(d, 10);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 1024>>>(d, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double instr = 148.0 * 2 * 1024 * 4.0 * iters;   // thread-level instructions
  printf("%s: %.3f ms  %.2f G thread-instr/s  %.2f G elements/s  (per SM per ns: %.2f elements)\n", name, ms,
         instr / ms / 1e6, instr * elems_per_instr / ms / 1e6, instr * elems_per_instr / ms / 1e6 / 148);
  cudaFree(d);
}

int main() {
  run<0>("ex2.f32", 1);
  run<1>("ex2.f16x2", 2);
  run<2>("ex2.bf16x2", 2);
  return 0;
}

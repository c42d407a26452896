// Micro-benchmark: issue cost and throughput of tcgen05.mma (kind::f16, bf16 -> fp32, cta_group::1, M = 128) for the
// instruction shapes the attention kernel uses.  One CTA per SM, one issuing thread; reports, per MMA, the cycles the
// issuing thread spends issuing and the cycles until the last MMA has completed (tcgen05.commit -> mbarrier).
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I dit4sr_b200/csrc -I include tools/micro/mma_bench.cu -o /tmp/mma_bench -lcuda
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include "ptx.cuh"

struct Res { long long issue, total; };

// mode 0: SS N=128 | 1: SS N=256 | 2: SS N=64 | 3: TS N=64 (B MN-major) | 4: TS N=128 (B MN-major)
// mode 5: attention pattern per slot: 8 x TS N=64 then 4 x SS N=128 | mode 6: same but one commit per group
// CONT: contention from 8 extra warps: 0 none | 1 softmax-like (TMEM ld + ex2 + TMEM st) | 2 math only | 3 TMEM ld/st only
template <int MODE, int CONT>
__global__ void __launch_bounds__(384, 1) k(Res* out, int reps, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int done;
  if (threadIdx.x == 0) done = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::mbar_init(&bar2, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_ptr, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (warp == 1) {
    const uint64_t a = ptx::make_sdesc_sw128(ptx::smem_u32(smem));               // 128 x 64 bf16 (16 KB)
    const uint64_t b = ptx::make_sdesc_sw128(ptx::smem_u32(smem + 16384));       // up to 256 x 64 bf16 (32 KB)
    const uint64_t v = ptx::make_sdesc_sw128(ptx::smem_u32(smem + 49152));       // 128(k) x 64(n) MN-major (16 KB)
    constexpr uint32_t i128 = ptx::make_idesc_bf16(128, 128, 0, 0), i256 = ptx::make_idesc_bf16(128, 256, 0, 0),
                       i64 = ptx::make_idesc_bf16(128, 64, 0, 0), t64 = ptx::make_idesc_bf16(128, 64, 0, 1),
                       t128 = ptx::make_idesc_bf16(128, 128, 0, 1);
    long long t0 = clock64(), t1 = 0;
    int count = 0;
    for (int r = 0; r < reps; ++r) {
      if (ptx::elect_one()) {
        if (MODE == 0) { for (int kk = 0; kk < 4; ++kk) ptx::mma_ss(tm, a + 2 * kk, b + 2 * kk, i128, 1); }
        if (MODE == 1) { for (int kk = 0; kk < 4; ++kk) ptx::mma_ss(tm, a + 2 * kk, b + 2 * kk, i256, 1); }
        if (MODE == 2) { for (int kk = 0; kk < 4; ++kk) ptx::mma_ss(tm, a + 2 * kk, b + 2 * kk, i64, 1); }
        if (MODE == 3) { for (int kk = 0; kk < 8; ++kk) ptx::mma_ts(tm + 384, tm + 64 + 8 * kk, v + 128 * kk, t64, 1); }
        if (MODE == 4) { for (int kk = 0; kk < 8; ++kk) ptx::mma_ts(tm + 256, tm + 8 * kk, v + 128 * kk, t128, 1); }
        if (MODE == 5 || MODE == 6) {
          for (int kk = 0; kk < 8; ++kk) ptx::mma_ts(tm + 384, tm + 64 + 8 * kk, v + 128 * kk, t64, 1);
          if (MODE == 6) ptx::tc_commit(&bar2);
          for (int kk = 0; kk < 4; ++kk) ptx::mma_ss(tm + 256, a + 2 * kk, b + 2 * kk, i128, 1);
          if (MODE == 6) ptx::tc_commit(&bar2);
        }
      }
      __syncwarp();
      count += (MODE == 3 || MODE == 4) ? 8 : (MODE >= 5 ? 12 : 4);
    }
    t1 = clock64();
    if (ptx::elect_one()) ptx::tc_commit(&bar);
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) {
      out->issue = (t1 - t0);
      out->total = (t2 - t0);
      out[1].issue = count;
    }
    done = 1;
  } else if (warp >= 4 && CONT != 0) {
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tm + ((warp >= 8) ? 128 : 0) + lane_off;   // columns [0,256): not the accumulators' (256..511)
    float acc = 0.f;
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0x3f800000u + i + lane;
    while (!done) {
      for (int c = 0; c < 4; ++c) {
        if (CONT == 1 || CONT == 3) {
          ptx::tmem_ld32(tS + c * 32, v);
          ptx::tmem_wait_ld();
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = __uint_as_float(v[2 * i]), p1 = __uint_as_float(v[2 * i + 1]);
          if (CONT == 1 || CONT == 2) {
            p0 = ptx::ex2(p0 * 0.5f - 1.0f);
            p1 = ptx::ex2(p1 * 0.5f - 1.0f);
          }
          acc += p0 + p1;
          pk[i] = ptx::pack_bf16(p0, p1);
          if (CONT == 2) { v[2 * i] = __float_as_uint(p0); v[2 * i + 1] = __float_as_uint(p1); }
        }
        if (CONT == 1 || CONT == 3) {
          ptx::tmem_st16(tS + 64 + c * 16, pk);
          ptx::tmem_wait_st();
        }
      }
    }
    if (acc == 123.456f) sink[threadIdx.x] = acc;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tm, 512);
  }
}

template <int MODE, int CONT = 0>
void run(const char* name, double flop_per_mma_avg) {
  Res* d;
  cudaMalloc(&d, 2 * sizeof(Res));
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(k<MODE, CONT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 200;
  float* sink;
  cudaMalloc(&sink, 4096);
  k<MODE, CONT><<<148, 384, smem>>>(d, 10, sink);
  k<MODE, CONT><<<148, 384, smem>>>(d, reps, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  Res h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const double n = (double)h[1].issue;
  printf("%-44s issue %7.1f cyc/MMA   complete %7.1f cyc/MMA   -> %6.0f flop/clk/SM\n", name, h[0].issue / n,
         h[0].total / n, flop_per_mma_avg * n / h[0].total);
  cudaFree(d);
}

int main() {
  run<0>("SS M128 N128 K16", 2.0 * 128 * 128 * 16);
  run<1>("SS M128 N256 K16", 2.0 * 128 * 256 * 16);
  run<2>("SS M128 N64  K16", 2.0 * 128 * 64 * 16);
  run<3>("TS M128 N64  K16 (A tmem, B MN-major)", 2.0 * 128 * 64 * 16);
  run<4>("TS M128 N128 K16 (A tmem, B MN-major)", 2.0 * 128 * 128 * 16);
  run<5>("attention slot: 8 TS N64 + 4 SS N128", (8 * 2.0 * 128 * 64 * 16 + 4 * 2.0 * 128 * 128 * 16) / 12);
  run<5, 1>("  + 8 softmax-like warps (TMEM ld/st + ex2)", (8 * 2.0 * 128 * 64 * 16 + 4 * 2.0 * 128 * 128 * 16) / 12);
  run<5, 2>("  + 8 math-only warps (ex2/fma)", (8 * 2.0 * 128 * 64 * 16 + 4 * 2.0 * 128 * 128 * 16) / 12);
  run<5, 3>("  + 8 TMEM ld/st-only warps", (8 * 2.0 * 128 * 64 * 16 + 4 * 2.0 * 128 * 128 * 16) / 12);
  run<0, 3>("SS N128 + 8 TMEM ld/st-only warps", 2.0 * 128 * 128 * 16);
  run<3, 3>("TS N64 + 8 TMEM ld/st-only warps", 2.0 * 128 * 64 * 16);
  run<6>("  same, commit after each group", (8 * 2.0 * 128 * 64 * 16 + 4 * 2.0 * 128 * 128 * 16) / 12);
  return 0;
}

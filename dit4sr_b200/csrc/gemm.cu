// Batched-row bf16 GEMM on sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (128x256 tiles, fp32 accumulators
// in TMEM, double-buffered) -> fused epilogues.  Persistent: one CTA per SM walks the tile list.
//
// Warp roles (256 threads): warp 0 = TMA producer (1 thread), warp 1 = MMA issuer (1 thread), warp 2 = TMEM
// allocator, warps 4..7 = epilogue (each owns the 32 TMEM lanes of its lane quarter; thread = one output row).
#include "host_util.h"
#include "ptx.cuh"

namespace gemm {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int ACC_STAGES = 2;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NUM_THREADS = 256;
constexpr int TMEM_COLS = ACC_STAGES * BN;  // 512
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct Args {
  const float* bias;
  void* C;
  int64_t ldc, c_batch_stride;
  int c_f32;
  int batches, rows, N, K;
  int m_tiles, n_tiles, num_tiles, k_blocks;
  const float* gate;
  int64_t gate_batch_stride;
  const float* resid;
  int64_t ldr, r_batch_stride;
  const float* norm_w0;
  const float* norm_w1;
  int sec_cols;
  float eps;
  int64_t dbl_row_offset;
};

__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5 x (1 + tanh( sqrt(2/pi) (x + 0.044715 x^3) ))
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + ptx::tanh_approx(inner));
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float* f) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = ptx::pack_bf16(f[8 * i + 0], f[8 * i + 1]);
    u.y = ptx::pack_bf16(f[8 * i + 2], f[8 * i + 3]);
    u.z = ptx::pack_bf16(f[8 * i + 4], f[8 * i + 5]);
    u.w = ptx::pack_bf16(f[8 * i + 6], f[8 * i + 7]);
    d4[i] = u;
  }
}
__device__ __forceinline__ void store_f32x32(float* dst, const float* f) {
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int i = 0; i < 8; ++i) d4[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
}

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const Args args) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + STAGES;             // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;         // [ACC_STAGES]
  uint64_t* tempty_bar = tfull_bar + ACC_STAGES;   // [ACC_STAGES]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_tiles = args.num_tiles;
  const int k_blocks = args.k_blocks;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile % args.n_tiles;
        const int rest = tile / args.n_tiles;
        const int m_tile = rest % args.m_tiles;
        const int b = rest / args.m_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          ptx::mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          ptx::tma_load_3d(sa, &tmA, &full_bar[stage], kb * BK, m_tile * BM, b);
          ptx::tma_load_2d(sb, &tmW, &full_bar[stage], kb * BK, n_tile * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = ptx::make_sdesc_sw128(sa);
          const uint64_t bdesc = ptx::make_sdesc_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128B swizzle atom: +2 in the (addr>>4) field
            ptx::mma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          ptx::tc_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::tc_commit(&tfull_bar[as]);  // accumulator complete
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int n_tile = tile % args.n_tiles;
      const int rest = tile / args.n_tiles;
      const int m_tile = rest % args.m_tiles;
      const int b = rest / args.m_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ptx::mbar_wait(&tfull_bar[as], aphase);
      ptx::tc_fence_after();

      const int r = m_tile * BM + q * 32 + lane;
      const bool row_ok = r < args.rows;
      const uint32_t taddr = tmem_base + as * BN + ((uint32_t)(q * 32) << 16);
      const int n_base = n_tile * BN;
      const int64_t c_off = (int64_t)b * args.c_batch_stride + (int64_t)r * args.ldc;

      if constexpr (EPI == DIT4SR_EPI_QKV_NORM) {
        __nv_bfloat16* Cb = reinterpret_cast<__nv_bfloat16*>(args.C);
#pragma unroll 1
        for (int hgrp = 0; hgrp < BN / 64; ++hgrp) {
          const int n0 = n_base + hgrp * 64;
          if (n0 >= args.N) break;
          uint32_t v0[32], v1[32];
          ptx::tmem_ld32(taddr + hgrp * 64, v0);
          ptx::tmem_ld32(taddr + hgrp * 64 + 32, v1);
          ptx::tmem_wait_ld();
          float f[64];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            f[j] = __uint_as_float(v0[j]) + (args.bias ? __ldg(args.bias + n0 + j) : 0.f);
            f[32 + j] = __uint_as_float(v1[j]) + (args.bias ? __ldg(args.bias + n0 + 32 + j) : 0.f);
          }
          const int sec = n0 / args.sec_cols;
          const float* w = sec == 0 ? args.norm_w0 : (sec == 1 ? args.norm_w1 : nullptr);
          if (w != nullptr) {
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) ss += f[j] * f[j];
            const float inv = rsqrtf(ss * (1.0f / 64.0f) + args.eps);
#pragma unroll
            for (int j = 0; j < 64; ++j) f[j] = f[j] * inv * __ldg(w + j);
          }
          if (row_ok) {
            store_bf16x32(Cb + c_off + n0, f);
            store_bf16x32(Cb + c_off + n0 + 32, f + 32);
          }
          if (sec == 1 && args.dbl_row_offset != 0 && w != nullptr) {
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) ss += f[j] * f[j];
            const float inv = rsqrtf(ss * (1.0f / 64.0f) + args.eps);
#pragma unroll
            for (int j = 0; j < 64; ++j) f[j] = f[j] * inv * __ldg(w + j);
            if (row_ok) {
              __nv_bfloat16* d = Cb + c_off + args.dbl_row_offset * args.ldc + n0;
              store_bf16x32(d, f);
              store_bf16x32(d + 32, f + 32);
            }
          }
        }
      } else {
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          const int n0 = n_base + ch * 32;
          if (n0 >= args.N) break;
          uint32_t v[32];
          ptx::tmem_ld32(taddr + ch * 32, v);
          ptx::tmem_wait_ld();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + (args.bias ? __ldg(args.bias + n0 + j) : 0.f);
          if constexpr (EPI == DIT4SR_EPI_BIAS_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_tanh(f[j]);
          } else if constexpr (EPI == DIT4SR_EPI_BIAS_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = silu(f[j]);
          } else if constexpr (EPI == DIT4SR_EPI_GATE_RESID) {
            if (row_ok) {
              const float4* rs = reinterpret_cast<const float4*>(args.resid + (int64_t)b * args.r_batch_stride +
                                                                 (int64_t)r * args.ldr + n0);
              const float* g = args.gate ? args.gate + (int64_t)b * args.gate_batch_stride + n0 : nullptr;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 rv = rs[i];
                const float g0 = g ? __ldg(g + 4 * i + 0) : 1.f, g1 = g ? __ldg(g + 4 * i + 1) : 1.f;
                const float g2 = g ? __ldg(g + 4 * i + 2) : 1.f, g3 = g ? __ldg(g + 4 * i + 3) : 1.f;
                f[4 * i + 0] = rv.x + g0 * f[4 * i + 0];
                f[4 * i + 1] = rv.y + g1 * f[4 * i + 1];
                f[4 * i + 2] = rv.z + g2 * f[4 * i + 2];
                f[4 * i + 3] = rv.w + g3 * f[4 * i + 3];
              }
            }
          }
          if (row_ok) {
            if (args.c_f32) store_f32x32(reinterpret_cast<float*>(args.C) + c_off + n0, f);
            else store_bf16x32(reinterpret_cast<__nv_bfloat16*>(args.C) + c_off + n0, f);
          }
        }
      }
      // release the accumulator stage
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int EPI>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmW, const Args& a, cudaStream_t stream) {
  static bool configured = false;  // benign race: idempotent attribute set
  if (!configured) {
    DIT4SR_CUDA(cudaFuncSetAttribute(gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  int grid = host::sm_count();
  if (grid > a.num_tiles) grid = a.num_tiles;
  gemm_kernel<EPI><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmA, tmW, a);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gemm

extern "C" int dit4sr_gemm_bf16(const dit4sr_gemm_params* p, void* stream_) {
  using namespace gemm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(p != nullptr && p->A && p->W && p->C, DIT4SR_E_BADARG, "gemm: null pointer argument");
  DIT4SR_REQUIRE(p->batches > 0 && p->rows > 0 && p->N > 0 && p->K > 0, DIT4SR_E_BADARG,
                 "gemm: non-positive size (batches %d rows %d N %d K %d)", p->batches, p->rows, p->N, p->K);
  DIT4SR_REQUIRE(p->K % 8 == 0 && p->N % 32 == 0, DIT4SR_E_SHAPE, "gemm: need K %% 8 == 0 and N %% 32 == 0 (K %d N %d)",
                 p->K, p->N);
  DIT4SR_REQUIRE(p->lda % 8 == 0 && p->ldw % 8 == 0 && (p->a_batch_stride % 8 == 0), DIT4SR_E_ALIGN,
                 "gemm: lda/ldw/a_batch_stride must be multiples of 8 elements");
  DIT4SR_REQUIRE(p->ldc % 8 == 0 && p->c_batch_stride % 8 == 0 && host::aligned16(p->C), DIT4SR_E_ALIGN,
                 "gemm: C pointer/strides must be 16-byte aligned");
  DIT4SR_REQUIRE(p->epilogue >= DIT4SR_EPI_BIAS && p->epilogue <= DIT4SR_EPI_GATE_RESID, DIT4SR_E_BADARG,
                 "gemm: unknown epilogue %d", p->epilogue);
  if (p->epilogue == DIT4SR_EPI_QKV_NORM) {
    DIT4SR_REQUIRE(p->N % 64 == 0 && p->sec_cols > 0 && p->sec_cols % 64 == 0 && p->c_dtype == DIT4SR_DT_BF16,
                   DIT4SR_E_SHAPE, "gemm/QKV_NORM: N and sec_cols must be multiples of 64 and C bf16");
  }
  if (p->epilogue == DIT4SR_EPI_GATE_RESID) {
    DIT4SR_REQUIRE(p->resid != nullptr && p->c_dtype == DIT4SR_DT_F32 && p->ldr % 4 == 0 && p->r_batch_stride % 4 == 0,
                   DIT4SR_E_BADARG, "gemm/GATE_RESID: needs fp32 resid and fp32 C");
  }
  if (p->epilogue == DIT4SR_EPI_BIAS_GELU || p->epilogue == DIT4SR_EPI_BIAS_SILU) {
    DIT4SR_REQUIRE(p->c_dtype == DIT4SR_DT_BF16, DIT4SR_E_BADARG, "gemm: activation epilogues write bf16");
  }
  int rc = dit4sr_check_device();
  if (rc != 0) return rc;

  CUtensorMap tmA, tmW;
  {
    uint64_t dims[3] = {(uint64_t)p->K, (uint64_t)p->rows, (uint64_t)p->batches};
    uint64_t strides[2] = {(uint64_t)p->lda * 2, (uint64_t)(p->batches > 1 ? p->a_batch_stride : (int64_t)p->rows * p->lda) * 2};
    uint32_t box[3] = {BK, BM, 1};
    rc = host::make_tmap_bf16(&tmA, p->A, 3, dims, strides, box);
    if (rc != 0) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)p->K, (uint64_t)p->N};
    uint64_t strides[1] = {(uint64_t)p->ldw * 2};
    uint32_t box[2] = {BK, BN};
    rc = host::make_tmap_bf16(&tmW, p->W, 2, dims, strides, box);
    if (rc != 0) return rc;
  }
  Args a;
  a.bias = p->bias;
  a.C = p->C;
  a.ldc = p->ldc;
  a.c_batch_stride = p->c_batch_stride;
  a.c_f32 = p->c_dtype == DIT4SR_DT_F32;
  a.batches = p->batches;
  a.rows = p->rows;
  a.N = p->N;
  a.K = p->K;
  a.m_tiles = (p->rows + BM - 1) / BM;
  a.n_tiles = (p->N + BN - 1) / BN;
  a.num_tiles = a.m_tiles * a.n_tiles * p->batches;
  a.k_blocks = (p->K + BK - 1) / BK;
  a.gate = p->gate;
  a.gate_batch_stride = p->gate_batch_stride;
  a.resid = p->resid;
  a.ldr = p->ldr;
  a.r_batch_stride = p->r_batch_stride;
  a.norm_w0 = p->norm_w0;
  a.norm_w1 = p->norm_w1;
  a.sec_cols = p->sec_cols > 0 ? p->sec_cols : p->N;
  a.eps = p->eps;
  a.dbl_row_offset = p->dbl_row_offset;
  switch (p->epilogue) {
    case DIT4SR_EPI_BIAS: return launch<DIT4SR_EPI_BIAS>(tmA, tmW, a, stream);
    case DIT4SR_EPI_BIAS_GELU: return launch<DIT4SR_EPI_BIAS_GELU>(tmA, tmW, a, stream);
    case DIT4SR_EPI_BIAS_SILU: return launch<DIT4SR_EPI_BIAS_SILU>(tmA, tmW, a, stream);
    case DIT4SR_EPI_QKV_NORM: return launch<DIT4SR_EPI_QKV_NORM>(tmA, tmW, a, stream);
    default: return launch<DIT4SR_EPI_GATE_RESID>(tmA, tmW, a, stream);
  }
}

// LayerNorm (no affine) fused with adaLN modulation; HBM-bound.  One warp per token row, the whole row lives in
// registers (two-pass mean/variance in fp32), 16-byte coalesced loads, one or two bf16 outputs from a single read.
#include "host_util.h"
#include "ptx.cuh"

namespace lnmod {

constexpr int WARPS = 8;

// fp32-output variant (fp32-accurate mode): same statistics, one float4 store per 4 elements
template <int NV>
__global__ void __launch_bounds__(WARPS * 32)
ln_modulate_f32_kernel(const dit4sr_ln_modulate_params p) {
  const int lane = threadIdx.x & 31;
  const int64_t row_id = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int64_t total = (int64_t)p.batches * p.rows;
  if (row_id >= total) return;
  const int b = (int)(row_id / p.rows);
  const int r = (int)(row_id % p.rows);
  const float4* x = reinterpret_cast<const float4*>(p.x + (int64_t)b * p.x_batch_stride + (int64_t)r * p.ldx);
  float4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = __ldg(x + i * 32 + lane);
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv_dim = 1.0f / (float)(NV * 128);
  const float mean = sum * inv_dim;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = 1.0f / sqrtf(sq * inv_dim + p.eps);
  for (int which = 0; which < (p.out2 != nullptr ? 2 : 1); ++which) {
    const float4* sh = reinterpret_cast<const float4*>((which ? p.shift2 : p.shift) + (int64_t)b * p.mod_batch_stride);
    const float4* sc = reinterpret_cast<const float4*>((which ? p.scale2 : p.scale) + (int64_t)b * p.mod_batch_stride);
    float4* out = reinterpret_cast<float4*>(reinterpret_cast<float*>(which ? p.out2 : p.out) +
                                            (int64_t)b * p.o_batch_stride + (int64_t)r * p.ldo);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 s = __ldg(sc + i * 32 + lane), t = __ldg(sh + i * 32 + lane);
      out[i * 32 + lane] = make_float4(v[i].x * rstd * (1.f + s.x) + t.x, v[i].y * rstd * (1.f + s.y) + t.y,
                                       v[i].z * rstd * (1.f + s.z) + t.z, v[i].w * rstd * (1.f + s.w) + t.w);
    }
  }
}

template <int NV, bool DUAL>  // NV float4 per lane => dim = NV * 128
__global__ void __launch_bounds__(WARPS * 32)
ln_modulate_kernel(const dit4sr_ln_modulate_params p) {
  const int lane = threadIdx.x & 31;
  const int64_t row_id = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int64_t total = (int64_t)p.batches * p.rows;
  if (row_id >= total) return;
  const int b = (int)(row_id / p.rows);
  const int r = (int)(row_id % p.rows);
  const float4* x = reinterpret_cast<const float4*>(p.x + (int64_t)b * p.x_batch_stride + (int64_t)r * p.ldx);
  float4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = __ldg(x + i * 32 + lane);
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv_dim = 1.0f / (float)(NV * 128);
  const float mean = sum * inv_dim;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq * inv_dim + p.eps);

  const float4* sh = reinterpret_cast<const float4*>(p.shift + (int64_t)b * p.mod_batch_stride);
  const float4* sc = reinterpret_cast<const float4*>(p.scale + (int64_t)b * p.mod_batch_stride);
  uint2* out = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)b * p.o_batch_stride +
                                        (int64_t)r * p.ldo);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 s = __ldg(sc + i * 32 + lane), t = __ldg(sh + i * 32 + lane);
    const float n0 = v[i].x * rstd, n1 = v[i].y * rstd, n2 = v[i].z * rstd, n3 = v[i].w * rstd;
    uint2 o;
    o.x = ptx::pack_bf16(n0 * (1.f + s.x) + t.x, n1 * (1.f + s.y) + t.y);
    o.y = ptx::pack_bf16(n2 * (1.f + s.z) + t.z, n3 * (1.f + s.w) + t.w);
    out[i * 32 + lane] = o;
  }
  if constexpr (DUAL) {
    const float4* sh2 = reinterpret_cast<const float4*>(p.shift2 + (int64_t)b * p.mod_batch_stride);
    const float4* sc2 = reinterpret_cast<const float4*>(p.scale2 + (int64_t)b * p.mod_batch_stride);
    uint2* out2 = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + (int64_t)b * p.o_batch_stride +
                                           (int64_t)r * p.ldo);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 s = __ldg(sc2 + i * 32 + lane), t = __ldg(sh2 + i * 32 + lane);
      const float n0 = v[i].x * rstd, n1 = v[i].y * rstd, n2 = v[i].z * rstd, n3 = v[i].w * rstd;
      uint2 o;
      o.x = ptx::pack_bf16(n0 * (1.f + s.x) + t.x, n1 * (1.f + s.y) + t.y);
      o.y = ptx::pack_bf16(n2 * (1.f + s.z) + t.z, n3 * (1.f + s.w) + t.w);
      out2[i * 32 + lane] = o;
    }
  }
}

template <int NV>
static int launch(const dit4sr_ln_modulate_params& p, cudaStream_t stream) {
  const int64_t total = (int64_t)p.batches * p.rows;
  const unsigned grid = (unsigned)((total + WARPS - 1) / WARPS);
  if (p.out_f32) ln_modulate_f32_kernel<NV><<<grid, WARPS * 32, 0, stream>>>(p);
  else if (p.out2 != nullptr) ln_modulate_kernel<NV, true><<<grid, WARPS * 32, 0, stream>>>(p);
  else ln_modulate_kernel<NV, false><<<grid, WARPS * 32, 0, stream>>>(p);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace lnmod

extern "C" int dit4sr_ln_modulate(const dit4sr_ln_modulate_params* p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(p && p->x && p->shift && p->scale && p->out, DIT4SR_E_BADARG, "ln_modulate: null pointer argument");
  DIT4SR_REQUIRE(p->batches > 0 && p->rows > 0 && p->dim > 0, DIT4SR_E_BADARG, "ln_modulate: non-positive size");
  DIT4SR_REQUIRE((p->out2 == nullptr) || (p->shift2 && p->scale2), DIT4SR_E_BADARG,
                 "ln_modulate: out2 given without shift2/scale2");
  DIT4SR_REQUIRE(p->ldx % 4 == 0 && p->x_batch_stride % 4 == 0 && p->ldo % 4 == 0 && p->o_batch_stride % 4 == 0 &&
                     p->mod_batch_stride % 4 == 0 && host::aligned16(p->x) && host::aligned16(p->shift) &&
                     host::aligned16(p->scale) && (reinterpret_cast<uintptr_t>(p->out) & 7) == 0,
                 DIT4SR_E_ALIGN, "ln_modulate: pointers/strides not 16-byte aligned");
  switch (p->dim) {
    case 128: return lnmod::launch<1>(*p, stream);
    case 256: return lnmod::launch<2>(*p, stream);
    case 512: return lnmod::launch<4>(*p, stream);
    case 1024: return lnmod::launch<8>(*p, stream);
    case 1152: return lnmod::launch<9>(*p, stream);
    case 1536: return lnmod::launch<12>(*p, stream);
    case 2048: return lnmod::launch<16>(*p, stream);
    default:
      host::set_error("ln_modulate: dim %d unsupported (128,256,512,1024,1152,1536,2048)", p->dim);
      return DIT4SR_E_SHAPE;
  }
}

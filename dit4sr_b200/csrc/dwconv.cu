// Depthwise 3x3 (pad 1, bias) of the LR half added into the HQ half, in place, on the token-major (NHWC) layout.
// HBM-bound stencil: lane = 8 consecutive channels (16-byte vectors, a warp covers 512 contiguous bytes), warp = one x
// position of an 8-wide strip, each thread walks down a y segment keeping the 3x3 window of its channels in registers
// (the x-1/x/x+1 rows loaded by neighbouring warps of the CTA are L1 hits). Weights live in registers.
#include "host_util.h"
#include "ptx.cuh"

namespace dwconv {

constexpr int XT = 8;    // x positions per CTA (one per warp)
constexpr int TY = 8;    // y rows per CTA

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__global__ void __launch_bounds__(XT * 32)
dwconv3x3_kernel(const dit4sr_dwconv3x3_params p, int x_tiles) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 32 + lane) * 8;
  const int xt = blockIdx.y % x_tiles;
  const int ys = blockIdx.y / x_tiles;
  const int b = blockIdx.z;
  const int x = xt * XT + warp;
  const int G = p.grid;
  if (c0 >= p.channels || x >= G) return;

  float w[9][8], bias[8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.weight9 + (int64_t)t * p.channels + c0));
    const float4 c = __ldg(reinterpret_cast<const float4*>(p.weight9 + (int64_t)t * p.channels + c0) + 1);
    w[t][0] = a.x; w[t][1] = a.y; w[t][2] = a.z; w[t][3] = a.w;
    w[t][4] = c.x; w[t][5] = c.y; w[t][6] = c.z; w[t][7] = c.w;
  }
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
    const float4 c = __ldg(reinterpret_cast<const float4*>(p.bias + c0) + 1);
    bias[0] = a.x; bias[1] = a.y; bias[2] = a.z; bias[3] = a.w;
    bias[4] = c.x; bias[5] = c.y; bias[6] = c.z; bias[7] = c.w;
  }
  const __nv_bfloat16* lr = reinterpret_cast<const __nv_bfloat16*>(p.u_lr) + (int64_t)b * p.lr_batch_stride + c0;
  __nv_bfloat16* hq = reinterpret_cast<__nv_bfloat16*>(p.u_hq) + (int64_t)b * p.hq_batch_stride + c0;

  const uint4 zero = make_uint4(0, 0, 0, 0);
  auto load_row = [&](int y, uint4 (&row)[3]) {
    if (y < 0 || y >= G) {
      row[0] = row[1] = row[2] = zero;
      return;
    }
    const __nv_bfloat16* base = lr + ((int64_t)y * G + x) * p.ld;
    row[0] = x > 0 ? __ldg(reinterpret_cast<const uint4*>(base - p.ld)) : zero;
    row[1] = __ldg(reinterpret_cast<const uint4*>(base));
    row[2] = x + 1 < G ? __ldg(reinterpret_cast<const uint4*>(base + p.ld)) : zero;
  };

  const int y_begin = ys * TY;
  const int y_end = min(y_begin + TY, G);
  uint4 win[3][3];  // [row: y-1, y, y+1][col: x-1, x, x+1]
  load_row(y_begin - 1, win[0]);
  load_row(y_begin, win[1]);
  for (int y = y_begin; y < y_end; ++y) {
    load_row(y + 1, win[2]);
    uint4* dst = reinterpret_cast<uint4*>(hq + ((int64_t)y * G + x) * p.ld);
    const uint4 hv = *dst;
    float acc[8], t[8];
    unpack8(hv, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += bias[j];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        unpack8(win[dy][dx], t);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(w[dy * 3 + dx][j], t[j], acc[j]);
      }
    uint4 o;
    o.x = ptx::pack_bf16(acc[0], acc[1]);
    o.y = ptx::pack_bf16(acc[2], acc[3]);
    o.z = ptx::pack_bf16(acc[4], acc[5]);
    o.w = ptx::pack_bf16(acc[6], acc[7]);
    *dst = o;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      win[0][dx] = win[1][dx];
      win[1][dx] = win[2][dx];
    }
  }
}

}  // namespace dwconv

extern "C" int dit4sr_dwconv3x3_add(const dit4sr_dwconv3x3_params* p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(p && p->u_hq && p->u_lr && p->weight9 && p->bias, DIT4SR_E_BADARG, "dwconv: null pointer argument");
  DIT4SR_REQUIRE(p->batches > 0 && p->grid > 0 && p->channels > 0, DIT4SR_E_BADARG, "dwconv: non-positive size");
  DIT4SR_REQUIRE(p->channels % 8 == 0, DIT4SR_E_SHAPE, "dwconv: channels %d not a multiple of 8", p->channels);
  DIT4SR_REQUIRE(p->ld % 8 == 0 && p->hq_batch_stride % 8 == 0 && p->lr_batch_stride % 8 == 0 &&
                     host::aligned16(p->u_hq) && host::aligned16(p->u_lr) && host::aligned16(p->weight9) &&
                     host::aligned16(p->bias),
                 DIT4SR_E_ALIGN, "dwconv: pointers/strides not 16-byte aligned");
  DIT4SR_REQUIRE(p->batches <= 65535, DIT4SR_E_SHAPE, "dwconv: batches > 65535");
  const int x_tiles = (p->grid + dwconv::XT - 1) / dwconv::XT;
  const int y_segs = (p->grid + dwconv::TY - 1) / dwconv::TY;
  dim3 grid((p->channels / 8 + 31) / 32, x_tiles * y_segs, p->batches);
  dwconv::dwconv3x3_kernel<<<grid, dwconv::XT * 32, 0, stream>>>(*p, x_tiles);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

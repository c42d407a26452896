// Depthwise 3x3 (pad 1, bias) of the LR half added into the HQ half, in place, on the token-major (NHWC) layout.
// HBM-bound stencil, shared-memory staged:
//   * one CTA = a TY x TX pixel tile x 64 channels; ONE 4-D TMA box load brings the (TY+2) x (TX+2) x 64 halo tile of
//     the LR stream into shared memory -- TMA's out-of-bounds zero fill IS the conv's zero padding;
//   * several CTAs are resident per SM (26 KB of smem each), so many tiles are in flight per SM (memory-level
//     parallelism comes from the TMA engine, not from per-thread loads);
//   * thread = 8 channels (16 B) x one pixel column: it slides a 3x3 register window down the tile rows, weights in
//     registers, HQ read-modify-written with 16-byte accesses issued before the TMA wait.
#include "host_util.h"
#include "ptx.cuh"
#include <stdlib.h>

namespace dwconv {

constexpr int CH = 64;        // channels per CTA (128 bytes per pixel)
constexpr int TX = 32;        // pixel columns per CTA
constexpr int TY = 4;         // pixel rows per CTA
constexpr int BX = TX + 2, BY = TY + 2;
constexpr int THREADS = 128;  // 8 channel groups x 16 pixel columns; each thread does columns x and x+16
constexpr int TILE_BYTES = BY * BX * CH * 2;

struct Args {
  __nv_bfloat16* u_hq;
  int64_t ld, hq_batch_stride;
  const float* weight9;
  const float* bias;
  int grid_h, grid_w, lr_row_offset, channels, x_tiles, y_tiles, batches;
  int reverse;   // pipelined form: walk the tiles from the last one down (the producing GEMM wrote the rows ascending)
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__global__ void __launch_bounds__(THREADS, 3)
dwconv3x3_kernel(const __grid_constant__ CUtensorMap tmLR, const Args a) {
  __shared__ __align__(128) uint8_t tile[TILE_BYTES];
  __shared__ __align__(8) uint64_t bar;

  const int tid = threadIdx.x;
  const int cg = tid & 7;           // channel group (8 channels)
  const int px = tid >> 3;          // 0..15
  const int c_base = blockIdx.x * CH;
  const int xt = blockIdx.y % a.x_tiles;
  const int yt = blockIdx.y / a.x_tiles;
  const int b = blockIdx.z;
  const int x0 = xt * TX, y0 = yt * TY;
  const int c0 = c_base + cg * 8;

  if (tid == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  ptx::pdl_wait();                 // (programmatic dependent launch: everything below reads the previous kernel's output)
  ptx::pdl_launch_dependents();
  if (tid == 0) {
    ptx::mbar_expect_tx(&bar, TILE_BYTES);
    // negative / past-the-end coordinates read as zero: the conv's zero padding
    ptx::tma_load_4d(tile, &tmLR, &bar, c_base, x0 - 1, y0 - 1 + a.lr_row_offset, b);
  }

  // weights + bias for this thread's 8 channels as packed fp32 pairs (L2-resident, 221 KB for F=6144)
  uint64_t w2[9][4], bias2[4];
  const bool c_ok = c0 < a.channels;
  if (c_ok) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4* p = reinterpret_cast<const float4*>(a.weight9 + (int64_t)t * a.channels + c0);
      const float4 lo = __ldg(p), hi = __ldg(p + 1);
      w2[t][0] = ptx::pack2(lo.x, lo.y); w2[t][1] = ptx::pack2(lo.z, lo.w);
      w2[t][2] = ptx::pack2(hi.x, hi.y); w2[t][3] = ptx::pack2(hi.z, hi.w);
    }
    const float4* p = reinterpret_cast<const float4*>(a.bias + c0);
    const float4 lo = __ldg(p), hi = __ldg(p + 1);
    bias2[0] = ptx::pack2(lo.x, lo.y); bias2[1] = ptx::pack2(lo.z, lo.w);
    bias2[2] = ptx::pack2(hi.x, hi.y); bias2[3] = ptx::pack2(hi.z, hi.w);
  }
  __syncthreads();   // barrier init visible to all waiters

  __nv_bfloat16* hq = a.u_hq + (int64_t)b * a.hq_batch_stride + c0;
  const int GH = a.grid_h, GW = a.grid_w;

  // bf16 pair -> packed fp32 pair: (lo << 16, hi & 0xffff0000)
  auto unpack2 = [](uint32_t u) { return ptx::pack2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); };

#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const int lx = px + half * 16;         // column inside the tile
    const int x = x0 + lx;
    const bool col_ok = c_ok && x < GW;
    // HQ values of this column (issued before the TMA wait)
    uint4 hv[TY];
#pragma unroll
    for (int r = 0; r < TY; ++r) {
      const int y = y0 + r;
      hv[r] = (col_ok && y < GH) ? *reinterpret_cast<const uint4*>(hq + ((int64_t)y * GW + x) * a.ld) : make_uint4(0, 0, 0, 0);
    }
    if (half == 0) ptx::mbar_wait(&bar, 0);
    if (!col_ok) continue;
    // smem tile layout: [BY][BX][CH] bf16; this thread reads 16 B at channel offset cg*16.
    // Every LR row of the halo tile is read and unpacked ONCE and scattered into the (up to three) output rows it
    // contributes to; all arithmetic is packed fp32x2 (FFMA2): 9 taps x 4 pairs per output pixel.
    const uint8_t* sbase = tile + (size_t)lx * (CH * 2) + cg * 16;
    uint64_t acc[TY][4];
#pragma unroll
    for (int r = 0; r < TY; ++r) {
      const uint32_t* h = reinterpret_cast<const uint32_t*>(&hv[r]);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[r][k] = ptx::add2(unpack2(h[k]), bias2[k]);
    }
#pragma unroll
    for (int ri = 0; ri < BY; ++ri) {
      uint64_t in[3][4];
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const uint4 u = *reinterpret_cast<const uint4*>(sbase + ((size_t)ri * BX + dx) * (CH * 2));
        in[dx][0] = unpack2(u.x); in[dx][1] = unpack2(u.y); in[dx][2] = unpack2(u.z); in[dx][3] = unpack2(u.w);
      }
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int r = ri - dy;              // output row fed by this input row through tap row dy
        if (r >= 0 && r < TY) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[r][k] = ptx::fma2(w2[dy * 3 + dx][k], in[dx][k], acc[r][k]);
        }
      }
      const int r = ri - 2;                 // this output row is complete
      if (r >= 0) {
        const int y = y0 + r;
        if (y < GH) {
          uint4 o;
          float f0, f1;
          ptx::unpack2(acc[r][0], f0, f1); o.x = ptx::pack_bf16(f0, f1);
          ptx::unpack2(acc[r][1], f0, f1); o.y = ptx::pack_bf16(f0, f1);
          ptx::unpack2(acc[r][2], f0, f1); o.z = ptx::pack_bf16(f0, f1);
          ptx::unpack2(acc[r][3], f0, f1); o.w = ptx::pack_bf16(f0, f1);
          *reinterpret_cast<uint4*>(hq + ((int64_t)y * GW + x) * a.ld) = o;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Second form (default since round 2b): persistent CTAs with every byte of global traffic on the TMA engine.
// ncu of the kernel above (profiles/r02b_ncu_full.json): DRAM 46 % busy, issue slots 34 %, 11 warps per SM - a CTA loads,
// then computes, then stores, and three such CTAs per SM do not keep enough bytes in flight.  Here a CTA owns one
// 64-channel block and walks tiles (x tile, y tile, batch) with a two-stage shared-memory ring: while tile i is computed,
// the halo box of the LR stream AND the HQ tile of tile i+1 are in flight as 4-D TMA loads, results are written back into
// the HQ tile in shared memory and leave as one TMA store (box elements outside the image are dropped on the way out, as
// they are zero-filled on the way in), weights and bias are loaded once per CTA.  Two CTAs per SM (85 KB each).
constexpr int LR_BYTES = TILE_BYTES;                       // (TY+2) x (TX+2) x 64 ch
constexpr int HQ_BYTES = TY * TX * CH * 2;                 // TY x TX x 64 ch
constexpr int STAGE_BYTES = LR_BYTES + HQ_BYTES;
static_assert(LR_BYTES % 128 == 0 && HQ_BYTES % 128 == 0, "TMA destinations stay 128-byte aligned");
constexpr int PIPE_SMEM = 2 * STAGE_BYTES + 128 + 64;

__global__ void __launch_bounds__(THREADS, 2)
dwconv3x3_pipe_kernel(const __grid_constant__ CUtensorMap tmLR, const __grid_constant__ CUtensorMap tmHQ, const Args a) {
  extern __shared__ uint8_t dw_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dw_smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * STAGE_BYTES);

  const int tid = threadIdx.x;
  const int cg = tid & 7;           // channel group (8 channels)
  const int px = tid >> 3;          // 0..15
  const int c_base = blockIdx.x * CH;
  const int c0 = c_base + cg * 8;
  const int tiles_per_image = a.x_tiles * a.y_tiles;
  const int total = tiles_per_image * a.batches;
  const int first = blockIdx.y, stride = gridDim.y;

  if (tid == 0) {
    ptx::prefetch_tmap(&tmLR);
    ptx::prefetch_tmap(&tmHQ);
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::fence_barrier_init();
  }
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  __syncthreads();

  auto coords = [&](int t, int& x0, int& y0, int& b) {
    if (a.reverse) t = total - 1 - t;
    b = t / tiles_per_image;
    const int r = t - b * tiles_per_image;
    const int yt = r / a.x_tiles;
    x0 = (r - yt * a.x_tiles) * TX;
    y0 = yt * TY;
  };
  auto issue = [&](int t, int s) {   // both boxes of tile t into stage s (negative / past-the-end coordinates read as zero)
    int x0, y0, b;
    coords(t, x0, y0, b);
    uint8_t* lr = smem + s * STAGE_BYTES;
    ptx::mbar_expect_tx(&bar[s], STAGE_BYTES);
    ptx::tma_load_4d(lr, &tmLR, &bar[s], c_base, x0 - 1, y0 - 1 + a.lr_row_offset, b);
    ptx::tma_load_4d(lr + LR_BYTES, &tmHQ, &bar[s], c_base, x0, y0, b);
  };
  if (tid == 0 && first < total) issue(first, 0);

  // weights + bias for this thread's 8 channels as packed fp32 pairs, once per CTA
  uint64_t w2[9][4], bias2[4];
  if (c0 < a.channels) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4* p = reinterpret_cast<const float4*>(a.weight9 + (int64_t)t * a.channels + c0);
      const float4 lo = __ldg(p), hi = __ldg(p + 1);
      w2[t][0] = ptx::pack2(lo.x, lo.y); w2[t][1] = ptx::pack2(lo.z, lo.w);
      w2[t][2] = ptx::pack2(hi.x, hi.y); w2[t][3] = ptx::pack2(hi.z, hi.w);
    }
    const float4* p = reinterpret_cast<const float4*>(a.bias + c0);
    const float4 lo = __ldg(p), hi = __ldg(p + 1);
    bias2[0] = ptx::pack2(lo.x, lo.y); bias2[1] = ptx::pack2(lo.z, lo.w);
    bias2[2] = ptx::pack2(hi.x, hi.y); bias2[3] = ptx::pack2(hi.z, hi.w);
  } else {   // channel block past the end (channels % 64 != 0): the boxes are zero there and the store drops them
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k) w2[t][k] = 0ull;
#pragma unroll
    for (int k = 0; k < 4; ++k) bias2[k] = 0ull;
  }
  auto unpack2 = [](uint32_t u) { return ptx::pack2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); };

  int it = 0;
#pragma unroll 1
  for (int t = first; t < total; t += stride, ++it) {
    const int s = it & 1;
    if (tid == 0 && t + stride < total) {
      ptx::bulk_wait_read<0>();        // the store issued from the other stage one iteration ago has read it
      issue(t + stride, s ^ 1);
    }
    ptx::mbar_wait(&bar[s], (it >> 1) & 1);
    uint8_t* lr = smem + s * STAGE_BYTES;
    uint8_t* hq = lr + LR_BYTES;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int lx = px + half * 16;         // column inside the tile
      const uint8_t* sbase = lr + (size_t)lx * (CH * 2) + cg * 16;
      uint8_t* hbase = hq + (size_t)lx * (CH * 2) + cg * 16;
      uint64_t acc[TY][4];
#pragma unroll
      for (int r = 0; r < TY; ++r) {
        const uint4 h = *reinterpret_cast<const uint4*>(hbase + (size_t)r * TX * (CH * 2));
        acc[r][0] = ptx::add2(unpack2(h.x), bias2[0]);
        acc[r][1] = ptx::add2(unpack2(h.y), bias2[1]);
        acc[r][2] = ptx::add2(unpack2(h.z), bias2[2]);
        acc[r][3] = ptx::add2(unpack2(h.w), bias2[3]);
      }
#pragma unroll
      for (int ri = 0; ri < BY; ++ri) {
        uint64_t in[3][4];
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const uint4 u = *reinterpret_cast<const uint4*>(sbase + ((size_t)ri * BX + dx) * (CH * 2));
          in[dx][0] = unpack2(u.x); in[dx][1] = unpack2(u.y); in[dx][2] = unpack2(u.z); in[dx][3] = unpack2(u.w);
        }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int r = ri - dy;              // output row fed by this input row through tap row dy
          if (r >= 0 && r < TY) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
              for (int k = 0; k < 4; ++k) acc[r][k] = ptx::fma2(w2[dy * 3 + dx][k], in[dx][k], acc[r][k]);
          }
        }
        const int r = ri - 2;                 // this output row is complete: back into the HQ tile
        if (r >= 0) {
          uint4 o;
          float f0, f1;
          ptx::unpack2(acc[r][0], f0, f1); o.x = ptx::pack_bf16(f0, f1);
          ptx::unpack2(acc[r][1], f0, f1); o.y = ptx::pack_bf16(f0, f1);
          ptx::unpack2(acc[r][2], f0, f1); o.z = ptx::pack_bf16(f0, f1);
          ptx::unpack2(acc[r][3], f0, f1); o.w = ptx::pack_bf16(f0, f1);
          *reinterpret_cast<uint4*>(hbase + (size_t)r * TX * (CH * 2)) = o;
        }
      }
    }
    ptx::fence_proxy_async();          // the tile written through the generic proxy is read by the TMA store
    __syncthreads();
    if (tid == 0) {
      int x0, y0, b;
      coords(t, x0, y0, b);
      ptx::tma_store_4d(&tmHQ, hq, c_base, x0, y0, b);
      ptx::bulk_commit();
    }
  }
  if (tid == 0) ptx::bulk_wait_all();
}

}  // namespace dwconv

extern "C" int dit4sr_dwconv3x3_add(const dit4sr_dwconv3x3_params* p, void* stream_) {
  using namespace dwconv;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(p && p->u_hq && p->u_lr && p->weight9 && p->bias, DIT4SR_E_BADARG, "dwconv: null pointer argument");
  DIT4SR_REQUIRE(p->batches > 0 && p->grid > 0 && p->channels > 0, DIT4SR_E_BADARG, "dwconv: non-positive size");
  DIT4SR_REQUIRE(p->channels % 8 == 0, DIT4SR_E_SHAPE, "dwconv: channels %d not a multiple of 8", p->channels);
  DIT4SR_REQUIRE(p->ld % 8 == 0 && p->hq_batch_stride % 8 == 0 && p->lr_batch_stride % 8 == 0 &&
                     host::aligned16(p->u_hq) && host::aligned16(p->u_lr) && host::aligned16(p->weight9) &&
                     host::aligned16(p->bias),
                 DIT4SR_E_ALIGN, "dwconv: pointers/strides not 16-byte aligned");
  DIT4SR_REQUIRE(p->batches <= 65535, DIT4SR_E_SHAPE, "dwconv: batches > 65535");
  int rc = dit4sr_check_device();
  if (rc != 0) return rc;

  const int gw = p->grid_w > 0 ? p->grid_w : p->grid;
  const int lr_h = p->lr_grid_h > 0 ? p->lr_grid_h : p->grid;
  DIT4SR_REQUIRE(p->grid_w >= 0 && p->lr_grid_h >= 0, DIT4SR_E_BADARG, "dwconv: negative grid extents");
  CUtensorMap tm;
  const uint64_t dims[4] = {(uint64_t)p->channels, (uint64_t)gw, (uint64_t)lr_h, (uint64_t)p->batches};
  const uint64_t strides[3] = {(uint64_t)p->ld * 2, (uint64_t)p->ld * 2 * gw,
                               (uint64_t)(p->batches > 1 ? p->lr_batch_stride : p->ld * (int64_t)gw * lr_h) * 2};
  const uint32_t box[4] = {CH, BX, BY, 1};
  if ((rc = host::make_tmap_bf16_noswizzle(&tm, p->u_lr, 4, dims, strides, box)) != 0) return rc;

  Args a;
  a.u_hq = reinterpret_cast<__nv_bfloat16*>(p->u_hq);
  a.ld = p->ld;
  a.hq_batch_stride = p->hq_batch_stride;
  a.weight9 = p->weight9;
  a.bias = p->bias;
  a.grid_h = p->grid;
  a.grid_w = gw;
  a.lr_row_offset = p->lr_row_offset;
  a.channels = p->channels;
  a.x_tiles = (gw + TX - 1) / TX;
  a.y_tiles = (p->grid + TY - 1) / TY;
  a.batches = p->batches;
  static int reverse = -1;   // DIT4SR_DWCONV_REVERSE=0: ascending tile order
  if (reverse < 0) { const char* e = getenv("DIT4SR_DWCONV_REVERSE"); reverse = (e != nullptr && e[0] == '0') ? 0 : 1; }
  a.reverse = reverse;
  static int variant = -1;   // developer switch DIT4SR_DWCONV_VARIANT: 0 = first form (one tile per CTA), 1 = pipelined (default)
  if (variant < 0) { const char* e = getenv("DIT4SR_DWCONV_VARIANT"); variant = e ? atoi(e) : 1; }
  if (variant == 1) {
    CUtensorMap tmh;
    const uint64_t hdims[4] = {(uint64_t)p->channels, (uint64_t)gw, (uint64_t)p->grid, (uint64_t)p->batches};
    const uint64_t hstrides[3] = {(uint64_t)p->ld * 2, (uint64_t)p->ld * 2 * gw,
                                  (uint64_t)(p->batches > 1 ? p->hq_batch_stride : p->ld * (int64_t)gw * p->grid) * 2};
    const uint32_t hbox[4] = {CH, TX, TY, 1};
    if ((rc = host::make_tmap_bf16_noswizzle(&tmh, p->u_hq, 4, hdims, hstrides, hbox)) != 0) return rc;
    static unsigned long long configured = 0;   // per (instance, device)
    if (host::first_use_on_device(&configured))
      DIT4SR_CUDA(cudaFuncSetAttribute(dwconv3x3_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM));
    const int cblocks = (p->channels + CH - 1) / CH;
    const int tiles = a.x_tiles * a.y_tiles * p->batches;
    int per_block = 2 * host::sm_count() / cblocks;          // CTAs per channel block: about two per SM in total
    if (per_block < 1) per_block = 1;
    if (per_block > tiles) per_block = tiles;
    DIT4SR_CUDA(host::launch_pdl(dwconv3x3_pipe_kernel, dim3(cblocks, per_block), dim3(THREADS), (size_t)PIPE_SMEM, stream, tm,
                                 tmh, a));
    return 0;
  }
  dim3 grid((p->channels + CH - 1) / CH, a.x_tiles * a.y_tiles, p->batches);
  DIT4SR_CUDA(host::launch_pdl(dwconv3x3_kernel, grid, dim3(THREADS), (size_t)0, stream, tm, a));
  return 0;
}

// LayerNorm (no affine) fused with adaLN modulation; HBM-bound.  One warp per token row, the whole row lives in
// registers (two-pass mean/variance in fp32), 16-byte coalesced loads, one or two bf16 outputs from a single read.
#include "host_util.h"
#include "ptx.cuh"
#include <stdlib.h>

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



// TMA-staged variant (bf16 outputs): the register kernel below keeps DRAM only ~45 % busy (ncu: long-scoreboard bound,
// 23 resident warps x one row of loads each).  Here every warp is persistent and keeps LN_STAGES rows in flight as 1-D
// bulk copies (cp.async.bulk -> its own shared-memory slots, one mbarrier per slot), which costs no registers; a row is
// read back from shared memory with a lane owning 8 consecutive elements per 256-element chunk, so each output store
// is 16 bytes per lane (a full 512-byte segment per warp instruction), and the slot is refilled with the row
// LN_STAGES ahead as soon as the row sits in registers.
constexpr int LN_STAGES = 3;
constexpr int LN_WARPS = 8;

template <int NC, bool DUAL>  // NC chunks of 256 elements => dim = NC * 256
__global__ void __launch_bounds__(LN_WARPS * 32, 1)
ln_modulate_tma_kernel(const dit4sr_ln_modulate_params p) {
  extern __shared__ uint8_t ln_smem_raw[];
  constexpr int ROW_BYTES = NC * 256 * 4;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ln_smem_raw) + 127) & ~uintptr_t(127));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* slots = smem + (size_t)warp * LN_STAGES * ROW_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)LN_WARPS * LN_STAGES * ROW_BYTES) + warp * LN_STAGES;
  const int64_t total = (int64_t)p.batches * p.rows;
  const int64_t nwarps = (int64_t)gridDim.x * LN_WARPS;
  const int64_t first = (int64_t)blockIdx.x * LN_WARPS + warp;
  auto row_ptr = [&](int64_t rid) {
    const int b = (int)(rid / p.rows);
    const int r = (int)(rid % p.rows);
    return p.x + (int64_t)b * p.x_batch_stride + (int64_t)r * p.ldx;
  };
  if (lane == 0) {
    for (int s = 0; s < LN_STAGES; ++s) ptx::mbar_init(&bars[s], 1);
    ptx::fence_barrier_init();
  }
  ptx::pdl_wait();                 // (programmatic dependent launch: barrier set-up ran under the previous kernel's tail)
  ptx::pdl_launch_dependents();
  if (lane == 0) {
    for (int s = 0; s < LN_STAGES; ++s) {
      const int64_t rid = first + s * nwarps;
      if (rid < total) {
        ptx::mbar_expect_tx(&bars[s], ROW_BYTES);
        ptx::bulk_load(slots + s * ROW_BYTES, row_ptr(rid), ROW_BYTES, &bars[s]);
      }
    }
  }
  __syncwarp();
  const float inv_dim = 1.0f / (float)(NC * 256);
  int k = 0;
  for (int64_t rid = first; rid < total; rid += nwarps, ++k) {
    const int s = k % LN_STAGES;
    ptx::mbar_wait(&bars[s], (k / LN_STAGES) & 1);
    const float4* src = reinterpret_cast<const float4*>(slots + s * ROW_BYTES) + lane * 2;
    float4 v[2 * NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      v[2 * i] = src[i * 64];
      v[2 * i + 1] = src[i * 64 + 1];
    }
    __syncwarp();
    if (lane == 0) {   // refill the slot with the row LN_STAGES ahead
      const int64_t nrid = rid + (int64_t)LN_STAGES * nwarps;
      if (nrid < total) {
        ptx::mbar_expect_tx(&bars[s], ROW_BYTES);
        ptx::bulk_load(slots + s * ROW_BYTES, row_ptr(nrid), ROW_BYTES, &bars[s]);
      }
    }
    const int b = (int)(rid / p.rows);
    const int r = (int)(rid % p.rows);
    // statistics in one pass with packed fp32x2 adds / fmas (FADD2 / FFMA2): var = E[x^2] - mean^2 in fp32 (the
    // residual stream has |mean| << std, so the cancellation costs ~1e-7 relative)
    uint64_t s2[2] = {ptx::pack2(0.f, 0.f), ptx::pack2(0.f, 0.f)}, q2[2] = {s2[0], s2[0]};
#pragma unroll
    for (int i = 0; i < 2 * NC; ++i) {
      const uint64_t lo = ptx::pack2(v[i].x, v[i].y), hi = ptx::pack2(v[i].z, v[i].w);
      s2[0] = ptx::add2(s2[0], lo);
      s2[1] = ptx::add2(s2[1], hi);
      q2[0] = ptx::fma2(lo, lo, q2[0]);
      q2[1] = ptx::fma2(hi, hi, q2[1]);
    }
    float sa, sb, qa, qb;
    ptx::unpack2(ptx::add2(s2[0], s2[1]), sa, sb);
    ptx::unpack2(ptx::add2(q2[0], q2[1]), qa, qb);
    float sum = sa + sb, sq = qa + qb;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    const float mean = sum * inv_dim;
    const float rstd = rsqrtf(fmaxf(sq * inv_dim - mean * mean, 0.f) + p.eps);
    const uint64_t rstd2 = ptx::pack2(rstd, rstd), nmean2 = ptx::pack2(-mean, -mean);
#pragma unroll
    for (int which = 0; which < (DUAL ? 2 : 1); ++which) {
      const float4* sh = reinterpret_cast<const float4*>((which ? p.shift2 : p.shift) + (int64_t)b * p.mod_batch_stride) + lane * 2;
      const float4* sc = reinterpret_cast<const float4*>((which ? p.scale2 : p.scale) + (int64_t)b * p.mod_batch_stride) + lane * 2;
      uint4* out = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(which ? p.out2 : p.out) +
                                            (int64_t)b * p.o_batch_stride + (int64_t)r * p.ldo) + lane;
      // out = (x - mean) rstd (1 + s) + t = x a + (t - mean a),  a = rstd s + rstd: three packed fmas per two elements
      auto mod2 = [&](float x0, float x1, float sc0, float sc1, float t0, float t1) {
        const uint64_t a2 = ptx::fma2(ptx::pack2(sc0, sc1), rstd2, rstd2);
        const uint64_t c2 = ptx::fma2(nmean2, a2, ptx::pack2(t0, t1));
        float o0, o1;
        ptx::unpack2(ptx::fma2(ptx::pack2(x0, x1), a2, c2), o0, o1);
        return ptx::pack_bf16(o0, o1);
      };
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const float4 s0 = __ldg(sc + i * 64), s1 = __ldg(sc + i * 64 + 1);
        const float4 t0 = __ldg(sh + i * 64), t1 = __ldg(sh + i * 64 + 1);
        const float4 a = v[2 * i], c = v[2 * i + 1];
        uint4 o;
        o.x = mod2(a.x, a.y, s0.x, s0.y, t0.x, t0.y);
        o.y = mod2(a.z, a.w, s0.z, s0.w, t0.z, t0.w);
        o.z = mod2(c.x, c.y, s1.x, s1.y, t1.x, t1.y);
        o.w = mod2(c.z, c.w, s1.z, s1.w, t1.z, t1.w);
        out[i * 32] = o;
      }
    }
  }
}

// Second form of the TMA-staged kernel (default since round 2b).  ncu of the first form (profiles/r02b_ncu_full.json): DRAM
// 45 % busy with l1tex at 61 %: per 6 KB row the SM moved 6 KB (TMA fill) + 12 KB (row read back with 32-byte lane stride:
// two-way bank conflicts) + 12 KB (scale / shift re-read from L1 for EVERY row) + 3 KB (stores) through its L1.  Here
//  * a lane owns float4 number i*32 + lane of the row (conflict-free 16-byte shared-memory reads, coalesced 8-byte stores),
//  * the single-output instance keeps (1 + scale) and shift of the CURRENT batch in registers (96 per lane for dim 1536)
//    and reloads them only when a row of another batch arrives (rows are visited in batch order), so the steady state
//    touches L1 for the row itself only.  The dual-output instance has no registers left for that and reads the four
//    vectors per row as before (coalesced).
template <int NV, bool DUAL>  // NV float4 per lane => dim = NV * 128
__global__ void __launch_bounds__(LN_WARPS * 32, 1)
ln_modulate_tma2_kernel(const dit4sr_ln_modulate_params p, const int reverse) {
  extern __shared__ uint8_t ln_smem_raw[];
  constexpr int ROW_BYTES = NV * 128 * 4;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ln_smem_raw) + 127) & ~uintptr_t(127));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* slots = smem + (size_t)warp * LN_STAGES * ROW_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)LN_WARPS * LN_STAGES * ROW_BYTES) + warp * LN_STAGES;
  // rows [0, total_a) are the main problem, [total_a, total) the companion (ABI v5: the text stream's rows ride in the
  // image stream's launch; always single-output)
  const int64_t total_a = (int64_t)p.batches * p.rows;
  const int64_t total = total_a + (int64_t)p.batches * p.rows_b;
  const int64_t nwarps = (int64_t)gridDim.x * LN_WARPS;
  const int64_t first = (int64_t)blockIdx.x * LN_WARPS + warp;
  // reverse: rows are visited from the last one down - the kernel that produced x wrote them in ascending order, so the
  // rows written last are the ones most likely still in L2
  auto row_ptr = [&](int64_t rid) {
    if (reverse) rid = total - 1 - rid;
    if (rid < total_a) {
      const int b = (int)(rid / p.rows);
      const int r = (int)(rid % p.rows);
      return p.x + (int64_t)b * p.x_batch_stride + (int64_t)r * p.ldx;
    }
    rid -= total_a;
    const int b = (int)(rid / p.rows_b);
    const int r = (int)(rid % p.rows_b);
    return p.x_b + (int64_t)b * p.x_b_batch_stride + (int64_t)r * p.ldx_b;
  };
  if (lane == 0) {
    for (int s = 0; s < LN_STAGES; ++s) ptx::mbar_init(&bars[s], 1);
    ptx::fence_barrier_init();
  }
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  if (lane == 0) {
    for (int s = 0; s < LN_STAGES; ++s) {
      const int64_t rid = first + s * nwarps;
      if (rid < total) {
        ptx::mbar_expect_tx(&bars[s], ROW_BYTES);
        ptx::bulk_load(slots + s * ROW_BYTES, row_ptr(rid), ROW_BYTES, &bars[s]);
      }
    }
  }
  __syncwarp();
  const float inv_dim = 1.0f / (float)(NV * 128);
  float4 ca[DUAL ? 1 : NV], ct[DUAL ? 1 : NV];   // (1 + scale), shift of batch `cached_b` (single-output instance)
  int cached_b = -1;                             // companion rows use the keys batches + b
  int k = 0;
  for (int64_t rid = first; rid < total; rid += nwarps, ++k) {
    const int s = k % LN_STAGES;
    const int64_t row = reverse ? total - 1 - rid : rid;
    const bool comp = row >= total_a;
    const int64_t lrid = comp ? row - total_a : row;
    const int prow = comp ? p.rows_b : p.rows;
    const int b = (int)(lrid / prow);
    const int r = (int)(lrid % prow);
    const float* shift_v = (comp ? p.shift_b : p.shift) + (int64_t)b * p.mod_batch_stride;
    const float* scale_v = (comp ? p.scale_b : p.scale) + (int64_t)b * p.mod_batch_stride;
    __nv_bfloat16* out_row = comp ? reinterpret_cast<__nv_bfloat16*>(p.out_b) + (int64_t)b * p.o_b_batch_stride + (int64_t)r * p.ldo_b
                                  : reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)b * p.o_batch_stride + (int64_t)r * p.ldo;
    if constexpr (!DUAL) {
      const int key = comp ? p.batches + b : b;
      if (key != cached_b) {
        const float4* sh = reinterpret_cast<const float4*>(shift_v) + lane;
        const float4* sc = reinterpret_cast<const float4*>(scale_v) + lane;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float4 a = __ldg(sc + i * 32);
          ca[i] = make_float4(1.f + a.x, 1.f + a.y, 1.f + a.z, 1.f + a.w);
          ct[i] = __ldg(sh + i * 32);
        }
        cached_b = key;
      }
    }
    ptx::mbar_wait(&bars[s], (k / LN_STAGES) & 1);
    const float4* src = reinterpret_cast<const float4*>(slots + s * ROW_BYTES) + lane;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = src[i * 32];
    __syncwarp();
    if (lane == 0) {   // refill the slot with the row LN_STAGES ahead
      const int64_t nrid = rid + (int64_t)LN_STAGES * nwarps;
      if (nrid < total) {
        ptx::mbar_expect_tx(&bars[s], ROW_BYTES);
        ptx::bulk_load(slots + s * ROW_BYTES, row_ptr(nrid), ROW_BYTES, &bars[s]);
      }
    }
    // one-pass statistics in packed fp32x2 (see the first form)
    uint64_t s2[2] = {ptx::pack2(0.f, 0.f), ptx::pack2(0.f, 0.f)}, q2[2] = {s2[0], s2[0]};
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint64_t lo = ptx::pack2(v[i].x, v[i].y), hi = ptx::pack2(v[i].z, v[i].w);
      s2[0] = ptx::add2(s2[0], lo);
      s2[1] = ptx::add2(s2[1], hi);
      q2[0] = ptx::fma2(lo, lo, q2[0]);
      q2[1] = ptx::fma2(hi, hi, q2[1]);
    }
    float sa, sb, qa, qb;
    ptx::unpack2(ptx::add2(s2[0], s2[1]), sa, sb);
    ptx::unpack2(ptx::add2(q2[0], q2[1]), qa, qb);
    float sum = sa + sb, sq = qa + qb;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    const float mean = sum * inv_dim;
    const float rstd = rsqrtf(fmaxf(sq * inv_dim - mean * mean, 0.f) + p.eps);
    const uint64_t rstd2 = ptx::pack2(rstd, rstd), nmean2 = ptx::pack2(-mean, -mean);
    // out = (x - mean) rstd (1 + s) + t = x a + (t - mean a),  a = rstd (1 + s)
    auto mod2 = [&](float x0, float x1, uint64_t a1_2, float t0, float t1) {
      uint64_t a2;
      asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(a2) : "l"(a1_2), "l"(rstd2));
      const uint64_t c2 = ptx::fma2(nmean2, a2, ptx::pack2(t0, t1));
      float o0, o1;
      ptx::unpack2(ptx::fma2(ptx::pack2(x0, x1), a2, c2), o0, o1);
      return ptx::pack_bf16(o0, o1);
    };
    if constexpr (!DUAL) {
      uint2* out = reinterpret_cast<uint2*>(out_row) + lane;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        uint2 o;
        o.x = mod2(v[i].x, v[i].y, ptx::pack2(ca[i].x, ca[i].y), ct[i].x, ct[i].y);
        o.y = mod2(v[i].z, v[i].w, ptx::pack2(ca[i].z, ca[i].w), ct[i].z, ct[i].w);
        out[i * 32] = o;
      }
    } else {
      const int n_out = comp ? 1 : 2;
      for (int which = 0; which < n_out; ++which) {
        const float4* sh = reinterpret_cast<const float4*>(which ? p.shift2 + (int64_t)b * p.mod_batch_stride : shift_v) + lane;
        const float4* sc = reinterpret_cast<const float4*>(which ? p.scale2 + (int64_t)b * p.mod_batch_stride : scale_v) + lane;
        uint2* out = reinterpret_cast<uint2*>(which ? reinterpret_cast<__nv_bfloat16*>(p.out2) + (int64_t)b * p.o_batch_stride +
                                                          (int64_t)r * p.ldo
                                                    : out_row) + lane;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float4 a = __ldg(sc + i * 32), t = __ldg(sh + i * 32);
          uint2 o;
          o.x = mod2(v[i].x, v[i].y, ptx::pack2(1.f + a.x, 1.f + a.y), t.x, t.y);
          o.y = mod2(v[i].z, v[i].w, ptx::pack2(1.f + a.z, 1.f + a.w), t.z, t.w);
          out[i * 32] = o;
        }
      }
    }
  }
}

// W warps per CTA; ONEPASS: sum and sum of squares reduced together (one shuffle tree instead of two in sequence)
template <int NV, bool DUAL, int W, bool ONEPASS>  // NV float4 per lane => dim = NV * 128
__global__ void __launch_bounds__(W * 32)
ln_modulate_kernel(const dit4sr_ln_modulate_params p) {
  const int lane = threadIdx.x & 31;
  const int64_t row_id = (int64_t)blockIdx.x * W + (threadIdx.x >> 5);
  const int64_t total = (int64_t)p.batches * p.rows;
  if (row_id >= total) return;
  const int b = (int)(row_id / p.rows);
  const int r = (int)(row_id % p.rows);
  const float4* x = reinterpret_cast<const float4*>(p.x + (int64_t)b * p.x_batch_stride + (int64_t)r * p.ldx);
  float4 v[NV];
  float sum = 0.f, sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = __ldg(x + i * 32 + lane);
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    if (ONEPASS) sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float inv_dim = 1.0f / (float)(NV * 128);
  float mean, rstd;
  if (ONEPASS) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    mean = sum * inv_dim;
    rstd = rsqrtf(fmaxf(sq * inv_dim - mean * mean, 0.f) + p.eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) { v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean; }
  } else {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mean = sum * inv_dim;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    rstd = rsqrtf(sq * inv_dim + p.eps);
  }

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

template <int NV, int W, bool ONEPASS>
static void launch_bf16(const dit4sr_ln_modulate_params& p, cudaStream_t stream) {
  const int64_t total = (int64_t)p.batches * p.rows;
  const unsigned grid = (unsigned)((total + W - 1) / W);
  if (p.out2 != nullptr) ln_modulate_kernel<NV, true, W, ONEPASS><<<grid, W * 32, 0, stream>>>(p);
  else ln_modulate_kernel<NV, false, W, ONEPASS><<<grid, W * 32, 0, stream>>>(p);
}

template <int NV>
static int launch(const dit4sr_ln_modulate_params& p, cudaStream_t stream) {
  const int64_t total = (int64_t)p.batches * p.rows + (int64_t)p.batches * p.rows_b;
  static int variant = -1;   // developer switch DIT4SR_LN_VARIANT: warps per CTA (4|8) + 10 for one-pass statistics
  if (variant < 0) {
    const char* e = getenv("DIT4SR_LN_VARIANT");
    variant = e ? atoi(e) : 9;   // 9 = TMA-staged kernel, second form (default); 8 = first form; 1 = register kernel
  }
  {
    constexpr int SMEM2 = LN_WARPS * LN_STAGES * NV * 128 * 4 + LN_WARPS * LN_STAGES * 8 + 128;
    const bool comp_ok = p.rows_b == 0 || ((reinterpret_cast<uintptr_t>(p.out_b) & 7) == 0 && p.ldo_b % 4 == 0 &&
                                           p.o_b_batch_stride % 4 == 0);
    if (!p.out_f32 && variant == 9 && SMEM2 <= 220 * 1024 && (reinterpret_cast<uintptr_t>(p.out) & 7) == 0 && comp_ok &&
        p.ldo % 4 == 0 && p.o_batch_stride % 4 == 0 && (p.out2 == nullptr || (reinterpret_cast<uintptr_t>(p.out2) & 7) == 0)) {
      static unsigned long long configured2 = 0;   // per (instance, device)
      if (host::first_use_on_device(&configured2)) {
        DIT4SR_CUDA(cudaFuncSetAttribute(ln_modulate_tma2_kernel<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2));
        DIT4SR_CUDA(cudaFuncSetAttribute(ln_modulate_tma2_kernel<NV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2));
      }
      unsigned grid = (unsigned)((total + LN_WARPS - 1) / LN_WARPS);
      const unsigned sms = (unsigned)host::sm_count();
      if (grid > sms) grid = sms;
      static int reverse = -1;   // DIT4SR_LN_REVERSE=0: visit the rows in ascending order
      if (reverse < 0) { const char* e = getenv("DIT4SR_LN_REVERSE"); reverse = (e != nullptr && e[0] == '0') ? 0 : 1; }
      if (p.out2 != nullptr)
        DIT4SR_CUDA(host::launch_pdl(ln_modulate_tma2_kernel<NV, true>, dim3(grid), dim3(LN_WARPS * 32), (size_t)SMEM2, stream, p, reverse));
      else
        DIT4SR_CUDA(host::launch_pdl(ln_modulate_tma2_kernel<NV, false>, dim3(grid), dim3(LN_WARPS * 32), (size_t)SMEM2, stream, p, reverse));
      return 0;
    }
  }
  if (p.rows_b > 0) {   // the other kernels serve one problem per launch: main problem, then the companion as its own
    dit4sr_ln_modulate_params a = p, c = p;
    a.rows_b = 0;
    c.rows_b = 0;
    c.x = p.x_b; c.ldx = p.ldx_b; c.x_batch_stride = p.x_b_batch_stride;
    c.shift = p.shift_b; c.scale = p.scale_b; c.shift2 = nullptr; c.scale2 = nullptr;
    c.out = p.out_b; c.out2 = nullptr; c.ldo = p.ldo_b; c.o_batch_stride = p.o_b_batch_stride;
    c.rows = p.rows_b;
    const int rc = launch<NV>(a, stream);
    return rc != 0 ? rc : launch<NV>(c, stream);
  }
  if constexpr (NV % 2 == 0) {
    constexpr int SMEM = LN_WARPS * LN_STAGES * NV * 128 * 4 + LN_WARPS * LN_STAGES * 8 + 128;
    if (!p.out_f32 && (variant == 8 || variant == 9) && SMEM <= 220 * 1024 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 &&
        p.ldo % 8 == 0 && p.o_batch_stride % 8 == 0 && (variant == 8 || variant == 9) &&
        (p.out2 == nullptr || (reinterpret_cast<uintptr_t>(p.out2) & 15) == 0)) {
      static unsigned long long configured = 0;   // per (instance, device)
      if (host::first_use_on_device(&configured)) {
        DIT4SR_CUDA(cudaFuncSetAttribute(ln_modulate_tma_kernel<NV / 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        DIT4SR_CUDA(cudaFuncSetAttribute(ln_modulate_tma_kernel<NV / 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      }
      unsigned grid = (unsigned)((total + LN_WARPS - 1) / LN_WARPS);
      const unsigned sms = (unsigned)host::sm_count();
      if (grid > sms) grid = sms;
      if (p.out2 != nullptr)
        DIT4SR_CUDA(host::launch_pdl(ln_modulate_tma_kernel<NV / 2, true>, dim3(grid), dim3(LN_WARPS * 32), (size_t)SMEM, stream, p));
      else
        DIT4SR_CUDA(host::launch_pdl(ln_modulate_tma_kernel<NV / 2, false>, dim3(grid), dim3(LN_WARPS * 32), (size_t)SMEM, stream, p));
      return 0;
    }
  }
  if (p.out_f32) {
    const unsigned grid = (unsigned)((total + WARPS - 1) / WARPS);
    ln_modulate_f32_kernel<NV><<<grid, WARPS * 32, 0, stream>>>(p);
  } else if (variant == 4) launch_bf16<NV, 4, false>(p, stream);
  else if (variant == 14) launch_bf16<NV, 4, true>(p, stream);
  else if (variant == 18) launch_bf16<NV, 8, true>(p, stream);
  else if (variant == 2) launch_bf16<NV, 2, false>(p, stream);
  else if (variant == 12) launch_bf16<NV, 2, true>(p, stream);
  else launch_bf16<NV, 8, false>(p, stream);
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
  if (p->rows_b > 0) {   // companion problem (ABI v5)
    DIT4SR_REQUIRE(p->x_b && p->shift_b && p->scale_b && p->out_b, DIT4SR_E_BADARG,
                   "ln_modulate: companion problem (rows_b > 0) with a null pointer");
    DIT4SR_REQUIRE(p->ldx_b % 4 == 0 && p->x_b_batch_stride % 4 == 0 && p->ldo_b % 4 == 0 && p->o_b_batch_stride % 4 == 0 &&
                       host::aligned16(p->x_b) && host::aligned16(p->shift_b) && host::aligned16(p->scale_b) &&
                       (reinterpret_cast<uintptr_t>(p->out_b) & 7) == 0,
                   DIT4SR_E_ALIGN, "ln_modulate: companion pointers/strides not 16-byte aligned");
  } else {
    DIT4SR_REQUIRE(p->rows_b == 0, DIT4SR_E_BADARG, "ln_modulate: negative rows_b");
  }
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

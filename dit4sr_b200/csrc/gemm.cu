// Grouped batched-row bf16 GEMM on sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (128x256 tiles, fp32
// accumulators in TMEM, double-buffered) -> fused epilogues.  Persistent: one CTA per SM walks the tile list of up to
// MAX_PROBLEMS independent problems (e.g. the HQ, LR and text projections of one attention) in a single launch, which
// removes the per-launch fill/drain bubbles and the wave-quantisation tail of the small problems.
//
// Warp roles (256 threads): warp 0 = TMA producer (1 thread), warp 1 = MMA issuer (1 thread), warp 2 = TMEM
// allocator, warps 4..7 = epilogue (each owns the 32 TMEM lanes of its lane quarter; thread = one output row).
#include "host_util.h"
#include "ptx.cuh"
#include <stdlib.h>

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
constexpr int EPI_ROW_BYTES = 144;                       // 128 B of payload + 16 B pad: conflict-free 16 B accesses
constexpr int RED_BOX_BYTES = 32 * 32 * 4;               // 32 rows x 32 fp32 columns, 128-byte swizzled (TMA reduce source)
constexpr int EPI_STAGE_BYTES = 2 * RED_BOX_BYTES;       // per epilogue warp: two reduce boxes, or one padded staging tile
static_assert(32 * EPI_ROW_BYTES <= EPI_STAGE_BYTES, "staging tile must fit");
// layout: [operand ring][4 epilogue staging areas (1024-byte aligned)][barriers]
constexpr int SMEM_BYTES = 1024 /*align*/ + STAGES * STAGE_BYTES + 4 * EPI_STAGE_BYTES + 256 /*barriers*/;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory per CTA");
constexpr int MAX_PROBLEMS = DIT4SR_GEMM_MAX_GROUP;

struct Problem {
  const float* bias;
  void* C;
  int64_t ldc, c_batch_stride;
  int c_f32;
  int rows, N;
  int m_tiles, n_tiles, k_blocks;
  int tile_begin;  // first global tile index of this problem
  const float* gate;
  int64_t gate_batch_stride;
  const float* resid;
  int64_t ldr, r_batch_stride;
  const float* norm_w0;
  const float* norm_w1;
  int sec_cols;
  float eps;
  int64_t dbl_row_offset;
  int c_reduce;   // GATE_RESID with resid == C: the epilogue adds gate*y into C with TMA reduce-adds (maps.c)
  // Ulysses scatter (dit4sr_gemm_sp): bf16 output heads go straight to the attention-side buffers of the owning ranks
  int sp_world, sp_rank, sp_own_only, sp_D, sp_col0;
  int64_t sp_rows_dst, sp_row0;
  // implicit-GEMM convolution (dit4sr_gemm_params.conv_*): A is an NHWC image read through a 4-D tensor map
  int conv_w, conv_cblocks, conv_kw, conv_org;
  const __nv_bfloat16* resid_b;   // BIAS epilogue, bf16 C: added to the output
  int64_t ld_rb, rb_batch_stride;
};

// Tail splitting: with T tiles on G CTAs the last T mod G tiles would occupy a whole extra wave with most SMs idle
// (e.g. 912 = 6 x 148 + 24 tiles for the N = 1536 projections: 7 waves for 6.16 waves of work).  Those last tiles are
// cut into `split` column slices of BN/split columns (MMA N = 64 or 128, W loaded through a second tensor map with a
// matching box), so the remainder spreads over split x as many CTAs and costs a fraction of a wave.
struct Group {
  int num_problems;
  int num_tiles;      // work items of the launch = split_begin + (tiles - split_begin) * split
  int split_begin;    // first tile (in schedule order) that is cut into column slices
  int split;          // 1 (no splitting), 2 or 4 slices per 256-column tile
  int pair;           // 1: CTA pairs (cta_group::2) work on 256-row x 256-column tiles, m_tiles count 256-row tiles
  Problem p[MAX_PROBLEMS];
  void* sp_peer[DIT4SR_SP_MAX_RANKS];   // peer-mapped [B, rows_dst, 3, D/world] buffers (problems with sp_world > 0)
};

struct Maps {
  CUtensorMap a[MAX_PROBLEMS];
  CUtensorMap w[MAX_PROBLEMS];
  CUtensorMap ws[MAX_PROBLEMS];   // W with a BN/split-row box (split > 1), or a BN/2-row box (pair: each CTA loads half of W)
  CUtensorMap c[MAX_PROBLEMS];    // fp32 C as (N, rows, batches) with a 32x32 box (only valid when c_reduce)
  CUtensorMap wt[MAX_PROBLEMS];   // pair + split > 1: W with a BN/(2 split)-row box (half of a tail slice per CTA)
};

struct TileCoord {
  int pi, b, m_tile;
  int col0;   // first output column of the work item
  int bn;     // its width: BN, or BN/split for a slice of a tail tile
  int mma_n;  // MMA N of the item: bn, or what is left of a problem whose N ends inside the tile (N = 128 / 32 layers of
              // the VAE: an MMA over 256 columns would spend 2-8x the tensor time on zero padding)
};

__host__ __device__ __forceinline__ TileCoord decode_tile(const Group& g, int item) {
  int tile = item, part = 0, bn = BN;
  if (item >= g.split_begin) {
    const int idx = item - g.split_begin;
    tile = g.split_begin + idx / g.split;
    part = idx % g.split;
    bn = BN / g.split;
  }
  int pi = 0;
#pragma unroll
  for (int i = 1; i < MAX_PROBLEMS; ++i)
    if (i < g.num_problems && tile >= g.p[i].tile_begin) pi = i;
  const Problem& p = g.p[pi];
  const int local = tile - p.tile_begin;
  TileCoord c;
  c.pi = pi;
  c.col0 = (local % p.n_tiles) * BN + part * bn;
  c.bn = bn;
  c.mma_n = min(bn, p.N - c.col0);   // N % 32 == 0 (checked at the entry point): a legal tcgen05 N for M = 128
  const int rest = local / p.n_tiles;
  c.m_tile = rest % p.m_tiles;
  c.b = rest / p.m_tiles;
  return c;
}

__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5 x (1 + tanh( sqrt(2/pi) (x + 0.044715 x^3) ))
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + ptx::tanh_approx(inner));
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }
// full-precision variants for fp32 outputs (the fp32-accurate mode)
__device__ __forceinline__ float gelu_tanh_precise(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.0f + tanhf(k0 * (x + k1 * x * x * x)));
}
__device__ __forceinline__ float silu_precise(float x) { return x / (1.0f + expf(-x)); }

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

// RED (GATE_RESID only): every problem of the launch is an in-place gated residual and the epilogue adds into C with TMA
// reduce-adds (Problem::c_reduce); a separate instance so that neither epilogue pays for the other's registers.
// PAIR: two CTAs of a cluster (one TPC) run ONE 256-row x 256-column tile with tcgen05.mma.cta_group::2: each CTA
// loads its own 128 rows of A and HALF of the W tile, the leader (cluster rank 0) issues M = 256 MMAs that read A from
// each CTA's shared memory and W from both, and each CTA's TMEM receives the accumulators of its 128 rows.  Per SM the
// shared-memory operand traffic falls from 96 to 64 bytes per clock and the L2 -> SM traffic of W is halved.
//   both CTAs' TMA loads (cta_group::2 form) complete on the LEADER's full_bar, which expects the bytes of the pair;
//   the leader's commits are multicast to both CTAs' empty_bar / tfull_bar;  both CTAs' epilogue warps arrive on the
//   leader's tempty_bar (8 arrivals, the non-leader's through remote mbarrier arrives).
template <int EPI, bool RED = false, bool PAIR = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Group grp) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // operand ring: 4 stages of (A 16 KB + W 32 KB), or for CTA pairs 6 stages of (A 16 KB + half of W 16 KB) - the same
  // 192 KB; the pair needs the deeper ring because a stage is only consumable after the peer's relay hop
  constexpr int NST = PAIR ? 6 : STAGES;
  constexpr int STG = PAIR ? (A_BYTES + B_BYTES / 2) : STAGE_BYTES;
  static_assert(NST * STG == STAGES * STAGE_BYTES, "ring size");
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + 4 * EPI_STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [NST]
  uint64_t* empty_bar = bars + NST;                // [NST]
  uint64_t* tfull_bar = bars + 2 * NST;            // [ACC_STAGES]
  uint64_t* tempty_bar = tfull_bar + ACC_STAGES;   // [ACC_STAGES]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = PAIR ? (int)ptx::cluster_ctarank() : 0;          // position in the CTA pair
  const int cta_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int cta_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int BMT = PAIR ? 2 * BM : BM;                            // rows of a work item

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < grp.num_problems; ++i) {
      ptx::prefetch_tmap(&maps.a[i]);
      ptx::prefetch_tmap(&maps.w[i]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NST; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], PAIR ? 8 : 4);  // one arrive per epilogue warp (of both CTAs)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      ptx::tmem_alloc2(tmem_ptr, TMEM_COLS);
      ptx::tmem_relinquish2();
    } else {
      ptx::tmem_alloc(tmem_ptr, TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) ptx::cluster_sync();   // the peer's barriers exist before anything arrives on them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // programmatic dependent launch: the prologue above ran under the previous kernel's tail; nothing of global memory has
  // been touched yet.  From here on the previous kernels are complete, and the next kernel's CTAs may be scheduled.
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();

  const int num_tiles = grp.num_tiles;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cta_first; tile < num_tiles; tile += cta_stride) {
        const TileCoord tc = decode_tile(grp, tile);
        const Problem& PP = grp.p[tc.pi];
        const int k_blocks = PP.k_blocks;
        const CUtensorMap* ma = &maps.a[tc.pi];
        const bool conv = PP.conv_w > 0;
        int cy0 = 0, cx0 = 0;
        if (conv) {   // first output pixel of this CTA's 128 rows (a whole number of image rows, or a 128-pixel run of one)
          const int pix0 = tc.m_tile * BMT + rank * BM;
          cy0 = pix0 / PP.conv_w;
          cx0 = pix0 - cy0 * PP.conv_w;
        }
        const CUtensorMap* mw = PAIR ? (tc.bn == BN ? &maps.ws[tc.pi] : &maps.wt[tc.pi])
                                     : (tc.bn == BN ? &maps.w[tc.pi] : &maps.ws[tc.pi]);
        const int w_rows = PAIR ? tc.bn / 2 : tc.bn;                  // W rows this CTA loads
        const uint32_t tx_bytes = A_BYTES + w_rows * BK * 2;
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STG;
          uint8_t* sb = sa + A_BYTES;
          if constexpr (PAIR) {
            // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of the pair
            if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * tx_bytes);
            const uint32_t lead_bar = ptx::mapa_u32(&full_bar[stage], 0);
            if (conv) {
              const int tap = kb / PP.conv_cblocks, cb = kb - tap * PP.conv_cblocks, ty = tap / PP.conv_kw;
              ptx::tma_load_4d_2sm(sa, ma, lead_bar, cb * BK, cx0 + (tap - ty * PP.conv_kw) + PP.conv_org,
                                   cy0 + ty + PP.conv_org, tc.b);
            } else {
              ptx::tma_load_3d_2sm(sa, ma, lead_bar, kb * BK, tc.m_tile * BMT + rank * BM, tc.b);
            }
            ptx::tma_load_2d_2sm(sb, mw, lead_bar, kb * BK, tc.col0 + rank * w_rows);
          } else {
            ptx::mbar_expect_tx(&full_bar[stage], tx_bytes);
            if (conv) {
              const int tap = kb / PP.conv_cblocks, cb = kb - tap * PP.conv_cblocks, ty = tap / PP.conv_kw;
              ptx::tma_load_4d(sa, ma, &full_bar[stage], cb * BK, cx0 + (tap - ty * PP.conv_kw) + PP.conv_org,
                               cy0 + ty + PP.conv_org, tc.b);
            } else {
              ptx::tma_load_3d(sa, ma, &full_bar[stage], kb * BK, tc.m_tile * BM, tc.b);
            }
            ptx::tma_load_2d(sb, mw, &full_bar[stage], kb * BK, tc.col0);
          }
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (PAIR: the leader CTA only) =====================
    // The whole warp runs the loop (warp-uniform barrier addresses / descriptors); one elected lane issues.
    const uint64_t desc_base = ptx::make_sdesc_sw128(ptx::smem_u32(smem));
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = cta_first; tile < num_tiles; tile += cta_stride, ++it) {
      const TileCoord tc = decode_tile(grp, tile);
      const int k_blocks = grp.p[tc.pi].k_blocks;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const uint32_t idesc = ptx::make_idesc_bf16(BMT, tc.mma_n, 0, 0);   // N = 256, 128 / 64 for a tail slice, or a ragged end
      ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < k_blocks; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint64_t adesc = desc_base + (uint64_t)(stage * (STG >> 4));
        const uint64_t bdesc = adesc + (uint64_t)(A_BYTES >> 4);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128B swizzle atom: +2 in the (addr>>4) field
            if constexpr (PAIR) ptx::mma_ss2(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            else ptx::mma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          // frees the smem slot (of both CTAs) when these MMAs have read it
          if constexpr (PAIR) ptx::tc_commit_mc(&empty_bar[stage], 3);
          else ptx::tc_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
      if (ptx::elect_one()) {   // accumulator complete
        if constexpr (PAIR) ptx::tc_commit_mc(&tfull_bar[as], 3);
        else ptx::tc_commit(&tfull_bar[as]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // The accumulator row of a thread (TMEM lane) gets bias / activation / RMSNorm applied in registers, then goes
    // through a per-warp shared-memory staging tile so that every global access of the warp is a few whole contiguous
    // row segments (128 B per row per instruction) instead of 32 scattered 16-byte pieces.
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint8_t* stg = smem + STAGES * STAGE_BYTES + q * EPI_STAGE_BYTES;
    const int sub = lane >> 3, seg = lane & 7;
    int it = 0;
    for (int tile = cta_first; tile < num_tiles; tile += cta_stride, ++it) {
      const TileCoord tc = decode_tile(grp, tile);
      const Problem& P = grp.p[tc.pi];
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;

      const int row_base = tc.m_tile * BMT + rank * BM + q * 32;       // first row of this warp
      const uint32_t taddr = tmem_base + as * BN + ((uint32_t)(q * 32) << 16);
      const int n_base = tc.col0;
      const int n_valid = min(tc.bn, P.N - n_base);       // multiple of 32
      const int64_t c_tile_off = (int64_t)tc.b * P.c_batch_stride + n_base;

      if (EPI == DIT4SR_EPI_QKV_NORM || !P.c_f32) {
        // ---------------- bf16 outputs: 64 columns (one head) per pass ----------------
        __nv_bfloat16* Cb = reinterpret_cast<__nv_bfloat16*>(P.C) + c_tile_off;
        ptx::mbar_wait(&tfull_bar[as], aphase);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int pass = 0; pass < BN / 64; ++pass) {
          const int c0 = pass * 64;                      // column offset inside the tile
          if (c0 < n_valid) {
            const int cols = min(64, n_valid - c0);      // 32 or 64
            // where this head goes: the local C, or (sequence parallelism) the q|k|v buffer of the rank owning the head
            __nv_bfloat16* dst_base = Cb + c0;
            int64_t dst_ld = P.ldc;
            bool dst_skip = false;
            if (P.sp_world > 0) {
              const int col = P.sp_col0 + n_base + c0;
              const int sec = col / P.sp_D, within = col - sec * P.sp_D;
              const int Dl = P.sp_D / P.sp_world;
              const int owner = within / Dl;
              dst_skip = P.sp_own_only != 0 && owner != P.sp_rank;     // replicated rows: keep my own heads only
              dst_base = reinterpret_cast<__nv_bfloat16*>(grp.sp_peer[owner]) +
                         (((int64_t)tc.b * P.sp_rows_dst + P.sp_row0) * 3 + sec) * Dl + (within - owner * Dl);
              dst_ld = 3 * (int64_t)Dl;
            }
            uint32_t v[64];
            ptx::tmem_ld32(taddr + c0, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
            ptx::tmem_ld32(taddr + c0 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
            ptx::tmem_wait_ld();
            float f[64];
#pragma unroll
            for (int j4 = 0; j4 < 16; ++j4) {
              float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
              if (P.bias != nullptr && 4 * j4 < cols) bb = __ldg(reinterpret_cast<const float4*>(P.bias + n_base + c0) + j4);
              f[4 * j4 + 0] = __uint_as_float(v[4 * j4 + 0]) + bb.x;
              f[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) + bb.y;
              f[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) + bb.z;
              f[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) + bb.w;
            }
            bool dbl = false;
            const float* w = nullptr;
            if constexpr (EPI == DIT4SR_EPI_BIAS_GELU) {
#pragma unroll
              for (int j = 0; j < 64; ++j) f[j] = gelu_tanh(f[j]);
            } else if constexpr (EPI == DIT4SR_EPI_BIAS_SILU) {
#pragma unroll
              for (int j = 0; j < 64; ++j) f[j] = silu(f[j]);
            } else if constexpr (EPI == DIT4SR_EPI_QKV_NORM) {
              const int sec = (n_base + c0) / P.sec_cols;
              w = sec == 0 ? P.norm_w0 : (sec == 1 ? P.norm_w1 : nullptr);
              dbl = sec == 1 && P.dbl_row_offset != 0 && w != nullptr;
            }
#pragma unroll 1
            for (int rep = 0; rep < 2; ++rep) {          // rep 1 = the second normalisation of the LR keys (quirk Q1)
              if (rep == 1 && !dbl) break;
              if (w != nullptr) {
                float ss0 = 0.f, ss1 = 0.f, ss2 = 0.f, ss3 = 0.f;
#pragma unroll
                for (int j = 0; j < 64; j += 4) {
                  ss0 = fmaf(f[j], f[j], ss0);
                  ss1 = fmaf(f[j + 1], f[j + 1], ss1);
                  ss2 = fmaf(f[j + 2], f[j + 2], ss2);
                  ss3 = fmaf(f[j + 3], f[j + 3], ss3);
                }
                const float inv = rsqrtf(((ss0 + ss1) + (ss2 + ss3)) * (1.0f / 64.0f) + P.eps);
#pragma unroll
                for (int j4 = 0; j4 < 16; ++j4) {
                  const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + j4);
                  f[4 * j4 + 0] *= inv * ww.x;
                  f[4 * j4 + 1] *= inv * ww.y;
                  f[4 * j4 + 2] *= inv * ww.z;
                  f[4 * j4 + 3] *= inv * ww.w;
                }
              }
              const int64_t row_shift = rep == 1 ? P.dbl_row_offset : 0;
              if (P.c_f32) {
                // fp32 output (accuracy mode): two staged passes of 32 floats per row
                float* Cf = reinterpret_cast<float*>(P.C) + c_tile_off;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                  if (hf * 32 < cols) {
                    __syncwarp();
                    float4* wr4 = reinterpret_cast<float4*>(stg + lane * EPI_ROW_BYTES);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                      wr4[i] = make_float4(f[hf * 32 + 4 * i], f[hf * 32 + 4 * i + 1], f[hf * 32 + 4 * i + 2],
                                           f[hf * 32 + 4 * i + 3]);
                    __syncwarp();
#pragma unroll
                    for (int r4 = 0; r4 < 8; ++r4) {
                      const int rl = r4 * 4 + sub;
                      const float4 a4 = *reinterpret_cast<const float4*>(stg + rl * EPI_ROW_BYTES + seg * 16);
                      const int rg = row_base + rl;
                      if (rg < P.rows)
                        *reinterpret_cast<float4*>(Cf + (rg + row_shift) * P.ldc + c0 + hf * 32 + seg * 4) = a4;
                    }
                  }
                }
                continue;
              }
              // stage this thread's row (64 bf16 = 128 B), then store 4 whole rows per warp instruction
              __syncwarp();
              uint4* wr = reinterpret_cast<uint4*>(stg + lane * EPI_ROW_BYTES);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                uint4 u;
                u.x = ptx::pack_bf16(f[8 * i + 0], f[8 * i + 1]);
                u.y = ptx::pack_bf16(f[8 * i + 2], f[8 * i + 3]);
                u.z = ptx::pack_bf16(f[8 * i + 4], f[8 * i + 5]);
                u.w = ptx::pack_bf16(f[8 * i + 6], f[8 * i + 7]);
                wr[i] = u;
              }
              __syncwarp();
              if (seg * 8 < cols && !dst_skip) {
                const __nv_bfloat16* rb = nullptr;
                if constexpr (EPI == DIT4SR_EPI_BIAS) {
                  if (P.resid_b != nullptr) rb = P.resid_b + (int64_t)tc.b * P.rb_batch_stride + n_base + c0 + seg * 8;
                }
#pragma unroll
                for (int r4 = 0; r4 < 8; ++r4) {
                  const int rl = r4 * 4 + sub;
                  uint4 u = *reinterpret_cast<const uint4*>(stg + rl * EPI_ROW_BYTES + seg * 16);
                  const int rg = row_base + rl;
                  if (rg < P.rows) {
                    if constexpr (EPI == DIT4SR_EPI_BIAS) {
                      if (rb != nullptr) {   // skip connection (VAE resnet / attention block), added in fp32
                        const uint4 r = *reinterpret_cast<const uint4*>(rb + (int64_t)rg * P.ld_rb);
                        const __nv_bfloat162* ua = reinterpret_cast<const __nv_bfloat162*>(&u);
                        const __nv_bfloat162* ra = reinterpret_cast<const __nv_bfloat162*>(&r);
                        uint32_t o[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                          const float2 x = __bfloat1622float2(ua[k]), y = __bfloat1622float2(ra[k]);
                          o[k] = ptx::pack_bf16(x.x + y.x, x.y + y.y);
                        }
                        u = make_uint4(o[0], o[1], o[2], o[3]);
                      }
                    }
                    *reinterpret_cast<uint4*>(dst_base + (rg + row_shift) * dst_ld + seg * 8) = u;
                  }
                }
              }
            }
          }
        }
      } else {
        // ---------------- fp32 outputs (plain bias, or gate * y + residual): 32 columns per pass ----------------
        float* Cf = reinterpret_cast<float*>(P.C) + c_tile_off;
        const float* Rf = nullptr;
        const float* g_row = nullptr;
        if constexpr (EPI == DIT4SR_EPI_GATE_RESID) {
          Rf = P.resid + (int64_t)tc.b * P.r_batch_stride + n_base;
          g_row = P.gate ? P.gate + (int64_t)tc.b * P.gate_batch_stride + n_base : nullptr;
        }
        if constexpr (EPI == DIT4SR_EPI_GATE_RESID && RED) {
          {
            // ---- in-place gated residual: C += gate * (acc + bias) as TMA reduce-adds of 32x32 fp32 boxes.  The SM
            // never reads the residual (the addition happens at L2) and the stores are asynchronous bulk operations, so
            // this epilogue is a TMEM read, 32 FMAs per row and eight shared-memory stores per pass.
            const CUtensorMap* mc = &maps.c[tc.pi];
            const float* b_row = P.bias != nullptr ? P.bias + n_base : nullptr;
            ptx::mbar_wait(&tfull_bar[as], aphase);
            ptx::tc_fence_after();
            uint32_t v[2][32];
            ptx::tmem_ld32(taddr, v[0]);
#pragma unroll
            for (int ch = 0; ch < BN / 32; ++ch) {
              const int c0 = ch * 32;
              if (c0 < n_valid) {
                ptx::tmem_wait_ld();
                if (c0 + 32 < n_valid) ptx::tmem_ld32(taddr + c0 + 32, v[(ch + 1) & 1]);
                uint8_t* box = stg + (ch & 1) * RED_BOX_BYTES;
                if (lane == 0) ptx::bulk_wait_read<1>();   // the reduce issued from this box two passes ago has read it
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f), gg = make_float4(1.f, 1.f, 1.f, 1.f);
                  if (b_row != nullptr) bb = __ldg(reinterpret_cast<const float4*>(b_row + c0) + i);
                  if (g_row != nullptr) gg = __ldg(reinterpret_cast<const float4*>(g_row + c0) + i);
                  float4 a;
                  a.x = gg.x * (__uint_as_float(v[ch & 1][4 * i + 0]) + bb.x);
                  a.y = gg.y * (__uint_as_float(v[ch & 1][4 * i + 1]) + bb.y);
                  a.z = gg.z * (__uint_as_float(v[ch & 1][4 * i + 2]) + bb.z);
                  a.w = gg.w * (__uint_as_float(v[ch & 1][4 * i + 3]) + bb.w);
                  // 128-byte swizzle: 16-byte chunk i of row r lives at chunk i ^ (r & 7)
                  *reinterpret_cast<float4*>(box + lane * 128 + ((i ^ (lane & 7)) << 4)) = a;
                }
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0 && row_base < P.rows) {
                  ptx::tma_reduce_add_3d(mc, box, n_base + c0, row_base, tc.b);   // rows >= P.rows are clipped
                  ptx::bulk_commit();
                }
              }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR && rank == 1) ptx::mbar_arrive_remote(&tempty_bar[as], 0);
              else ptx::mbar_arrive(&tempty_bar[as]);
            }
            continue;
          }
        }
        if constexpr (RED) continue;   // (not reached: the block above ends the iteration)
        // The residual rows are the only cold global reads of this epilogue and each warp has just one 32-column pass
        // of them in flight, so pull the NEXT tile's residual rows into L2 now (a whole tile ahead of their use).
        if constexpr (EPI == DIT4SR_EPI_GATE_RESID) {
          const int nt = tile + cta_stride;
          if (nt < num_tiles) {
            const TileCoord nc = decode_tile(grp, nt);
            const Problem& NP = grp.p[nc.pi];
            const int nr = nc.m_tile * BMT + rank * BM + q * 32 + lane;
            if (nr < NP.rows) {
              const float* np = NP.resid + (int64_t)nc.b * NP.r_batch_stride + (int64_t)nr * NP.ldr + nc.col0;
              const int ncols = min(nc.bn, NP.N - nc.col0);
#pragma unroll
              for (int i = 0; i < BN / 32; ++i)
                if (i * 32 < ncols) ptx::prefetch_l2(np + i * 32);
            }
          }
        }
        // the residual does not depend on the accumulator: fetch the first pass before waiting for the MMAs
        float4 rv[2][8];
        if constexpr (EPI == DIT4SR_EPI_GATE_RESID) {
#pragma unroll
          for (int r4 = 0; r4 < 8; ++r4) {
            const int rg = row_base + r4 * 4 + sub;
            if (rg < P.rows) rv[0][r4] = *reinterpret_cast<const float4*>(Rf + (int64_t)rg * P.ldr + seg * 4);
          }
        }
        ptx::mbar_wait(&tfull_bar[as], aphase);
        ptx::tc_fence_after();
        uint32_t v[2][32];
        ptx::tmem_ld32(taddr, v[0]);
#pragma unroll
        for (int ch = 0; ch < BN / 32; ++ch) {
          const int c0 = ch * 32;
          if (c0 < n_valid) {
            ptx::tmem_wait_ld();
            if (c0 + 32 < n_valid) {   // software pipeline: next pass's TMEM + residual loads in flight
              ptx::tmem_ld32(taddr + c0 + 32, v[(ch + 1) & 1]);
              if constexpr (EPI == DIT4SR_EPI_GATE_RESID) {
#pragma unroll
                for (int r4 = 0; r4 < 8; ++r4) {
                  const int rg = row_base + r4 * 4 + sub;
                  if (rg < P.rows)
                    rv[(ch + 1) & 1][r4] = *reinterpret_cast<const float4*>(Rf + (int64_t)rg * P.ldr + c0 + 32 + seg * 4);
                }
              }
            }
            __syncwarp();
            float4* wr = reinterpret_cast<float4*>(stg + lane * EPI_ROW_BYTES);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              wr[i] = make_float4(__uint_as_float(v[ch & 1][4 * i]), __uint_as_float(v[ch & 1][4 * i + 1]),
                                  __uint_as_float(v[ch & 1][4 * i + 2]), __uint_as_float(v[ch & 1][4 * i + 3]));
            __syncwarp();
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f), gg = make_float4(1.f, 1.f, 1.f, 1.f);
            if (P.bias != nullptr) bb = __ldg(reinterpret_cast<const float4*>(P.bias + n_base + c0) + seg);
            if constexpr (EPI == DIT4SR_EPI_GATE_RESID) {
              if (g_row != nullptr) gg = __ldg(reinterpret_cast<const float4*>(g_row + c0) + seg);
            }
#pragma unroll
            for (int r4 = 0; r4 < 8; ++r4) {
              const int rl = r4 * 4 + sub;
              const int rg = row_base + rl;
              float4 a = *reinterpret_cast<const float4*>(stg + rl * EPI_ROW_BYTES + seg * 16);
              a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
              if constexpr (EPI == DIT4SR_EPI_BIAS_GELU) {
                a.x = gelu_tanh_precise(a.x); a.y = gelu_tanh_precise(a.y);
                a.z = gelu_tanh_precise(a.z); a.w = gelu_tanh_precise(a.w);
              } else if constexpr (EPI == DIT4SR_EPI_BIAS_SILU) {
                a.x = silu_precise(a.x); a.y = silu_precise(a.y); a.z = silu_precise(a.z); a.w = silu_precise(a.w);
              }
              if constexpr (EPI == DIT4SR_EPI_GATE_RESID) {
                const float4 rr = rv[ch & 1][r4];
                a.x = fmaf(gg.x, a.x, rr.x);
                a.y = fmaf(gg.y, a.y, rr.y);
                a.z = fmaf(gg.z, a.z, rr.z);
                a.w = fmaf(gg.w, a.w, rr.w);
              }
              if (rg < P.rows) *reinterpret_cast<float4*>(Cf + (int64_t)rg * P.ldc + c0 + seg * 4) = a;
            }
          }
        }
      }
      // release the accumulator stage (PAIR: the leader's issuer counts the epilogue warps of both CTAs)
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && rank == 1) ptx::mbar_arrive_remote(&tempty_bar[as], 0);
        else ptx::mbar_arrive(&tempty_bar[as]);
      }
    }
  }

  if (RED && warp >= 4 && lane == 0) ptx::bulk_wait_all();   // outstanding reduce-adds
  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) ptx::cluster_sync();   // nobody leaves while the peer may still read its shared memory / barriers
  if (warp == 2) {
    ptx::tc_fence_after();
    if constexpr (PAIR) ptx::tmem_dealloc2(tmem_base, TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int EPI, bool RED = false>
static int launch(const Maps& maps, const Group& g, cudaStream_t stream) {
  if (g.pair) {   // clusters of two CTAs (one TPC each), one 256x256 tile per pair at a time
    static unsigned long long configured2 = 0;
    if (host::first_use_on_device(&configured2))
      DIT4SR_CUDA(cudaFuncSetAttribute(gemm_kernel<EPI, RED, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    int pairs = host::sm_count() / 2;
    if (pairs > g.num_tiles) pairs = g.num_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DIT4SR_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<EPI, RED, true>, maps, g));
    return 0;
  }
  static unsigned long long configured = 0;   // per (instance, device)
  if (host::first_use_on_device(&configured))
    DIT4SR_CUDA(cudaFuncSetAttribute(gemm_kernel<EPI, RED>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  int grid = host::sm_count();
  if (grid > g.num_tiles) grid = g.num_tiles;
  DIT4SR_CUDA(host::launch_pdl(gemm_kernel<EPI, RED>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, maps, g));
  return 0;
}

static int check_problem(const dit4sr_gemm_params* p, int epilogue) {
  DIT4SR_REQUIRE(p->A && p->W && p->C, DIT4SR_E_BADARG, "gemm: null pointer argument");
  DIT4SR_REQUIRE(p->batches > 0 && p->rows > 0 && p->N > 0 && p->K > 0, DIT4SR_E_BADARG,
                 "gemm: non-positive size (batches %d rows %d N %d K %d)", p->batches, p->rows, p->N, p->K);
  DIT4SR_REQUIRE(p->K % 8 == 0 && p->N % 32 == 0, DIT4SR_E_SHAPE, "gemm: need K %% 8 == 0 and N %% 32 == 0 (K %d N %d)",
                 p->K, p->N);
  DIT4SR_REQUIRE(p->lda % 8 == 0 && p->ldw % 8 == 0 && (p->a_batch_stride % 8 == 0), DIT4SR_E_ALIGN,
                 "gemm: lda/ldw/a_batch_stride must be multiples of 8 elements");
  DIT4SR_REQUIRE(p->ldc % 8 == 0 && p->c_batch_stride % 8 == 0 && host::aligned16(p->C), DIT4SR_E_ALIGN,
                 "gemm: C pointer/strides must be 16-byte aligned");
  DIT4SR_REQUIRE(p->bias == nullptr || host::aligned16(p->bias), DIT4SR_E_ALIGN, "gemm: bias must be 16-byte aligned");
  DIT4SR_REQUIRE(p->epilogue == epilogue, DIT4SR_E_BADARG, "gemm: all problems of a group must share the epilogue");
  if (epilogue == DIT4SR_EPI_QKV_NORM) {
    DIT4SR_REQUIRE(p->N % 64 == 0 && p->sec_cols >= 0 && p->sec_cols % 64 == 0, DIT4SR_E_SHAPE,
                   "gemm/QKV_NORM: N and sec_cols must be multiples of 64");
    DIT4SR_REQUIRE((p->norm_w0 == nullptr || host::aligned16(p->norm_w0)) &&
                       (p->norm_w1 == nullptr || host::aligned16(p->norm_w1)),
                   DIT4SR_E_ALIGN, "gemm/QKV_NORM: norm weights must be 16-byte aligned");
  }
  if (p->conv_w != 0) {
    DIT4SR_REQUIRE(p->conv_w > 0 && p->conv_h > 0 && p->conv_kw > 0 && p->conv_taps > 0 && p->conv_taps <= 16 &&
                       p->K % p->conv_taps == 0 && (p->K / p->conv_taps) % BK == 0 &&
                       p->rows == p->conv_w * p->conv_h && (BM % p->conv_w == 0 || p->conv_w % BM == 0) &&
                       p->conv_w_valid >= 0 && p->conv_w_valid <= p->conv_w,
                   DIT4SR_E_SHAPE, "gemm/conv: need rows == conv_h*conv_w, C_in = K/conv_taps a multiple of 64, conv_w "
                   "dividing 128 or a multiple of 128 (w %d h %d taps %d K %d rows %d)", p->conv_w, p->conv_h, p->conv_taps,
                   p->K, p->rows);
    DIT4SR_REQUIRE(p->sp == nullptr, DIT4SR_E_BADARG, "gemm/conv: no sp scatter");
  }
  if (p->resid_bf16 != nullptr) {
    DIT4SR_REQUIRE(epilogue == DIT4SR_EPI_BIAS && p->c_dtype == DIT4SR_DT_BF16 && host::aligned16(p->resid_bf16) &&
                       p->ld_rb % 8 == 0 && p->rb_batch_stride % 8 == 0 && p->sp == nullptr,
                   DIT4SR_E_BADARG, "gemm: resid_bf16 needs the BIAS epilogue, bf16 C and 16-byte aligned rows");
  }
  if (epilogue == DIT4SR_EPI_GATE_RESID) {
    DIT4SR_REQUIRE(p->resid != nullptr && p->c_dtype == DIT4SR_DT_F32 && p->ldr % 4 == 0 && p->r_batch_stride % 4 == 0 &&
                       host::aligned16(p->resid),
                   DIT4SR_E_BADARG, "gemm/GATE_RESID: needs 16-byte aligned fp32 resid and fp32 C");
    DIT4SR_REQUIRE(p->gate == nullptr || (host::aligned16(p->gate) && p->gate_batch_stride % 4 == 0), DIT4SR_E_ALIGN,
                   "gemm/GATE_RESID: gate must be 16-byte aligned");
  }
  return 0;
}

// Tile schedule of a group: pure host arithmetic shared by the launch path and dit4sr_gemm_schedule (so that the
// schedule the CPU tests enumerate is the one the kernel walks).  Fills the shape fields of g.p[0..count) and the
// work-item counts; does not touch pointers.
static void plan_tiles(const dit4sr_gemm_params* probs, int count, int sms, bool tail_split, Group& g, bool pair = false) {
  g.num_problems = count;
  g.pair = pair ? 1 : 0;
  const int bm = pair ? 2 * BM : BM;
  if (pair) sms /= 2;   // work items are handed to CTA pairs
  int tiles = 0;
  bool whole = true;
  for (int i = 0; i < count; ++i) {
    const dit4sr_gemm_params* p = &probs[i];
    Problem& P = g.p[i];
    P.rows = p->rows;
    P.N = p->N;
    P.m_tiles = (p->rows + bm - 1) / bm;
    P.n_tiles = (p->N + BN - 1) / BN;
    P.k_blocks = (p->K + BK - 1) / BK;
    P.tile_begin = tiles;
    tiles += P.m_tiles * P.n_tiles * p->batches;
    whole = whole && (p->N % BN == 0);
  }
  // tail splitting (see Group): only when every problem's N is a whole number of 256-column tiles
  g.num_tiles = tiles;
  g.split_begin = tiles;
  g.split = 1;
  const int rem = tiles % sms;
  if (tail_split && whole && tiles > sms && rem > 0) {
    const int split = 4 * rem <= sms ? 4 : (2 * rem <= sms ? 2 : 1);
    if (split > 1) {
      g.split = split;
      g.split_begin = tiles - rem;
      g.num_tiles = g.split_begin + rem * split;
    }
  }
}

static bool pair_enabled() {   // DIT4SR_GEMM_PAIR=0: never use CTA pairs (cta_group::2)
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("DIT4SR_GEMM_PAIR"); enabled = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return enabled != 0;
}

static bool tail_split_enabled() {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("DIT4SR_GEMM_TAIL_SPLIT"); enabled = (e == nullptr || e[0] != '0') ? 1 : 0; }
  return enabled != 0;
}

}  // namespace gemm

extern "C" int dit4sr_gemm_bf16_grouped(const dit4sr_gemm_params* probs, int count, void* stream_) {
  using namespace gemm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(probs != nullptr && count >= 1 && count <= MAX_PROBLEMS, DIT4SR_E_BADARG,
                 "gemm: group size must be 1..%d", MAX_PROBLEMS);
  const int epilogue = probs[0].epilogue;
  DIT4SR_REQUIRE(epilogue >= DIT4SR_EPI_BIAS && epilogue <= DIT4SR_EPI_GATE_RESID, DIT4SR_E_BADARG,
                 "gemm: unknown epilogue %d", epilogue);
  int rc;
  for (int i = 0; i < count; ++i)
    if ((rc = check_problem(&probs[i], epilogue)) != 0) return rc;
  if ((rc = dit4sr_check_device()) != 0) return rc;

  Maps maps;
  Group g;
  // CTA pairs (cta_group::2): every problem a whole number of 256-column tiles and enough 256-row tiles to go round
  bool pair = pair_enabled() && host::sm_count() % 2 == 0;
  {
    int64_t pair_tiles = 0;
    for (int i = 0; i < count; ++i) {
      pair = pair && probs[i].N % BN == 0;
      pair_tiles += (int64_t)((probs[i].rows + 2 * BM - 1) / (2 * BM)) * (probs[i].N / BN) * probs[i].batches;
    }
    pair = pair && pair_tiles >= host::sm_count() / 2;
  }
  plan_tiles(probs, count, host::sm_count(), tail_split_enabled(), g, pair);
  bool have_sp = false;
  for (int r = 0; r < DIT4SR_SP_MAX_RANKS; ++r) g.sp_peer[r] = nullptr;
  for (int i = 0; i < count; ++i) {
    const dit4sr_gemm_params* p = &probs[i];
    if (p->conv_w > 0) {   // NHWC image: (channels, x, y, batch); a box = 128 output pixels x 64 input channels
      const int bw = p->conv_w < BM ? p->conv_w : BM;
      const int wv = p->conv_w_valid > 0 ? p->conv_w_valid : p->conv_w;     // pitched image: the map clips at the real width
      uint64_t dims[4] = {(uint64_t)(p->K / p->conv_taps), (uint64_t)wv, (uint64_t)p->conv_h, (uint64_t)p->batches};
      uint64_t strides[3] = {(uint64_t)p->lda * 2, (uint64_t)p->lda * 2 * p->conv_w,
                             (uint64_t)(p->batches > 1 ? p->a_batch_stride : (int64_t)p->rows * p->lda) * 2};
      uint32_t box[4] = {BK, (uint32_t)bw, (uint32_t)(BM / bw), 1};
      if ((rc = host::make_tmap_bf16(&maps.a[i], p->A, 4, dims, strides, box)) != 0) return rc;
    } else {
      uint64_t dims[3] = {(uint64_t)p->K, (uint64_t)p->rows, (uint64_t)p->batches};
      uint64_t strides[2] = {(uint64_t)p->lda * 2,
                             (uint64_t)(p->batches > 1 ? p->a_batch_stride : (int64_t)p->rows * p->lda) * 2};
      uint32_t box[3] = {BK, BM, 1};
      if ((rc = host::make_tmap_bf16(&maps.a[i], p->A, 3, dims, strides, box)) != 0) return rc;
    }
    {
      uint64_t dims[2] = {(uint64_t)p->K, (uint64_t)p->N};
      uint64_t strides[1] = {(uint64_t)p->ldw * 2};
      uint32_t box[2] = {BK, BN};
      if ((rc = host::make_tmap_bf16(&maps.w[i], p->W, 2, dims, strides, box)) != 0) return rc;
    }
    maps.ws[i] = maps.w[i];
    if (g.split > 1 || g.pair) {   // W through a box of BN/split rows (column slices of the tail tiles) or BN/2 (pairs)
      uint64_t dims[2] = {(uint64_t)p->K, (uint64_t)p->N};
      uint64_t strides[1] = {(uint64_t)p->ldw * 2};
      uint32_t box[2] = {BK, (uint32_t)(g.pair ? BN / 2 : BN / g.split)};
      if ((rc = host::make_tmap_bf16(&maps.ws[i], p->W, 2, dims, strides, box)) != 0) return rc;
    }
    maps.wt[i] = maps.ws[i];
    if (g.pair && g.split > 1) {
      uint64_t dims[2] = {(uint64_t)p->K, (uint64_t)p->N};
      uint64_t strides[1] = {(uint64_t)p->ldw * 2};
      uint32_t box[2] = {BK, (uint32_t)(BN / (2 * g.split))};
      if ((rc = host::make_tmap_bf16(&maps.wt[i], p->W, 2, dims, strides, box)) != 0) return rc;
    }
    Problem& P = g.p[i];   // shape fields were set by plan_tiles
    P.bias = p->bias;
    P.C = p->C;
    P.ldc = p->ldc;
    P.c_batch_stride = p->c_batch_stride;
    P.c_f32 = p->c_dtype == DIT4SR_DT_F32;
    P.gate = p->gate;
    P.gate_batch_stride = p->gate_batch_stride;
    P.resid = p->resid;
    P.ldr = p->ldr;
    P.r_batch_stride = p->r_batch_stride;
    P.norm_w0 = p->norm_w0;
    P.norm_w1 = p->norm_w1;
    P.sec_cols = p->sec_cols > 0 ? p->sec_cols : p->N;
    P.eps = p->eps;
    P.dbl_row_offset = p->dbl_row_offset;
    P.conv_w = p->conv_w;
    P.conv_cblocks = p->conv_w > 0 ? p->K / p->conv_taps / BK : 1;
    P.conv_kw = p->conv_w > 0 ? p->conv_kw : 1;
    P.conv_org = p->conv_org;
    P.resid_b = reinterpret_cast<const __nv_bfloat16*>(p->resid_bf16);
    P.ld_rb = p->ld_rb;
    P.rb_batch_stride = p->rb_batch_stride;
    P.sp_world = 0;
    if (p->sp != nullptr) {
      const dit4sr_gemm_sp* sp = p->sp;
      DIT4SR_REQUIRE(sp->world >= 1 && sp->world <= DIT4SR_SP_MAX_RANKS && sp->rank >= 0 && sp->rank < sp->world &&
                         sp->inner_dim > 0 && sp->inner_dim % (64 * sp->world) == 0 && sp->col0 >= 0 && sp->col0 % 64 == 0 &&
                         sp->col0 + p->N <= 3 * sp->inner_dim && sp->rows_dst > 0 && sp->row0 >= 0,
                     DIT4SR_E_BADARG, "gemm: bad sp scatter description");
      DIT4SR_REQUIRE(p->c_dtype == DIT4SR_DT_BF16 && p->N % 64 == 0 &&
                         (epilogue == DIT4SR_EPI_QKV_NORM || epilogue == DIT4SR_EPI_BIAS),
                     DIT4SR_E_BADARG, "gemm: sp scatter needs a bf16 q|k|v projection (BIAS or QKV_NORM epilogue)");
      for (int r = 0; r < sp->world; ++r) {
        DIT4SR_REQUIRE(sp->peer[r] != nullptr && host::aligned16(sp->peer[r]), DIT4SR_E_BADARG, "gemm: sp peer %d missing", r);
        DIT4SR_REQUIRE(!have_sp || g.sp_peer[r] == sp->peer[r], DIT4SR_E_BADARG,
                       "gemm: the problems of one launch must scatter into the same peer buffers");
        g.sp_peer[r] = sp->peer[r];
      }
      have_sp = true;
      P.sp_world = sp->world;
      P.sp_rank = sp->rank;
      P.sp_own_only = sp->own_heads_only;
      P.sp_D = sp->inner_dim;
      P.sp_col0 = sp->col0;
      P.sp_rows_dst = sp->rows_dst;
      P.sp_row0 = sp->row0;
    }
    // in-place gated residual -> TMA reduce-add epilogue (DIT4SR_GEMM_REDUCE=0 keeps the load/add/store epilogue)
    static int reduce_enabled = -1;
    if (reduce_enabled < 0) { const char* e = getenv("DIT4SR_GEMM_REDUCE"); reduce_enabled = (e == nullptr || e[0] != '0') ? 1 : 0; }
    P.c_reduce = reduce_enabled && epilogue == DIT4SR_EPI_GATE_RESID && p->resid == p->C && p->ldr == p->ldc &&
                 (p->batches == 1 || p->r_batch_stride == p->c_batch_stride);
    maps.c[i] = maps.a[i];
    if (P.c_reduce) {
      uint64_t dims[3] = {(uint64_t)p->N, (uint64_t)p->rows, (uint64_t)p->batches};
      uint64_t strides[2] = {(uint64_t)p->ldc * 4,
                             (uint64_t)(p->batches > 1 ? p->c_batch_stride : (int64_t)p->rows * p->ldc) * 4};
      uint32_t box[3] = {32, 32, 1};
      if ((rc = host::make_tmap_f32(&maps.c[i], p->C, 3, dims, strides, box)) != 0) return rc;
    }
  }
  for (int i = count; i < MAX_PROBLEMS; ++i) {
    g.p[i] = g.p[0];
    g.p[i].tile_begin = 0x7fffffff;
    maps.a[i] = maps.a[0];
    maps.w[i] = maps.w[0];
    maps.ws[i] = maps.ws[0];
    maps.c[i] = maps.c[0];
    maps.wt[i] = maps.wt[0];
  }
  switch (epilogue) {
    case DIT4SR_EPI_BIAS: return launch<DIT4SR_EPI_BIAS>(maps, g, stream);
    case DIT4SR_EPI_BIAS_GELU: return launch<DIT4SR_EPI_BIAS_GELU>(maps, g, stream);
    case DIT4SR_EPI_BIAS_SILU: return launch<DIT4SR_EPI_BIAS_SILU>(maps, g, stream);
    case DIT4SR_EPI_QKV_NORM: return launch<DIT4SR_EPI_QKV_NORM>(maps, g, stream);
    default: {
      bool all_reduce = true;
      for (int i = 0; i < count; ++i) all_reduce = all_reduce && g.p[i].c_reduce;
      return all_reduce ? launch<DIT4SR_EPI_GATE_RESID, true>(maps, g, stream)
                        : launch<DIT4SR_EPI_GATE_RESID, false>(maps, g, stream);
    }
  }
}

extern "C" int dit4sr_gemm_bf16(const dit4sr_gemm_params* p, void* stream) {
  DIT4SR_REQUIRE(p != nullptr, DIT4SR_E_BADARG, "gemm: null pointer argument");
  return dit4sr_gemm_bf16_grouped(p, 1, stream);
}

extern "C" int dit4sr_gemm_schedule(const dit4sr_gemm_params* probs, int count, int sm_count, int* items,
                                    int max_items) {
  using namespace gemm;
  DIT4SR_REQUIRE(probs != nullptr && count >= 1 && count <= MAX_PROBLEMS, DIT4SR_E_BADARG,
                 "gemm: group size must be 1..%d", MAX_PROBLEMS);
  DIT4SR_REQUIRE(sm_count >= 1 && max_items >= 0 && (items != nullptr || max_items == 0), DIT4SR_E_BADARG,
                 "gemm_schedule: bad sm_count / items");
  for (int i = 0; i < count; ++i) {
    const dit4sr_gemm_params* p = &probs[i];
    DIT4SR_REQUIRE(p->batches > 0 && p->rows > 0 && p->N > 0 && p->K > 0, DIT4SR_E_BADARG,
                   "gemm: non-positive size (batches %d rows %d N %d K %d)", p->batches, p->rows, p->N, p->K);
    DIT4SR_REQUIRE(p->K % 8 == 0 && p->N % 32 == 0, DIT4SR_E_SHAPE, "gemm: need K %% 8 == 0 and N %% 32 == 0 (K %d N %d)",
                   p->K, p->N);
  }
  Group g;
  plan_tiles(probs, count, sm_count, tail_split_enabled(), g);
  for (int item = 0; item < g.num_tiles && item < max_items; ++item) {
    const TileCoord tc = decode_tile(g, item);
    int* r = items + 5 * item;
    r[0] = tc.pi;
    r[1] = tc.b;
    r[2] = tc.m_tile;
    r[3] = tc.col0;
    r[4] = tc.bn;
  }
  return g.num_tiles;
}

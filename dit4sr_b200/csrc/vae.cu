// VAE-side kernels (AutoencoderKL encode / decode around the sampling loop, pipeline_dit4sr.py:1001-1002,
// 1602-1605).  The convolutions and linear layers run on the tcgen05 GEMM (gemm.cu, implicit-GEMM convolution mode);
// this file holds the HBM-bound rest on the channels-last (NHWC, bf16) activation layout:
//   GroupNorm (+ SiLU), nearest 2x upsampling, space-to-depth (turns the stride-2 3x3 down-sampling convolution into a
//   2x2 stride-1 one), fp32 row softmax -> bf16 probabilities (single-head mid-block attention with head_dim 512),
//   NCHW <-> padded NHWC layout changes at the two ends.
#include "host_util.h"
#include "ptx.cuh"

namespace vae {

constexpr int THREADS = 256;

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// eight consecutive channels of a bf16 or an fp32 activation tensor as floats
template <typename T>
__device__ __forceinline__ void load8(const T* p, float* f);
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
  unpack8(__ldg(reinterpret_cast<const uint4*>(p)), f);
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(ptx::pack_bf16(f[0], f[1]), ptx::pack_bf16(f[2], f[3]), ptx::pack_bf16(f[4], f[5]), ptx::pack_bf16(f[6], f[7]));
}

// ---- GroupNorm statistics, deterministic (no atomics): grid (chunks, batches); a CTA sums a contiguous range of pixels.
// Thread = one 8-channel vector (16 B) of every (THREADS / vecs)-th pixel; the per-thread partial sums are laid out in
// shared memory [lane][channel] and summed in a fixed order per channel, then per group (double), and written to
// part[b][chunk][g] = {sum, sum of squares}.  gn_finalize_kernel adds the chunks in order.
constexpr int GN_MAX_CHUNKS = 256;

template <typename T>
__global__ void __launch_bounds__(THREADS)
gn_stats_kernel(const T* __restrict__ x, int64_t ld, int64_t batch_stride, int pixels, int width, int pitch, int channels,
                int groups, int pixels_per_cta, double* __restrict__ part) {
  extern __shared__ float gn_sm[];   // [2][lanes][channels], then [2][channels]
  const int vecs = channels >> 3;
  const int lanes = THREADS / vecs;
  const int v = threadIdx.x % vecs, pl = threadIdx.x / vecs;
  const int b = blockIdx.y;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f, q[i] = 0.f;
  const int p0 = blockIdx.x * pixels_per_cta;
  const int p1 = min(pixels, p0 + pixels_per_cta);
  float* sm_s = gn_sm;
  float* sm_q = gn_sm + lanes * channels;
  float* ch_s = gn_sm + 2 * lanes * channels;
  float* ch_q = ch_s + channels;
  if (pl < lanes) {
    const T* base = x + (int64_t)b * batch_stride + v * 8;
    for (int p = p0 + pl; p < p1; p += lanes) {
      if (pitch != width && p % pitch >= width) continue;      // pad columns of a pitched image hold no data
      float f[8];
      load8(base + (int64_t)p * ld, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f[i];
        q[i] = fmaf(f[i], f[i], q[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sm_s[pl * channels + v * 8 + i] = s[i];
      sm_q[pl * channels + v * 8 + i] = q[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < channels; c += THREADS) {
    float a = 0.f, d = 0.f;
    for (int l = 0; l < lanes; ++l) {
      a += sm_s[l * channels + c];
      d += sm_q[l * channels + c];
    }
    ch_s[c] = a;
    ch_q[c] = d;
  }
  __syncthreads();
  const int cpg = channels / groups;
  for (int g = threadIdx.x; g < groups; g += THREADS) {
    double ds = 0.0, dq = 0.0;
    for (int c = 0; c < cpg; ++c) {
      ds += (double)ch_s[g * cpg + c];
      dq += (double)ch_q[g * cpg + c];
    }
    double* dst = part + (((int64_t)b * GN_MAX_CHUNKS + blockIdx.x) * groups + g) * 2;
    dst[0] = ds;
    dst[1] = dq;
  }
}

// stats[b][g] = sum over the chunks in a fixed order (deterministic): one warp per (group, batch); lane l adds chunks l,
// l + 32, ... in order, then the 32 lane sums meet in a shuffle tree.  (The first form - one thread per group walking up to
// 256 chunks with dependent double adds - took 42 us per call, a fifth of a 512px decode: profiles/r02b_vae_launches.csv.)
__global__ void __launch_bounds__(32) gn_finalize_kernel(const double* __restrict__ part, int chunks, int groups,
                                                         double* __restrict__ stats) {
  const int g = blockIdx.x, b = blockIdx.y, lane = threadIdx.x;
  double ds = 0.0, dq = 0.0;
  for (int c = lane; c < chunks; c += 32) {
    const double* src = part + (((int64_t)b * GN_MAX_CHUNKS + c) * groups + g) * 2;
    ds += __ldg(src);
    dq += __ldg(src + 1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dq += __shfl_xor_sync(0xffffffffu, dq, o);
  }
  if (lane == 0) {
    stats[((int64_t)b * groups + g) * 2] = ds;
    stats[((int64_t)b * groups + g) * 2 + 1] = dq;
  }
}

// ---- GroupNorm apply (+ SiLU): y = (x - mean_g) rstd_g gamma_c + beta_c = x a_c + b_c; one 8-channel vector per thread
// and step.  The grid stride is a multiple of the vectors per pixel (THREADS % vecs == 0), so a thread keeps the same eight
// channels for its whole loop and forms a_c, b_c once (the statistics are finished in double, once per thread).
template <typename T>
__global__ void __launch_bounds__(THREADS)
gn_apply_kernel(const T* __restrict__ x, int64_t ldx, int64_t x_batch_stride, __nv_bfloat16* __restrict__ y,
                int64_t ldy, int64_t y_batch_stride, int pixels, int valid_pixels, int channels, int groups,
                const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                int silu) {
  const int vecs = channels >> 3;
  const int b = blockIdx.y;
  const int cpg = channels / groups;
  const double inv_n = 1.0 / ((double)valid_pixels * cpg);
  const int64_t total = (int64_t)pixels * vecs;
  const int64_t first = (int64_t)blockIdx.x * THREADS + threadIdx.x;
  const int v = (int)(first % vecs);
  float ca[8], cb[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = v * 8 + k;
    const int g = c / cpg;
    const double m = stats[((int64_t)b * groups + g) * 2] * inv_n;
    const double var = fmax(stats[((int64_t)b * groups + g) * 2 + 1] * inv_n - m * m, 0.0);
    const float rstd = rsqrtf((float)var + eps);
    ca[k] = rstd * __ldg(gamma + c);
    cb[k] = __ldg(beta + c) - (float)m * ca[k];
  }
  const T* xb = x + (int64_t)b * x_batch_stride + v * 8;
  __nv_bfloat16* yb = y + (int64_t)b * y_batch_stride + v * 8;
  for (int64_t i = first; i < total; i += (int64_t)gridDim.x * THREADS) {
    const int64_t p = i / vecs;
    float f[8];
    load8(xb + p * ldx, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float t = fmaf(f[k], ca[k], cb[k]);
      if (silu) t = t / (1.0f + __expf(-t));
      f[k] = t;
    }
    *reinterpret_cast<uint4*>(yb + p * ldy) = pack8(f);
  }
}

// Images may be PITCHED: row y of a (H x W) image starts at pixel y * pitch, pitch >= W (the implicit-GEMM convolution
// needs a pitch that divides 128 or is a multiple of it); pad pixels hold no data and are written as zeros here.

// ---- nearest-neighbour 2x upsampling, NHWC (diffusers Upsample2D: F.interpolate(scale_factor=2.0, mode="nearest"))
template <typename T>
__global__ void __launch_bounds__(THREADS)
upsample2x_kernel(const T* __restrict__ x, uint4* __restrict__ y, int H, int W, int pitch_in, int pitch_out, int vecs,
                  int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * THREADS) {
    const int v = (int)(i % vecs);
    int64_t r = i / vecs;
    const int ox = (int)(r % pitch_out);
    r /= pitch_out;
    const int oy = (int)(r % (2 * H));
    const int64_t b = r / (2 * H);
    if (ox >= 2 * W) { y[i] = make_uint4(0, 0, 0, 0); continue; }
    float f[8];
    load8(x + (((b * H + (oy >> 1)) * pitch_in + (ox >> 1)) * vecs + v) * 8, f);
    y[i] = pack8(f);
  }
}

// ---- space to depth: y[b, Y, X, (py*2+px)*C + c] = x[b, 2Y+py, 2X+px, c]  (H, W even)
template <typename T>
__global__ void __launch_bounds__(THREADS)
space_to_depth2_kernel(const T* __restrict__ x, uint4* __restrict__ y, int H, int W, int pitch_in, int pitch_out, int vecs,
                       int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * THREADS) {
    const int v = (int)(i % vecs);
    int64_t r = i / vecs;
    const int ph = (int)(r % 4);
    r /= 4;
    const int X = (int)(r % pitch_out);
    r /= pitch_out;
    const int Y = (int)(r % (H / 2));
    const int64_t b = r / (H / 2);
    if (X >= W / 2) { y[i] = make_uint4(0, 0, 0, 0); continue; }
    float f[8];
    load8(x + (((b * H + 2 * Y + (ph >> 1)) * pitch_in + 2 * X + (ph & 1)) * vecs + v) * 8, f);
    y[i] = pack8(f);
  }
}

// ---- fp32 score rows -> bf16 probabilities: p = softmax(scale * s) over the first `valid` columns, zeros in the rest (key
// padding to the GEMM's granularity); one CTA per row, three passes over the row (it stays in L1 / L2: <= 256 KB).
__global__ void __launch_bounds__(THREADS)
softmax_rows_kernel(const float* __restrict__ s, int64_t lds, __nv_bfloat16* __restrict__ p, int64_t ldp, int cols, int valid,
                    float scale_log2) {
  __shared__ float red[THREADS / 32];
  const float* row = s + (int64_t)blockIdx.x * lds;
  __nv_bfloat16* out = p + (int64_t)blockIdx.x * ldp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto load4 = [&](int c, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(row + c);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (c + i >= valid) v[i] = -INFINITY;
  };
  float m = -INFINITY;
  for (int c = threadIdx.x * 4; c < cols; c += THREADS * 4) {
    float v[4];
    load4(c, v);
    m = fmaxf(m, fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < THREADS / 32; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  const float nm = -m * scale_log2;
  float sum = 0.f;
  for (int c = threadIdx.x * 4; c < cols; c += THREADS * 4) {
    float v[4];
    load4(c, v);
    sum += (ptx::ex2(fmaf(v[0], scale_log2, nm)) + ptx::ex2(fmaf(v[1], scale_log2, nm))) +
           (ptx::ex2(fmaf(v[2], scale_log2, nm)) + ptx::ex2(fmaf(v[3], scale_log2, nm)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) sum += red[w];
  const float inv = 1.0f / sum;
  for (int c = threadIdx.x * 4; c < cols; c += THREADS * 4) {
    float v[4];
    load4(c, v);
    uint2 o;
    o.x = ptx::pack_bf16(ptx::ex2(fmaf(v[0], scale_log2, nm)) * inv, ptx::ex2(fmaf(v[1], scale_log2, nm)) * inv);
    o.y = ptx::pack_bf16(ptx::ex2(fmaf(v[2], scale_log2, nm)) * inv, ptx::ex2(fmaf(v[3], scale_log2, nm)) * inv);
    *reinterpret_cast<uint2*>(out + c) = o;
  }
}

// ---- NCHW (fp32 / bf16) [B, C, H, W] -> NHWC bf16 [B, H, pitch, c_pad] with the channels zero-padded to c_pad, y = a x + b
template <typename T>
__global__ void __launch_bounds__(THREADS)
nchw_to_nhwc_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ y, int C, int c_pad, int H, int W, int pitch, float a,
                    float bb) {
  const int b = blockIdx.y;
  const int64_t hw = (int64_t)H * W, hp = (int64_t)H * pitch;
  for (int64_t p = (int64_t)blockIdx.x * THREADS + threadIdx.x; p < hp; p += (int64_t)gridDim.x * THREADS) {
    const int yy = (int)(p / pitch), xx = (int)(p % pitch);
    const bool ok = xx < W;
    const int64_t src = (int64_t)yy * W + xx;
    __nv_bfloat16* dst = y + ((int64_t)b * hp + p) * c_pad;
    for (int c0 = 0; c0 < c_pad; c0 += 8) {
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        float f[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = c0 + k + e;
          f[e] = (ok && c < C) ? fmaf((float)x[((int64_t)b * C + c) * hw + src], a, bb) : 0.f;
        }
        o[k >> 1] = ptx::pack_bf16(f[0], f[1]);
      }
      *reinterpret_cast<uint4*>(dst + c0) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ---- fp32 NHWC [B, H, pitch, ld] (first C columns) -> NCHW fp32 [B, C, H, W]
__global__ void __launch_bounds__(THREADS)
nhwc_to_nchw_kernel(const float* __restrict__ x, int64_t ld, float* __restrict__ y, int C, int H, int W, int pitch) {
  const int b = blockIdx.y;
  const int64_t hw = (int64_t)H * W;
  for (int64_t p = (int64_t)blockIdx.x * THREADS + threadIdx.x; p < hw; p += (int64_t)gridDim.x * THREADS) {
    const int yy = (int)(p / W), xx = (int)(p % W);
    const float* src = x + (((int64_t)b * H + yy) * pitch + xx) * ld;
    for (int c = 0; c < C; ++c) y[((int64_t)b * C + c) * hw + p] = src[c];
  }
}

static unsigned grid_for(int64_t work_items) {
  int64_t g = (work_items + THREADS - 1) / THREADS;
  const int64_t cap = (int64_t)host::sm_count() * 16;
  if (g > cap) g = cap;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace vae

extern "C" int64_t dit4sr_groupnorm_workspace_bytes(int batches, int groups) {
  if (batches < 1 || groups < 1) return 0;
  return (int64_t)batches * (vae::GN_MAX_CHUNKS + 1) * groups * 2 * (int64_t)sizeof(double);
}

extern "C" int dit4sr_groupnorm_nhwc(const dit4sr_groupnorm_params* p, void* stream_) {
  using namespace vae;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(p && p->x && p->y && p->gamma && p->beta && p->stats, DIT4SR_E_BADARG, "groupnorm: null pointer argument");
  DIT4SR_REQUIRE(p->batches > 0 && p->batches <= 65535 && p->pixels > 0 && p->channels > 0 && p->groups > 0,
                 DIT4SR_E_BADARG, "groupnorm: non-positive size");
  DIT4SR_REQUIRE(p->x_dtype == DIT4SR_DT_BF16 || p->x_dtype == DIT4SR_DT_F32, DIT4SR_E_BADARG, "groupnorm: unknown x_dtype");
  DIT4SR_REQUIRE(p->width >= 0 && p->pitch >= p->width && (p->width == 0 || p->pixels % p->pitch == 0), DIT4SR_E_BADARG,
                 "groupnorm: pitched image needs 0 < width <= pitch and pixels a multiple of pitch");
  const int width = p->width > 0 ? p->width : 1, pitch = p->width > 0 ? p->pitch : 1;
  const int valid_pixels = p->width > 0 ? p->pixels / p->pitch * p->width : p->pixels;
  DIT4SR_REQUIRE(p->channels % 8 == 0 && p->channels % p->groups == 0 && p->channels <= 8 * THREADS &&
                     THREADS % (p->channels / 8) == 0,
                 DIT4SR_E_SHAPE, "groupnorm: channels %d must be 8 * a divisor of %d and a multiple of groups %d", p->channels,
                 THREADS, p->groups);
  DIT4SR_REQUIRE(p->stats_bytes >= dit4sr_groupnorm_workspace_bytes(p->batches, p->groups), DIT4SR_E_BADARG,
                 "groupnorm: stats scratch of %lld bytes, need %lld", (long long)p->stats_bytes,
                 (long long)dit4sr_groupnorm_workspace_bytes(p->batches, p->groups));
  DIT4SR_REQUIRE(p->ldx % 8 == 0 && p->ldy % 8 == 0 && p->x_batch_stride % 8 == 0 && p->y_batch_stride % 8 == 0 &&
                     host::aligned16(p->x) && host::aligned16(p->y) && (reinterpret_cast<uintptr_t>(p->stats) & 7) == 0,
                 DIT4SR_E_ALIGN, "groupnorm: pointers/strides not 16-byte aligned");
  int rc = dit4sr_check_device();
  if (rc != 0) return rc;
  const int sms = host::sm_count();
  int chunks = (4 * sms + p->batches - 1) / p->batches;
  if (chunks > GN_MAX_CHUNKS) chunks = GN_MAX_CHUNKS;
  int per = (p->pixels + chunks - 1) / chunks;
  if (per < 64) per = 64;
  chunks = (p->pixels + per - 1) / per;
  const int64_t items = (int64_t)p->pixels * (p->channels / 8);
  const unsigned gx = grid_for(items);
  const size_t sm = (2 * (THREADS / (p->channels / 8)) + 2) * p->channels * sizeof(float);   // <= 16 KB + 2 rows
  double* part = reinterpret_cast<double*>(p->stats);
  double* stats = part + (size_t)p->batches * GN_MAX_CHUNKS * p->groups * 2;
  __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p->y);
  if (p->x_dtype == DIT4SR_DT_F32) {
    const float* x = reinterpret_cast<const float*>(p->x);
    gn_stats_kernel<<<dim3(chunks, p->batches), THREADS, sm, stream>>>(x, p->ldx, p->x_batch_stride, p->pixels, width, pitch,
                                                                       p->channels, p->groups, per, part);
    gn_finalize_kernel<<<dim3(p->groups, p->batches), 32, 0, stream>>>(part, chunks, p->groups, stats);
    gn_apply_kernel<<<dim3(gx, p->batches), THREADS, 0, stream>>>(x, p->ldx, p->x_batch_stride, y, p->ldy, p->y_batch_stride,
                                                                 p->pixels, valid_pixels, p->channels, p->groups, stats,
                                                                 p->gamma, p->beta, p->eps, p->silu);
  } else {
    const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p->x);
    gn_stats_kernel<<<dim3(chunks, p->batches), THREADS, sm, stream>>>(x, p->ldx, p->x_batch_stride, p->pixels, width, pitch,
                                                                       p->channels, p->groups, per, part);
    gn_finalize_kernel<<<dim3(p->groups, p->batches), 32, 0, stream>>>(part, chunks, p->groups, stats);
    gn_apply_kernel<<<dim3(gx, p->batches), THREADS, 0, stream>>>(x, p->ldx, p->x_batch_stride, y, p->ldy, p->y_batch_stride,
                                                                 p->pixels, valid_pixels, p->channels, p->groups, stats,
                                                                 p->gamma, p->beta, p->eps, p->silu);
  }
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dit4sr_upsample2x_nhwc(const void* x, int x_dtype, void* y, int batches, int h, int w, int pitch_in,
                                      int pitch_out, int channels, void* stream_) {
  using namespace vae;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(x && y, DIT4SR_E_BADARG, "upsample2x: null pointer argument");
  DIT4SR_REQUIRE(batches > 0 && h > 0 && w > 0 && channels > 0 && channels % 8 == 0 && pitch_in >= w && pitch_out >= 2 * w,
                 DIT4SR_E_SHAPE, "upsample2x: sizes must be positive, channels a multiple of 8, pitches >= the widths");
  DIT4SR_REQUIRE(host::aligned16(x) && host::aligned16(y), DIT4SR_E_ALIGN, "upsample2x: pointers not 16-byte aligned");
  const int64_t total = (int64_t)batches * 2 * h * pitch_out * (channels / 8);
  if (x_dtype == DIT4SR_DT_F32)
    upsample2x_kernel<<<grid_for(total), THREADS, 0, stream>>>(reinterpret_cast<const float*>(x), reinterpret_cast<uint4*>(y), h,
                                                              w, pitch_in, pitch_out, channels / 8, total);
  else
    upsample2x_kernel<<<grid_for(total), THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                              reinterpret_cast<uint4*>(y), h, w, pitch_in, pitch_out,
                                                              channels / 8, total);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dit4sr_space_to_depth2_nhwc(const void* x, int x_dtype, void* y, int batches, int h, int w, int pitch_in,
                                           int pitch_out, int channels, void* stream_) {
  using namespace vae;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(x && y, DIT4SR_E_BADARG, "space_to_depth2: null pointer argument");
  DIT4SR_REQUIRE(batches > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0 && channels > 0 && channels % 8 == 0 &&
                     pitch_in >= w && pitch_out >= w / 2,
                 DIT4SR_E_SHAPE, "space_to_depth2: h, w must be even, channels a multiple of 8, pitches >= the widths");
  DIT4SR_REQUIRE(host::aligned16(x) && host::aligned16(y), DIT4SR_E_ALIGN, "space_to_depth2: pointers not 16-byte aligned");
  const int64_t total = (int64_t)batches * (h / 2) * pitch_out * 4 * (channels / 8);
  if (x_dtype == DIT4SR_DT_F32)
    space_to_depth2_kernel<<<grid_for(total), THREADS, 0, stream>>>(reinterpret_cast<const float*>(x), reinterpret_cast<uint4*>(y),
                                                                   h, w, pitch_in, pitch_out, channels / 8, total);
  else
    space_to_depth2_kernel<<<grid_for(total), THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                   reinterpret_cast<uint4*>(y), h, w, pitch_in, pitch_out,
                                                                   channels / 8, total);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dit4sr_softmax_rows_to_bf16(const float* s, int64_t lds, void* p, int64_t ldp, int rows, int cols, int valid_cols,
                                           float scale, void* stream_) {
  using namespace vae;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(s && p, DIT4SR_E_BADARG, "softmax_rows_to_bf16: null pointer argument");
  DIT4SR_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0 && valid_cols > 0 && valid_cols <= cols, DIT4SR_E_SHAPE,
                 "softmax_rows_to_bf16: cols must be a multiple of 4, 0 < valid_cols <= cols");
  DIT4SR_REQUIRE(lds % 4 == 0 && ldp % 4 == 0 && host::aligned16(s) && (reinterpret_cast<uintptr_t>(p) & 7) == 0, DIT4SR_E_ALIGN,
                 "softmax_rows_to_bf16: pointers/strides not aligned");
  softmax_rows_kernel<<<rows, THREADS, 0, stream>>>(s, lds, reinterpret_cast<__nv_bfloat16*>(p), ldp, cols, valid_cols,
                                                   scale * 1.4426950408889634f);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dit4sr_nchw_to_nhwc_bf16(const void* x, int x_dtype, void* y, int batches, int channels, int c_pad, int h, int w,
                                        int pitch, float a, float b, void* stream_) {
  using namespace vae;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(x && y, DIT4SR_E_BADARG, "nchw_to_nhwc: null pointer argument");
  DIT4SR_REQUIRE(batches > 0 && batches <= 65535 && channels > 0 && c_pad >= channels && c_pad % 8 == 0 && h > 0 && w > 0 &&
                     pitch >= w,
                 DIT4SR_E_SHAPE, "nchw_to_nhwc: bad sizes (c_pad must be a multiple of 8 and >= channels, pitch >= w)");
  DIT4SR_REQUIRE(host::aligned16(y), DIT4SR_E_ALIGN, "nchw_to_nhwc: output not 16-byte aligned");
  DIT4SR_REQUIRE(x_dtype == DIT4SR_DT_BF16 || x_dtype == DIT4SR_DT_F32, DIT4SR_E_BADARG,
                 "nchw_to_nhwc: unknown input dtype %d", x_dtype);
  const dim3 grid(grid_for((int64_t)h * pitch), batches);
  if (x_dtype == DIT4SR_DT_F32)
    nchw_to_nhwc_kernel<<<grid, THREADS, 0, stream>>>(reinterpret_cast<const float*>(x), reinterpret_cast<__nv_bfloat16*>(y),
                                                     channels, c_pad, h, w, pitch, a, b);
  else
    nchw_to_nhwc_kernel<<<grid, THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                     reinterpret_cast<__nv_bfloat16*>(y), channels, c_pad, h, w, pitch, a, b);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dit4sr_nhwc_to_nchw_f32(const float* x, int64_t ld, float* y, int batches, int channels, int h, int w, int pitch,
                                       void* stream_) {
  using namespace vae;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(x && y, DIT4SR_E_BADARG, "nhwc_to_nchw: null pointer argument");
  DIT4SR_REQUIRE(batches > 0 && batches <= 65535 && channels > 0 && ld >= channels && h > 0 && w > 0 && pitch >= w,
                 DIT4SR_E_SHAPE, "nhwc_to_nchw: bad sizes");
  nhwc_to_nchw_kernel<<<dim3(grid_for((int64_t)h * w), batches), THREADS, 0, stream>>>(x, ld, y, channels, h, w, pitch);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

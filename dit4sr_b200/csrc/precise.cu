// Helper kernels of the fp32-accurate ("bf16x3") mode.  In that mode every activation between kernels is fp32 and each
// tensor-core GEMM C = A W^T is evaluated as  Ah Wh^T + Ah Wl^T + Al Wh^T  (hi/lo bf16 splits, the lo*lo term dropped:
// relative error ~2^-16) by concatenating the three terms along K:  A' = [Ah | Ah | Al],  W' = [Wh | Wl | Wh].  The unchanged
// tcgen05 GEMM kernel then runs on K' = 3K.  Attention is evaluated unfused per head with the same GEMM (scores, then
// probabilities x values) and an fp32 row softmax in between.
#include "host_util.h"
#include "ptx.cuh"

namespace precise {

__device__ __forceinline__ void split(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// out[(b*rows_pad + r), :] = pattern 0: [hi | hi | lo], pattern 1: [hi | lo | hi]; rows r >= rows are zero-filled
__global__ void split3_kernel(const float* __restrict__ x, int64_t ldx, int64_t xbs, __nv_bfloat16* __restrict__ out,
                              int batches, int rows, int rows_pad, int cols, int pattern) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)batches * rows_pad * cols;
  if (i >= total) return;
  const int c = (int)(i % cols);
  const int r = (int)((i / cols) % rows_pad);
  const int b = (int)(i / ((int64_t)cols * rows_pad));
  __nv_bfloat16 hi = __float2bfloat16_rn(0.f), lo = hi;
  if (r < rows) split(x[(int64_t)b * xbs + (int64_t)r * ldx + c], hi, lo);
  __nv_bfloat16* o = out + ((int64_t)b * rows_pad + r) * (3 * (int64_t)cols);
  o[c] = hi;
  o[cols + c] = pattern == 0 ? hi : lo;
  o[2 * cols + c] = pattern == 0 ? lo : hi;
}

// V [rows(kv), cols(dh)] fp32 (row stride ldv) -> out [cols, 3*rows_pad] = [Vh^T | Vl^T | Vh^T] (the W operand of P V)
__global__ void split3_transposed_kernel(const float* __restrict__ v, int64_t ldv, __nv_bfloat16* __restrict__ out,
                                         int rows, int rows_pad, int cols) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows_pad * cols) return;
  const int r = (int)(i % rows_pad);   // kv index (fast -> coalesced writes)
  const int c = (int)(i / rows_pad);   // head-dim index
  __nv_bfloat16 hi = __float2bfloat16_rn(0.f), lo = hi;
  if (r < rows) split(v[(int64_t)r * ldv + c], hi, lo);
  __nv_bfloat16* o = out + (int64_t)c * (3 * (int64_t)rows_pad);
  o[r] = hi;
  o[rows_pad + r] = lo;
  o[2 * rows_pad + r] = hi;
}

// in-place softmax(scale * s) over the first cols_valid columns of each row; columns [cols_valid, cols_total) are set to 0
__global__ void softmax_rows_kernel(float* __restrict__ s, int64_t ld, int rows, int cols_valid, int cols_total,
                                    float scale) {
  const int row = blockIdx.x;
  if (row >= rows) return;
  float* p = s + (int64_t)row * ld;
  __shared__ float red[32];
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < cols_valid; c += blockDim.x) mx = fmaxf(mx, p[c]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < cols_valid; c += blockDim.x) {
    const float e = expf((p[c] - mx) * scale);
    p[c] = e;
    sum += e;
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum += red[w];
  const float inv = 1.0f / sum;
  for (int c = threadIdx.x; c < cols_total; c += blockDim.x) p[c] = c < cols_valid ? p[c] * inv : 0.f;
}

__global__ void add_f32_kernel(float* __restrict__ dst, int64_t drs, int64_t dbs, const float* __restrict__ src,
                               int64_t srs, int64_t sbs, int B, int R, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * R * C) return;
  const int c = (int)(i % C);
  const int r = (int)((i / C) % R);
  const int b = (int)(i / ((int64_t)C * R));
  dst[b * dbs + r * drs + c] += src[b * sbs + r * srs + c];
}

// fp32 depthwise 3x3 (same contract as dwconv3x3_kernel, plain global loads; accuracy mode only)
__global__ void dwconv3x3_f32_kernel(float* __restrict__ hq, const float* __restrict__ lr, int64_t ld, int64_t hbs,
                                     int64_t lbs, const float* __restrict__ w9, const float* __restrict__ bias, int B,
                                     int GH, int GW, int LH, int off, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * GH * GW * C) return;
  const int c = (int)(i % C);
  const int x = (int)((i / C) % GW);
  const int y = (int)((i / ((int64_t)C * GW)) % GH);
  const int b = (int)(i / ((int64_t)C * GW * GH));
  float acc = bias[c];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int yy = y + off + dy - 1, xx = x + dx - 1;
      if (yy >= 0 && yy < LH && xx >= 0 && xx < GW)
        acc = fmaf(w9[(dy * 3 + dx) * (int64_t)C + c], lr[b * lbs + ((int64_t)yy * GW + xx) * ld + c], acc);
    }
  hq[b * hbs + ((int64_t)y * GW + x) * ld + c] += acc;
}


// fp32 head kernels of the accuracy mode (same math as misc.cu, fp32 outputs)
__global__ void patchify_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int C, int H, int W) {
  const int Hp = H / 2, Wp = W / 2;
  const int64_t total = (int64_t)B * Hp * Wp * C;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t tok = i / C;
  const int wp = (int)(tok % Wp);
  const int hp = (int)((tok / Wp) % Hp);
  const int b = (int)(tok / ((int64_t)Wp * Hp));
  const float* src = x + (((int64_t)b * C + c) * H + 2 * hp) * W + 2 * wp;
  *reinterpret_cast<float4*>(out + tok * (C * 4) + c * 4) = make_float4(src[0], src[1], src[W], src[W + 1]);
}

__global__ void timestep_features_f32_kernel(const float* __restrict__ t, float* __restrict__ out, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 128) return;
  const int b = i / 128, k = i % 128;
  const float freq = expf(-9.210340371976184f * (float)k / 128.0f);
  const float a = t[b] * freq;
  out[b * 256 + k] = cosf(a);
  out[b * 256 + 128 + k] = sinf(a);
}

__global__ void silu_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  y[i] = v / (1.0f + expf(-v));
}

static inline unsigned blocks(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace precise

extern "C" {

int dit4sr_split3(const float* x, int64_t ldx, int64_t x_batch_stride, void* out_bf16, int batches, int rows,
                  int rows_pad, int cols, int pattern, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(x && out_bf16 && batches > 0 && rows > 0 && cols > 0 && rows_pad >= rows && (pattern == 0 || pattern == 1),
                 DIT4SR_E_BADARG, "split3: bad argument");
  const int64_t n = (int64_t)batches * rows_pad * cols;
  precise::split3_kernel<<<precise::blocks(n, 256), 256, 0, stream>>>(x, ldx, x_batch_stride, (__nv_bfloat16*)out_bf16,
                                                                      batches, rows, rows_pad, cols, pattern);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_split3_transposed(const float* v, int64_t ldv, void* out_bf16, int rows, int rows_pad, int cols,
                             void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(v && out_bf16 && rows > 0 && cols > 0 && rows_pad >= rows, DIT4SR_E_BADARG,
                 "split3_transposed: bad argument");
  const int64_t n = (int64_t)rows_pad * cols;
  precise::split3_transposed_kernel<<<precise::blocks(n, 256), 256, 0, stream>>>(v, ldv, (__nv_bfloat16*)out_bf16, rows,
                                                                                 rows_pad, cols);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_softmax_rows_f32(float* s, int64_t ld, int rows, int cols_valid, int cols_total, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(s && rows > 0 && cols_valid > 0 && cols_total >= cols_valid, DIT4SR_E_BADARG,
                 "softmax_rows: bad argument");
  precise::softmax_rows_kernel<<<rows, 256, 0, stream>>>(s, ld, rows, cols_valid, cols_total, scale);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_add_f32(float* dst, int64_t dst_row_stride, int64_t dst_batch_stride, const float* src,
                   int64_t src_row_stride, int64_t src_batch_stride, int batches, int rows, int cols, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(dst && src && batches > 0 && rows > 0 && cols > 0, DIT4SR_E_BADARG, "add_f32: bad argument");
  const int64_t n = (int64_t)batches * rows * cols;
  precise::add_f32_kernel<<<precise::blocks(n, 256), 256, 0, stream>>>(dst, dst_row_stride, dst_batch_stride, src,
                                                                       src_row_stride, src_batch_stride, batches, rows,
                                                                       cols);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_dwconv3x3_add_f32(float* u_hq, const float* u_lr, int64_t ld, int64_t hq_batch_stride,
                             int64_t lr_batch_stride, const float* weight9, const float* bias, int batches, int grid_h,
                             int grid_w, int lr_grid_h, int lr_row_offset, int channels, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(u_hq && u_lr && weight9 && bias && batches > 0 && grid_h > 0 && grid_w > 0 && lr_grid_h > 0 &&
                     channels > 0,
                 DIT4SR_E_BADARG, "dwconv_f32: bad argument");
  const int64_t n = (int64_t)batches * grid_h * grid_w * channels;
  precise::dwconv3x3_f32_kernel<<<precise::blocks(n, 256), 256, 0, stream>>>(
      u_hq, u_lr, ld, hq_batch_stride, lr_batch_stride, weight9, bias, batches, grid_h, grid_w, lr_grid_h, lr_row_offset,
      channels);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_patchify_f32(const float* x, float* patches, int batches, int channels, int height, int width, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(x && patches && batches > 0 && channels > 0, DIT4SR_E_BADARG, "patchify_f32: bad argument");
  DIT4SR_REQUIRE(height % 2 == 0 && width % 2 == 0 && height > 0 && width > 0, DIT4SR_E_SHAPE,
                 "patchify_f32: height/width must be even");
  DIT4SR_REQUIRE(host::aligned16(patches), DIT4SR_E_ALIGN, "patchify_f32: output not 16-byte aligned");
  const int64_t total = (int64_t)batches * (height / 2) * (width / 2) * channels;
  precise::patchify_f32_kernel<<<precise::blocks(total, 256), 256, 0, stream>>>(x, patches, batches, channels, height,
                                                                               width);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_timestep_features_f32(const float* timesteps, float* out, int batches, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(timesteps && out && batches > 0, DIT4SR_E_BADARG, "timestep_features_f32: bad argument");
  precise::timestep_features_f32_kernel<<<precise::blocks((int64_t)batches * 128, 128), 128, 0, stream>>>(timesteps, out,
                                                                                                          batches);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_silu_f32(const float* x, float* y, int64_t n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(x && y && n > 0, DIT4SR_E_BADARG, "silu_f32: bad argument");
  precise::silu_f32_kernel<<<precise::blocks(n, 256), 256, 0, stream>>>(x, y, n);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}
}

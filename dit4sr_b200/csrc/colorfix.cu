// Colour-fix post-processing of the SR result against the LQ input (reference utils/wavelet_color_fix.py, called from
// test/test.py:381-383 after the VAE decode): wavelet reconstruction (5-level a-trous decomposition with the 3x3 kernel
// [1 2 1]^T [1 2 1] / 16, replicate padding, dilation 2^i) and adaptive instance normalisation, on fp32 [planes, H, W]
// images (planes = batch x 3).  Small HBM-bound kernels: one thread per pixel, coalesced along the row.
#include "host_util.h"

namespace colorfix {

// low = blur(image, radius) (utils/wavelet_color_fix.py:76-94); if high != nullptr: high += image - low (:104-108)
__global__ void wavelet_blur_kernel(const float* __restrict__ image, float* __restrict__ low, float* __restrict__ high,
                                    int planes, int H, int W, int radius) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)planes * H * W;
  if (i >= total) return;
  const int x = (int)(i % W);
  const int y = (int)((i / W) % H);
  const float* pl = image + (i - (int64_t)y * W - x);
  const int xm = max(x - radius, 0), xp = min(x + radius, W - 1);     // replicate padding = clamped taps
  const int ym = max(y - radius, 0), yp = min(y + radius, H - 1);
  const float* r0 = pl + (int64_t)ym * W;
  const float* r1 = pl + (int64_t)y * W;
  const float* r2 = pl + (int64_t)yp * W;
  float acc = 0.0625f * r0[xm];
  acc = fmaf(0.125f, r0[x], acc);
  acc = fmaf(0.0625f, r0[xp], acc);
  acc = fmaf(0.125f, r1[xm], acc);
  const float c = r1[x];
  acc = fmaf(0.25f, c, acc);
  acc = fmaf(0.125f, r1[xp], acc);
  acc = fmaf(0.0625f, r2[xm], acc);
  acc = fmaf(0.125f, r2[x], acc);
  acc = fmaf(0.0625f, r2[xp], acc);
  low[i] = acc;
  if (high != nullptr) high[i] += c - acc;
}

// per plane: mean and unbiased variance (torch .var default), two passes, double accumulators
__global__ void __launch_bounds__(1024) plane_stats_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ mean,
                                                           float* __restrict__ var) {
  __shared__ double red[32];
  __shared__ double s_mean;
  const float* p = x + (int64_t)blockIdx.x * n;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)p[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    s_mean = t / (double)n;
  }
  __syncthreads();
  const double m = s_mean;
  double q = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = (double)p[i] - m;
    q += d * d;
  }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    mean[blockIdx.x] = (float)m;
    var[blockIdx.x] = (float)(t / (double)(n > 1 ? n - 1 : 1));
  }
}

// out = (content - mean_c) / sqrt(var_c + eps) * sqrt(var_s + eps) + mean_s   (utils/wavelet_color_fix.py:47-74)
__global__ void adain_apply_kernel(const float* __restrict__ content, float* __restrict__ out, int64_t n, int64_t total,
                                   const float* __restrict__ c_mean, const float* __restrict__ c_var,
                                   const float* __restrict__ s_mean, const float* __restrict__ s_var, float eps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int pl = (int)(i / n);
  const float cs = sqrtf(c_var[pl] + eps), ss = sqrtf(s_var[pl] + eps);
  out[i] = (content[i] - c_mean[pl]) / cs * ss + s_mean[pl];
}

static inline unsigned blocks(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace colorfix

extern "C" {

int dit4sr_wavelet_blur_f32(const float* image, float* low, float* high_accum, int planes, int height, int width,
                            int radius, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(image && low && planes > 0 && height > 0 && width > 0 && radius > 0, DIT4SR_E_BADARG,
                 "wavelet_blur: bad argument");
  DIT4SR_REQUIRE(image != low, DIT4SR_E_BADARG, "wavelet_blur: cannot run in place");
  const int64_t total = (int64_t)planes * height * width;
  colorfix::wavelet_blur_kernel<<<colorfix::blocks(total, 256), 256, 0, stream>>>(image, low, high_accum, planes, height,
                                                                                 width, radius);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_plane_stats_f32(const float* x, int planes, int64_t plane_elems, float* mean, float* var_unbiased,
                           void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(x && mean && var_unbiased && planes > 0 && plane_elems > 0, DIT4SR_E_BADARG, "plane_stats: bad argument");
  colorfix::plane_stats_kernel<<<planes, 1024, 0, stream>>>(x, plane_elems, mean, var_unbiased);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_adain_apply_f32(const float* content, float* out, int planes, int64_t plane_elems, const float* c_mean,
                           const float* c_var, const float* s_mean, const float* s_var, float eps, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(content && out && c_mean && c_var && s_mean && s_var && planes > 0 && plane_elems > 0, DIT4SR_E_BADARG,
                 "adain_apply: bad argument");
  const int64_t total = (int64_t)planes * plane_elems;
  colorfix::adain_apply_kernel<<<colorfix::blocks(total, 256), 256, 0, stream>>>(content, out, plane_elems, total, c_mean,
                                                                                c_var, s_mean, s_var, eps);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}
}

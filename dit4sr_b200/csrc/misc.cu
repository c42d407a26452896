// Small data-movement kernels of the model head and tail (transformer_sd3.py:326-332,379-391): patchify (im2col for the
// k=2,s=2 PatchEmbed conv), unpatchify (+ optional fused flow-match Euler update), timestep features, casts.
#include "host_util.h"
#include "ptx.cuh"

namespace misc {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// one thread per (b, token, c): writes 4 contiguous bf16 (ph, pw)
template <typename T>
__global__ void patchify_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int C, int H, int W) {
  const int Hp = H / 2, Wp = W / 2;
  const int64_t total = (int64_t)B * Hp * Wp * C;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t tok = i / C;
  const int wp = (int)(tok % Wp);
  const int hp = (int)((tok / Wp) % Hp);
  const int b = (int)(tok / ((int64_t)Wp * Hp));
  const T* src = x + (((int64_t)b * C + c) * H + 2 * hp) * W + 2 * wp;
  uint2 o;
  o.x = ptx::pack_bf16(to_f<T>(src[0]), to_f<T>(src[1]));
  o.y = ptx::pack_bf16(to_f<T>(src[W]), to_f<T>(src[W + 1]));
  *reinterpret_cast<uint2*>(out + tok * (C * 4) + c * 4) = o;
}

// one thread per output element (b, d, y, x); tok layout [B, Hp*Wp, (pH pW d)]
template <typename TO>
__global__ void unpatchify_kernel(const float* __restrict__ tok, TO* __restrict__ out, const TO* __restrict__ latents,
                                  float dsigma, int B, int C, int H, int W) {
  const int64_t total = (int64_t)B * C * H * W;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W);
  const int y = (int)((i / W) % H);
  const int d = (int)((i / ((int64_t)W * H)) % C);
  const int b = (int)(i / ((int64_t)W * H * C));
  const int Wp = W / 2;
  const int64_t t = ((int64_t)b * (H / 2) + y / 2) * Wp + x / 2;
  float v = tok[t * (4 * C) + ((y & 1) * 2 + (x & 1)) * C + d];
  if (latents != nullptr) v = to_f<TO>(latents[i]) + dsigma * v;
  if constexpr (sizeof(TO) == 2) out[i] = __float2bfloat16_rn(v);
  else out[i] = v;
}

// classifier-free guidance (pipeline_dit4sr.py:1561-1563) fused with unpatchify and the Euler update: tok holds the
// velocity tokens of the doubled batch [uncond (B) | text (B)]; v = v_u + g (v_t - v_u)
template <typename TO>
__global__ void unpatchify_cfg_kernel(const float* __restrict__ tok, TO* __restrict__ out, const TO* __restrict__ latents,
                                      float dsigma, float guidance, int B, int C, int H, int W) {
  const int64_t total = (int64_t)B * C * H * W;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W);
  const int y = (int)((i / W) % H);
  const int d = (int)((i / ((int64_t)W * H)) % C);
  const int b = (int)(i / ((int64_t)W * H * C));
  const int Wp = W / 2, Hp = H / 2;
  const int64_t t = ((int64_t)b * Hp + y / 2) * Wp + x / 2;
  const int64_t e = ((y & 1) * 2 + (x & 1)) * C + d;
  const float vu = tok[t * (4 * C) + e];
  const float vt = tok[(t + (int64_t)B * Hp * Wp) * (4 * C) + e];
  float v = vu + guidance * (vt - vu);
  if (latents != nullptr) v = to_f<TO>(latents[i]) + dsigma * v;
  if constexpr (sizeof(TO) == 2) out[i] = __float2bfloat16_rn(v);
  else out[i] = v;
}

// Per-block feature output for the text-spotting module (pipeline_dit4sr.py:1214-1245): the rows [row0, row0 + n*Hp*Wp) of
// the fp32 residual stream, 'b (N H W) (pH pW d) -> b (N d) (H pH) (W pW)' to fp32 NCHW, and optionally the token-major
// copy the reference returns as `extract_feat` (attention.py:363-374), in one pass.
// One CTA = one (b, n, h) token-grid row and one chunk of 32 channels d: coalesced 128-byte reads along d, shared-memory
// transpose, coalesced row writes along x.
template <typename TT>
__global__ void feature_extract_kernel(const float* __restrict__ x, int64_t ldx, int64_t xbs, int row0, int n_concat, int Hp,
                                       int Wp, int d, float* __restrict__ nchw, TT* __restrict__ tok) {
  extern __shared__ float tile[];   // [2 ph][32 dd][2*Wp + 1]
  const int dd0 = blockIdx.x * 32;
  const int h = blockIdx.y % Hp, n = blockIdx.y / Hp;
  const int b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int W2 = 2 * Wp, pitch = W2 + 1;
  const int D = 4 * d;
  for (int item = warp; item < Wp * 4; item += nwarp) {     // (w, ph, pw)
    const int w = item >> 2, pp = item & 3;
    const int64_t trow = (int64_t)row0 + ((int64_t)n * Hp + h) * Wp + w;
    const float v = x[b * xbs + trow * ldx + pp * d + dd0 + lane];
    tile[((pp >> 1) * 32 + lane) * pitch + 2 * w + (pp & 1)] = v;
    if (tok != nullptr) {
      const int64_t orow = ((int64_t)n * Hp + h) * Wp + w;
      const int64_t o = ((int64_t)b * n_concat * Hp * Wp + orow) * D + pp * d + dd0 + lane;
      if constexpr (sizeof(TT) == 2) tok[o] = __float2bfloat16_rn(v);
      else tok[o] = v;
    }
  }
  __syncthreads();
  if (nchw != nullptr) {
    for (int item = warp; item < 64; item += nwarp) {          // (ph, dd)
      const int ph = item >> 5, dl = item & 31;
      float* dst = nchw + (((int64_t)b * n_concat * d + (int64_t)n * d + dd0 + dl) * (2 * Hp) + 2 * h + ph) * W2;
      for (int xx = lane; xx < W2; xx += 32) dst[xx] = tile[(ph * 32 + dl) * pitch + xx];
    }
  }
}

__global__ void timestep_features_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 128) return;
  const int b = i / 128, k = i % 128;
  // freq = exp(-ln(10000) * k / 128), fp32 as in diffusers get_timestep_embedding
  const float freq = expf(-9.210340371976184f * (float)k / 128.0f);
  const float a = t[b] * freq;
  out[b * 256 + k] = __float2bfloat16_rn(cosf(a));        // flip_sin_to_cos: [cos | sin]
  out[b * 256 + 128 + k] = __float2bfloat16_rn(sinf(a));
}

__global__ void silu_cast_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  y[i] = __float2bfloat16_rn(v / (1.0f + expf(-v)));
}
__global__ void cast_f2b_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __float2bfloat16_rn(x[i]);
}
__global__ void cast_b2f_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __bfloat162float(x[i]);
}

// one thread per 8 bf16 (16 bytes)
__global__ void add_bf16_kernel(__nv_bfloat16* __restrict__ dst, int64_t drs, int64_t dbs,
                                const __nv_bfloat16* __restrict__ src, int64_t srs, int64_t sbs, int B, int R, int C8) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * R * C8) return;
  const int c = (int)(i % C8);
  const int r = (int)((i / C8) % R);
  const int b = (int)(i / ((int64_t)C8 * R));
  uint4* d = reinterpret_cast<uint4*>(dst + b * dbs + r * drs) + c;
  const uint4 s4 = *(reinterpret_cast<const uint4*>(src + b * sbs + r * srs) + c);
  uint4 d4 = *d;
  __nv_bfloat162* dp = reinterpret_cast<__nv_bfloat162*>(&d4);
  const __nv_bfloat162* sp = reinterpret_cast<const __nv_bfloat162*>(&s4);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 a = __bfloat1622float2(dp[k]), bb = __bfloat1622float2(sp[k]);
    dp[k] = __floats2bfloat162_rn(a.x + bb.x, a.y + bb.y);
  }
  *d = d4;
}

// x_next = x + dsigma * v (flow-match Euler step in fp32, result cast back to the latent dtype)
template <typename T>
__global__ void euler_update_kernel(const T* __restrict__ x, const T* __restrict__ v, T* __restrict__ out, float dsigma,
                                    int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float r = to_f<T>(x[i]) + dsigma * to_f<T>(v[i]);
  if constexpr (sizeof(T) == 2) out[i] = __float2bfloat16_rn(r);
  else out[i] = r;
}

static inline unsigned blocks(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace misc

extern "C" {

int dit4sr_patchify(const void* x, int x_dtype, void* patches, int batches, int channels, int height, int width,
                    void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(x && patches && batches > 0 && channels > 0, DIT4SR_E_BADARG, "patchify: bad argument");
  DIT4SR_REQUIRE(height % 2 == 0 && width % 2 == 0 && height > 0 && width > 0, DIT4SR_E_SHAPE,
                 "patchify: height/width must be even");
  DIT4SR_REQUIRE((reinterpret_cast<uintptr_t>(patches) & 7) == 0, DIT4SR_E_ALIGN, "patchify: output not 8-byte aligned");
  const int64_t total = (int64_t)batches * (height / 2) * (width / 2) * channels;
  if (x_dtype == DIT4SR_DT_F32)
    misc::patchify_kernel<float><<<misc::blocks(total, 256), 256, 0, stream>>>(
        (const float*)x, (__nv_bfloat16*)patches, batches, channels, height, width);
  else
    misc::patchify_kernel<__nv_bfloat16><<<misc::blocks(total, 256), 256, 0, stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)patches, batches, channels, height, width);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_unpatchify(const float* tok, void* out, int out_dtype, const void* latents, float dsigma, int batches,
                      int channels, int height, int width, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(tok && out && batches > 0 && channels > 0, DIT4SR_E_BADARG, "unpatchify: bad argument");
  DIT4SR_REQUIRE(height % 2 == 0 && width % 2 == 0 && height > 0 && width > 0, DIT4SR_E_SHAPE,
                 "unpatchify: height/width must be even");
  const int64_t total = (int64_t)batches * channels * height * width;
  if (out_dtype == DIT4SR_DT_F32)
    misc::unpatchify_kernel<float><<<misc::blocks(total, 256), 256, 0, stream>>>(
        tok, (float*)out, (const float*)latents, dsigma, batches, channels, height, width);
  else
    misc::unpatchify_kernel<__nv_bfloat16><<<misc::blocks(total, 256), 256, 0, stream>>>(
        tok, (__nv_bfloat16*)out, (const __nv_bfloat16*)latents, dsigma, batches, channels, height, width);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_unpatchify_cfg(const float* tok, void* out, int out_dtype, const void* latents, float dsigma, float guidance,
                          int batches, int channels, int height, int width, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(tok && out && batches > 0 && channels > 0, DIT4SR_E_BADARG, "unpatchify_cfg: bad argument");
  DIT4SR_REQUIRE(height % 2 == 0 && width % 2 == 0 && height > 0 && width > 0, DIT4SR_E_SHAPE,
                 "unpatchify_cfg: height/width must be even");
  const int64_t total = (int64_t)batches * channels * height * width;
  if (out_dtype == DIT4SR_DT_F32)
    misc::unpatchify_cfg_kernel<float><<<misc::blocks(total, 256), 256, 0, stream>>>(
        tok, (float*)out, (const float*)latents, dsigma, guidance, batches, channels, height, width);
  else
    misc::unpatchify_cfg_kernel<__nv_bfloat16><<<misc::blocks(total, 256), 256, 0, stream>>>(
        tok, (__nv_bfloat16*)out, (const __nv_bfloat16*)latents, dsigma, guidance, batches, channels, height, width);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_feature_extract(const float* x, int64_t ldx, int64_t x_batch_stride, int row0, int batches, int n_concat,
                           int grid_h, int grid_w, int dim, float* nchw_f32, void* tok, int tok_dtype, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(x && (nchw_f32 || tok) && batches > 0 && n_concat > 0 && grid_h > 0 && grid_w > 0 && row0 >= 0,
                 DIT4SR_E_BADARG, "feature_extract: bad argument");
  DIT4SR_REQUIRE(dim > 0 && dim % 128 == 0, DIT4SR_E_SHAPE, "feature_extract: dim must be a multiple of 128 (4 x 32)");
  const int d = dim / 4;
  const size_t smem = (size_t)2 * 32 * (2 * grid_w + 1) * sizeof(float);
  DIT4SR_REQUIRE(smem <= 48 * 1024 && (int64_t)n_concat * grid_h <= 65535 && batches <= 65535, DIT4SR_E_SHAPE,
                 "feature_extract: token grid too large");
  dim3 grid(d / 32, n_concat * grid_h, batches);
  if (tok_dtype == DIT4SR_DT_F32)
    misc::feature_extract_kernel<float><<<grid, 256, smem, stream>>>(x, ldx, x_batch_stride, row0, n_concat, grid_h, grid_w, d,
                                                                     nchw_f32, (float*)tok);
  else
    misc::feature_extract_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(x, ldx, x_batch_stride, row0, n_concat, grid_h,
                                                                             grid_w, d, nchw_f32, (__nv_bfloat16*)tok);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_timestep_features(const float* timesteps, void* out_bf16, int batches, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(timesteps && out_bf16 && batches > 0, DIT4SR_E_BADARG, "timestep_features: bad argument");
  misc::timestep_features_kernel<<<misc::blocks((int64_t)batches * 128, 128), 128, 0, stream>>>(
      timesteps, (__nv_bfloat16*)out_bf16, batches);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_silu_cast(const float* x, void* y_bf16, int64_t n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(x && y_bf16 && n > 0, DIT4SR_E_BADARG, "silu_cast: bad argument");
  misc::silu_cast_kernel<<<misc::blocks(n, 256), 256, 0, stream>>>(x, (__nv_bfloat16*)y_bf16, n);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_add_bf16(void* dst, int64_t dst_row_stride, int64_t dst_batch_stride, const void* src, int64_t src_row_stride,
                    int64_t src_batch_stride, int batches, int rows, int cols, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(dst && src && batches > 0 && rows > 0 && cols > 0, DIT4SR_E_BADARG, "add_bf16: bad argument");
  DIT4SR_REQUIRE(cols % 8 == 0 && dst_row_stride % 8 == 0 && src_row_stride % 8 == 0 && dst_batch_stride % 8 == 0 &&
                     src_batch_stride % 8 == 0 && host::aligned16(dst) && host::aligned16(src),
                 DIT4SR_E_ALIGN, "add_bf16: 16-byte alignment required");
  const int64_t n = (int64_t)batches * rows * (cols / 8);
  misc::add_bf16_kernel<<<misc::blocks(n, 256), 256, 0, stream>>>((__nv_bfloat16*)dst, dst_row_stride, dst_batch_stride,
                                                                  (const __nv_bfloat16*)src, src_row_stride,
                                                                  src_batch_stride, batches, rows, cols / 8);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_euler_update(const void* x, const void* v, void* out, int dtype, float dsigma, int64_t n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(x && v && out && n > 0, DIT4SR_E_BADARG, "euler_update: bad argument");
  if (dtype == DIT4SR_DT_F32)
    misc::euler_update_kernel<float><<<misc::blocks(n, 256), 256, 0, stream>>>((const float*)x, (const float*)v, (float*)out,
                                                                               dsigma, n);
  else
    misc::euler_update_kernel<__nv_bfloat16><<<misc::blocks(n, 256), 256, 0, stream>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)v, (__nv_bfloat16*)out, dsigma, n);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_cast(const void* x, void* y, int64_t n, int dir, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DIT4SR_REQUIRE(x && y && n > 0, DIT4SR_E_BADARG, "cast: bad argument");
  if (dir == 0) misc::cast_f2b_kernel<<<misc::blocks(n, 256), 256, 0, stream>>>((const float*)x, (__nv_bfloat16*)y, n);
  else misc::cast_b2f_kernel<<<misc::blocks(n, 256), 256, 0, stream>>>((const __nv_bfloat16*)x, (float*)y, n);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}
}

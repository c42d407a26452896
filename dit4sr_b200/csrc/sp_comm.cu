// Ulysses sequence parallelism over NVLink peer memory (BASELINE config C4: one 2048px image over P GPUs of one box).
//
// One process per GPU.  Every exchange buffer lives in memory that each rank has exported (CUDA IPC) and every peer has
// mapped, so the head <-> sequence all-to-all of joint attention (the exchange around model_dit4sr/attention.py:528) and
// the 1-row halo exchange of the depthwise stencil (attention.py:168-177) are plain CUDA kernels that STORE straight into
// the destination rank's buffer through NVSwitch - no pack / unpack copies, no library collective, capturable in a
// CUDA graph.  Arrival is signalled with per-(channel, source) monotonic counters in peer memory:
//   writer : data stores ... kernel boundary ... fence.sc.sys ; st.release.sys  flags[dst][ch][me] = epoch
//   reader : spin on ld.acquire.sys flags[me][ch][src] >= epoch for every src, then the stream continues
// The generation counters live in device memory (a graph replay advances them like a direct launch does).
// All ranks must issue the same sequence of exchanges.
#include "host_util.h"
#include <cuda_bf16.h>

namespace spc {

constexpr int MAXR = DIT4SR_SP_MAX_RANKS;
constexpr int CH = DIT4SR_SP_CHANNELS;

struct Peers {
  void* p[MAXR];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// One CTA of 32 threads: bump this rank's generation of `channel`, tell every peer, wait for every peer.
__global__ void barrier_kernel(dit4sr_sp_comm c, int channel) {
  __shared__ uint32_t epoch_s;
  if (threadIdx.x == 0) {
    const uint32_t e = c.epoch[channel] + 1;
    c.epoch[channel] = e;
    epoch_s = e;
  }
  __syncthreads();
  const uint32_t e = epoch_s;
  const int r = threadIdx.x;
  if (r < c.world) {
    __threadfence_system();   // everything this rank stored before the barrier (earlier kernels included) is ordered first
    st_release_sys(c.flags[r] + channel * MAXR + c.rank, e);
    const uint32_t* mine = c.flags[c.rank] + channel * MAXR + r;
    const uint64_t t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
      if (globaltimer_ns() - t0 > 20ull * 1000000000ull) {   // a peer died or the ranks diverged: fail, do not hang
        printf("dit4sr sp_barrier: rank %d channel %d epoch %u: no arrival from rank %d after 20 s\n", c.rank, channel, e, r);
        __trap();
      }
    }
  }
}

// q|k|v of MY image tokens, all heads  ->  every rank's attention-side buffer (its heads, all tokens, source-rank-major),
// plus my own head slice of the (replicated) text tokens.  One thread per 16 bytes of the source.
//   src  [B, I + T, 3 * D]  (row stride src_rs, batch stride src_bs; D = heads * 64)
//   dst  [B, P * I + T, 3, Dl] on every rank (contiguous; Dl = D / P)
__global__ void a2a_qkv_kernel(const __nv_bfloat16* __restrict__ src, int64_t src_rs, int64_t src_bs, Peers dst, int B,
                               int I, int T, int D, int world, int rank) {
  const int D8 = D / 8, Dl = D / world, Dl8 = Dl / 8;
  const int64_t img_units = (int64_t)B * I * 3 * D8;
  const int64_t txt_units = (int64_t)B * T * 3 * Dl8;
  const int64_t rows_dst = (int64_t)world * I + T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < img_units + txt_units;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (i < img_units) {
      const int c8 = (int)(i % D8);
      const int sec = (int)((i / D8) % 3);
      const int s = (int)((i / (3 * D8)) % I);
      const int b = (int)(i / ((int64_t)3 * D8 * I));
      const uint4 v = *reinterpret_cast<const uint4*>(src + b * src_bs + s * src_rs + sec * D + c8 * 8);
      const int p = c8 / Dl8, w8 = c8 - p * Dl8;
      __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst.p[p]);
      *reinterpret_cast<uint4*>(d + ((b * rows_dst + (int64_t)rank * I + s) * 3 + sec) * Dl + w8 * 8) = v;
    } else {
      const int64_t k = i - img_units;
      const int w8 = (int)(k % Dl8);
      const int sec = (int)((k / Dl8) % 3);
      const int t = (int)((k / (3 * Dl8)) % T);
      const int b = (int)(k / ((int64_t)3 * Dl8 * T));
      const uint4 v = *reinterpret_cast<const uint4*>(src + b * src_bs + (int64_t)(I + t) * src_rs + sec * D + rank * Dl + w8 * 8);
      __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst.p[rank]);
      *reinterpret_cast<uint4*>(d + ((b * rows_dst + (int64_t)world * I + t) * 3 + sec) * Dl + w8 * 8) = v;
    }
  }
}

// attention output of MY heads, all tokens (source-rank-major)  ->  every rank's token-side buffer (its tokens, all heads);
// the text rows go to every rank.   src [B, P*I + T, Dl] contiguous;  dst [B, I + T, D] (row stride dst_rs, batch dst_bs)
__global__ void a2a_out_kernel(const __nv_bfloat16* __restrict__ src, Peers dst, int64_t dst_rs, int64_t dst_bs, int B, int I,
                               int T, int D, int world, int rank) {
  const int Dl = D / world, Dl8 = Dl / 8;
  const int64_t rows_src = (int64_t)world * I + T;
  const int64_t units = (int64_t)B * rows_src * Dl8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < units; i += (int64_t)gridDim.x * blockDim.x) {
    const int w8 = (int)(i % Dl8);
    const int64_t row = (i / Dl8) % rows_src;
    const int b = (int)(i / (Dl8 * rows_src));
    const uint4 v = *reinterpret_cast<const uint4*>(src + (b * rows_src + row) * Dl + w8 * 8);
    if (row < (int64_t)world * I) {
      const int r = (int)(row / I), s = (int)(row - (int64_t)r * I);
      __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst.p[r]);
      *reinterpret_cast<uint4*>(d + b * dst_bs + s * dst_rs + rank * Dl + w8 * 8) = v;
    } else {
      const int t = (int)(row - (int64_t)world * I);
      for (int r = 0; r < world; ++r) {
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst.p[r]);
        *reinterpret_cast<uint4*>(d + b * dst_bs + (int64_t)(I + t) * dst_rs + rank * Dl + w8 * 8) = v;
      }
    }
  }
}

// my first / last LR token-grid row -> the halo rows of the neighbours' hidden-state buffers.
//   u [B, rows_total, F] on every rank (row stride rs, batch stride bs, same layout everywhere):
//   token rows [top_halo, top_halo + gw) = halo above, [first, first + n_lr) = my LR rows, [bot_halo, bot_halo+gw) = below
__global__ void halo_kernel(Peers u, int64_t rs, int64_t bs, int B, int F8, int gw, int64_t first, int64_t n_lr,
                            int64_t top_halo, int64_t bot_halo, int world, int rank) {
  const int64_t per = (int64_t)B * gw * F8;
  const __nv_bfloat16* mine = reinterpret_cast<const __nv_bfloat16*>(u.p[rank]);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * per; i += (int64_t)gridDim.x * blockDim.x) {
    const bool down = i >= per;             // my LAST row goes to rank+1's top halo; my FIRST row to rank-1's bottom halo
    const int64_t k = down ? i - per : i;
    const int f8 = (int)(k % F8);
    const int x = (int)((k / F8) % gw);
    const int b = (int)(k / ((int64_t)F8 * gw));
    const int peer = down ? rank + 1 : rank - 1;
    if (peer < 0 || peer >= world) continue;
    const int64_t srow = down ? first + n_lr - gw + x : first + x;
    const int64_t drow = down ? top_halo + x : bot_halo + x;
    const uint4 v = *reinterpret_cast<const uint4*>(mine + b * bs + srow * rs + f8 * 8);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(u.p[peer]) + b * bs + drow * rs + f8 * 8) = v;
  }
}

static int check_comm(const dit4sr_sp_comm* c, const char* what) {
  DIT4SR_REQUIRE(c != nullptr, DIT4SR_E_BADARG, "%s: null communicator", what);
  DIT4SR_REQUIRE(c->world >= 1 && c->world <= MAXR && c->rank >= 0 && c->rank < c->world, DIT4SR_E_BADARG,
                 "%s: bad world/rank (%d/%d)", what, c->world, c->rank);
  DIT4SR_REQUIRE(c->epoch != nullptr, DIT4SR_E_BADARG, "%s: communicator without epoch counters", what);
  for (int r = 0; r < c->world; ++r)
    DIT4SR_REQUIRE(c->flags[r] != nullptr, DIT4SR_E_BADARG, "%s: flags of rank %d not mapped", what, r);
  return 0;
}

static int grid_for(int64_t units) {
  int64_t g = (units + 255) / 256;
  const int64_t cap = (int64_t)host::sm_count() * 8;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace spc

extern "C" {

int dit4sr_peer_alloc(size_t bytes, void** ptr) {
  DIT4SR_REQUIRE(ptr != nullptr && bytes > 0, DIT4SR_E_BADARG, "peer_alloc: bad argument");
  DIT4SR_CUDA(cudaMalloc(ptr, bytes));
  DIT4SR_CUDA(cudaMemset(*ptr, 0, bytes));
  DIT4SR_CUDA(cudaDeviceSynchronize());
  return 0;
}

int dit4sr_peer_free(void* ptr) {
  if (ptr != nullptr) DIT4SR_CUDA(cudaFree(ptr));
  return 0;
}

int dit4sr_peer_export(const void* ptr, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  DIT4SR_REQUIRE(ptr != nullptr && handle != nullptr, DIT4SR_E_BADARG, "peer_export: null argument");
  cudaIpcMemHandle_t h;
  DIT4SR_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle, &h, 64);
  return 0;
}

int dit4sr_peer_open(const unsigned char handle[64], void** ptr) {
  DIT4SR_REQUIRE(ptr != nullptr && handle != nullptr, DIT4SR_E_BADARG, "peer_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  DIT4SR_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int dit4sr_peer_close(void* ptr) {
  if (ptr != nullptr) DIT4SR_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int dit4sr_sp_barrier(const dit4sr_sp_comm* comm, int channel, void* stream_) {
  int rc = spc::check_comm(comm, "sp_barrier");
  if (rc != 0) return rc;
  DIT4SR_REQUIRE(channel >= 0 && channel < spc::CH, DIT4SR_E_BADARG, "sp_barrier: channel %d out of range", channel);
  if ((rc = dit4sr_check_device()) != 0) return rc;
  spc::barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(*comm, channel);
  DIT4SR_CUDA(cudaGetLastError());
  return 0;
}

int dit4sr_a2a_qkv(const dit4sr_sp_comm* comm, const dit4sr_a2a_params* p, void* stream_) {
  int rc = spc::check_comm(comm, "a2a_qkv");
  if (rc != 0) return rc;
  DIT4SR_REQUIRE(p, DIT4SR_E_BADARG, "a2a_qkv: null argument");
  if (p->src == nullptr)   // the projections' epilogues already stored into the peers (dit4sr_gemm_params.sp)
    return p->no_barrier ? 0 : dit4sr_sp_barrier(comm, DIT4SR_SP_CH_QKV, stream_);
  DIT4SR_REQUIRE(p->batches > 0 && p->image_rows > 0 && p->text_rows >= 0 && p->heads > 0 && p->heads % comm->world == 0,
                 DIT4SR_E_BADARG, "a2a_qkv: bad sizes (heads %d must divide over %d ranks)", p->heads, comm->world);
  DIT4SR_REQUIRE(p->src_row_stride % 8 == 0 && p->src_batch_stride % 8 == 0 && host::aligned16(p->src), DIT4SR_E_ALIGN,
                 "a2a_qkv: 16-byte alignment required");
  spc::Peers peers;
  for (int r = 0; r < spc::MAXR; ++r) peers.p[r] = r < comm->world ? p->peer[r] : nullptr;
  for (int r = 0; r < comm->world; ++r)
    DIT4SR_REQUIRE(peers.p[r] != nullptr && host::aligned16(peers.p[r]), DIT4SR_E_BADARG, "a2a_qkv: peer buffer %d missing", r);
  if ((rc = dit4sr_check_device()) != 0) return rc;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int D = p->heads * 64;
  const int64_t units = (int64_t)p->batches * 3 * ((int64_t)p->image_rows * (D / 8) + (int64_t)p->text_rows * (D / comm->world / 8));
  spc::a2a_qkv_kernel<<<spc::grid_for(units), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(p->src), p->src_row_stride, p->src_batch_stride, peers, p->batches,
      p->image_rows, p->text_rows, D, comm->world, comm->rank);
  DIT4SR_CUDA(cudaGetLastError());
  return p->no_barrier ? 0 : dit4sr_sp_barrier(comm, DIT4SR_SP_CH_QKV, stream_);
}

int dit4sr_a2a_out(const dit4sr_sp_comm* comm, const dit4sr_a2a_params* p, void* stream_) {
  int rc = spc::check_comm(comm, "a2a_out");
  if (rc != 0) return rc;
  DIT4SR_REQUIRE(p, DIT4SR_E_BADARG, "a2a_out: null argument");
  if (p->src != nullptr) {   // src == NULL: the attention kernel already stored into the peers (sp_out): barrier only
    DIT4SR_REQUIRE(p->batches > 0 && p->image_rows > 0 && p->text_rows >= 0 && p->heads > 0 && p->heads % comm->world == 0,
                   DIT4SR_E_BADARG, "a2a_out: bad sizes");
    DIT4SR_REQUIRE(p->dst_row_stride % 8 == 0 && p->dst_batch_stride % 8 == 0 && host::aligned16(p->src), DIT4SR_E_ALIGN,
                   "a2a_out: 16-byte alignment required");
    spc::Peers peers;
    for (int r = 0; r < spc::MAXR; ++r) peers.p[r] = r < comm->world ? p->peer[r] : nullptr;
    for (int r = 0; r < comm->world; ++r)
      DIT4SR_REQUIRE(peers.p[r] != nullptr && host::aligned16(peers.p[r]), DIT4SR_E_BADARG, "a2a_out: peer buffer %d missing", r);
    if ((rc = dit4sr_check_device()) != 0) return rc;
    const int D = p->heads * 64;
    const int64_t units = (int64_t)p->batches * ((int64_t)comm->world * p->image_rows + p->text_rows) * (D / comm->world / 8);
    spc::a2a_out_kernel<<<spc::grid_for(units), 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<const __nv_bfloat16*>(p->src), peers, p->dst_row_stride, p->dst_batch_stride, p->batches,
        p->image_rows, p->text_rows, D, comm->world, comm->rank);
    DIT4SR_CUDA(cudaGetLastError());
  }
  return p->no_barrier ? 0 : dit4sr_sp_barrier(comm, DIT4SR_SP_CH_OUT, stream_);
}

int dit4sr_halo_rows(const dit4sr_sp_comm* comm, const dit4sr_halo_params* p, void* stream_) {
  int rc = spc::check_comm(comm, "halo_rows");
  if (rc != 0) return rc;
  DIT4SR_REQUIRE(p, DIT4SR_E_BADARG, "halo_rows: null argument");
  DIT4SR_REQUIRE(p->batches > 0 && p->channels > 0 && p->channels % 8 == 0 && p->grid_w > 0 && p->lr_rows >= p->grid_w &&
                     p->row_stride % 8 == 0 && p->batch_stride % 8 == 0,
                 DIT4SR_E_BADARG, "halo_rows: bad sizes");
  spc::Peers peers;
  for (int r = 0; r < spc::MAXR; ++r) peers.p[r] = r < comm->world ? p->peer[r] : nullptr;
  for (int r = 0; r < comm->world; ++r)
    DIT4SR_REQUIRE(peers.p[r] != nullptr && host::aligned16(peers.p[r]), DIT4SR_E_BADARG, "halo_rows: peer buffer %d missing", r);
  if ((rc = dit4sr_check_device()) != 0) return rc;
  const int64_t units = 2ll * p->batches * p->grid_w * (p->channels / 8);
  spc::halo_kernel<<<spc::grid_for(units), 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      peers, p->row_stride, p->batch_stride, p->batches, p->channels / 8, p->grid_w, p->first_lr_row, p->lr_rows,
      p->top_halo_row, p->bottom_halo_row, comm->world, comm->rank);
  DIT4SR_CUDA(cudaGetLastError());
  return p->no_barrier ? 0 : dit4sr_sp_barrier(comm, DIT4SR_SP_CH_HALO, stream_);
}

}  // extern "C"

/* dit4sr_b200 -- C ABI of the B200 (sm_100a) kernels behind DiT4SR's denoising hot path.
 *
 * This is the drop-in boundary: plain C entry points, plain pointers and sizes, no torch types.
 * The Python host (dit4sr_b200/ops.py) binds them with ctypes; INTEGRATION.md shows the binding a reference
 * maintainer would add.  Each entry point names the reference code it replaces (paths relative to the
 * jinlovespho/DiT4SR repository).
 *
 * Conventions
 *   - All pointers are DEVICE pointers owned by the caller (e.g. PyTorch's caching allocator); the library never
 *     allocates or frees device memory and keeps no pointer after a call returns.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*); no hidden synchronisation.
 *   - Return value: 0 = success; negative = DIT4SR_E_* (bad argument / unsupported device); positive = cudaError_t.
 *     dit4sr_last_error() returns a thread-local, human-readable description of the last failure.
 *   - "bf16" = __nv_bfloat16 (uint16_t storage); all strides/leading dimensions are in ELEMENTS.
 *   - Re-entrant and thread-safe (no mutable global state besides one-time function-attribute setup).
 */
#ifndef DIT4SR_B200_H_
#define DIT4SR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIT4SR_ABI_VERSION 5

enum {
  DIT4SR_OK = 0,
  DIT4SR_E_BADARG = -1,      /* null pointer, non-positive size */
  DIT4SR_E_ALIGN = -2,       /* pointer / stride alignment not met */
  DIT4SR_E_SHAPE = -3,       /* shape not supported by the kernel (documented per entry point) */
  DIT4SR_E_ARCH = -4,        /* device is not compute capability 10.x */
  DIT4SR_E_DRIVER = -5       /* cuTensorMapEncodeTiled unavailable / failed */
};

int dit4sr_abi_version(void);
const char* dit4sr_last_error(void);
/* 0 if the current device can run the kernels (compute capability 10.x), DIT4SR_E_ARCH otherwise. */
int dit4sr_check_device(void);

/* ---------------------------------------------------------------------------------------------------------------
 * LayerNorm (no affine, eps) fused with adaLN modulation:  out = LN(x) * (1 + scale[b]) + shift[b]
 * Replaces: diffusers AdaLayerNormZero / SD35AdaLayerNormZeroX / AdaLayerNormContinuous normalisation + modulation
 * as called at model_dit4sr/attention.py:286-302,322-323,384-385 and model_dit4sr/transformer_sd3.py:382.
 * One read of x, one or two bf16 outputs (the dual-attention blocks need two modulations of the same LN(x)).
 * x: fp32 [batches, rows, dim] (row stride ldx, batch stride x_batch_stride). dim % 128 == 0, dim <= 4096.
 * shift/scale: fp32, one vector of `dim` per batch (stride mod_batch_stride).  out2/shift2/scale2 may be NULL.
 */
typedef struct {
  const float* x; int64_t ldx; int64_t x_batch_stride;
  const float* shift; const float* scale;
  const float* shift2; const float* scale2;
  int64_t mod_batch_stride;
  void* out; void* out2;               /* bf16 */
  int64_t ldo; int64_t o_batch_stride;
  int batches; int rows; int dim;
  float eps;
  int out_f32;                         /* 1: out/out2 are fp32 (fp32-accurate mode) instead of bf16 */
  /* ABI v5.  Optional companion problem served by the same launch (rows_b > 0): a second fp32 [batches, rows_b, dim] tensor
   * with its own modulation vectors (same mod_batch_stride) and one output of the same kind as `out`.  The engine passes
   * the text stream here (norm1_context / the text half of norm2, attention.py:294-302,384-385) next to the image stream,
   * which halves the number of LayerNorm launches per block; a few thousand text rows alone cannot fill the GPU. */
  const float* x_b; int64_t ldx_b; int64_t x_b_batch_stride;
  const float* shift_b; const float* scale_b;
  void* out_b; int64_t ldo_b; int64_t o_b_batch_stride;
  int rows_b;
} dit4sr_ln_modulate_params;
int dit4sr_ln_modulate(const dit4sr_ln_modulate_params* p, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Depthwise 3x3 convolution (zero pad 1, bias) of the LR half of the MLP hidden state, added into the HQ half,
 * in place, directly on the token-major layout (tokens = rows of a grid x grid image, channels innermost):
 *     u_hq[b, y, x, c] += bias[c] + sum_{dy,dx} w[c, dy, dx] * u_lr[b, y+dy-1, x+dx-1, c]
 * Replaces: FeedForwardControl's reshape/permute + control_conv + add + cat, model_dit4sr/attention.py:168-177.
 * u_hq / u_lr: bf16 [batches, grid*grid, channels] (row stride ld, batch stride given per pointer).
 * weight: fp32 repacked tap-major [9, channels] (tap = dy*3+dx); bias fp32 [channels]. channels % 8 == 0.
 */
typedef struct {
  void* u_hq; const void* u_lr;
  int64_t ld; int64_t hq_batch_stride; int64_t lr_batch_stride;
  const float* weight9; const float* bias;
  int batches; int grid; int channels;
  /* Rectangular / row-sharded grids (sequence parallelism: a rank owns `grid` rows of `grid_w` columns and holds the LR
   * stream with halo rows). All zero = square grid x grid image without halo.
   *   grid_w        columns per row (0 -> grid)
   *   lr_grid_h     rows present in u_lr (0 -> grid)
   *   lr_row_offset u_lr row that lies under u_hq row 0 (rows outside [0, lr_grid_h) read as zero padding) */
  int grid_w; int lr_grid_h; int lr_row_offset;
} dit4sr_dwconv3x3_params;
int dit4sr_dwconv3x3_add(const dit4sr_dwconv3x3_params* p, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Batched-row GEMM on tcgen05 tensor cores:  C[b, r, :] = epilogue( A[b, r, :] . W^T )
 *   A: bf16 [batches, rows, K]  (lda, a_batch_stride)     W: bf16 [N, K] (ldw; nn.Linear weight layout)
 * Replaces every nn.Linear on the path (model_dit4sr/attention.py:463-465,475-477,494-496,539,548,552, the
 * GELU/Linear pair of FeedForward(Control) :94-100,:165-166, the adaLN `linear`s, context_embedder / proj_out /
 * PatchEmbed.proj / time_text_embed in transformer_sd3.py:326-331,382-383) together with the elementwise work
 * that follows it in the reference, selected by `epilogue`:
 *   DIT4SR_EPI_BIAS       C = acc + bias                                   (C bf16 or fp32)
 *   DIT4SR_EPI_BIAS_GELU  C = gelu_tanh(acc + bias)                        (diffusers GELU(approximate="tanh"); tanh.approx for
 *                         bf16 C, tanhf for fp32 C)
 *   DIT4SR_EPI_BIAS_SILU  C = silu(acc + bias)
 *   DIT4SR_EPI_QKV_NORM   per 64-wide head: y = acc + bias; columns [0,sec_cols) use RMSNorm weight norm_w0,
 *                         [sec_cols,2*sec_cols) norm_w1, the rest none (q|k|v fused projection followed by
 *                         attn.norm_q / attn.norm_k, attention.py:484-489,502-505). If dbl_row_offset != 0 the k
 *                         section is normalised a second time and also written dbl_row_offset rows further down
 *                         (the reference's `key_control = attn.norm_k(key)` quirk, attention.py:489).
 *   DIT4SR_EPI_GATE_RESID C(fp32) = resid + gate[b,:] * (acc + bias)        (gated residual, attention.py:314-320,
 *                         329-347,381-393; gate NULL means 1; resid may alias C or be a broadcast tensor)
 * Constraints: K % 8 == 0, N % 32 == 0 (N % 64 == 0 for QKV_NORM), 16-byte aligned pointers / strides.
 */
enum {
  DIT4SR_EPI_BIAS = 0,
  DIT4SR_EPI_BIAS_GELU = 1,
  DIT4SR_EPI_BIAS_SILU = 2,
  DIT4SR_EPI_QKV_NORM = 3,
  DIT4SR_EPI_GATE_RESID = 4
};
enum { DIT4SR_DT_BF16 = 0, DIT4SR_DT_F32 = 1 };
#define DIT4SR_SP_MAX_RANKS 8
/* Ulysses sequence parallelism: a q|k|v projection (BIAS / QKV_NORM epilogue, bf16 C) may store its output heads
 * straight into the attention-side buffers of the ranks that own them instead of C (the sequence -> head all-to-all fused
 * into the GEMM epilogue; NVLink peer stores).  Column c of this problem is column col0 + c of the [3 * inner_dim] q|k|v
 * row: section s = col / inner_dim, head owner p = (col % inner_dim) / (inner_dim / world).  Row r of batch b lands in
 *   peer[p] [b, row0 + r, s, col % (inner_dim / world)]   of a bf16 [batches, rows_dst, 3, inner_dim / world] buffer.
 * own_heads_only: the rows are replicated on every rank (text tokens): store only this rank's own heads (into peer[rank]).
 * C / ldc still describe a valid local tensor (shape checks) but are not written.  Follow with dit4sr_a2a_qkv(src = NULL)
 * (arrival barrier).  All problems of one grouped launch must name the same peer buffers. */
typedef struct {
  int world; int rank; int own_heads_only; int inner_dim; int col0;
  int64_t rows_dst; int64_t row0;
  void* peer[DIT4SR_SP_MAX_RANKS];
} dit4sr_gemm_sp;
typedef struct {
  const void* A; int64_t lda; int64_t a_batch_stride;
  const void* W; int64_t ldw;
  const float* bias;                       /* [N] or NULL */
  void* C; int64_t ldc; int64_t c_batch_stride; int c_dtype;
  int batches; int rows; int N; int K;
  int epilogue;
  const float* gate; int64_t gate_batch_stride;           /* GATE_RESID */
  const float* resid; int64_t ldr; int64_t r_batch_stride;
  const float* norm_w0; const float* norm_w1;             /* QKV_NORM: fp32 [64] each (NULL = no norm) */
  int sec_cols; float eps; int64_t dbl_row_offset;
  const dit4sr_gemm_sp* sp;                               /* NULL: plain store into C */
  /* ABI v4.  Implicit-GEMM convolution (the VAE's 3x3 / strided convolutions, row N4): conv_w > 0 makes A an NHWC image
   * bf16 [batches, conv_h, conv_w, C_in] (pixel stride lda, C_in % 64 == 0), rows = conv_h * conv_w output pixels in
   * raster order, K = conv_taps * C_in with W laid out [N, tap, C_in].  K-block kb belongs to tap t = kb / (C_in / 64) and
   * is read from the image shifted by (t / conv_kw + conv_org) rows and (t % conv_kw + conv_org) columns; pixels outside
   * the image read as zero (TMA out-of-bounds fill = the convolution's zero padding).  3x3, pad 1: conv_kw 3, conv_taps 9,
   * conv_org -1.  conv_w must divide 128 or be a multiple of 128 (other widths: pitched images, conv_w_valid). */
  int conv_w; int conv_h; int conv_kw; int conv_taps; int conv_org;
  int conv_w_valid;   /* 0 = conv_w; else the image is pitched: rows of conv_w pixels of which the first conv_w_valid exist */
  /* BIAS epilogue with bf16 C: C = acc + bias + resid_bf16 (same row/column indexing as C; NULL = none) */
  const void* resid_bf16; int64_t ld_rb; int64_t rb_batch_stride;
} dit4sr_gemm_params;
int dit4sr_gemm_bf16(const dit4sr_gemm_params* p, void* stream);
/* Up to DIT4SR_GEMM_MAX_GROUP independent problems sharing one epilogue kind, run by ONE persistent launch (e.g. the HQ,
 * LR and text projections of one attention: attention.py:463-465,475-477,494-496). Shapes, pointers and epilogue operands
 * are per problem; tiles of problem i are scheduled before those of problem i+1. */
#define DIT4SR_GEMM_MAX_GROUP 6
int dit4sr_gemm_bf16_grouped(const dit4sr_gemm_params* problems, int count, void* stream);
/* Host-only query (no device, no driver; the pointers inside `problems` are not dereferenced): the work-item schedule
 * dit4sr_gemm_bf16_grouped walks on a GPU with `sm_count` SMs - 128-row x 256-column tiles in problem order, the last
 * (tiles mod sm_count) of them cut into 2 or 4 column slices when that fills the final wave better.  Writes up to
 * max_items records of five ints {problem, batch, row_tile, first_column, columns} in schedule order (work item i runs
 * on CTA i mod min(sm_count, items)) and returns the number of work items, or a negative DIT4SR_E_* code. */
int dit4sr_gemm_schedule(const dit4sr_gemm_params* problems, int count, int sm_count, int* items, int max_items);

/* ---------------------------------------------------------------------------------------------------------------
 * Joint attention, flash-style on tcgen05:  O = softmax(Q K^T * scale) V, non-causal, no mask, head_dim 64.
 * Replaces F.scaled_dot_product_attention over the concatenated [HQ | LR | text] sequence
 * (model_dit4sr/attention.py:508-530) plus the LR-stream residual add at :545.
 * q/k/v: bf16, element (b, s, h, d) at  ptr + b*batch_stride + s*row_stride + h*64 + d   (the fused projection
 * buffer [B, S, 3*H*64] is addressed with row_stride = 3*H*64 and q/k/v pointers 0 / H*64 / 2*H*64 apart).
 * o: bf16 [B, Sq, H*64] (row stride o_row_stride).  Only the first seq_q query rows are computed (the last block
 * drops the text queries).  If lr_res != NULL, rows [lr_row_begin, lr_row_end) of o get lr_res[b, row, :] added.
 *
 * mode (ABI v3; selection is per call, nothing is read from the environment):
 *   DIT4SR_ATTN_ONLINE     online softmax with the running row max as reference: valid for any q, k.
 *   DIT4SR_ATTN_PRESCALED  the caller has folded scale * log2(e) into q (`scale` is ignored) and guarantees
 *                          q.k <= score_bound <= DIT4SR_ATTN_MAX_SCORE_BOUND for every (query, key) pair; the kernel
 *                          then takes p = 2^(q.k) with no reference, no row max and no rescale of O.  With QK-RMSNorm
 *                          (attention.py:484-489,502-505) every normed row has |x| <= 8 max|w|, so the bound follows
 *                          from the checkpoint's norm weights; the host refuses the mode when it does not hold.
 * exp_offload: 0 = default; 1, 2, 3 = 0/8, 2/8, 4/8 of the exponentials evaluated by a polynomial on the FMA pipe instead
 *              of MUFU.EX2 (tuning knob: same results to within bf16 rounding of P).
 * variant:     5 = 4 with two softmax warpgroups per tile, each owning half of every 96-key score row (pre-scaled form only;
 *              an online-mode call asking for 5 runs 4);
 *              4 = one 128-row query tile per CTA, two CTAs per SM, double-buffered 96-key score tiles; 3 = the same tiles,
 *              two query tiles per CTA sharing K/V; 2 = three 128-key score buffers rotating over the two query tiles of
 *              a CTA (round 1's pipeline); 0 = default: the cheapest of the three under a fitted cost model (waves x
 *              (KV steps + per-CTA overhead) x step time): 4 (5 with pre-scaled scores) for short and medium sequences, 3 for very
 *              long ones.
 * workspace:   optional scratch of dit4sr_attn_workspace_bytes() bytes (contents irrelevant).  With it, work items that would
 *              only partly fill the last wave of SMs are cut into 2-4 slices along the keys; each slice publishes its
 *              unnormalised row + row sum (+ reference max) and a small second kernel merges them (e.g. 390 items on 148
 *              SMs: 2.67 instead of 3 waves).  Without it, or with split_kv < 0, every item runs over all keys.
 *              split_kv > 0 forces that many slices (tests).
 * sp_out (Ulysses sequence parallelism, NULL otherwise): the kernel's epilogue stores each output row straight into the
 *              token-side buffer of the rank that owns the token, through NVLink peer memory (the head -> sequence
 *              all-to-all fused into the attention kernel).  Query row i < world*rows_per_rank belongs to rank
 *              i / rows_per_rank and lands in row i % rows_per_rank of o_peer[rank]; the remaining (text) rows land in
 *              rows rows_per_rank.. of EVERY o_peer.  Columns: col_offset + head*64 (col_offset = my first head * 64).
 *              o / o_row_stride / o_batch_stride describe the (identical) layout of every o_peer buffer; lr_res must be
 *              NULL.  Follow with dit4sr_a2a_out(src = NULL) (arrival barrier) before reading the buffers.
 */
enum { DIT4SR_ATTN_ONLINE = 0, DIT4SR_ATTN_PRESCALED = 1 };
#define DIT4SR_ATTN_MAX_SCORE_BOUND 60.0f
typedef struct {
  int world; int rows_per_rank; int col_offset;
  void* o_peer[DIT4SR_SP_MAX_RANKS];
} dit4sr_attn_sp_out;
typedef struct {
  const void* q; const void* k; const void* v;
  int64_t qkv_row_stride; int64_t qkv_batch_stride;
  void* o; int64_t o_row_stride; int64_t o_batch_stride;
  const void* lr_res; int64_t lr_row_stride; int64_t lr_batch_stride;   /* bf16, indexed by the same row ids as o */
  int lr_row_begin; int lr_row_end;
  int batches; int heads; int seq_q; int seq_kv;
  float scale;
  int mode; float score_bound; int exp_offload; int variant;             /* ABI v3 */
  const dit4sr_attn_sp_out* sp_out;
  void* workspace; int64_t workspace_bytes; int split_kv;
} dit4sr_attn_params;
int dit4sr_attn_fwd(const dit4sr_attn_params* p, void* stream);
int64_t dit4sr_attn_workspace_bytes(void);   /* for the current device (148 SMs assumed when there is none) */
/* Host-only query (no device; pointers are only compared with NULL): the work decomposition dit4sr_attn_fwd would use on a
 * GPU with sm_count SMs: plan = {kernel variant, KV slices per tail item, whole items, tail items, KV tiles, KV tiles per
 * slice}.  workspace_bytes is compared with 4 * 256 * 68 * 4 bytes per SM. */
int dit4sr_attn_plan(const dit4sr_attn_params* p, int sm_count, int plan[6]);
/* Diagnostics: {kernel variant, KV slices per tail item, whole items, tail items} the latest dit4sr_attn_fwd call of this
 * process chose (host-side bookkeeping, not thread-safe). */
int dit4sr_attn_last_plan(int plan[4]);

/* ---------------------------------------------------------------------------------------------------------------
 * Ulysses sequence parallelism over NVLink peer memory (BASELINE config C4; the reference has no counterpart: it runs
 * F.scaled_dot_product_attention on one GPU, model_dit4sr/attention.py:528).  One process per GPU of one box.  Every rank
 * owns grid/P rows of the token grid of each image stream; around each joint attention the q|k|v of its tokens are
 * scattered to the ranks that own the heads and the outputs come back, and the depthwise stencil (attention.py:168-177)
 * needs one halo row from each neighbour.  All exchange buffers are allocated with dit4sr_peer_alloc on every rank,
 * exported (CUDA IPC) and mapped by the peers; the kernels store straight into the destination's memory and signal
 * arrival through per-(channel, source) counters, so an exchange is one or two kernel launches on the caller's stream:
 * no pack/unpack copies, no library collective, CUDA-graph capturable.  Every rank must issue the same sequence of calls.
 *
 * The communicator is passed in explicitly (no global state):
 *   flags[r]  rank r's arrival counters as mapped in THIS process: uint32 [DIT4SR_SP_CHANNELS][DIT4SR_SP_MAX_RANKS],
 *             zero-initialised (dit4sr_peer_alloc does that); flags[rank] is the local allocation itself
 *   epoch     local device memory, uint32 [DIT4SR_SP_CHANNELS], zero-initialised: generation per channel
 */
#define DIT4SR_SP_CHANNELS 4
enum { DIT4SR_SP_CH_QKV = 0, DIT4SR_SP_CH_OUT = 1, DIT4SR_SP_CH_HALO = 2, DIT4SR_SP_CH_USER = 3 };
typedef struct {
  int world; int rank;
  uint32_t* flags[DIT4SR_SP_MAX_RANKS];
  uint32_t* epoch;
} dit4sr_sp_comm;
/* Peer-visible device memory: allocate (zero-filled), export a 64-byte handle for the other processes, map a peer's
 * handle (enables peer access lazily), unmap, free.  Host-side setup calls; they synchronise the device. */
int dit4sr_peer_alloc(size_t bytes, void** ptr);
int dit4sr_peer_free(void* ptr);
int dit4sr_peer_export(const void* ptr, unsigned char handle[64]);
int dit4sr_peer_open(const unsigned char handle[64], void** ptr);
int dit4sr_peer_close(void* ptr);
/* Signal every peer on `channel` and wait for every peer's signal (also orders all earlier stores of this rank's stream
 * to peer memory before the signal).  A peer that never arrives traps the kernel after 20 s instead of hanging. */
int dit4sr_sp_barrier(const dit4sr_sp_comm* comm, int channel, void* stream);
/* Head <-> sequence exchange.
 *  dit4sr_a2a_qkv: src = this rank's fused projection buffer, bf16 [batches, image_rows + text_rows, 3 * heads*64]
 *     (src_row_stride / src_batch_stride in elements).  Image rows go to EVERY rank p as
 *     peer[p][b, rank*image_rows + s, {q,k,v}, Dl] with the columns of p's heads (Dl = heads*64 / world); the text rows are
 *     replicated, so only this rank's own head slice is copied, to peer[rank][b, world*image_rows + t, ...].
 *     peer[p]: bf16 [batches, world*image_rows + text_rows, 3, Dl] contiguous, the same shape on every rank.
 *     src == NULL: the projections already stored into the peers (dit4sr_gemm_params.sp): arrival barrier only.
 *  dit4sr_a2a_out: the way back.  src = bf16 [batches, world*image_rows + text_rows, Dl] contiguous (my heads, all tokens)
 *     -> peer[r][b, s, rank*Dl ...] for the image rows rank r owns, and the text rows to every peer at rows image_rows.. ;
 *     peer[r]: bf16 [batches, >= image_rows + text_rows, heads*64] with dst_row_stride / dst_batch_stride.
 *     src == NULL: nothing to move (dit4sr_attn_fwd already stored into the peers through sp_out): arrival barrier only.
 *  Unless no_barrier is set each call ends with dit4sr_sp_barrier on its own channel, i.e. when the call's work completes
 *  on the stream, everything the PEERS sent to this rank in their matching call has arrived too. */
typedef struct {
  const void* src; int64_t src_row_stride; int64_t src_batch_stride;
  void* peer[DIT4SR_SP_MAX_RANKS];
  int64_t dst_row_stride; int64_t dst_batch_stride;      /* a2a_out only */
  int batches; int image_rows; int text_rows; int heads;
  int no_barrier;
} dit4sr_a2a_params;
int dit4sr_a2a_qkv(const dit4sr_sp_comm* comm, const dit4sr_a2a_params* p, void* stream);
int dit4sr_a2a_out(const dit4sr_sp_comm* comm, const dit4sr_a2a_params* p, void* stream);
/* Halo exchange of the LR hidden state for the depthwise stencil.  peer[r] = rank r's MLP hidden-state buffer, bf16
 * [batches, rows, channels] (row_stride / batch_stride in elements, the same layout on every rank), which holds this
 * rank's LR token rows at [first_lr_row, first_lr_row + lr_rows) and one grid row (grid_w tokens) of halo above at
 * top_halo_row and below at bottom_halo_row.  Stores my first LR grid row into rank-1's bottom halo and my last into
 * rank+1's top halo; the outer halos of the first / last rank are never written (they stay zero = the conv's padding). */
typedef struct {
  void* peer[DIT4SR_SP_MAX_RANKS];
  int64_t row_stride; int64_t batch_stride;
  int batches; int channels; int grid_w;
  int64_t first_lr_row; int64_t lr_rows; int64_t top_halo_row; int64_t bottom_halo_row;
  int no_barrier;
} dit4sr_halo_params;
int dit4sr_halo_rows(const dit4sr_sp_comm* comm, const dit4sr_halo_params* p, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Small data-movement kernels of the model head / tail (transformer_sd3.py:326-332,379-391).
 */
/* im2col for PatchEmbed's k=2,s=2 conv: x [B, C, H, W] (fp32 or bf16) -> patches bf16 [B, (H/2)*(W/2), C*4],
 * inner order (c, ph, pw) = the flattened conv weight [D, C, 2, 2]. */
int dit4sr_patchify(const void* x, int x_dtype, void* patches, int batches, int channels, int height, int width,
                    void* stream);
/* unpatchify 'b (H W) (pH pW d) -> b d (H pH) (W pW)': tok fp32 [B, (H/2)*(W/2), 4*C] -> out [B, C, H, W].
 * If `latents` != NULL performs the flow-match Euler update instead of a plain store:
 *   out = latents + dsigma * velocity   (pipelines/pipeline_dit4sr.py:1567, scheduler.step) */
int dit4sr_unpatchify(const float* tok, void* out, int out_dtype, const void* latents, float dsigma, int batches,
                      int channels, int height, int width, void* stream);
/* The same for a classifier-free-guidance step (pipelines/pipeline_dit4sr.py:1092-1108,1561-1563): tok holds the velocity
 * tokens of the doubled batch [uncond (batches) | text (batches)]; v = v_uncond + guidance * (v_text - v_uncond), then the
 * plain store or the Euler update as above.  out / latents: [batches, C, H, W]. */
int dit4sr_unpatchify_cfg(const float* tok, void* out, int out_dtype, const void* latents, float dsigma, float guidance,
                          int batches, int channels, int height, int width, void* stream);
/* Per-block feature outputs for the text-spotting module: rows [row0, row0 + n_concat*grid_h*grid_w) of the fp32 residual
 * stream x [batches, rows, dim] (ldx, x_batch_stride) are written as
 *   tok       [batches, n_concat*grid_h*grid_w, dim]  (tok_dtype; the `extract_feat` slice of model_dit4sr/attention.py:363-374)
 *   nchw_f32  [batches, n_concat*dim/4, 2*grid_h, 2*grid_w] fp32 = rearrange('b (N H W) (pH pW d) -> b (N d) (H pH) (W pW)')
 *             followed by .to(float32), i.e. what pipelines/pipeline_dit4sr.py:1214-1245 feeds to the ts_module.
 * Either output may be NULL. dim % 128 == 0. */
int dit4sr_feature_extract(const float* x, int64_t ldx, int64_t x_batch_stride, int row0, int batches, int n_concat,
                           int grid_h, int grid_w, int dim, float* nchw_f32, void* tok, int tok_dtype, void* stream);
/* sinusoidal timestep features [cos | sin] (diffusers Timesteps(256, flip_sin_to_cos=True, shift 0)) -> bf16 [B, 256] */
int dit4sr_timestep_features(const float* timesteps, void* out_bf16, int batches, void* stream);
/* y = silu(x) (fp32 -> bf16), n elements */
int dit4sr_silu_cast(const float* x, void* y_bf16, int64_t n, void* stream);
/* dst[b, r, :] += src[b, r, :] for bf16 [batches, rows, cols] views (cols % 8 == 0, strides in elements): the LR-stream
 * residual add of attention.py:545 when the attention output arrives through the sequence-parallel all-to-all. */
int dit4sr_add_bf16(void* dst, int64_t dst_row_stride, int64_t dst_batch_stride, const void* src, int64_t src_row_stride,
                    int64_t src_batch_stride, int batches, int rows, int cols, void* stream);
/* ---------------------------------------------------------------------------------------------------------------
 * fp32-accurate mode ("bf16x3"): activations stay fp32 between kernels and every tensor-core product A W^T is evaluated
 * as Ah Wh^T + Ah Wl^T + Al Wh^T (hi/lo bf16 splits) by concatenating the three terms along K, so the unchanged
 * dit4sr_gemm_bf16 runs on A' = [Ah|Ah|Al], W' = [Wh|Wl|Wh] with K' = 3K (fp32 C is accepted by every epilogue).
 * Attention is evaluated unfused per (batch, head): scores GEMM -> fp32 row softmax -> probabilities x values GEMM.
 */
/* x fp32 [batches, rows, cols] (ldx, x_batch_stride) -> out bf16 [batches*rows_pad, 3*cols]; pattern 0 = [hi|hi|lo]
 * (A operand), 1 = [hi|lo|hi] (W operand); rows [rows, rows_pad) are zero-filled. */
int dit4sr_split3(const float* x, int64_t ldx, int64_t x_batch_stride, void* out_bf16, int batches, int rows,
                  int rows_pad, int cols, int pattern, void* stream);
/* v fp32 [rows, cols] (ldv) -> out bf16 [cols, 3*rows_pad] = [Vh^T | Vl^T | Vh^T] (W operand of P V). */
int dit4sr_split3_transposed(const float* v, int64_t ldv, void* out_bf16, int rows, int rows_pad, int cols, void* stream);
/* in-place softmax(scale * s) over the first cols_valid columns of each row of s [rows, cols_total] (ld); the padding
 * columns are set to 0 (model_dit4sr/attention.py:528, the softmax inside F.scaled_dot_product_attention). */
int dit4sr_softmax_rows_f32(float* s, int64_t ld, int rows, int cols_valid, int cols_total, float scale, void* stream);
int dit4sr_add_f32(float* dst, int64_t dst_row_stride, int64_t dst_batch_stride, const float* src,
                   int64_t src_row_stride, int64_t src_batch_stride, int batches, int rows, int cols, void* stream);
/* fp32 version of dit4sr_dwconv3x3_add (same semantics, explicit arguments). */
int dit4sr_dwconv3x3_add_f32(float* u_hq, const float* u_lr, int64_t ld, int64_t hq_batch_stride,
                             int64_t lr_batch_stride, const float* weight9, const float* bias, int batches, int grid_h,
                             int grid_w, int lr_grid_h, int lr_row_offset, int channels, void* stream);
/* fp32 variants of the head kernels (PatchEmbed im2col, timestep sinusoid, SiLU) for the fp32-accurate mode */
int dit4sr_patchify_f32(const float* x, float* patches, int batches, int channels, int height, int width, void* stream);
int dit4sr_timestep_features_f32(const float* timesteps, float* out, int batches, void* stream);
int dit4sr_silu_f32(const float* x, float* y, int64_t n, void* stream);
/* out = x + dsigma * v over n contiguous elements of dtype (DIT4SR_DT_*), fp32 arithmetic: the flow-match Euler step
 * `scheduler.step` of pipelines/pipeline_dit4sr.py:1567 when the velocity is not taken straight from the unpatchify
 * kernel (sequence-parallel sampling, where the velocity strips are gathered first). */
int dit4sr_euler_update(const void* x, const void* v, void* out, int dtype, float dsigma, int64_t n, void* stream);
/* ---------------------------------------------------------------------------------------------------------------
 * Colour fix of the SR result against the LQ input (reference utils/wavelet_color_fix.py, called
 * from test/test.py:381-383), fp32 images [planes = batch*3, height, width].
 * wavelet_blur: low = 3x3 [1 2 1]x[1 2 1]/16 blur of `image` with dilation `radius` and replicate padding
 * (wavelet_color_fix.py:76-94); if high_accum != NULL also high_accum += image - low (one level of
 * wavelet_decomposition, :96-110).  plane_stats: per-plane mean and UNBIASED variance (calc_mean_std, :47-60).
 * adain_apply: out = (content - c_mean) / sqrt(c_var + eps) * sqrt(s_var + eps) + s_mean (:62-74). */
int dit4sr_wavelet_blur_f32(const float* image, float* low, float* high_accum, int planes, int height, int width,
                            int radius, void* stream);
int dit4sr_plane_stats_f32(const float* x, int planes, int64_t plane_elems, float* mean, float* var_unbiased,
                           void* stream);
int dit4sr_adain_apply_f32(const float* content, float* out, int planes, int64_t plane_elems, const float* c_mean,
                           const float* c_var, const float* s_mean, const float* s_var, float eps, void* stream);
/* y = x (fp32 -> bf16 or bf16 -> fp32), n elements; dir 0 = f32->bf16, 1 = bf16->f32 */
int dit4sr_cast(const void* x, void* y, int64_t n, int dir, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * VAE encode / decode around the sampling loop: diffusers AutoencoderKL as the reference calls it at
 * pipelines/pipeline_dit4sr.py:1001-1002 (`vae.encode(control_image).latent_dist.sample()`) and :1602-1605
 * (`vae.decode(latents)`), whose layer sequence utils/vaehook.py:279-358 walks for its tiled variant.  Activations are
 * channels-last bf16 ([batches, H*W, C], C % 64 == 0); every convolution / linear layer is dit4sr_gemm_bf16 (3x3: conv_*
 * fields); the skip stream of the residual blocks is fp32 (GATE_RESID epilogue without a gate, added in place with TMA
 * reduce-adds); the entry points below are the layers in between.
 */
/* GroupNorm(groups, eps, affine) (+ SiLU when silu != 0) over NHWC x (x_dtype: bf16, or the fp32 skip stream) -> bf16 y: statistics per (batch, group) over pixels x
 * (channels / groups), fixed summation order (bit-reproducible; per-CTA partial sums in `stats`, a caller-owned scratch of
 * dit4sr_groupnorm_workspace_bytes(batches, groups) bytes, contents irrelevant, combined in double), then
 * y = (x - mean) rstd gamma + beta.  ResnetBlock2D.norm1/norm2, Attention.group_norm, conv_norm_out. */
typedef struct {
  const void* x; int64_t ldx; int64_t x_batch_stride;
  void* y; int64_t ldy; int64_t y_batch_stride;
  const float* gamma; const float* beta;
  void* stats; int64_t stats_bytes;
  int batches; int pixels; int channels; int groups;
  float eps; int silu;
  int x_dtype;                              /* DIT4SR_DT_BF16 / DIT4SR_DT_F32 (strides in elements of that type) */
  int width; int pitch;                     /* pitched image (see below): `pixels` = rows * pitch, only the first `width`
                                               pixels of every `pitch` enter the statistics; 0, 0 = every pixel counts */
} dit4sr_groupnorm_params;
int dit4sr_groupnorm_nhwc(const dit4sr_groupnorm_params* p, void* stream);
int64_t dit4sr_groupnorm_workspace_bytes(int batches, int groups);
/* PITCHED images: the convolution mode of the GEMM needs an image width that divides 128 or is a multiple of 128, so an
 * (h x w) activation is stored with a row pitch >= w of that kind (row y starts at pixel y * pitch; conv_w = pitch,
 * conv_w_valid = w: the tensor map clips at w, so whatever the pad pixels hold is read as the zero padding).  The layer
 * kernels below take the width and the pitches of their input and output and write zeros into pad pixels. */
/* y[b, 2h, 2w, c] = x[b, h, w, c] nearest-neighbour repeat (Upsample2D before its 3x3 convolution); NHWC, x of
 * x_dtype (DIT4SR_DT_*) [batches, h, pitch_in, c], y bf16 [batches, 2h, pitch_out, c] */
int dit4sr_upsample2x_nhwc(const void* x, int x_dtype, void* y, int batches, int h, int w, int pitch_in, int pitch_out,
                           int channels, void* stream);
/* y[b, Y, X, (py*2+px)*C + c] = x[b, 2Y+py, 2X+px, c] (y bf16 [batches, h/2, pitch_out, 4C]): with it Downsample2D
 * (pad (0,1,0,1), 3x3, stride 2) is a 2x2 stride-1 convolution over 4C channels (conv_kw 2, conv_taps 4, conv_org 0;
 * weights repacked by the host) */
int dit4sr_space_to_depth2_nhwc(const void* x, int x_dtype, void* y, int batches, int h, int w, int pitch_in, int pitch_out,
                                int channels, void* stream);
/* p[r, c] = softmax(scale * s[r, :valid_cols])[c] for c < valid_cols, 0 for valid_cols <= c < cols: fp32 score rows -> bf16
 * probabilities (the mid-block's single-head attention with head_dim = channels: scores and P.V are plain GEMMs whose key
 * dimension is padded to a multiple of 32) */
int dit4sr_softmax_rows_to_bf16(const float* s, int64_t lds, void* p, int64_t ldp, int rows, int cols, int valid_cols,
                                float scale, void* stream);
/* NCHW (x_dtype DIT4SR_DT_*) [batches, channels, h, w] -> NHWC bf16 [batches, h, pitch, c_pad] (channels >= `channels`
 * and pad pixels zero), y = a x + b on the way (latents: a = 1 / scaling_factor, b = shift_factor) */
int dit4sr_nchw_to_nhwc_bf16(const void* x, int x_dtype, void* y, int batches, int channels, int c_pad, int h, int w, int pitch,
                             float a, float b, void* stream);
/* fp32 NHWC [batches, h, pitch, ld] (first `channels` columns) -> NCHW fp32 [batches, channels, h, w] */
int dit4sr_nhwc_to_nchw_f32(const float* x, int64_t ld, float* y, int batches, int channels, int h, int w, int pitch,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIT4SR_B200_H_ */

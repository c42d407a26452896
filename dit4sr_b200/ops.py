"""Tensor-level wrappers over the C-ABI (include/dit4sr_b200.h): they only translate torch tensors into device
pointers / strides and launch on torch's current CUDA stream.  No math happens in Python and there is no fallback."""
import ctypes as C

import torch

from . import _lib
from ._lib import (A2aParams, AttnParams, AttnSpOut, DwconvParams, GemmParams, GroupNormParams, HaloParams, LnModulateParams,
                   check)

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_SILU, EPI_QKV_NORM, EPI_GATE_RESID = 0, 1, 2, 3, 4
DT_BF16, DT_F32 = 0, 1

_launches = 0  # number of kernel launches issued through this module (bench.py reports it as gpu_launches)


def launch_count():
    return _launches


def _count(n=1):
    global _launches
    _launches += n


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _req(t, dtype, name):
    if not t.is_cuda:
        raise _lib.Dit4srError(f"{name}: expected a CUDA tensor (dit4sr_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.Dit4srError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise _lib.Dit4srError(f"{name}: innermost dimension must be contiguous")


def _as3(t):
    """View [M, K] as [1, M, K]; keep [B, R, K]."""
    return t.unsqueeze(0) if t.dim() == 2 else t


def ln_modulate(x, shift, scale, out, shift2=None, scale2=None, out2=None, eps=1e-6, companion=None):
    """out = LN(x) * (1 + scale[b]) + shift[b]  (and optionally out2 with shift2/scale2). x fp32 [B,R,D]; out bf16.
    companion = (x_b, shift_b, scale_b, out_b): a second tensor [B,R_b,D] (the text stream) normalised by the same launch."""
    x = _as3(x)
    out = _as3(out)
    _req(x, torch.float32, "x")
    out_dtype = out.dtype
    if out_dtype not in (torch.bfloat16, torch.float32):
        raise _lib.Dit4srError("ln_modulate: out must be bf16 or fp32")
    _req(out, out_dtype, "out")
    for n, t in (("shift", shift), ("scale", scale)):
        _req(t, torch.float32, n)
    B, R, D = x.shape
    p = LnModulateParams()
    p.x, p.ldx, p.x_batch_stride = _ptr(x), x.stride(1), x.stride(0)
    p.shift, p.scale = _ptr(shift), _ptr(scale)
    p.mod_batch_stride = shift.stride(0) if shift.dim() == 2 else 0
    if scale.dim() == 2 and scale.stride(0) != p.mod_batch_stride:
        raise _lib.Dit4srError("ln_modulate: shift and scale must share the batch stride")
    p.out, p.ldo, p.o_batch_stride = _ptr(out), out.stride(1), out.stride(0)
    if out2 is not None:
        out2 = _as3(out2)
        _req(out2, out_dtype, "out2")
        if out2.stride() != out.stride():
            raise _lib.Dit4srError("ln_modulate: out and out2 must share strides")
        p.out2, p.shift2, p.scale2 = _ptr(out2), _ptr(shift2), _ptr(scale2)
    p.batches, p.rows, p.dim, p.eps = B, R, D, eps
    p.out_f32 = 1 if out_dtype == torch.float32 else 0
    if companion is not None:
        xb, shb, scb, ob = companion
        xb, ob = _as3(xb), _as3(ob)
        _req(xb, torch.float32, "companion x")
        _req(ob, out_dtype, "companion out")
        for n, t in (("companion shift", shb), ("companion scale", scb)):
            _req(t, torch.float32, n)
        if xb.shape[0] != B or xb.shape[2] != D or ob.shape != xb.shape:
            raise _lib.Dit4srError("ln_modulate: companion must be [batches, rows_b, dim] like x, with an output of its shape")
        mbs = shb.stride(0) if shb.dim() == 2 else 0
        if mbs != p.mod_batch_stride or (scb.stride(0) if scb.dim() == 2 else 0) != mbs:
            raise _lib.Dit4srError("ln_modulate: companion shift/scale must share the batch stride of shift/scale")
        p.x_b, p.ldx_b, p.x_b_batch_stride = _ptr(xb), xb.stride(1), xb.stride(0)
        p.shift_b, p.scale_b = _ptr(shb), _ptr(scb)
        p.out_b, p.ldo_b, p.o_b_batch_stride = _ptr(ob), ob.stride(1), ob.stride(0)
        p.rows_b = xb.shape[1]
    check(_lib.load().dit4sr_ln_modulate(C.byref(p), _stream()), "dit4sr_ln_modulate")
    _count()


def dwconv3x3_add(u_hq, u_lr, weight9, bias, grid, grid_w=0, lr_row_offset=0):
    """u_hq[b, y*W+x, c] += bias[c] + dwconv3x3(u_lr)[...]; u_* bf16 views of the MLP hidden state, channels innermost.

    Square case: u_hq, u_lr [B, grid*grid, C]. Row-sharded case (sequence parallelism): u_hq [B, grid*grid_w, C] holds
    `grid` rows of `grid_w` columns, u_lr [B, lr_rows*grid_w, C] holds the LR rows including halo rows, and u_hq row 0 lies
    over u_lr row `lr_row_offset`."""
    _req(u_hq, torch.bfloat16, "u_hq")
    _req(u_lr, torch.bfloat16, "u_lr")
    _req(weight9, torch.float32, "weight9")
    _req(bias, torch.float32, "bias")
    B, N, Cc = u_hq.shape
    W = grid_w or grid
    if N != grid * W or u_lr.shape[1] % W or u_lr.shape[0] != B or u_lr.shape[2] != Cc or u_hq.stride(1) != u_lr.stride(1):
        raise _lib.Dit4srError("dwconv3x3_add: shape mismatch (token grid must be rows x columns, attention.py:169-172)")
    p = DwconvParams()
    p.u_hq, p.u_lr, p.ld = _ptr(u_hq), _ptr(u_lr), u_hq.stride(1)
    p.hq_batch_stride, p.lr_batch_stride = u_hq.stride(0), u_lr.stride(0)
    p.weight9, p.bias = _ptr(weight9), _ptr(bias)
    p.batches, p.grid, p.channels = B, grid, Cc
    p.grid_w, p.lr_grid_h, p.lr_row_offset = W, u_lr.shape[1] // W, lr_row_offset
    check(_lib.load().dit4sr_dwconv3x3_add(C.byref(p), _stream()), "dit4sr_dwconv3x3_add")
    _count()


def add_bf16(dst, src):
    """dst += src for bf16 [B, R, C] views (innermost contiguous)."""
    _req(dst, torch.bfloat16, "dst")
    _req(src, torch.bfloat16, "src")
    if dst.shape != src.shape or dst.dim() != 3:
        raise _lib.Dit4srError("add_bf16: shapes must match ([B, R, C])")
    B, Rr, Cc = dst.shape
    check(_lib.load().dit4sr_add_bf16(_ptr(dst), dst.stride(1), dst.stride(0), _ptr(src), src.stride(1), src.stride(0),
                                      B, Rr, Cc, _stream()), "dit4sr_add_bf16")
    _count()


def gemm_sp(world, rank, inner_dim, rows_dst, peers, row0, col0=0, own_heads_only=False):
    """dit4sr_gemm_sp: scatter description of one q|k|v projection under sequence parallelism (kept alive by the caller)."""
    sp = _lib.GemmSp()
    sp.world, sp.rank, sp.own_heads_only, sp.inner_dim, sp.col0 = world, rank, 1 if own_heads_only else 0, inner_dim, col0
    sp.rows_dst, sp.row0 = rows_dst, row0
    for r, ptr in enumerate(peers):
        sp.peer[r] = ptr
    return sp


def _gemm_params(p, A, W, bias, out, epilogue=EPI_BIAS, gate=None, resid=None, norm_w0=None, norm_w1=None, sec_cols=0,
                 eps=1e-6, dbl_row_offset=0, sp=None, conv=None, resid_bf16=None):
    """conv = (H, W, kernel_width, taps, origin[, valid width]): A [B, H*W, C_in] is an NHWC image (pitched when a valid width
    < W is given) and W is [N, taps * C_in] (implicit-GEMM convolution); resid_bf16 [B, R, N] is added to a bf16 output."""
    A = _as3(A)
    out = _as3(out)
    _req(A, torch.bfloat16, "A")
    _req(W, torch.bfloat16, "W")
    if out.dtype not in (torch.bfloat16, torch.float32):
        raise _lib.Dit4srError("gemm: out must be bf16 or fp32")
    _req(out, out.dtype, "out")
    B, R, K = A.shape
    N = W.shape[0]
    if conv is not None:
        ch, cw, ckw, ctaps, corg = conv[:5]
        if R != ch * cw:
            raise _lib.Dit4srError(f"gemm/conv: A has {R} rows for a {ch}x{cw} image")
        K = K * ctaps
        p.conv_h, p.conv_w, p.conv_kw, p.conv_taps, p.conv_org = ch, cw, ckw, ctaps, corg
        p.conv_w_valid = conv[5] if len(conv) > 5 else 0
    if W.shape[1] != K or out.shape[0] != B or out.shape[1] != R or out.shape[2] != N:
        raise _lib.Dit4srError(f"gemm: shape mismatch A{tuple(A.shape)} W{tuple(W.shape)} out{tuple(out.shape)}")
    p.A, p.lda, p.a_batch_stride = _ptr(A), A.stride(1), A.stride(0)
    p.W, p.ldw = _ptr(W), W.stride(0)
    if bias is not None:
        _req(bias, torch.float32, "bias")
    p.bias = _ptr(bias)
    p.C, p.ldc, p.c_batch_stride = _ptr(out), out.stride(1), out.stride(0)
    p.c_dtype = DT_F32 if out.dtype == torch.float32 else DT_BF16
    p.batches, p.rows, p.N, p.K, p.epilogue = B, R, N, K, epilogue
    if epilogue == EPI_GATE_RESID:
        resid = out if resid is None else _as3(resid)
        _req(resid, torch.float32, "resid")
        p.resid, p.ldr = _ptr(resid), resid.stride(1)
        p.r_batch_stride = resid.stride(0) if resid.shape[0] == B else 0
        if gate is not None:
            _req(gate, torch.float32, "gate")
            p.gate = _ptr(gate)
            p.gate_batch_stride = gate.stride(0) if gate.dim() == 2 else 0
    if epilogue == EPI_QKV_NORM:
        for n, t in (("norm_w0", norm_w0), ("norm_w1", norm_w1)):
            if t is not None:
                _req(t, torch.float32, n)
        p.norm_w0, p.norm_w1 = _ptr(norm_w0), _ptr(norm_w1)
        p.sec_cols, p.dbl_row_offset = sec_cols or N, dbl_row_offset
    p.eps = eps
    if sp is not None:
        p.sp = C.pointer(sp)
    if resid_bf16 is not None:
        resid_bf16 = _as3(resid_bf16)
        _req(resid_bf16, torch.bfloat16, "resid_bf16")
        if resid_bf16.shape != out.shape:
            raise _lib.Dit4srError("gemm: resid_bf16 must have the shape of out")
        p.resid_bf16, p.ld_rb, p.rb_batch_stride = _ptr(resid_bf16), resid_bf16.stride(1), resid_bf16.stride(0)


def gemm(A, W, bias, out, epilogue=EPI_BIAS, **kw):
    """out[b, r, :] = epilogue(A[b, r, :] @ W.T).  A bf16 [B,R,K] or [M,K]; W bf16 [N,K]; out bf16/fp32 same rows."""
    p = GemmParams()
    _gemm_params(p, A, W, bias, out, epilogue, **kw)
    check(_lib.load().dit4sr_gemm_bf16(C.byref(p), _stream()), "dit4sr_gemm_bf16")
    _count()


GEMM_MAX_GROUP = 6


def gemm_grouped(problems):
    """Several independent GEMMs with the same epilogue kind in ONE persistent launch.
    problems: list of dicts with the keyword arguments of `gemm` (A, W, bias, out, epilogue, ...)."""
    n = len(problems)
    if not 1 <= n <= GEMM_MAX_GROUP:
        raise _lib.Dit4srError(f"gemm_grouped: 1..{GEMM_MAX_GROUP} problems per launch")
    arr = (GemmParams * n)()
    for p, kw in zip(arr, problems):
        _gemm_params(p, **kw)
    check(_lib.load().dit4sr_gemm_bf16_grouped(arr, n, _stream()), "dit4sr_gemm_bf16_grouped")
    _count()


def gemm_schedule(shapes, sm_count=148):
    """Host-only: the work items (problem, batch, row_tile, first_column, columns) one grouped launch walks on a GPU with
    `sm_count` SMs, in schedule order.  shapes: list of (batches, rows, N, K).  Needs no device."""
    import ctypes
    n = len(shapes)
    arr = (GemmParams * n)()
    for p, (b, r, N, K) in zip(arr, shapes):
        p.batches, p.rows, p.N, p.K = b, r, N, K
    lib = _lib.load()
    count = lib.dit4sr_gemm_schedule(arr, n, sm_count, None, 0)
    if count < 0:
        check(count, "dit4sr_gemm_schedule")
    items = (ctypes.c_int * (5 * count))()
    check(0 if lib.dit4sr_gemm_schedule(arr, n, sm_count, items, count) == count else -1, "dit4sr_gemm_schedule")
    return [tuple(items[5 * i:5 * i + 5]) for i in range(count)]


def attention_workspace_bytes():
    return int(_lib.load().dit4sr_attn_workspace_bytes())


def attention_plan(batches, heads, seq_q, seq_kv, sm_count=148, workspace=True, split_kv=0, variant=0, prescaled=False):
    """Host-only: (kernel variant, KV slices, whole items, tail items, KV tiles, KV tiles per slice) dit4sr_attn_fwd would use
    for this shape on a GPU with `sm_count` SMs.  Needs no device."""
    p = AttnParams()
    p.batches, p.heads, p.seq_q, p.seq_kv, p.split_kv, p.variant = batches, heads, seq_q, seq_kv, split_kv, variant
    p.mode = 1 if prescaled else 0      # (ATTN_PRESCALED: the split-row form of the one-tile kernel is the default there)
    if workspace:
        p.workspace, p.workspace_bytes = C.c_void_p(16), 1 << 40      # only compared with NULL / the size needed
    plan = (C.c_int * 6)()
    check(_lib.load().dit4sr_attn_plan(C.byref(p), sm_count, plan), "dit4sr_attn_plan")
    return tuple(plan)


def attention_last_plan():
    """(kernel variant, KV slices per tail item, whole items, tail items) of the latest attention call (diagnostics)."""
    plan = (C.c_int * 4)()
    check(_lib.load().dit4sr_attn_last_plan(plan), "dit4sr_attn_last_plan")
    return tuple(plan)


ATTN_ONLINE, ATTN_PRESCALED = 0, 1
ATTN_MAX_SCORE_BOUND = 60.0


def attention(q, k, v, o, heads, seq_q=None, scale=None, lr_res=None, lr_rows=(0, 0), score_bound=None, exp_offload=0, variant=0,
              sp_out=None, workspace=None, split_kv=0):
    """o = softmax(q k^T * scale) v over [B, S, heads*64] bf16 views (row stride may be 3*heads*64 for a fused buffer).

    Only the first seq_q query rows are produced. lr_res (bf16 [B, >=lr_rows[1], heads*64]) is added to rows lr_rows.
    score_bound (log2 units): q already carries scale*log2(e) and q.k <= score_bound for every pair -> the kernel takes
    p = 2^(q.k) without a reference max (DIT4SR_ATTN_PRESCALED); None = online softmax (any q, k).
    sp_out (an AttnSpOut from `attn_sp_out`): the rows are stored into the owning ranks' peer-mapped buffers instead of o,
    which then only supplies the common strides."""
    for n, t in (("q", q), ("k", k), ("v", v), ("o", o)):
        _req(t, torch.bfloat16, n)
    B, S, D = k.shape
    if D != heads * 64:
        raise _lib.Dit4srError("attention: head_dim must be 64")
    if not (q.stride() == k.stride() == v.stride()):
        raise _lib.Dit4srError("attention: q/k/v must share strides")
    seq_q = q.shape[1] if seq_q is None else seq_q
    p = AttnParams()
    p.q, p.k, p.v = _ptr(q), _ptr(k), _ptr(v)
    p.qkv_row_stride, p.qkv_batch_stride = k.stride(1), k.stride(0)
    p.o, p.o_row_stride, p.o_batch_stride = _ptr(o), o.stride(1), o.stride(0)
    if lr_res is not None:
        _req(lr_res, torch.bfloat16, "lr_res")
        p.lr_res, p.lr_row_stride, p.lr_batch_stride = _ptr(lr_res), lr_res.stride(1), lr_res.stride(0)
        p.lr_row_begin, p.lr_row_end = lr_rows
    p.batches, p.heads, p.seq_q, p.seq_kv = B, heads, seq_q, S
    p.scale = (1.0 / 8.0) if scale is None else scale
    if score_bound is not None:
        p.mode, p.score_bound = ATTN_PRESCALED, float(score_bound)
    p.exp_offload, p.variant = exp_offload, variant
    if sp_out is not None:
        p.sp_out = C.pointer(sp_out)
    if workspace is not None:     # zero-filled uint8 scratch of attention_workspace_bytes(): enables the split-KV tail
        p.workspace, p.workspace_bytes = _ptr(workspace), workspace.numel() * workspace.element_size()
    p.split_kv = split_kv
    check(_lib.load().dit4sr_attn_fwd(C.byref(p), _stream()), "dit4sr_attn_fwd")
    _count()


# ---------------------------------------------------------------------------------------------------------------------
# Ulysses sequence parallelism over NVLink peer memory (csrc/sp_comm.cu)
def attn_sp_out(world, rows_per_rank, col_offset, o_peers):
    so = AttnSpOut()
    so.world, so.rows_per_rank, so.col_offset = world, rows_per_rank, col_offset
    for r, ptr in enumerate(o_peers):
        so.o_peer[r] = ptr
    return so


def sp_barrier(comm, channel):
    check(_lib.load().dit4sr_sp_barrier(C.byref(comm), channel, _stream()), "dit4sr_sp_barrier")
    _count()


def a2a_qkv(comm, src, image_rows, text_rows, heads, peers, no_barrier=False):
    """src bf16 [B, image_rows + text_rows, 3*heads*64] (my tokens, all heads) -> every rank's [B, P*image_rows +
    text_rows, 3, heads*64/P] buffer (`peers`: device pointers as mapped here), then the arrival barrier.
    src=None: the projections already stored into the peers (gemm `sp`), arrival barrier only."""
    p = A2aParams()
    if src is not None:
        _req(src, torch.bfloat16, "src")
        p.src, p.src_row_stride, p.src_batch_stride = _ptr(src), src.stride(1), src.stride(0)
        p.batches = src.shape[0]
    for r, ptr in enumerate(peers):
        p.peer[r] = ptr
    p.image_rows, p.text_rows, p.heads = image_rows, text_rows, heads
    p.no_barrier = 1 if no_barrier else 0
    check(_lib.load().dit4sr_a2a_qkv(C.byref(comm), C.byref(p), _stream()), "dit4sr_a2a_qkv")
    _count((0 if src is None else 1) + (0 if no_barrier else 1))


def a2a_out(comm, src, image_rows, text_rows, heads, peers, dst_row_stride, dst_batch_stride, no_barrier=False):
    """The way back: src bf16 [B, P*image_rows + text_rows, heads*64/P] contiguous -> the owning ranks' [B, >= image_rows +
    text_rows, heads*64] buffers. src=None: the attention kernel already stored through sp_out, arrival barrier only."""
    p = A2aParams()
    if src is not None:
        _req(src, torch.bfloat16, "src")
        if not src.is_contiguous():
            raise _lib.Dit4srError("a2a_out: src must be contiguous")
        p.src = _ptr(src)
        p.batches = src.shape[0]
    for r, ptr in enumerate(peers):
        p.peer[r] = ptr
    p.dst_row_stride, p.dst_batch_stride = dst_row_stride, dst_batch_stride
    p.image_rows, p.text_rows, p.heads = image_rows, text_rows, heads
    p.no_barrier = 1 if no_barrier else 0
    check(_lib.load().dit4sr_a2a_out(C.byref(comm), C.byref(p), _stream()), "dit4sr_a2a_out")
    _count((0 if src is None else 1) + (0 if no_barrier else 1))


def halo_rows(comm, u, peers, grid_w, first_lr_row, lr_rows, top_halo_row, bottom_halo_row, no_barrier=False):
    """u bf16 [B, rows, F]: my MLP hidden-state buffer (same layout on every rank, `peers` = its mappings): my first / last
    LR grid row -> the neighbours' halo rows, then the arrival barrier."""
    _req(u, torch.bfloat16, "u")
    p = HaloParams()
    for r, ptr in enumerate(peers):
        p.peer[r] = ptr
    p.row_stride, p.batch_stride = u.stride(1), u.stride(0)
    p.batches, p.channels, p.grid_w = u.shape[0], u.shape[2], grid_w
    p.first_lr_row, p.lr_rows, p.top_halo_row, p.bottom_halo_row = first_lr_row, lr_rows, top_halo_row, bottom_halo_row
    p.no_barrier = 1 if no_barrier else 0
    check(_lib.load().dit4sr_halo_rows(C.byref(comm), C.byref(p), _stream()), "dit4sr_halo_rows")
    _count(1 if no_barrier else 2)


def patchify(x, out):
    """x [B,C,H,W] fp32/bf16 -> out bf16 [B, (H/2)(W/2), 4C] (im2col of PatchEmbed's k=2,s=2 conv)."""
    if not x.is_contiguous() or not out.is_contiguous():
        raise _lib.Dit4srError("patchify: tensors must be contiguous")
    B, Cc, H, W = x.shape
    if out.dtype == torch.float32:      # fp32-accurate mode
        _req(x, torch.float32, "x")
        check(_lib.load().dit4sr_patchify_f32(_ptr(x), _ptr(out), B, Cc, H, W, _stream()), "dit4sr_patchify_f32")
        _count()
        return
    _req(out, torch.bfloat16, "out")
    dt = DT_F32 if x.dtype == torch.float32 else DT_BF16
    check(_lib.load().dit4sr_patchify(_ptr(x), dt, _ptr(out), B, Cc, H, W, _stream()), "dit4sr_patchify")
    _count()


def unpatchify(tok, out, latents=None, dsigma=0.0, guidance=None):
    """tok fp32 [B, (H/2)(W/2), 4C] -> out [B,C,H,W]; with latents: out = latents + dsigma * velocity (Euler step).
    guidance (classifier-free guidance): tok is the doubled batch [uncond | text] (2B rows) and the velocity is
    v_uncond + guidance * (v_text - v_uncond)."""
    _req(tok, torch.float32, "tok")
    if not tok.is_contiguous() or not out.is_contiguous():
        raise _lib.Dit4srError("unpatchify: tensors must be contiguous")
    B, Cc, H, W = out.shape
    dt = DT_F32 if out.dtype == torch.float32 else DT_BF16
    if latents is not None and (latents.dtype != out.dtype or not latents.is_contiguous()):
        raise _lib.Dit4srError("unpatchify: latents must match out")
    if guidance is None:
        if tok.shape[0] != B:
            raise _lib.Dit4srError("unpatchify: batch mismatch")
        check(_lib.load().dit4sr_unpatchify(_ptr(tok), _ptr(out), dt, _ptr(latents), float(dsigma), B, Cc, H, W,
                                            _stream()), "dit4sr_unpatchify")
    else:
        if tok.shape[0] != 2 * B:
            raise _lib.Dit4srError("unpatchify: guidance needs the doubled batch [uncond | text]")
        check(_lib.load().dit4sr_unpatchify_cfg(_ptr(tok), _ptr(out), dt, _ptr(latents), float(dsigma), float(guidance), B,
                                                Cc, H, W, _stream()), "dit4sr_unpatchify_cfg")
    _count()


def feature_extract(x, row0, n_concat, grid_h, grid_w, nchw=None, tok=None):
    """Rows [row0, row0 + n_concat*grid_h*grid_w) of the fp32 residual stream x [B, rows, D] -> tok [B, n*gh*gw, D] (bf16 or
    fp32, the reference's `extract_feat`) and / or nchw fp32 [B, n*D/4, 2gh, 2gw] (what the pipeline feeds the ts_module)."""
    _req(x, torch.float32, "x")
    B, _, D = x.shape
    if nchw is not None and (nchw.dtype != torch.float32 or not nchw.is_contiguous()
                             or tuple(nchw.shape) != (B, n_concat * D // 4, 2 * grid_h, 2 * grid_w)):
        raise _lib.Dit4srError("feature_extract: nchw must be contiguous fp32 [B, n*D/4, 2*gh, 2*gw]")
    if tok is not None and (tok.dtype not in (torch.float32, torch.bfloat16) or not tok.is_contiguous()
                            or tuple(tok.shape) != (B, n_concat * grid_h * grid_w, D)):
        raise _lib.Dit4srError("feature_extract: tok must be contiguous bf16/fp32 [B, n*gh*gw, D]")
    td = DT_F32 if (tok is not None and tok.dtype == torch.float32) else DT_BF16
    check(_lib.load().dit4sr_feature_extract(_ptr(x), x.stride(1), x.stride(0), row0, B, n_concat, grid_h, grid_w, D,
                                             _ptr(nchw), _ptr(tok), td, _stream()), "dit4sr_feature_extract")
    _count()


def euler_update(x, v, out, dsigma):
    """out = x + dsigma * v (contiguous tensors of one dtype, fp32 or bf16)."""
    if not (x.is_contiguous() and v.is_contiguous() and out.is_contiguous()):
        raise _lib.Dit4srError("euler_update: tensors must be contiguous")
    if not (x.dtype == v.dtype == out.dtype) or x.dtype not in (torch.float32, torch.bfloat16):
        raise _lib.Dit4srError("euler_update: x, v, out must share dtype fp32 or bf16")
    if not (x.numel() == v.numel() == out.numel()):
        raise _lib.Dit4srError("euler_update: size mismatch")
    for n_, t_ in (("x", x), ("v", v), ("out", out)):
        _req(t_, x.dtype, n_)
    dt = DT_F32 if x.dtype == torch.float32 else DT_BF16
    check(_lib.load().dit4sr_euler_update(_ptr(x), _ptr(v), _ptr(out), dt, float(dsigma), x.numel(), _stream()),
          "dit4sr_euler_update")
    _count()


def timestep_features(t, out):
    _req(t, torch.float32, "timestep")
    if out.dtype == torch.float32:      # fp32-accurate mode
        check(_lib.load().dit4sr_timestep_features_f32(_ptr(t), _ptr(out), t.shape[0], _stream()),
              "dit4sr_timestep_features_f32")
        _count()
        return
    _req(out, torch.bfloat16, "out")
    check(_lib.load().dit4sr_timestep_features(_ptr(t), _ptr(out), t.shape[0], _stream()), "dit4sr_timestep_features")
    _count()


def silu_cast(x, out):
    _req(x, torch.float32, "x")
    if out.dtype == torch.float32:      # fp32-accurate mode
        check(_lib.load().dit4sr_silu_f32(_ptr(x), _ptr(out), x.numel(), _stream()), "dit4sr_silu_f32")
        _count()
        return
    _req(out, torch.bfloat16, "out")
    check(_lib.load().dit4sr_silu_cast(_ptr(x), _ptr(out), x.numel(), _stream()), "dit4sr_silu_cast")
    _count()


def cast(x, out):
    if not x.is_contiguous() or not out.is_contiguous():
        raise _lib.Dit4srError("cast: tensors must be contiguous")
    d = 0 if x.dtype == torch.float32 else 1
    check(_lib.load().dit4sr_cast(_ptr(x), _ptr(out), x.numel(), d, _stream()), "dit4sr_cast")
    _count()


# ---------------------------------------------------------------------------------------------------------------
# fp32-accurate ("bf16x3") mode helpers
# ---------------------------------------------------------------------------------------------------------------
def split3(x, out, pattern=0, rows_pad=None):
    """x fp32 [B, R, C] (or [R, C]) -> out bf16 [B*rows_pad, 3C]: pattern 0 = [hi|hi|lo] (A), 1 = [hi|lo|hi] (W)."""
    x = _as3(x)
    _req(x, torch.float32, "x")
    _req(out, torch.bfloat16, "out")
    B, R, Cc = x.shape
    rp = R if rows_pad is None else rows_pad
    if not out.is_contiguous() or out.numel() != B * rp * 3 * Cc:
        raise _lib.Dit4srError("split3: out must be contiguous [B*rows_pad, 3*cols]")
    check(_lib.load().dit4sr_split3(_ptr(x), x.stride(1), x.stride(0), _ptr(out), B, R, rp, Cc, pattern, _stream()),
          "dit4sr_split3")
    _count()


def split3_transposed(v, out, rows_pad):
    """v fp32 [R, C] -> out bf16 [C, 3*rows_pad] = [Vh^T | Vl^T | Vh^T]."""
    _req(v, torch.float32, "v")
    _req(out, torch.bfloat16, "out")
    R, Cc = v.shape
    if not out.is_contiguous() or out.numel() != Cc * 3 * rows_pad:
        raise _lib.Dit4srError("split3_transposed: out must be contiguous [cols, 3*rows_pad]")
    check(_lib.load().dit4sr_split3_transposed(_ptr(v), v.stride(0), _ptr(out), R, rows_pad, Cc, _stream()),
          "dit4sr_split3_transposed")
    _count()


def softmax_rows(s, cols_valid, scale):
    """in-place fp32 softmax(scale * s[:, :cols_valid]) per row; padding columns zeroed."""
    _req(s, torch.float32, "s")
    R, Ct = s.shape
    check(_lib.load().dit4sr_softmax_rows_f32(_ptr(s), s.stride(0), R, cols_valid, Ct, float(scale), _stream()),
          "dit4sr_softmax_rows_f32")
    _count()


def add_f32(dst, src):
    _req(dst, torch.float32, "dst")
    _req(src, torch.float32, "src")
    if dst.shape != src.shape or dst.dim() != 3:
        raise _lib.Dit4srError("add_f32: shapes must match ([B, R, C])")
    B, Rr, Cc = dst.shape
    check(_lib.load().dit4sr_add_f32(_ptr(dst), dst.stride(1), dst.stride(0), _ptr(src), src.stride(1), src.stride(0),
                                     B, Rr, Cc, _stream()), "dit4sr_add_f32")
    _count()


def dwconv3x3_add_f32(u_hq, u_lr, weight9, bias, grid, grid_w=0, lr_row_offset=0):
    for n, t in (("u_hq", u_hq), ("u_lr", u_lr), ("weight9", weight9), ("bias", bias)):
        _req(t, torch.float32, n)
    B, N, Cc = u_hq.shape
    W = grid_w or grid
    if N != grid * W or u_lr.shape[1] % W or u_hq.stride(1) != u_lr.stride(1):
        raise _lib.Dit4srError("dwconv3x3_add_f32: shape mismatch")
    check(_lib.load().dit4sr_dwconv3x3_add_f32(_ptr(u_hq), _ptr(u_lr), u_hq.stride(1), u_hq.stride(0), u_lr.stride(0),
                                               _ptr(weight9), _ptr(bias), B, grid, W, u_lr.shape[1] // W, lr_row_offset,
                                               Cc, _stream()), "dit4sr_dwconv3x3_add_f32")
    _count()


# ---------------------------------------------------------------------------------------------------------------- VAE (row N4)
def conv_pitch(w):
    """Row pitch (in pixels) an image of width w is stored with for the implicit-GEMM convolution: the smallest value
    >= w that divides 128 or is a multiple of 128."""
    if w >= 128:
        return (w + 127) // 128 * 128
    p = 1
    while p < w:
        p *= 2
    return p


def conv3x3(x, H, W, weight, bias, out, resid_bf16=None, resid=None, width=0):
    """3x3 convolution, zero pad 1, on NHWC bf16 x [B, H*W, C_in] with weight bf16 [N, 9 * C_in] (tap-major) -> out [B, H*W, N]
    bf16 (+ resid_bf16) or fp32 (+ resid fp32; resid may be out itself: added in place with TMA reduce-adds).
    width: the image is pitched - rows of W pixels of which the first `width` exist (0 = all W)."""
    if resid is not None:
        gemm(x, weight, bias, out, EPI_GATE_RESID, resid=resid, conv=(H, W, 3, 9, -1, width))
    else:
        gemm(x, weight, bias, out, EPI_BIAS, conv=(H, W, 3, 9, -1, width), resid_bf16=resid_bf16)


def conv2x2_s2d(x, H, W, weight, bias, out, width=0):
    """The 2x2 stride-1 convolution (taps at offsets 0 / +1, zero beyond the image) that a space-to-depth input turns
    Downsample2D's padded stride-2 3x3 convolution into: x [B, H*W, 4C], weight [N, 4 * 4C]."""
    gemm(x, weight, bias, out, EPI_BIAS, conv=(H, W, 2, 4, 0, width))


def groupnorm_nhwc(x, gamma, beta, out, groups=32, eps=1e-6, silu=False, stats=None, width=0, pitch=0):
    """out (bf16) = [silu](GroupNorm(x)) on NHWC x [B, pixels, C] (bf16 or fp32); stats: scratch of
    dit4sr_groupnorm_workspace_bytes(B, groups) bytes (allocated if None).  width / pitch: pitched image (pixels = rows *
    pitch, only the first `width` of every `pitch` pixels enter the statistics)."""
    _req(x, x.dtype if x.dtype == torch.float32 else torch.bfloat16, "x")
    _req(out, torch.bfloat16, "out")
    _req(gamma, torch.float32, "gamma")
    _req(beta, torch.float32, "beta")
    B, P, Cc = x.shape
    if out.shape != x.shape or gamma.numel() != Cc or beta.numel() != Cc:
        raise _lib.Dit4srError("groupnorm_nhwc: shape mismatch")
    if x.dtype not in (torch.bfloat16, torch.float32):
        raise _lib.Dit4srError("groupnorm_nhwc: x must be bf16 or fp32")
    need = int(_lib.load().dit4sr_groupnorm_workspace_bytes(B, groups))
    if stats is None:
        stats = torch.empty(need, device=x.device, dtype=torch.uint8)
    p = GroupNormParams()
    p.x, p.ldx, p.x_batch_stride = _ptr(x), x.stride(1), x.stride(0)
    p.y, p.ldy, p.y_batch_stride = _ptr(out), out.stride(1), out.stride(0)
    p.gamma, p.beta, p.stats, p.stats_bytes = _ptr(gamma), _ptr(beta), _ptr(stats), stats.numel() * stats.element_size()
    p.batches, p.pixels, p.channels, p.groups, p.eps, p.silu = B, P, Cc, groups, eps, 1 if silu else 0
    p.x_dtype = DT_F32 if x.dtype == torch.float32 else DT_BF16
    if width and pitch and width != pitch:
        p.width, p.pitch = width, pitch
    check(_lib.load().dit4sr_groupnorm_nhwc(C.byref(p), _stream()), "dit4sr_groupnorm_nhwc")
    _count(3)


def _dt(x):
    return DT_F32 if x.dtype == torch.float32 else DT_BF16


def upsample2x_nhwc(x, H, W, out, pitch_in=0, pitch_out=0):
    """out [B, 2H*pitch_out, C] (bf16) = nearest-neighbour repeat of the (H x W) image x [B, H*pitch_in, C] (contiguous, bf16
    or fp32); pitches default to the widths."""
    _req(x, x.dtype if x.dtype == torch.float32 else torch.bfloat16, "x")
    _req(out, torch.bfloat16, "out")
    pi, po = pitch_in or W, pitch_out or 2 * W
    B, P, Cc = x.shape
    if P != H * pi or tuple(out.shape) != (B, 2 * H * po, Cc) or not x.is_contiguous() or not out.is_contiguous():
        raise _lib.Dit4srError("upsample2x_nhwc: contiguous [B, H*pitch_in, C] -> [B, 2H*pitch_out, C]")
    check(_lib.load().dit4sr_upsample2x_nhwc(_ptr(x), _dt(x), _ptr(out), B, H, W, pi, po, Cc, _stream()), "dit4sr_upsample2x_nhwc")
    _count()


def space_to_depth2_nhwc(x, H, W, out, pitch_in=0, pitch_out=0):
    """out [B, (H/2)*pitch_out, 4C] (bf16) with channel (py*2+px)*C + c = x[b, (2Y+py)*pitch_in + 2X+px, c] (contiguous, bf16
    or fp32); pitches default to the widths."""
    _req(x, x.dtype if x.dtype == torch.float32 else torch.bfloat16, "x")
    _req(out, torch.bfloat16, "out")
    pi, po = pitch_in or W, pitch_out or W // 2
    B, P, Cc = x.shape
    if P != H * pi or tuple(out.shape) != (B, (H // 2) * po, 4 * Cc) or not x.is_contiguous() or not out.is_contiguous():
        raise _lib.Dit4srError("space_to_depth2_nhwc: contiguous [B, H*pitch_in, C] -> [B, (H/2)*pitch_out, 4C]")
    check(_lib.load().dit4sr_space_to_depth2_nhwc(_ptr(x), _dt(x), _ptr(out), B, H, W, pi, po, Cc, _stream()),
          "dit4sr_space_to_depth2_nhwc")
    _count()


def softmax_rows_to_bf16(scores, probs, scale, valid_cols=None):
    """probs[r, :valid] = softmax(scale * scores[r, :valid]), zeros beyond; scores fp32 [R, C], probs bf16 [R, C]."""
    _req(scores, torch.float32, "scores")
    _req(probs, torch.bfloat16, "probs")
    if scores.shape != probs.shape or scores.dim() != 2:
        raise _lib.Dit4srError("softmax_rows_to_bf16: [R, C] fp32 -> [R, C] bf16")
    valid = scores.shape[1] if valid_cols is None else valid_cols
    check(_lib.load().dit4sr_softmax_rows_to_bf16(_ptr(scores), scores.stride(0), _ptr(probs), probs.stride(0), scores.shape[0],
                                                  scores.shape[1], valid, float(scale), _stream()), "dit4sr_softmax_rows_to_bf16")
    _count()


def nchw_to_nhwc_bf16(x, c_pad, a=1.0, b=0.0, pitch=0):
    """x [B, C, H, W] (fp32 / bf16, contiguous) -> bf16 [B, H*pitch, c_pad] = a x + b, channels >= C and pad pixels zero."""
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    x = x.contiguous()
    _req(x, x.dtype, "x")
    B, Cc, H, W = x.shape
    pitch = pitch or W
    out = torch.empty(B, H * pitch, c_pad, device=x.device, dtype=torch.bfloat16)
    check(_lib.load().dit4sr_nchw_to_nhwc_bf16(_ptr(x), _dt(x), _ptr(out), B, Cc, c_pad, H, W, pitch, float(a), float(b),
                                               _stream()), "dit4sr_nchw_to_nhwc_bf16")
    _count()
    return out


def nhwc_to_nchw_f32(x, channels, H, W, pitch=0):
    """fp32 [B, H*pitch, ld] -> fp32 [B, channels, H, W] (the first `channels` columns of the first W pixels per row)."""
    _req(x, torch.float32, "x")
    pitch = pitch or W
    B, P, _ = x.shape
    if P != H * pitch or x.stride(0) != P * x.stride(1):
        raise _lib.Dit4srError("nhwc_to_nchw_f32: expected [B, H*pitch, ld] with dense batches")
    out = torch.empty(B, channels, H, W, device=x.device, dtype=torch.float32)
    check(_lib.load().dit4sr_nhwc_to_nchw_f32(_ptr(x), x.stride(1), _ptr(out), B, channels, H, W, pitch, _stream()),
          "dit4sr_nhwc_to_nchw_f32")
    _count()
    return out

"""Host-side schedule of the denoiser forward: packs the module's weights once, owns the activation workspace, and
issues the CUDA kernels (through the C-ABI, `ops.py`) in the order of SURVEY.md 3.4.  No arithmetic in Python.

Dataflow of one step (reference: transformer_sd3.py:326-391 and attention.py:280-398, 445-559):
  head    timestep/pooled MLPs -> temb; ONE GEMM produces every block's adaLN vectors (temb is block-invariant);
          patchify + GEMM(+pos-embed epilogue) for HQ and LR latents -> x[B,2N,D] fp32; context_embedder -> ctx fp32
  block   LN+modulate (1 read, 1-2 bf16 writes) -> fused q|k|v GEMM with RMSNorm epilogue writing straight into the
          joint [HQ|LR|text] buffer (and the double-normed LR keys, quirk Q1) -> LR chain GEMMs -> text q|k|v GEMM ->
          joint attention (+LR residual, Q2) -> out-projections with gate*y+residual epilogue (in place on x/ctx)
          -> [attn2] -> LN+modulate -> FF1 (GELU-tanh epilogue) -> depthwise stencil -> FF2 (gated residual) -> text MLP
  tail    LN+modulate (norm_out) -> proj_out -> unpatchify (optionally fused with the Euler update)

Precision: tensor-core operands bf16, accumulation / residual streams (x, ctx) / LN+RMS statistics / softmax fp32.
"""
import math
import os
from typing import Dict, Optional, Tuple

import torch

from . import ops


def _bf16(t):
    return t.detach().to(torch.bfloat16).contiguous()


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


def _pack_w(w, precise):
    """bf16 [N, K] tensor-core operand; fp32-accurate mode: [N, 3K] = [hi | lo | hi] bf16 split of the fp32 weight
    (pairs with activations packed [hi | hi | lo], so one K'=3K GEMM evaluates Ah Wh + Ah Wl + Al Wh)."""
    if not precise:
        return _bf16(w)
    w = _f32(w)
    out = torch.empty(w.shape[0], 3 * w.shape[1], device=w.device, dtype=torch.bfloat16)
    ops.split3(w, out, pattern=1)
    return out


class _Lin:
    """bf16 weight [N,K] + fp32 bias [N] (possibly a concatenation of several nn.Linear)."""

    def __init__(self, *linears, precise=False):
        self.w = _pack_w(torch.cat([l.weight for l in linears], dim=0) if len(linears) > 1 else linears[0].weight,
                         precise)
        self.b = _f32(torch.cat([l.bias for l in linears], dim=0)) if len(linears) > 1 else _f32(linears[0].bias)


class _Folded:
    """Exact-algebra fold of a chained pair of Linears  y = Wc (W x + b) + bc  =  (Wc W) x + (Wc b + bc)  (quirk Q3:
    `to_x_control(to_x(lr))`, attention.py:475-477). Product formed in fp32, stored as one bf16 weight."""

    def __init__(self, *pairs, precise=False):
        """pairs: (inner, outer) Linears; several pairs are concatenated along the output dimension."""
        if len(pairs) == 2 and not isinstance(pairs[0], (tuple, list)):
            pairs = (pairs,)
        ws, bs = [], []
        for inner, outer in pairs:
            # load-time weight preparation; fp64 in the fp32-accurate mode so the fold itself adds no fp32 rounding
            dt = torch.float64 if precise else torch.float32
            wc, w = outer.weight.detach().to(dt), inner.weight.detach().to(dt)
            ws.append((wc @ w).float())
            bs.append((wc @ inner.bias.detach().to(dt) + outer.bias.detach().to(dt)).float())
        self.w = _pack_w(torch.cat(ws, 0), precise)
        self.b = torch.cat(bs, 0).contiguous()


LOG2E = 1.4426950408889634
PRESCALE_MAX_BOUND = 60.0   # log2 units: 2^60 * 8525 keys * |v| is far inside fp32 (and bf16 P) range


def prescaled_q_norm_weights(wq, wk, waq=None, wak=None):
    """Fold the softmax scale and log2(e) into the q RMSNorm weights, so that the joint attention kernel can take
    p = 2^score without a multiply, a reference max or a rescale (ops.ATTN_PRESCALED).

    Valid because QK-RMSNorm bounds every score: each normed row has norm <= sqrt(dh) * max|w| (attention.py:484-489,
    502-505; the twice-normed LR keys included), hence score * scale * log2(e) <= dh * max|w_q| * max|w_k| * scale * log2(e).
    Returns (wq', waq', bound in log2 units); raises if the bound is too large to drop the reference."""
    dh = wq.numel()
    f = LOG2E / math.sqrt(dh)
    qmax = max(float(wq.abs().max()), float(waq.abs().max()) if waq is not None else 0.0)
    kmax = max(float(wk.abs().max()), float(wak.abs().max()) if wak is not None else 0.0)
    bound = dh * qmax * kmax * f
    if not bound <= PRESCALE_MAX_BOUND:
        raise ValueError(f"score bound {bound:.1f} (log2 units) exceeds {PRESCALE_MAX_BOUND}: this checkpoint needs the "
                         "online-max softmax")
    return wq * f, (waq * f if waq is not None else None), bound


def _prescale_allowed():
    """DIT4SR_ATTN_PRESCALE=0 keeps the online softmax for every attention (read once per Engine, at pack time)."""
    return os.environ.get("DIT4SR_ATTN_PRESCALE", "1") != "0"


class _AttnPack:
    def __init__(self, attn, has_text: bool, fold: bool, precise: bool = False, prescale: bool = False):
        from functools import partial
        _Lin, _Folded = partial(globals()["_Lin"], precise=precise), partial(globals()["_Folded"], precise=precise)
        self.qk_norm = attn.norm_q is not None
        self.fold = fold
        self.qkv = _Lin(attn.to_q, attn.to_k, attn.to_v)
        if self.qk_norm:
            # Q1: with qk_norm the LR keys are norm_k(norm_k(k_hq)); to_k_control(to_k(lr)) is dead code
            self.wq, self.wk = _f32(attn.norm_q.weight), _f32(attn.norm_k.weight)
        if fold:
            self.qf, self.vf = _Folded(attn.to_q, attn.to_q_control), _Folded(attn.to_v, attn.to_v_control)
            if not self.qk_norm:
                self.kf = _Folded(attn.to_k, attn.to_k_control)
        else:
            if self.qk_norm:
                self.lr1 = _Lin(attn.to_q, attn.to_v)
            else:
                self.lr1 = self.qkv
                self.kc = _Lin(attn.to_k_control)
            self.qc, self.vc = _Lin(attn.to_q_control), _Lin(attn.to_v_control)
        self.out, self.outc = _Lin(attn.to_out[0]), _Lin(attn.to_out_control)
        self.ocr = getattr(attn, "ocr_branch", False)
        if self.ocr:   # third stream: to_x_ocr(to_x(ocr)) chains, always folded; its keys get a live norm_k
            self.of = _Folded((attn.to_q, attn.to_q_ocr), (attn.to_k, attn.to_k_ocr), (attn.to_v, attn.to_v_ocr))
            self.outo = _Lin(attn.to_out_ocr)
        self.has_text = has_text
        if has_text:
            self.add = _Lin(attn.add_q_proj, attn.add_k_proj, attn.add_v_proj)
            if self.qk_norm:
                self.waq, self.wak = _f32(attn.norm_added_q.weight), _f32(attn.norm_added_k.weight)
            self.add_out = _Lin(attn.to_add_out) if hasattr(attn, "to_add_out") else None
        # Bounded scores (QK-RMSNorm with small enough weights): fold scale*log2(e) into the q norm weights and let the
        # attention kernel drop the running max; any other checkpoint (qk_norm=None, large norm weights) keeps the online
        # softmax.  Decided per attention module at pack time; the kernel gets the mode with every call (ABI v3).
        self.score_bound = None
        if prescale and self.qk_norm and not precise:
            try:
                wq, waq, bound = prescaled_q_norm_weights(self.wq, self.wk, getattr(self, "waq", None),
                                                          getattr(self, "wak", None))
            except ValueError:
                pass
            else:
                self.wq, self.score_bound = wq, bound
                if waq is not None:
                    self.waq = waq


class _BlockPack:
    def __init__(self, blk, fold, precise=False, prescale=False):
        from functools import partial
        _Lin = partial(globals()["_Lin"], precise=precise)
        self.dual = blk.use_dual_attention
        self.last = blk.context_pre_only
        self.attn = _AttnPack(blk.attn, True, fold, precise, prescale)
        self.attn2 = _AttnPack(blk.attn2, False, fold, precise, prescale) if self.dual else None
        self.ff1, self.ff2 = _Lin(blk.ff.net[0].proj), _Lin(blk.ff.net[2])
        cw = blk.ff.control_conv.weight
        self.conv_w9 = _f32(cw.reshape(cw.shape[0], 9).t())           # tap-major [9, F]
        self.conv_b = _f32(blk.ff.control_conv.bias)
        self.ocr = getattr(blk, "ocr_branch", False)
        if self.ocr:
            def taps(conv):
                return _f32(conv.weight.reshape(conv.weight.shape[0], 9).t()), _f32(conv.bias)
            self.conv_o_w9, self.conv_o_b = taps(blk.ff.ocr_to_hidden_conv)
            self.conv_h_w9, self.conv_h_b = taps(blk.ff.hidden_to_ocr_conv)
        if not self.last:
            self.ffc1, self.ffc2 = _Lin(blk.ff_context.net[0].proj), _Lin(blk.ff_context.net[2])


class Engine:
    def __init__(self, model, fold_lr_chains: Optional[bool] = None, sp=None, precise: Optional[bool] = None,
                 prescale_attention: Optional[bool] = None):
        """fold_lr_chains: fold `to_x_control(to_x(.))` into one GEMM per projection (exact algebra, one bf16 rounding
        of the product weight instead of a bf16 intermediate activation). Default on; DIT4SR_FOLD_LR=0 turns it off.

        precise: the fp32-accurate mode (velocity within 1e-4 of the fp32 reference, BASELINE north_star): activations
        stay fp32 between kernels, every tensor-core product runs on bf16 hi/lo splits (three bf16 products per fp32
        product, K' = 3K on the same tcgen05 GEMM kernel) and attention is evaluated unfused per head (scores GEMM ->
        fp32 softmax -> P V GEMM).  Default: on iff the module's parameters are fp32 (DIT4SR_FP32_MODE=bf16 keeps the
        fast bf16 path for fp32 modules)."""
        import os
        if fold_lr_chains is None:
            fold_lr_chains = os.environ.get("DIT4SR_FOLD_LR", "1") != "0"
        self.fold_lr_chains = fold_lr_chains
        if precise is None:
            precise = (next(model.parameters()).dtype == torch.float32
                       and os.environ.get("DIT4SR_FP32_MODE", "fp32") != "bf16")
        self.precise = bool(precise)
        if self.precise and sp is not None:
            raise NotImplementedError("the fp32-accurate mode is single-GPU (no sequence parallelism)")
        self._scratch: Dict = {}
        self.sp = sp    # parallel.SequenceParallel or None: Ulysses sequence parallelism over the token-grid rows
        self.streams = 3 if getattr(model, "_ocr_branch", False) else 2   # image streams: [HQ | LR] or [HQ | LR | OCR]
        cfg = model.config
        self.cfg = cfg
        self.D = cfg.num_attention_heads * cfg.attention_head_dim
        self.H = cfg.num_attention_heads
        if cfg.attention_head_dim != 64:
            raise ValueError("dit4sr_b200: attention_head_dim must be 64")
        if cfg.patch_size != 2:
            raise ValueError("dit4sr_b200: patch_size must be 2")
        if cfg.caption_projection_dim != self.D:
            raise ValueError("dit4sr_b200: caption_projection_dim must equal the inner dim (joint attention)")
        D = self.D
        from functools import partial
        _Lin = partial(globals()["_Lin"], precise=self.precise)
        if prescale_attention is None:
            prescale_attention = _prescale_allowed()
        self.prescale_attention = bool(prescale_attention) and not self.precise
        self.blocks = [_BlockPack(b, fold_lr_chains, self.precise, self.prescale_attention)
                       for b in model.transformer_blocks]
        # every adaLN linear of the model in one weight: temb is the same for all blocks, so one GEMM per step
        mods, self.mod_off = [], []
        off = 0
        for b in model.transformer_blocks:
            o1 = off
            mods.append(b.norm1.linear)
            off += b.norm1.linear.out_features
            o2 = off
            mods.append(b.norm1_context.linear)
            off += b.norm1_context.linear.out_features
            self.mod_off.append((o1, o2))
        self.mod_out_off = off
        mods.append(model.norm_out.linear)
        off += model.norm_out.linear.out_features
        self.mod_total = off
        self.mod = _Lin(*mods)
        pe = model.pos_embed.proj
        self.patch_w = _pack_w(pe.weight.reshape(D, -1), self.precise)
        self.patch_b = _f32(pe.bias)
        self.pos_table = model.pos_embed.pos_embed
        tt = model.time_text_embed
        self.t1, self.t2 = _Lin(tt.timestep_embedder.linear_1), _Lin(tt.timestep_embedder.linear_2)
        self.p1, self.p2 = _Lin(tt.text_embedder.linear_1), _Lin(tt.text_embedder.linear_2)
        self.ctx_embed = _Lin(model.context_embedder)
        self.proj_out = _Lin(model.proj_out)
        self._ws: Dict[Tuple, dict] = {}
        self._prof = None
        self._graphs: Dict[Tuple, tuple] = {}
        import os as _os
        self.use_cuda_graph = _os.environ.get("DIT4SR_CUDA_GRAPH", "1") != "0"
        # capturing NCCL collectives inside the graph hung on 2 GPUs in round 1: the torch.distributed path launches
        # directly; the NVLink peer-memory path consists of this library's kernels only and is captured
        self.sp_graph = _os.environ.get("DIT4SR_SP_GRAPH", "1" if (sp is not None and sp.peer) else "0") == "1"
        # peer path: q|k|v heads are stored into the owning ranks' buffers by the projection epilogues themselves
        # (DIT4SR_SP_FUSED_SCATTER=0: a separate scatter kernel after the projections, kept for A/B)
        self.sp_fused_scatter = _os.environ.get("DIT4SR_SP_FUSED_SCATTER", "1") != "0"
        self._pos: Dict[Tuple, torch.Tensor] = {}
        self._attn_ws: Dict = {}
        self._attn_split = _os.environ.get("DIT4SR_ATTN_SPLIT_KV", "1") != "0"

    # ------------------------------------------------------------------------------------------------ profiling
    def profile_kernel(self, pattern: str):
        """Record CUDA events around every launch of kernel class `pattern` ('*' = all) until profile_result()."""
        self._prof = (pattern, {})

    def profile_result(self):
        """{kernel class: [ms per launch]} for the launches recorded since profile_kernel(); stops recording."""
        if self._prof is None:
            return None
        torch.cuda.synchronize()
        res = {k: [a.elapsed_time(b) for a, b in v] for k, v in self._prof[1].items()}
        self._prof = None
        return res

    def _run(self, name, fn, *args, **kwargs):
        p = self._prof
        if p is not None and (p[0] == "*" or p[0] == name):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*args, **kwargs)
            e1.record()
            p[1].setdefault(name, []).append((e0, e1))
        else:
            fn(*args, **kwargs)

    # ------------------------------------------------------------------------------------------------ workspace
    def _pos_crop(self, hp, wp, device):
        key = (hp, wp)
        if key not in self._pos:
            mx = self.cfg.pos_embed_max_size
            if hp > mx or wp > mx:
                raise ValueError(f"latent grid {hp}x{wp} exceeds pos_embed_max_size {mx}")
            top, left = (mx - hp) // 2, (mx - wp) // 2
            pe = self.pos_table.reshape(1, mx, mx, -1)[:, top:top + hp, left:left + wp, :]
            self._pos[key] = pe.reshape(1, hp * wp, -1).to(device=device, dtype=torch.float32).contiguous()
        return self._pos[key]

    def _workspace(self, B, N, T, device, grid_w=0):
        key = (B, N, T, str(device), grid_w)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        D, F = self.D, 4 * self.D
        I = self.streams * N            # image tokens
        S = I + T
        f32 = torch.float32
        bf = f32 if self.precise else torch.bfloat16     # activation dtype between kernels

        def e(*shape, dtype=bf):
            return torch.empty(*shape, device=device, dtype=dtype)

        ws = dict(
            x=e(B, I, D, dtype=f32), ctx=e(B, T, D, dtype=f32),
            n1=e(B, I, D), n2=e(B, I, D), nc=e(B, T, D),
            qkv=e(B, S, 3 * D), o=e(B, S, D), lrtmp=e(B, I // 2, 3 * D),
            u=e(B, I, F), uc=e(B, T, F),
            mod=e(B, self.mod_total, dtype=f32),
            tf=e(B, 256), th=e(B, D), ph=e(B, D), temb=e(B, D, dtype=f32), st=e(B, D),
            pooled=e(B, self.cfg.pooled_projection_dim), enc=e(B, T, self.cfg.joint_attention_dim),
            patches=e(B, N, 4 * self.cfg.in_channels), tok=e(B, N, 4 * self.cfg.out_channels, dtype=f32),
            t=e(B, dtype=f32),
        )
        if self.sp is not None and not self.sp.peer:
            # attention-side buffers of the head<->sequence exchange + halo-padded LR hidden state (torch.distributed path)
            P = self.sp.world
            Dl = D // P
            ws.update(qkv_sp=e(B, P * I + T, 3, Dl), o_sp=e(B, P * I + T, Dl),
                      lr_pad=e(B, N + 2 * grid_w, F))
        elif self.sp is not None:
            # NVLink peer-memory path: the buffers the PEERS write into are peer-mapped allocations (collective set-up:
            # every rank builds the same workspaces in the same order).  q|k|v land on the attention side, attention
            # outputs come back into `o`, and the MLP hidden state carries one halo grid row above and below its LR rows:
            #   u rows: [0,N) HQ | [N,N+gw) halo above | [N+gw, 2N+gw) LR | [2N+gw, 2N+2gw) halo below
            sp, P = self.sp, self.sp.world
            Dl = D // P
            ws["qkv_sp"], ws["qkv_sp_peers"] = sp.shared_empty((B * (P * I + T) * 3 * Dl,))
            ws["o"], ws["o_peers"] = sp.shared_empty((B, S, D))
            ws["u"], ws["u_peers"] = sp.shared_empty((B, I + 2 * grid_w, F))
            ws["sp_out"] = ops.attn_sp_out(P, I, sp.rank * Dl, ws["o_peers"])
        self._ws[key] = ws
        return ws

    def _attn_scratch(self, device):
        """Zero-filled scratch of the attention kernel's split-KV tail (one per engine and device; the kernel keeps its
        counters zero between launches).  DIT4SR_ATTN_SPLIT_KV=0 turns the splitting off."""
        if not self._attn_split:
            return None
        t = self._attn_ws.get(device)
        if t is None:
            t = torch.zeros(ops.attention_workspace_bytes(), device=device, dtype=torch.uint8)
            self._attn_ws[device] = t
        return t

    def _feature_buffers(self, features, B, grid, device):
        """Static outputs of the per-block feature kernels for one feature spec (allocated once: graph replays refill them)."""
        if not hasattr(self, "_feat_bufs"):
            self._feat_bufs = {}
        bufs = self._feat_bufs.get(features)
        layers, row0, n_concat, tdt = features
        if bufs is None or bufs[0][0].shape[0] != B or bufs[0][1].shape[-1] != 2 * grid or bufs[0][0].device != device:
            bufs = [(torch.empty(B, n_concat * grid * grid, self.D, device=device, dtype=tdt),
                     torch.empty(B, n_concat * self.D // 4, 2 * grid, 2 * grid, device=device, dtype=torch.float32))
                    for _ in layers]
            self._feat_bufs[features] = bufs
        return bufs

    # ------------------------------------------------------------------------------------------------ GEMM dispatch
    def _scr(self, key, numel, dtype):
        """Grow-only scratch buffers of the fp32-accurate mode (operand splits, per-head scores)."""
        t = self._scratch.get(key)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(numel, device=self._dev, dtype=dtype)
            self._scratch[key] = t
        return t[:numel]

    def _split_a(self, A, slot):
        """fp32 activations [B,R,K] (any row/batch stride) -> contiguous bf16 [B,R,3K] = [hi | hi | lo]."""
        A3 = A.unsqueeze(0) if A.dim() == 2 else A
        B, Rr, K = A3.shape
        out = self._scr(("a3", slot), B * Rr * 3 * K, torch.bfloat16)
        self._run("split", ops.split3, A3, out, 0)
        out = out.view(B, Rr, 3 * K)
        return out[0] if A.dim() == 2 else out

    def _gg(self, name, problems):
        """One grouped GEMM launch; in the fp32-accurate mode the fp32 A operands are split first (K' = 3K)."""
        if self.precise:
            problems = [dict(p, A=self._split_a(p["A"], i)) for i, p in enumerate(problems)]
        self._run(name, ops.gemm_grouped, problems)

    def _g1(self, name, A, lin_w, bias, out, epilogue, **kw):
        self._gg(name, [dict(A=A, W=lin_w, bias=bias, out=out, epilogue=epilogue, **kw)])

    def _attention_core_precise(self, ws, S, seq_q, n_in, lr_rows):
        """softmax(q k^T / 8) v per (batch, head) in the fp32-accurate mode (attention.py:528): scores and P V on the
        GEMM kernel with split operands, fp32 row softmax in between; then the Q2 residual on the LR rows."""
        D, H, R = self.D, self.H, self._run
        qkv, o = ws["qkv"], ws["o"]
        B = qkv.shape[0]
        Sp = (S + 127) // 128 * 128
        bf = torch.bfloat16
        q3 = self._scr("q3", H * seq_q * 192, bf).view(H, seq_q, 192)
        k3 = self._scr("k3", H * Sp * 192, bf).view(H, Sp, 192)
        sc = self._scr("sc", H * seq_q * Sp, torch.float32).view(H, seq_q, Sp)
        p3 = self._scr("p3", H * seq_q * 3 * Sp, bf).view(H, seq_q, 3 * Sp)
        vt = self._scr("vt", H * 64 * 3 * Sp, bf).view(H, 64, 3 * Sp)
        G = ops.GEMM_MAX_GROUP
        for b in range(B):
            def heads(c0, rows):   # [H, rows, 64] view of one q/k/v section of image b
                return qkv[b, :rows, c0:c0 + D].unflatten(-1, (H, 64)).permute(1, 0, 2)
            R("split", ops.split3, heads(0, seq_q), q3.view(-1), 0)
            R("split", ops.split3, heads(D, S), k3.view(-1), 1, Sp)
            for h0 in range(0, H, G):
                R("attn", ops.gemm_grouped, [dict(A=q3[h], W=k3[h], bias=None, out=sc[h], epilogue=ops.EPI_BIAS)
                                             for h in range(h0, min(H, h0 + G))])
            R("attn", ops.softmax_rows, sc.view(H * seq_q, Sp), S, 0.125)
            R("split", ops.split3, sc.view(1, H * seq_q, Sp), p3.view(-1), 0)
            for h in range(H):
                R("split", ops.split3_transposed, qkv[b, :S, 2 * D + h * 64: 2 * D + (h + 1) * 64], vt[h].view(-1), Sp)
            for h0 in range(0, H, G):
                R("attn", ops.gemm_grouped, [dict(A=p3[h], W=vt[h], bias=None, out=o[b, :seq_q, h * 64:(h + 1) * 64],
                                                  epilogue=ops.EPI_BIAS) for h in range(h0, min(H, h0 + G))])
        r0, r1 = lr_rows
        if r1 > r0:
            R("attn", ops.add_f32, o[:, r0:r1], n_in[:, r0:r1])

    # ------------------------------------------------------------------------------------------------ attention
    def _attention(self, ap: _AttnPack, ws, n_in, nc, N, T, gate, c_gate, last):
        """JointAttnProcessor2_0 (attention.py:445-559) + the gated residual adds (attention.py:314-320,381-382).

        Three grouped GEMM launches + one attention launch: {HQ q|k|v, LR stage-1, text q|k|v}, {LR q_c, (k_c,) v_c},
        attention, {to_out, to_out_control, to_add_out}."""
        D, H, R = self.D, self.H, self._run
        x, ctx, qkv, o = ws["x"], ws["ctx"], ws["qkv"], ws["o"]
        lrtmp = ws["lrtmp"][:, :N]        # sized for the larger segment (attn2 of the 3-stream model: 1.5N rows)
        I = (3 if ap.ocr else 2) * N      # image rows seen by this processor (N = rows per segment)
        S = I + (T if ap.has_text else 0)
        hq, lr = n_in[:, :N], n_in[:, N:2 * N]
        E = ops.EPI_QKV_NORM if ap.qk_norm else ops.EPI_BIAS
        nq = dict(norm_w0=ap.wq, sec_cols=D) if ap.qk_norm else {}
        g1 = [dict(A=hq, W=ap.qkv.w, bias=ap.qkv.b, out=qkv[:, :N], epilogue=E,
                   **(dict(norm_w0=ap.wq, norm_w1=ap.wk, sec_cols=D, dbl_row_offset=N) if ap.qk_norm else {}))]
        #          ^ Q1: with qk_norm this launch also emits K_lr = norm_k(norm_k(k_hq)) N rows further down
        g2 = []
        if ap.fold:       # LR projections straight from the LR tokens with the folded weights
            g1.append(dict(A=lr, W=ap.qf.w, bias=ap.qf.b, out=qkv[:, N:2 * N, :D], epilogue=E, **nq))
            g1.append(dict(A=lr, W=ap.vf.w, bias=ap.vf.b, out=qkv[:, N:2 * N, 2 * D:], epilogue=E))
            if not ap.qk_norm:
                g2.append(dict(A=lr, W=ap.kf.w, bias=ap.kf.b, out=qkv[:, N:2 * N, D:2 * D], epilogue=E))
        elif ap.qk_norm:  # as written: stage 1 = to_q|to_v of the LR stream, stage 2 = the control projections
            t2 = lrtmp[:, :, :2 * D]
            g1.append(dict(A=lr, W=ap.lr1.w, bias=ap.lr1.b, out=t2, epilogue=E))
            g2 = [dict(A=t2[:, :, :D], W=ap.qc.w, bias=ap.qc.b, out=qkv[:, N:2 * N, :D], epilogue=E, **nq),
                  dict(A=t2[:, :, D:], W=ap.vc.w, bias=ap.vc.b, out=qkv[:, N:2 * N, 2 * D:], epilogue=E)]
        else:
            g1.append(dict(A=lr, W=ap.qkv.w, bias=ap.qkv.b, out=lrtmp, epilogue=E))
            g2 = [dict(A=lrtmp[:, :, :D], W=ap.qc.w, bias=ap.qc.b, out=qkv[:, N:2 * N, :D], epilogue=E),
                  dict(A=lrtmp[:, :, D:2 * D], W=ap.kc.w, bias=ap.kc.b, out=qkv[:, N:2 * N, D:2 * D], epilogue=E),
                  dict(A=lrtmp[:, :, 2 * D:], W=ap.vc.w, bias=ap.vc.b, out=qkv[:, N:2 * N, 2 * D:], epilogue=E)]
        if ap.ocr:        # OCR stream: folded q|k|v chains; norm_q on q, a live norm_k on k (ocr file :585,:589)
            g1.append(dict(A=n_in[:, 2 * N:I], W=ap.of.w, bias=ap.of.b, out=qkv[:, 2 * N:I], epilogue=E,
                           **(dict(norm_w0=ap.wq, norm_w1=ap.wk, sec_cols=D) if ap.qk_norm else {})))
        if ap.has_text:
            g1.append(dict(A=nc, W=ap.add.w, bias=ap.add.b, out=qkv[:, I:S], epilogue=E,
                           **(dict(norm_w0=ap.waq, norm_w1=ap.wak, sec_cols=D) if ap.qk_norm else {})))
        fused_scatter = self.sp is not None and self.sp.peer and self.sp_fused_scatter
        if fused_scatter:
            # sequence -> head all-to-all fused into the projection epilogues: every problem that writes into the joint
            # q|k|v buffer stores its heads straight into the owning ranks' attention-side buffers instead
            P, rk = self.sp.world, self.sp.rank
            T_ = T if ap.has_text else 0
            base = qkv.data_ptr()
            row_bytes, sec_bytes = qkv.stride(1) * 2, D * 2
            keep = ws.setdefault("_sp_desc", [])
            keep.clear()
            for pr in g1 + g2:
                off = pr["out"].data_ptr() - base
                if not 0 <= off < qkv.stride(0) * 2:
                    continue                                   # a scratch output (as-written LR stage 1): stays local
                row, col = divmod(off, row_bytes)
                text = row >= I
                d = ops.gemm_sp(P, rk, D, P * I + T_, ws["qkv_sp_peers"], row0=(P * I + row - I) if text else rk * I + row,
                                col0=col // 2, own_heads_only=text)
                keep.append(d)
                pr["sp"] = d
        self._gg("qkv_group", g1)
        if g2:
            self._gg("lr2_group", g2)
        # the last block's text queries are discarded by the reference (context_pre_only): do not compute them
        if self.precise:
            self._attention_core_precise(ws, S, I if (last or not ap.has_text) else S, n_in, (N, I))
        elif self.sp is None:
            seq_q = I if (last or not ap.has_text) else S
            R("attn", ops.attention, qkv[:, :S, :D], qkv[:, :S, D:2 * D], qkv[:, :S, 2 * D:], o, H, seq_q=seq_q,
              lr_res=n_in, lr_rows=(N, I),   # Q2: the LR (and OCR) rows get their normed+modulated input added back
              score_bound=ap.score_bound, workspace=self._attn_scratch(qkv.device))
        else:
            self._attention_sp(ap, ws, n_in, N, I, S, last)
        G = ops.EPI_GATE_RESID
        g3 = [dict(A=o[:, :N], W=ap.out.w, bias=ap.out.b, out=x[:, :N], epilogue=G, gate=gate),
              dict(A=o[:, N:2 * N], W=ap.outc.w, bias=ap.outc.b, out=x[:, N:2 * N], epilogue=G, gate=gate)]
        if ap.ocr:
            g3.append(dict(A=o[:, 2 * N:I], W=ap.outo.w, bias=ap.outo.b, out=x[:, 2 * N:I], epilogue=G, gate=gate))
        if ap.has_text and not last:
            g3.append(dict(A=o[:, I:S], W=ap.add_out.w, bias=ap.add_out.b, out=ctx, epilogue=G, gate=c_gate))
        self._gg("out_group", g3)

    def _attention_sp(self, ap, ws, n_in, N, I, S, last):
        """Ulysses: all-to-all (tokens -> heads), local attention on H/P heads over the whole joint sequence,
        all-to-all back, all-gather of the text rows over heads (parallel.SequenceParallel)."""
        sp, R, H = self.sp, self._run, self.H
        P = sp.world
        Hl = H // P
        Dl = Hl * 64
        qkv, o = ws["qkv"], ws["o"]
        B = qkv.shape[0]
        T = S - I
        if sp.peer:
            # one scatter kernel + arrival barrier, attention with the way back fused into its epilogue, arrival barrier
            Ssp = P * I + T
            qs = ws["qkv_sp"][:B * Ssp * 3 * Dl].view(B, Ssp, 3, Dl)
            R("a2a", ops.a2a_qkv, sp.comm, None if self.sp_fused_scatter else qkv[:, :S], I, T, H, ws["qkv_sp_peers"])
            seq_q = P * I if (last or not T) else Ssp
            R("attn", ops.attention, qs[:, :, 0], qs[:, :, 1], qs[:, :, 2], o, Hl, seq_q=seq_q,
              score_bound=ap.score_bound, sp_out=ws["sp_out"], workspace=self._attn_scratch(qkv.device))
            R("a2a", ops.a2a_out, sp.comm, None, I, T, H, ws["o_peers"], o.stride(1), o.stride(0))
            R("attn", ops.add_bf16, o[:, N:I], n_in[:, N:I])                # Q2 residual on the LR rows
            return
        qs, os_ = ws["qkv_sp"], ws["o_sp"]
        xs = sp.seq_to_heads(qkv[:, :I].reshape(B, I, 3, H, 64))           # [B, P*I, 3, Hl, 64] rank-major tokens
        qs[:, :P * I] = xs.reshape(B, P * I, 3, Dl)
        if T:
            h0, h1 = sp.head_slice(H)
            qs[:, P * I:P * I + T] = qkv[:, I:S].reshape(B, T, 3, H, 64)[:, :, :, h0:h1].reshape(B, T, 3, Dl)
        Ssp = P * I + T
        seq_q = P * I if (last or not T) else Ssp
        R("attn", ops.attention, qs[:, :Ssp, 0], qs[:, :Ssp, 1], qs[:, :Ssp, 2], os_, Hl, seq_q=seq_q,
          score_bound=ap.score_bound)
        o[:, :I] = sp.heads_to_seq(os_[:, :P * I])
        if T and not last:
            o[:, I:S] = sp.gather_heads(os_[:, P * I:Ssp])
        R("attn", ops.add_bf16, o[:, N:I], n_in[:, N:I])                    # Q2 residual on the LR (and OCR) rows

    def _conv_add(self, ws, dst, src, w9, b, rows, grid_w):
        """dst += dwconv3x3(src) + b on a (rows x grid_w) token grid; with sequence parallelism the source rows get a
        1-row halo from the neighbouring ranks first."""
        if self.precise:
            self._run("dwconv", ops.dwconv3x3_add_f32, dst, src, w9, b, rows, grid_w=grid_w)
        elif self.sp is None:
            self._run("dwconv", ops.dwconv3x3_add, dst, src, w9, b, rows, grid_w=grid_w)
        elif self.sp.peer:   # the neighbours store their boundary rows straight into this rank's halo rows of u
            u, N = ws["u"], dst.shape[1]
            self._run("halo", ops.halo_rows, self.sp.comm, u, ws["u_peers"], grid_w, N + grid_w, N, N, 2 * N + grid_w)
            self._run("dwconv", ops.dwconv3x3_add, dst, u[:, N:], w9, b, rows, grid_w=grid_w, lr_row_offset=1)
        else:
            padded = self.sp.halo_rows(src, grid_w, out=ws["lr_pad"])
            self._run("dwconv", ops.dwconv3x3_add, dst, padded, w9, b, rows, grid_w=grid_w, lr_row_offset=1)

    # ------------------------------------------------------------------------------------------------ CUDA graph
    @torch.no_grad()
    def step(self, hidden_states, controlnet_image, encoder_hidden_states, pooled_projections, timestep,
             out=None, euler: Optional[Tuple[torch.Tensor, float]] = None, sharded_io: bool = False,
             guidance: Optional[float] = None, features: Optional[tuple] = None, loop=None):
        """One denoiser call, replayed from a captured CUDA graph when possible (the ~300 kernel launches of a step are
        recorded once per input shape; replay removes the per-launch host cost, which matters for small batches).
        Same results as `forward` -- the same kernels run. Falls back to direct launches while profiling, if capture is
        unavailable, and under sequence parallelism over torch.distributed (the NVLink peer-memory path is captured).

        sharded_io (sequence parallelism): the image tensors are this rank's row strips [B, C, 2*grid/P, W] and so is the
        result - a sampling loop keeps its latents sharded and gathers once at the end (sampler.denoise).
        guidance: classifier-free guidance - the inputs are the doubled batch [uncond | text] (2B) and the result (velocity
        or Euler update of euler[0], B images) uses v_uncond + guidance * (v_text - v_uncond), fused into the last kernel.
        features: (layers, row0, n_concat, token dtype) - per-block feature outputs (see `forward`); with them the call
        returns (result, [ {extract_feat, extract_feat_nchw_f32}, ... ]) and still replays the graph.
        loop: (cache, i) from `prepare_loop` - step i of a sampling loop takes its adaLN vectors and text embedding from the
        loop's table (the head of the step is not recomputed)."""
        if features is not None and self.sp is not None:
            raise NotImplementedError("feature outputs are not available under sequence parallelism")
        if self.sp is not None and not sharded_io:
            # whole images in, whole image out: cut my strip, run the sharded step, gather the strips (outside any graph)
            g0, g1 = self.sp.row_slice(hidden_states.shape[2] // 2)
            loc = lambda t: t[:, :, 2 * g0:2 * g1].contiguous()
            res = self.step(loc(hidden_states), loc(controlnet_image), encoder_hidden_states, pooled_projections, timestep,
                            euler=None if euler is None else (loc(euler[0]), euler[1]), sharded_io=True, guidance=guidance,
                            loop=loop)
            gathered = self.sp.gather_rows(res, dim=2)
            if out is not None:
                out.copy_(gathered)
                return out
            return gathered
        if (not self.use_cuda_graph or self._prof is not None or self.precise
                or (self.sp is not None and not self.sp_graph)):
            res = self.forward(hidden_states, controlnet_image, encoder_hidden_states, pooled_projections, timestep,
                               out=out, euler=euler, sharded_io=sharded_io, guidance=guidance, features=features, loop=loop)
            return (res, self._feature_results(features)) if features is not None else res
        B = hidden_states.shape[0]
        key = (tuple(hidden_states.shape), hidden_states.dtype, tuple(encoder_hidden_states.shape),
               encoder_hidden_states.dtype, pooled_projections.dtype, str(hidden_states.device), guidance is not None,
               features, loop is not None)
        h, w = hidden_states.shape[-2:]
        ws = self._workspace(B, (h // 2) * (w // 2), encoder_hidden_states.shape[1], hidden_states.device,
                             grid_w=w // 2 if self.sp is not None else 0)
        if loop is not None:
            self._load_loop_slice(ws, loop)      # this step's adaLN vectors / text embedding -> the buffers the graph reads
        entry = self._graphs.get(key)
        if entry is None:
            st = dict(hs=hidden_states.clone().contiguous(), ci=controlnet_image.clone().contiguous(),
                      enc=encoder_hidden_states.clone().contiguous(), pooled=pooled_projections.clone().contiguous(),
                      t=torch.zeros(B, device=hidden_states.device, dtype=torch.float32))
            try:
                fw = lambda: self.forward(st["hs"], st["ci"], st["enc"], st["pooled"], st["t"], sharded_io=sharded_io,
                                          guidance=guidance, features=features,
                                          loop="preloaded" if loop is not None else None)
                fw()                                                                    # warm-up outside capture
                torch.cuda.synchronize()
                n0 = ops.launch_count()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    vel = fw()
                entry = (graph, st, vel, ops.launch_count() - n0)
            except Exception as e:   # capture not possible here: keep launching directly
                if self.sp is not None:
                    raise            # the ranks must not diverge (one replaying, another launching directly)
                import warnings
                warnings.warn(f"dit4sr_b200: CUDA graph capture failed ({e}); using direct launches")
                self.use_cuda_graph = False
                torch.cuda.synchronize()
                res = self.forward(hidden_states, controlnet_image, encoder_hidden_states, pooled_projections,
                                   timestep, out=out, euler=euler, guidance=guidance, features=features, loop=loop)
                return (res, self._feature_results(features)) if features is not None else res
            self._graphs[key] = entry
        graph, st, vel, n_launch = entry
        st["hs"].copy_(hidden_states)
        st["ci"].copy_(controlnet_image)
        if loop is None:             # (with a loop cache the graph reads neither: their products are in the workspace)
            st["enc"].copy_(encoder_hidden_states)
            st["pooled"].copy_(pooled_projections)
            st["t"].copy_(timestep.to(device=st["t"].device, dtype=torch.float32).expand(B))
        graph.replay()
        ops._count(n_launch)
        if euler is not None:        # x_next = x + dsigma * v straight from the velocity tokens the graph left behind
            if out is None:
                out = torch.empty_like(euler[0])
            ops.unpatchify(ws["tok"], out, latents=euler[0], dsigma=euler[1], guidance=guidance)
            res = out
        elif out is not None:
            out.copy_(vel)
            res = out
        else:
            res = vel.clone()
        return (res, self._feature_results(features)) if features is not None else res

    LOOP_TABLE_MAX_BYTES = 4 << 30

    @torch.no_grad()
    def prepare_loop(self, encoder_hidden_states, pooled_projections, timesteps):
        """What a sampling loop can form once instead of once per step (pipelines/pipeline_dit4sr.py:1078-1132 calls the model
        with the same prompt embeddings at every step and the schedule is known up front): the adaLN vectors of EVERY step
        - temb = MLP(sincos(t_i)) + MLP(pooled) and the [B, 325 D] modulation GEMM, one launch with steps x B rows, so its
        1.5 GB of weights are streamed once per loop instead of once per step - and the embedded text tokens
        (context_embedder).  Same kernels and operands as the per-step head, so the values are identical.  Returns the
        cache `step(..., loop=(cache, i))` takes, or None when the table would not be worth its memory (or in the
        fp32-accurate mode, which keeps the as-written order)."""
        B = encoder_hidden_states.shape[0]
        n = len(timesteps)
        if self.precise or n * B * self.mod_total * 4 > self.LOOP_TABLE_MAX_BYTES:
            return None
        dev, D, R, gm = encoder_hidden_states.device, self.D, self._run, self._g1
        bf, f32 = torch.bfloat16, torch.float32
        M = n * B

        def e(*shape, dtype=bf):
            return torch.empty(*shape, device=dev, dtype=dtype)

        t_all = torch.tensor([float(t) for t in timesteps], device=dev, dtype=f32).repeat_interleave(B)   # rows [step, batch]
        tf, th, ph, temb, st = e(M, 256), e(M, D), e(M, D), e(M, D, dtype=f32), e(M, D)
        R("head", ops.timestep_features, t_all, tf)
        gm("head", tf, self.t1.w, self.t1.b, th, ops.EPI_BIAS_SILU)
        gm("head", th, self.t2.w, self.t2.b, temb, ops.EPI_BIAS)
        pooled = pooled_projections.to(bf).repeat(n, 1).contiguous()
        gm("head", pooled, self.p1.w, self.p1.b, ph, ops.EPI_BIAS_SILU)
        gm("head", ph, self.p2.w, self.p2.b, temb, ops.EPI_GATE_RESID)
        R("head", ops.silu_cast, temb, st)
        mod_all = e(M, self.mod_total, dtype=f32)
        gm("mod_gemm", st, self.mod.w, self.mod.b, mod_all, ops.EPI_BIAS)
        enc = encoder_hidden_states.to(bf).contiguous()
        ctx0 = e(B, enc.shape[1], D, dtype=f32)
        gm("head", enc, self.ctx_embed.w, self.ctx_embed.b, ctx0, ops.EPI_BIAS)
        return dict(mod=mod_all.view(n, B, self.mod_total), ctx=ctx0, steps=n, batch=B)

    def _load_loop_slice(self, ws, loop):
        cache, i = loop
        if cache["batch"] != ws["mod"].shape[0] or not 0 <= i < cache["steps"]:
            raise ValueError("loop cache does not belong to this batch / step")
        if "ctx0" not in ws:
            ws["ctx0"] = torch.empty_like(ws["ctx"])
        ws["mod"].copy_(cache["mod"][i])
        ws["ctx0"].copy_(cache["ctx"])

    def _feature_results(self, features):
        """Fresh copies of the per-block feature buffers the last forward / replay filled (they are static workspace)."""
        return [dict(extract_feat=t.clone(), extract_feat_nchw_f32=n.clone()) for t, n in self._feat_bufs[features]]

    # ------------------------------------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, hidden_states, controlnet_image, encoder_hidden_states, pooled_projections, timestep,
                out=None, euler: Optional[Tuple[torch.Tensor, float]] = None, feat_hook=None, sharded_io: bool = False,
                guidance: Optional[float] = None, features: Optional[tuple] = None, loop=None):
        """loop = (cache of `prepare_loop`, step index): the adaLN vectors and the embedded text tokens come from the
        sampling loop's table instead of being recomputed (or "preloaded": `step` has put them into the workspace).
        features = (layers, row0, n_concat, token dtype): after each block in `layers` the rows [row0, row0 + n_concat *
        grid^2) of the residual stream are written (one kernel, also inside a captured graph) as the token-major slice the
        reference returns (attention.py:363-374) and as fp32 NCHW 'b (N H W) (pH pW d) -> b (N d) (H pH) (W pW)', what
        pipelines/pipeline_dit4sr.py:1214-1245 hands to the text-spotting module."""
        cfg, D, R = self.cfg, self.D, self._run
        B, Cin, h, w = hidden_states.shape
        sharded = sharded_io and self.sp is not None   # the image tensors are already this rank's row strips
        if w % 2 or h % 2 or (not sharded and h != w) or (sharded and h * self.sp.world != w):
            raise ValueError("dit4sr_b200: square, even latents only (reference asserts a square token grid, "
                             "attention.py:169-172)")
        grid = w // 2
        rows = h // 2                   # token-grid rows handled by this rank
        pos = self._pos_crop(grid, grid, hidden_states.device)
        full_out = out
        if self.sp is not None:         # Ulysses: this rank keeps grid/P contiguous rows of every image stream
            if self.streams == 3 and (self.sp.peer or any(b.dual for b in self.blocks)):
                # Q6: attn2 of the 3-stream model pairs token j of the first 1.5N tokens with token j of the last 1.5N;
                # under row sharding those partners live on different ranks. BASELINE config #5 is data-parallel.
                raise NotImplementedError("sequence parallelism is not supported for the OCR-branch model with dual "
                                          "attention blocks (use data-parallel sampling, BASELINE config #5)")
            if self.H % self.sp.world:
                raise ValueError(f"{self.H} heads do not divide over {self.sp.world} ranks")
            g0, g1 = self.sp.row_slice(grid)
            if not sharded:
                rows = g1 - g0
                hidden_states = hidden_states[:, :, 2 * g0:2 * g1].contiguous()
                controlnet_image = controlnet_image[:, :, 2 * g0:2 * g1].contiguous()
                out = None
                if euler is not None:
                    euler = (euler[0][:, :, 2 * g0:2 * g1].contiguous(), euler[1])
            pos = pos[:, g0 * grid:g1 * grid]
        N = rows * grid
        if self.streams == 3 and (3 * N) % 2:
            raise ValueError("3-stream model: attn2 halves the 3N image tokens, N must be even")
        T = encoder_hidden_states.shape[1]
        dev = hidden_states.device
        self._dev = dev
        ws = self._workspace(B, N, T, dev, grid_w=grid if self.sp is not None else 0)
        x, ctx, n1, n2, nc, mod = ws["x"], ws["ctx"], ws["n1"], ws["n2"], ws["nc"], ws["mod"]

        # ---- head: temb = MLP(sincos(t)) + MLP(pooled)  (CombinedTimestepTextProjEmbeddings, transformer_sd3.py:327)
        gm = self._g1
        if loop is None:
            ws["t"].copy_(timestep.to(device=dev, dtype=torch.float32).expand(B))
            R("head", ops.timestep_features, ws["t"], ws["tf"])
            gm("head", ws["tf"], self.t1.w, self.t1.b, ws["th"], ops.EPI_BIAS_SILU)
            gm("head", ws["th"], self.t2.w, self.t2.b, ws["temb"], ops.EPI_BIAS)
            ws["pooled"].copy_(pooled_projections)
            gm("head", ws["pooled"], self.p1.w, self.p1.b, ws["ph"], ops.EPI_BIAS_SILU)
            gm("head", ws["ph"], self.p2.w, self.p2.b, ws["temb"], ops.EPI_GATE_RESID)
            R("head", ops.silu_cast, ws["temb"], ws["st"])
            gm("mod_gemm", ws["st"], self.mod.w, self.mod.b, mod, ops.EPI_BIAS)   # all adaLN vectors of the step
        else:
            # sampling loop: timesteps, pooled and text embeddings are known before the loop starts, so the adaLN vectors of
            # every step and the embedded text tokens were formed once (prepare_loop); a step copies its slice
            if "ctx0" not in ws:
                ws["ctx0"] = torch.empty_like(ctx)
            if loop != "preloaded":
                self._load_loop_slice(ws, loop)
            ctx.copy_(ws["ctx0"])
        # ---- patch embed (+pos) for both streams, context embedder  (:326-332)
        streams = [(hidden_states, x[:, :N]), (controlnet_image, x[:, N:2 * N])]
        if self.streams == 3:   # the OCR stream starts as a second pos_embed(controlnet_image) (ocr model :334-337)
            streams.append((controlnet_image, x[:, 2 * N:]))
        for src, dst_rows in streams:
            src = src.contiguous()
            R("head", ops.patchify, src.float() if self.precise else src, ws["patches"])
            gm("head", ws["patches"], self.patch_w, self.patch_b, dst_rows, ops.EPI_GATE_RESID, resid=pos)
        if loop is None:
            ws["enc"].copy_(encoder_hidden_states)
            gm("head", ws["enc"], self.ctx_embed.w, self.ctx_embed.b, ctx, ops.EPI_BIAS)

        # ---- blocks (:336-370)
        for i, bp in enumerate(self.blocks):
            o1, o2 = self.mod_off[i]

            def ch(base, k):
                return mod[:, base + k * D: base + (k + 1) * D]

            shift_msa, scale_msa, gate_msa = ch(o1, 0), ch(o1, 1), ch(o1, 2)
            shift_mlp, scale_mlp, gate_mlp = ch(o1, 3), ch(o1, 4), ch(o1, 5)
            # the text stream's LayerNorm rides in the image stream's launch (companion problem of dit4sr_ln_modulate)
            if bp.last:   # AdaLayerNormContinuous: (scale, shift) order (Q12)
                c_scale, c_shift = ch(o2, 0), ch(o2, 1)
                c_gate_msa = None
                text_ln = (ctx, c_shift, c_scale, nc)
            else:
                c_shift_msa, c_scale_msa, c_gate_msa = ch(o2, 0), ch(o2, 1), ch(o2, 2)
                c_shift_mlp, c_scale_mlp, c_gate_mlp = ch(o2, 3), ch(o2, 4), ch(o2, 5)
                text_ln = (ctx, c_shift_msa, c_scale_msa, nc)
            if bp.dual:
                shift_msa2, scale_msa2, gate_msa2 = ch(o1, 6), ch(o1, 7), ch(o1, 8)
                R("ln_mod", ops.ln_modulate, x, shift_msa, scale_msa, n1, shift_msa2, scale_msa2, n2, companion=text_ln)
            else:
                R("ln_mod", ops.ln_modulate, x, shift_msa, scale_msa, n1, companion=text_ln)
            self._attention(bp.attn, ws, n1, nc, N, T, gate_msa, c_gate_msa, bp.last)
            if bp.dual:
                # 3-stream model: attn2 is still the 2-stream processor, so it halves the 3N tokens at 1.5N (Q6)
                n_seg = N if self.streams == 2 else (3 * N) // 2
                self._attention(bp.attn2, ws, n2, None, n_seg, T, gate_msa2, None, bp.last)
            # MLP with the LR->HQ depthwise conv (attention.py:322-347, FeedForwardControl :155-178)
            R("ln_mod", ops.ln_modulate, x, shift_mlp, scale_mlp, n1,
              companion=None if bp.last else (ctx, c_shift_mlp, c_scale_mlp, nc))
            if self.sp is not None and self.sp.peer:   # u keeps a halo grid row on either side of its LR rows
                u_hq, u_lr = ws["u"][:, :N], ws["u"][:, N + grid:2 * N + grid]
                ff1 = [dict(A=n1[:, :N], W=bp.ff1.w, bias=bp.ff1.b, out=u_hq, epilogue=ops.EPI_BIAS_GELU),
                       dict(A=n1[:, N:], W=bp.ff1.w, bias=bp.ff1.b, out=u_lr, epilogue=ops.EPI_BIAS_GELU)]
                ff2 = [dict(A=u_hq, W=bp.ff2.w, bias=bp.ff2.b, out=x[:, :N], epilogue=ops.EPI_GATE_RESID, gate=gate_mlp),
                       dict(A=u_lr, W=bp.ff2.w, bias=bp.ff2.b, out=x[:, N:], epilogue=ops.EPI_GATE_RESID, gate=gate_mlp)]
            else:
                u_hq, u_lr = ws["u"][:, :N], ws["u"][:, N:2 * N]
                ff1 = [dict(A=n1, W=bp.ff1.w, bias=bp.ff1.b, out=ws["u"], epilogue=ops.EPI_BIAS_GELU)]
                ff2 = [dict(A=ws["u"], W=bp.ff2.w, bias=bp.ff2.b, out=x, epilogue=ops.EPI_GATE_RESID, gate=gate_mlp)]
            if not bp.last:   # text MLP (attention.py:384-393) rides in the same launches
                ff1.append(dict(A=nc, W=bp.ffc1.w, bias=bp.ffc1.b, out=ws["uc"], epilogue=ops.EPI_BIAS_GELU))
                ff2.append(dict(A=ws["uc"], W=bp.ffc2.w, bias=bp.ffc2.b, out=ctx, epilogue=ops.EPI_GATE_RESID,
                                gate=c_gate_mlp))
            self._gg("ff1", ff1)
            u = ws["u"]
            self._conv_add(ws, u_hq, u_lr, bp.conv_w9, bp.conv_b, rows, grid)
            if bp.ocr:   # u_hq += conv_o(u_ocr); then u_ocr += conv_h(u_hq updated)  (ocr attention file :184-206)
                self._conv_add(ws, u[:, :N], u[:, 2 * N:], bp.conv_o_w9, bp.conv_o_b, rows, grid)
                self._conv_add(ws, u[:, 2 * N:], u[:, :N], bp.conv_h_w9, bp.conv_h_b, rows, grid)
            self._gg("ff2", ff2)
            if feat_hook is not None:
                feat_hook(i, x)
            if features is not None and i in features[0]:
                tokb, nchw = self._feature_buffers(features, B, grid, dev)[features[0].index(i)]
                R("feat", ops.feature_extract, x, features[1], features[2], grid, grid, nchw=nchw, tok=tokb)

        # ---- tail: drop the LR half, norm_out (scale, shift order), proj_out, unpatchify (:379-391)
        so = self.mod_out_off
        R("tail", ops.ln_modulate, x[:, :N], mod[:, so + D: so + 2 * D], mod[:, so: so + D], n1[:, :N])
        gm("tail", n1[:, :N], self.proj_out.w, self.proj_out.b, ws["tok"], ops.EPI_BIAS)
        if guidance is not None and B % 2:
            raise ValueError("guidance needs the doubled batch [uncond | text]")
        if out is None:
            out = torch.empty(B // 2 if guidance is not None else B, cfg.out_channels, 2 * rows, w, device=dev,
                              dtype=hidden_states.dtype)
        if euler is not None:
            latents, dsigma = euler
            R("tail", ops.unpatchify, ws["tok"], out, latents=latents, dsigma=dsigma, guidance=guidance)
        else:
            R("tail", ops.unpatchify, ws["tok"], out, guidance=guidance)
        if self.sp is not None and not sharded:   # every rank returns the whole image (row strips gathered)
            gathered = self.sp.gather_rows(out, dim=2)
            if full_out is not None:
                full_out.copy_(gathered)
                return full_out
            return gathered
        return out

"""TEST INFRASTRUCTURE ONLY -- CPU/plain-torch restatement of the DiT4SR denoiser hot path.

Parity status: the reference module cannot be imported as shipped (it needs `diffusers==0.31.0`,
requirements.txt:3, which is absent and not installable offline).  This restatement is pinned in two ways
(tests/test_oracle_vs_reference.py, tests/golden/make_golden.py):
  * against the reference's OWN first-party code (`/root/reference/model_dit4sr/{attention,transformer_sd3}.py`,
    imported unmodified on top of `oracle/diffusers_shim`, a minimal restatement of the twelve diffusers symbols
    it imports) -- outputs committed under tests/golden/;
  * against the only known-answer fixture the reference ships for this path: the 40 logged flow-match timesteps
    (result_val_dit4sr_3rd/.../final_pred_txt/*.txt).
The diffusers-0.31.0 semantics themselves (AdaLayerNormZero chunk order, RMSNorm, PatchEmbed crop, ...) are
restated from the published algorithm; they are "parity unpinned" w.r.t. a real diffusers install.

Everything here is functional and works on a plain `state_dict` (name -> tensor) whose key names/shapes are
exactly those the reference constructor creates.  The arithmetic runs in the dtype of the
tensors it is given, in the reference's op order, so calling it with bf16 tensors reproduces "the reference in
bf16" and with fp32/fp64 tensors gives the high-precision truth.

References (file:line are relative to /root/reference):
  model_dit4sr/transformer_sd3.py:63-119   constructor / state_dict layout
  model_dit4sr/transformer_sd3.py:267-409  SD3Transformer2DModel.forward
  model_dit4sr/attention.py:102-178        FeedForwardControl
  model_dit4sr/attention.py:186-398        JointTransformerBlock
  model_dit4sr/attention.py:401-436        AttentionZero / zero_module
  model_dit4sr/attention.py:438-559        JointAttnProcessor2_0
  pipelines/pipeline_dit4sr.py:1004-1034,1078-1132,1561-1567  sampling loop
"""
from __future__ import annotations

import math
from collections import OrderedDict
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------------------
# config
# --------------------------------------------------------------------------------------------------------------

# constructor defaults, transformer_sd3.py:64-81
CTOR_DEFAULTS = dict(
    sample_size=128, patch_size=2, in_channels=16, num_layers=18, attention_head_dim=64,
    num_attention_heads=18, joint_attention_dim=4096, caption_projection_dim=1152,
    pooled_projection_dim=2048, out_channels=16, pos_embed_max_size=96,
    dual_attention_layers=(), qk_norm=None,
)


def make_config(**kw) -> SimpleNamespace:
    c = dict(CTOR_DEFAULTS)
    c.update(kw)
    c["dual_attention_layers"] = tuple(c["dual_attention_layers"])
    return SimpleNamespace(**c)


def sd35_medium_config(**kw) -> SimpleNamespace:
    """The SD3.5-medium shape the reference loads (SURVEY.md 8d 'Model config')."""
    base = dict(num_layers=24, num_attention_heads=24, attention_head_dim=64, caption_projection_dim=1536,
                joint_attention_dim=4096, pooled_projection_dim=2048, patch_size=2, in_channels=16,
                out_channels=16, sample_size=128, pos_embed_max_size=384, qk_norm="rms_norm",
                dual_attention_layers=tuple(range(13)))
    base.update(kw)
    return make_config(**base)


def tiny_config(**kw) -> SimpleNamespace:
    """Small config for golden vectors: same structure (dual block, plain block, context_pre_only last block)."""
    base = dict(num_layers=3, num_attention_heads=2, attention_head_dim=64, caption_projection_dim=128,
                joint_attention_dim=96, pooled_projection_dim=80, patch_size=2, in_channels=16, out_channels=16,
                sample_size=32, pos_embed_max_size=24, qk_norm="rms_norm", dual_attention_layers=(0,))
    base.update(kw)
    return make_config(**base)


# --------------------------------------------------------------------------------------------------------------
# state_dict layout (transformer_sd3.py:63-119, attention.py:201-268,401-431)
# --------------------------------------------------------------------------------------------------------------

def state_dict_manifest(cfg, ocr: bool = False) -> "OrderedDict[str, Tuple[int, ...]]":
    """ocr=True: the 3-stream SD3Transformer2DModel_OCRBranch_OCR2HQ2OCR layout (extra `attn.to_*_ocr` and the two
    extra depthwise convs, attention_ocrbranch_ocr2hq2ocr.py:150-158,497-513); `attn2` stays the 2-stream AttentionZero."""
    D = cfg.num_attention_heads * cfg.attention_head_dim
    dh = cfg.attention_head_dim
    Fi = 4 * D
    p = cfg.patch_size
    m: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def lin(name, out_f, in_f):
        m[name + ".weight"] = (out_f, in_f)
        m[name + ".bias"] = (out_f,)

    m["pos_embed.pos_embed"] = (1, cfg.pos_embed_max_size ** 2, D)  # persistent buffer
    m["pos_embed.proj.weight"] = (D, cfg.in_channels, p, p)
    m["pos_embed.proj.bias"] = (D,)
    lin("time_text_embed.timestep_embedder.linear_1", D, 256)
    lin("time_text_embed.timestep_embedder.linear_2", D, D)
    lin("time_text_embed.text_embedder.linear_1", D, cfg.pooled_projection_dim)
    lin("time_text_embed.text_embedder.linear_2", D, D)
    lin("context_embedder", cfg.caption_projection_dim, cfg.joint_attention_dim)

    def attn(prefix, context_pre_only, with_ocr=False):
        # diffusers Attention ctor order, then AttentionZero's extra modules (attention.py:426-429)
        if cfg.qk_norm == "rms_norm":
            m[prefix + ".norm_q.weight"] = (dh,)
            m[prefix + ".norm_k.weight"] = (dh,)
        lin(prefix + ".to_q", D, D)
        lin(prefix + ".to_k", D, D)
        lin(prefix + ".to_v", D, D)
        lin(prefix + ".add_k_proj", D, D)
        lin(prefix + ".add_v_proj", D, D)
        lin(prefix + ".add_q_proj", D, D)
        lin(prefix + ".to_out.0", D, D)
        if not context_pre_only:
            lin(prefix + ".to_add_out", D, D)
        if cfg.qk_norm == "rms_norm":
            m[prefix + ".norm_added_q.weight"] = (dh,)
            m[prefix + ".norm_added_k.weight"] = (dh,)
        lin(prefix + ".to_q_control", D, D)
        lin(prefix + ".to_k_control", D, D)
        lin(prefix + ".to_v_control", D, D)
        lin(prefix + ".to_out_control", D, D)
        if with_ocr:
            lin(prefix + ".to_q_ocr", D, D)
            lin(prefix + ".to_k_ocr", D, D)
            lin(prefix + ".to_v_ocr", D, D)
            lin(prefix + ".to_out_ocr", D, D)

    for i in range(cfg.num_layers):
        b = f"transformer_blocks.{i}"
        last = i == cfg.num_layers - 1
        dual = i in cfg.dual_attention_layers
        lin(b + ".norm1.linear", (9 if dual else 6) * D, D)
        lin(b + ".norm1_context.linear", (2 if last else 6) * D, D)
        attn(b + ".attn", last, with_ocr=ocr)
        if dual:
            attn(b + ".attn2", last)  # Q5: owns unused add_*/to_add_out/norm_added_* parameters
        lin(b + ".ff.net.0.proj", Fi, D)
        lin(b + ".ff.net.2", D, Fi)
        m[b + ".ff.control_conv.weight"] = (Fi, 1, 3, 3)
        m[b + ".ff.control_conv.bias"] = (Fi,)
        if ocr:
            for cname in ("ocr_to_hidden_conv", "hidden_to_ocr_conv"):
                m[b + f".ff.{cname}.weight"] = (Fi, 1, 3, 3)
                m[b + f".ff.{cname}.bias"] = (Fi,)
        if not last:
            lin(b + ".ff_context.net.0.proj", Fi, D)
            lin(b + ".ff_context.net.2", D, Fi)
    lin("norm_out.linear", 2 * D, D)
    lin("proj_out", p * p * cfg.out_channels, D)
    return m


def sincos_pos_embed_2d(embed_dim: int, grid_size: int, base_size: int, interpolation_scale: float = 1.0) -> torch.Tensor:
    """diffusers get_2d_sincos_pos_embed (w-first meshgrid, [emb_h|emb_w], [sin|cos], float64 omega)."""
    gh = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size) / interpolation_scale
    gw = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size) / interpolation_scale
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid_size, grid_size])

    def one_d(dim, pos):
        omega = np.arange(dim // 2, dtype=np.float64)
        omega /= dim / 2.0
        omega = 1.0 / 10000 ** omega
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    emb = np.concatenate([one_d(embed_dim // 2, grid[0]), one_d(embed_dim // 2, grid[1])], axis=1)
    return torch.from_numpy(emb).float().unsqueeze(0)


def init_state_dict(cfg, seed: int = 0, dtype=torch.float32, ocr: bool = False) -> "OrderedDict[str, torch.Tensor]":
    """Random-init EVERY parameter, including the zero-initialised *_control / control_conv ones (Q4).

    Scales follow SURVEY.md 8d: Linear weights/biases N(0,0.02^2), conv weights N(0,0.1^2), RMSNorm weights
    1+0.1*N(0,1), pos_embed = the sincos formula.
    """
    g = torch.Generator().manual_seed(seed)
    D = cfg.num_attention_heads * cfg.attention_head_dim
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in state_dict_manifest(cfg, ocr=ocr).items():
        if name == "pos_embed.pos_embed":
            t = sincos_pos_embed_2d(D, cfg.pos_embed_max_size, cfg.sample_size // cfg.patch_size)
        elif ".norm_" in name and name.endswith(".weight") and len(shape) == 1 and shape[0] == cfg.attention_head_dim:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif "control_conv.weight" in name or "_conv.weight" in name:
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        sd[name] = t.to(dtype)
    return sd


def make_inputs(cfg, batch: int, height: int, width: int, text_tokens: int, seed: int = 1, dtype=torch.float32,
                timestep: float = 1000.0):
    g = torch.Generator().manual_seed(seed)
    hs = torch.randn(batch, cfg.in_channels, height, width, generator=g)
    lr = torch.randn(batch, cfg.in_channels, height, width, generator=g)
    enc = 0.5 * torch.randn(batch, text_tokens, cfg.joint_attention_dim, generator=g)
    pooled = torch.randn(batch, cfg.pooled_projection_dim, generator=g)
    t = torch.full((batch,), float(timestep))
    return hs.to(dtype), lr.to(dtype), enc.to(dtype), pooled.to(dtype), t


# --------------------------------------------------------------------------------------------------------------
# diffusers-0.31.0 building blocks, restated
# --------------------------------------------------------------------------------------------------------------

def _linear(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _layer_norm(x, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), None, None, eps)


def rms_norm(x, weight, eps=1e-6):
    """diffusers RMSNorm.forward: fp32 variance, cast to weight dtype if half, multiply by weight."""
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True) if x.dtype != torch.float64 else x.pow(2).mean(-1, keepdim=True)
    x = x * torch.rsqrt(var + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        x = x.to(weight.dtype)
    return x * weight


def timestep_sincos(t: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0): [cos|sin], fp32."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def time_text_embed(sd, timestep, pooled):
    """CombinedTimestepTextProjEmbeddings (transformer_sd3.py:95-97,327)."""
    p = "time_text_embed."
    tp = timestep_sincos(timestep.to(pooled.device)).to(pooled.dtype)
    te = _linear(sd, p + "timestep_embedder.linear_2", F.silu(_linear(sd, p + "timestep_embedder.linear_1", tp)))
    pe = _linear(sd, p + "text_embedder.linear_2", F.silu(_linear(sd, p + "text_embedder.linear_1", pooled)))
    return te + pe


def patch_embed(sd, cfg, latent):
    """PatchEmbed with pos_embed_max_size (transformer_sd3.py:87-94,326,331): conv k=s=patch, centre-cropped pos-embed."""
    h, w = latent.shape[-2:]
    x = F.conv2d(latent, sd["pos_embed.proj.weight"], sd["pos_embed.proj.bias"], stride=cfg.patch_size)
    x = x.flatten(2).transpose(1, 2)
    hp, wp = h // cfg.patch_size, w // cfg.patch_size
    mx = cfg.pos_embed_max_size
    top, left = (mx - hp) // 2, (mx - wp) // 2
    pe = sd["pos_embed.pos_embed"].reshape(1, mx, mx, -1)[:, top:top + hp, left:left + wp, :]
    pe = pe.reshape(1, -1, pe.shape[-1])
    return (x + pe).to(x.dtype)


# --------------------------------------------------------------------------------------------------------------
# the block (attention.py)
# --------------------------------------------------------------------------------------------------------------

def joint_attention(sd, cfg, prefix, hidden_states, encoder_hidden_states=None, context_pre_only=False,
                    as_written=True):
    """JointAttnProcessor2_0.__call__ (attention.py:445-559), including quirks Q1-Q3.

    as_written=True follows the file literally (computes the dead to_k_control chain when qk_norm is set).
    as_written=False elides it (the dataflow the CUDA path implements); results are identical.
    """
    H = cfg.num_attention_heads
    B = hidden_states.shape[0]
    residual = hidden_states
    hq, lr = hidden_states.chunk(2, dim=1)                                   # :459
    lr_res = lr                                                              # :460

    def heads(t):
        return t.view(B, -1, H, t.shape[-1] // H).transpose(1, 2)

    q = heads(_linear(sd, prefix + ".to_q", hq))                            # :463-472
    k = heads(_linear(sd, prefix + ".to_k", hq))
    v = heads(_linear(sd, prefix + ".to_v", hq))
    qc = heads(_linear(sd, prefix + ".to_q_control", _linear(sd, prefix + ".to_q", lr)))   # :475 (Q3)
    vc = heads(_linear(sd, prefix + ".to_v_control", _linear(sd, prefix + ".to_v", lr)))   # :477
    has_norm = cfg.qk_norm == "rms_norm"
    if as_written or not has_norm:
        kc = heads(_linear(sd, prefix + ".to_k_control", _linear(sd, prefix + ".to_k", lr)))  # :476
    if has_norm:
        q = rms_norm(q, sd[prefix + ".norm_q.weight"])                       # :485
        qc = rms_norm(qc, sd[prefix + ".norm_q.weight"])                     # :486
        k = rms_norm(k, sd[prefix + ".norm_k.weight"])                       # :488
        kc = rms_norm(k, sd[prefix + ".norm_k.weight"])                      # :489  Q1: norm of the normed HQ key
    if encoder_hidden_states is not None:
        tq = heads(_linear(sd, prefix + ".add_q_proj", encoder_hidden_states))      # :494-500
        tk = heads(_linear(sd, prefix + ".add_k_proj", encoder_hidden_states))
        tv = heads(_linear(sd, prefix + ".add_v_proj", encoder_hidden_states))
        if has_norm:
            tq = rms_norm(tq, sd[prefix + ".norm_added_q.weight"])                   # :503
            tk = rms_norm(tk, sd[prefix + ".norm_added_k.weight"])                   # :505
        Q = torch.cat([q, qc, tq], dim=2)                                            # :508-510
        K = torch.cat([k, kc, tk], dim=2)
        V = torch.cat([v, vc, tv], dim=2)
    else:
        Q = torch.cat([q, qc], dim=2)                                                # :520-522
        K = torch.cat([k, kc], dim=2)
        V = torch.cat([v, vc], dim=2)
    o = F.scaled_dot_product_attention(Q, K, V, dropout_p=0.0, is_causal=False)     # :528
    o = o.transpose(1, 2).reshape(B, -1, Q.shape[1] * Q.shape[-1]).to(Q.dtype)      # :529-530
    ctx_out = None
    if encoder_hidden_states is not None:
        o, ctx_out = o[:, : residual.shape[1]], o[:, residual.shape[1]:]            # :534-537
        if not context_pre_only:
            ctx_out = _linear(sd, prefix + ".to_add_out", ctx_out)                   # :539
    o_hq, o_lr = o.chunk(2, dim=1)                                                   # :543
    o_lr = o_lr + lr_res                                                             # :545 (Q2)
    o_hq = _linear(sd, prefix + ".to_out.0", o_hq)                                   # :548
    o_lr = _linear(sd, prefix + ".to_out_control", o_lr)                             # :552
    out = torch.cat([o_hq, o_lr], dim=1)                                             # :554
    if encoder_hidden_states is not None:
        return out, ctx_out
    return out


def feed_forward_control(sd, prefix, x):
    """FeedForwardControl.forward (attention.py:155-178): GELU-tanh MLP with the depthwise LR->HQ conv."""
    u = F.gelu(_linear(sd, prefix + ".net.0.proj", x), approximate="tanh")
    u_hq, u_lr = u.chunk(2, dim=1)                                                   # :168
    B, N, C = u_hq.shape
    h = w = int(np.sqrt(N))
    assert h * w == N                                                                # :172 (Q9)
    c = u_lr.reshape(B, h, w, C).permute(0, 3, 1, 2)                                 # :173
    c = F.conv2d(c, sd[prefix + ".control_conv.weight"], sd[prefix + ".control_conv.bias"], padding=1, groups=C)
    c = c.reshape(B, C, N).permute(0, 2, 1)                                          # :175
    u = torch.cat([u_hq + 1.0 * c, u_lr], dim=1)                                     # :176-177
    return _linear(sd, prefix + ".net.2", u)


def joint_attention_ocr(sd, cfg, prefix, hidden_states, encoder_hidden_states=None, context_pre_only=False):
    """JointAttnProcessor2_0_OCRBranch_OCR2HQ2OCR.__call__ (attention_ocrbranch_ocr2hq2ocr.py:520-664): three image
    streams [HQ | LR | OCR] + text. key_control keeps quirk Q1 (:588); key_ocr = norm_k(to_k_ocr(to_k(ocr))) is live (:589)."""
    H = cfg.num_attention_heads
    B = hidden_states.shape[0]
    residual = hidden_states
    hq, lr, oc = hidden_states.chunk(3, dim=1)                                         # :542

    def heads(t):
        return t.view(B, -1, H, t.shape[-1] // H).transpose(1, 2)

    def chain(x, a, b_):
        return heads(_linear(sd, prefix + b_, _linear(sd, prefix + a, x)))

    q, k, v = (heads(_linear(sd, prefix + n, hq)) for n in (".to_q", ".to_k", ".to_v"))   # :550-559
    qc, kc, vc = chain(lr, ".to_q", ".to_q_control"), chain(lr, ".to_k", ".to_k_control"), chain(lr, ".to_v", ".to_v_control")
    qo, ko, vo = chain(oc, ".to_q", ".to_q_ocr"), chain(oc, ".to_k", ".to_k_ocr"), chain(oc, ".to_v", ".to_v_ocr")
    if cfg.qk_norm == "rms_norm":
        wq, wk = sd[prefix + ".norm_q.weight"], sd[prefix + ".norm_k.weight"]
        q, qc, qo = rms_norm(q, wq), rms_norm(qc, wq), rms_norm(qo, wq)                 # :582-585
        k = rms_norm(k, wk)                                                              # :587
        kc = rms_norm(k, wk)                                                             # :588 (Q1)
        ko = rms_norm(ko, wk)                                                            # :589
    if encoder_hidden_states is not None:
        tq = heads(_linear(sd, prefix + ".add_q_proj", encoder_hidden_states))
        tk = heads(_linear(sd, prefix + ".add_k_proj", encoder_hidden_states))
        tv = heads(_linear(sd, prefix + ".add_v_proj", encoder_hidden_states))
        if cfg.qk_norm == "rms_norm":
            tq = rms_norm(tq, sd[prefix + ".norm_added_q.weight"])
            tk = rms_norm(tk, sd[prefix + ".norm_added_k.weight"])
        Q, K, V = torch.cat([q, qc, qo, tq], 2), torch.cat([k, kc, ko, tk], 2), torch.cat([v, vc, vo, tv], 2)   # :612-614
    else:
        Q, K, V = torch.cat([q, qc, qo], 2), torch.cat([k, kc, ko], 2), torch.cat([v, vc, vo], 2)
    o = F.scaled_dot_product_attention(Q, K, V, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(B, -1, Q.shape[1] * Q.shape[-1]).to(Q.dtype)
    ctx_out = None
    if encoder_hidden_states is not None:
        o, ctx_out = o[:, : residual.shape[1]], o[:, residual.shape[1]:]
        if not context_pre_only:
            ctx_out = _linear(sd, prefix + ".to_add_out", ctx_out)
    o_hq, o_lr, o_oc = o.chunk(3, dim=1)                                                # :642
    o_lr = o_lr + lr                                                                    # :644
    o_oc = o_oc + oc                                                                    # :645
    out = torch.cat([_linear(sd, prefix + ".to_out.0", o_hq), _linear(sd, prefix + ".to_out_control", o_lr),
                     _linear(sd, prefix + ".to_out_ocr", o_oc)], dim=1)                 # :648-656
    if encoder_hidden_states is not None:
        return out, ctx_out
    return out


def feed_forward_ocr2hq2ocr(sd, prefix, x):
    """FeedForwardControl_OCR2HQ2OCR.forward (attention_ocrbranch_ocr2hq2ocr.py:164-211): LR->HQ and OCR->HQ convs,
    then HQ(updated)->OCR conv."""
    u = F.gelu(_linear(sd, prefix + ".net.0.proj", x), approximate="tanh")
    u_hq, u_lr, u_oc = u.chunk(3, dim=1)                                                # :177
    B, N, C = u_hq.shape
    h = w = int(np.sqrt(N))
    assert h * w == N

    def conv(t, name):
        t = t.reshape(B, h, w, C).permute(0, 3, 1, 2)
        t = F.conv2d(t, sd[prefix + f".{name}.weight"], sd[prefix + f".{name}.bias"], padding=1, groups=C)
        return t.reshape(B, C, N).permute(0, 2, 1)

    u_hq = u_hq + 1.0 * conv(u_lr, "control_conv") + 1.0 * conv(u_oc, "ocr_to_hidden_conv")   # :184-196
    u_oc = u_oc + 1.0 * conv(u_hq, "hidden_to_ocr_conv")                                       # :200-206
    return _linear(sd, prefix + ".net.2", torch.cat([u_hq, u_lr, u_oc], dim=1))                # :210


def feed_forward(sd, prefix, x):
    """FeedForward.forward (attention.py:94-100), gelu-approximate."""
    return _linear(sd, prefix + ".net.2", F.gelu(_linear(sd, prefix + ".net.0.proj", x), approximate="tanh"))


def block_forward(sd, cfg, i, hidden_states, encoder_hidden_states, temb, as_written=True, ocr=False):
    """JointTransformerBlock.forward (attention.py:280-398). Returns (encoder_hidden_states, hidden_states).
    ocr=True: the 3-stream block (attention_ocrbranch_ocr2hq2ocr.py:317-433): `attn` is the 3-stream processor, `attn2`
    stays the 2-stream processor and therefore splits the 3N tokens at 1.5N (quirk Q6, :354,:686), `ff` has 3 convs."""
    b = f"transformer_blocks.{i}"
    dual = i in cfg.dual_attention_layers
    last = i == cfg.num_layers - 1
    emb = _linear(sd, b + ".norm1.linear", F.silu(temb))
    ln = _layer_norm(hidden_states)
    if dual:                                                                          # SD35AdaLayerNormZeroX, :286-289
        (shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp,
         shift_msa2, scale_msa2, gate_msa2) = emb.chunk(9, dim=1)
        n1 = ln * (1 + scale_msa[:, None]) + shift_msa[:, None]
        n2 = ln * (1 + scale_msa2[:, None]) + shift_msa2[:, None]
    else:                                                                             # AdaLayerNormZero, :291-293
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = emb.chunk(6, dim=1)
        n1 = ln * (1 + scale_msa[:, None]) + shift_msa[:, None]
    if last:                                                                          # AdaLayerNormContinuous, :296-298 (Q12)
        cemb = _linear(sd, b + ".norm1_context.linear", F.silu(temb).to(encoder_hidden_states.dtype))
        c_scale, c_shift = cemb.chunk(2, dim=1)
        nc = _layer_norm(encoder_hidden_states) * (1 + c_scale)[:, None, :] + c_shift[:, None, :]
    else:                                                                             # :300-302
        cemb = _linear(sd, b + ".norm1_context.linear", F.silu(temb))
        c_shift_msa, c_scale_msa, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = cemb.chunk(6, dim=1)
        nc = _layer_norm(encoder_hidden_states) * (1 + c_scale_msa[:, None]) + c_shift_msa[:, None]

    if ocr:
        attn_out, ctx_attn = joint_attention_ocr(sd, cfg, b + ".attn", n1, nc, context_pre_only=last)
    else:
        attn_out, ctx_attn = joint_attention(sd, cfg, b + ".attn", n1, nc, context_pre_only=last, as_written=as_written)
    hidden_states = hidden_states + gate_msa.unsqueeze(1) * attn_out                  # :314-315
    if dual:
        attn_out2 = joint_attention(sd, cfg, b + ".attn2", n2, None, as_written=as_written)   # :318
        hidden_states = hidden_states + gate_msa2.unsqueeze(1) * attn_out2           # :319-320
    m = _layer_norm(hidden_states) * (1 + scale_mlp[:, None]) + shift_mlp[:, None]   # :322-323
    ff = feed_forward_ocr2hq2ocr if ocr else feed_forward_control
    hidden_states = hidden_states + gate_mlp.unsqueeze(1) * ff(sd, b + ".ff", m)      # :328-347
    if last:
        encoder_hidden_states = None                                                  # :379
    else:
        encoder_hidden_states = encoder_hidden_states + c_gate_msa.unsqueeze(1) * ctx_attn      # :381-382
        mc = _layer_norm(encoder_hidden_states) * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None]
        encoder_hidden_states = encoder_hidden_states + c_gate_mlp.unsqueeze(1) * feed_forward(sd, b + ".ff_context", mc)
    return encoder_hidden_states, hidden_states


# --------------------------------------------------------------------------------------------------------------
# the model (transformer_sd3.py:267-409)
# --------------------------------------------------------------------------------------------------------------

def unpatchify(x, cfg, height, width):
    """einops 'b (H W) (pH pW d) -> b d (H pH) (W pW)' (transformer_sd3.py:391)."""
    p = cfg.patch_size
    B = x.shape[0]
    Hh, Ww = height // p, width // p
    x = x.reshape(B, Hh, Ww, p, p, cfg.out_channels)
    return x.permute(0, 5, 1, 3, 2, 4).reshape(B, cfg.out_channels, Hh * p, Ww * p)


def model_forward(sd, cfg, hidden_states, controlnet_image, encoder_hidden_states, pooled_projections, timestep,
                  as_written=True, feat_extract_layers=(), feat_extract=None, return_block_states=False, ocr=False):
    """SD3Transformer2DModel.forward. Returns velocity [B,C,h,w] (and optionally extras).
    ocr=True: SD3Transformer2DModel_OCRBranch_OCR2HQ2OCR.forward (transformer_sd3_ocrbranch_ocr2hq2ocr.py:269-410):
    the third stream is a second pos_embed(controlnet_image) (:334-337); chunk(3) at the end (:384)."""
    height, width = hidden_states.shape[-2:]
    x = patch_embed(sd, cfg, hidden_states)                                           # :326
    temb = time_text_embed(sd, timestep, pooled_projections)                          # :327
    ctx = _linear(sd, "context_embedder", encoder_hidden_states)                     # :328
    lr = patch_embed(sd, cfg, controlnet_image)                                       # :331
    x = torch.cat([x, lr, patch_embed(sd, cfg, controlnet_image)] if ocr else [x, lr], dim=-2)   # :332 / ocr :334-337
    feats: List[torch.Tensor] = []
    states = []
    for i in range(cfg.num_layers):                                                   # :336-370
        ctx, x = block_forward(sd, cfg, i, x, ctx, temb, as_written=as_written, ocr=ocr)
        if i in feat_extract_layers:                                                  # attention.py:353-374 (Q10)
            if feat_extract == "hq_feat":
                feats.append(x[:, 0:1024])
            elif feat_extract == "lq_feat":
                feats.append(x[:, 1024:2048])
            elif feat_extract == "hqlq_feat":
                feats.append(x[:, 0:2048])
            elif feat_extract == "ocr_feat" and ocr:
                feats.append(x[:, 2048:3072])
        if return_block_states:
            states.append((x.clone(), None if ctx is None else ctx.clone()))
    x = x.chunk(3 if ocr else 2, dim=1)[0]                                            # :379 / ocr :384
    emb = _linear(sd, "norm_out.linear", F.silu(temb).to(x.dtype))                    # :382 AdaLayerNormContinuous
    scale, shift = emb.chunk(2, dim=1)                                                # (scale, shift) order, Q12
    x = _layer_norm(x) * (1 + scale)[:, None, :] + shift[:, None, :]
    x = _linear(sd, "proj_out", x)                                                    # :383
    out = unpatchify(x, cfg, height, width)                                           # :386-391
    if return_block_states:
        return out, states
    if feat_extract_layers:
        return out, feats
    return out


# --------------------------------------------------------------------------------------------------------------
# sampling loop (pipelines/pipeline_dit4sr.py) -- FlowMatchEulerDiscreteScheduler(shift=3.0), guidance 1.0 (Q11)
# --------------------------------------------------------------------------------------------------------------

def flow_match_sigmas(num_steps: int, shift: float = 3.0, num_train_timesteps: int = 1000) -> np.ndarray:
    """set_timesteps(n): sigmas (n+1 values, last = 0). timesteps = 1000 * sigmas[:-1]."""
    t = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
    s = t / num_train_timesteps
    s = shift * s / (1 + (shift - 1) * s)
    sigma_max, sigma_min = float(s[0]), float(s[-1])
    ts = np.linspace(sigma_max * num_train_timesteps, sigma_min * num_train_timesteps, num_steps)
    sig = ts / num_train_timesteps
    sig = shift * sig / (1 + (shift - 1) * sig)
    return np.concatenate([sig, np.zeros(1)]).astype(np.float32)


def sample_loop(sd, cfg, latents, controlnet_image, encoder_hidden_states, pooled_projections, num_steps=40,
                model_fn=None):
    """pipeline_dit4sr.py:1078-1132,1567 with guidance_scale=1, start point 'noise', untiled branch.

    scheduler.step: x_prev = (x.float() + (sigma_next - sigma) * v.float()).to(v.dtype)
    """
    sig = flow_match_sigmas(num_steps)
    B = latents.shape[0]
    if model_fn is None:
        def model_fn(x, t):
            return model_forward(sd, cfg, x, controlnet_image, encoder_hidden_states, pooled_projections, t,
                                 as_written=False)
    x = latents
    for i in range(num_steps):
        t = torch.full((B,), float(sig[i]) * 1000.0)
        v = model_fn(x, t)
        x = (x.float() + (float(sig[i + 1]) - float(sig[i])) * v.float()).to(v.dtype)
    return x


def lre_start_point(latents, control_image, sigma0):
    """start_point == 'lr' (LRE strategy), pipelines/pipeline_dit4sr.py:1025-1034, as written: every image of the batch
    is blended with the FIRST control latent (`control_image[:1]`):  (1 - sigma_0) * control[:1] + sigma_0 * noise."""
    cond = control_image[:1]
    sigma = torch.as_tensor(sigma0, dtype=latents.dtype, device=latents.device).flatten()
    while sigma.dim() < 4:
        sigma = sigma.unsqueeze(-1)
    return (1.0 - sigma) * cond + sigma * latents


def sample_loop_cfg(sd, cfg, latents, controlnet_image, encoder_hidden_states, pooled_projections,
                    negative_prompt_embeds, negative_pooled_prompt_embeds, guidance_scale, num_steps=40,
                    start_point="noise", model_fn=None):
    """The loop with classifier-free guidance as pipelines/pipeline_dit4sr.py runs it (9 of the 36 shipped configs use
    guidance 8.0): doubled latents (:1085), control image doubled by prepare_image (:714-715), prompt embeddings
    [negative | positive] (:1106-1108), one model call on the 2B batch (:1122-1131), combine (:1561-1563)
    noise_pred = uncond + guidance_scale * (text - uncond) in the model dtype, scheduler.step in fp32 (:1567)."""
    sig = flow_match_sigmas(num_steps)
    B = latents.shape[0]
    ctrl2 = torch.cat([controlnet_image] * 2) if controlnet_image.shape[0] == B else controlnet_image
    enc2 = torch.cat([negative_prompt_embeds, encoder_hidden_states], dim=0)
    pooled2 = torch.cat([negative_pooled_prompt_embeds, pooled_projections], dim=0)
    if model_fn is None:
        def model_fn(x, c, e, p, t):
            return model_forward(sd, cfg, x, c, e, p, t, as_written=False)
    x = latents
    if start_point == "lr":
        x = lre_start_point(x, controlnet_image, float(sig[0]))
    for i in range(num_steps):
        t = torch.full((2 * B,), float(sig[i]) * 1000.0, device=x.device)
        v = model_fn(torch.cat([x] * 2), ctrl2, enc2, pooled2, t)
        vu, vt = v.chunk(2)
        v = vu + guidance_scale * (vt - vu)
        x = (x.float() + (float(sig[i + 1]) - float(sig[i])) * v.float()).to(v.dtype)
    return x


# --------------------------------------------------------------------------------------------------------------
# tiled branch of the sampling loop (pipelines/pipeline_dit4sr.py:746-761,1449-1558), restated literally
# --------------------------------------------------------------------------------------------------------------

def tiled_velocity_as_written(model_fn, latent_model_input, control_image, tile_size, tile_overlap):
    """One tile per model call, the reference's loops and offsets verbatim (batch_size 1)."""
    from numpy import pi, exp, sqrt
    h, w = latent_model_input.shape[-2:]
    tile_size = min(tile_size, min(h, w))
    var = 0.01
    midpoint = (tile_size - 1) / 2
    x_probs = [exp(-(x - midpoint) * (x - midpoint) / (tile_size * tile_size) / (2 * var)) / sqrt(2 * pi * var) for x in range(tile_size)]
    midpoint = tile_size / 2
    y_probs = [exp(-(y - midpoint) * (y - midpoint) / (tile_size * tile_size) / (2 * var)) / sqrt(2 * pi * var) for y in range(tile_size)]
    tile_weights = torch.tensor(np.outer(y_probs, x_probs)).to(torch.float32)              # :746-761
    grid_rows, cur_x = 0, 0
    while cur_x < latent_model_input.size(-1):                                             # :1454-1458
        cur_x = max(grid_rows * tile_size - tile_overlap * grid_rows, 0) + tile_size
        grid_rows += 1
    grid_cols, cur_y = 0, 0
    while cur_y < latent_model_input.size(-2):                                             # :1460-1464
        cur_y = max(grid_cols * tile_size - tile_overlap * grid_cols, 0) + tile_size
        grid_cols += 1

    def offsets(row, col):                                                                 # :1474-1481
        ofs_x = max(row * tile_size - tile_overlap * row, 0)
        ofs_y = max(col * tile_size - tile_overlap * col, 0)
        if row == grid_rows - 1:
            ofs_x = w - tile_size
        if col == grid_cols - 1:
            ofs_y = h - tile_size
        return ofs_x, ofs_y

    noise_preds = []
    for row in range(grid_rows):
        for col in range(grid_cols):
            ox, oy = offsets(row, col)
            noise_preds.append(model_fn(latent_model_input[:, :, oy:oy + tile_size, ox:ox + tile_size],
                                        control_image[:, :, oy:oy + tile_size, ox:ox + tile_size]).float())
    noise_pred = torch.zeros(latent_model_input.shape)                                     # :1535-1557
    contributors = torch.zeros(latent_model_input.shape)
    for row in range(grid_rows):
        for col in range(grid_cols):
            ox, oy = offsets(row, col)
            noise_pred[:, :, oy:oy + tile_size, ox:ox + tile_size] += noise_preds[row * grid_cols + col] * tile_weights
            contributors[:, :, oy:oy + tile_size, ox:ox + tile_size] += tile_weights
    return noise_pred / contributors

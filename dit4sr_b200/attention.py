"""Parameter containers mirroring `model_dit4sr/attention.py` of the reference (same attribute names, hence the same
state_dict keys and the same names for `finetune_layer_names` matching, initialize.py:498-515).

The reference builds these on diffusers' `Attention`, `AdaLayerNormZero`, ... classes; here they are plain
`nn.Module` holders: the arithmetic lives in the CUDA engine (`engine.py`), not in `forward` methods.
  JointTransformerBlock   attention.py:186-398
  AttentionZero           attention.py:401-431 (to_{q,k,v,out}_control on top of diffusers Attention's projections)
  FeedForwardControl      attention.py:102-178 (GELU-tanh MLP + depthwise `control_conv`)
  FeedForward             attention.py:42-100
"""
from typing import Optional

import torch
from torch import nn


class RMSNormWeight(nn.Module):
    """Holder for diffusers RMSNorm's `weight` (dim = head_dim, eps 1e-6)."""

    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))


class AdaNormLinear(nn.Module):
    """Holder for the `linear` of AdaLayerNormZero (6D) / SD35AdaLayerNormZeroX (9D) / AdaLayerNormContinuous (2D)."""

    def __init__(self, dim: int, chunks: int):
        super().__init__()
        self.chunks = chunks
        self.linear = nn.Linear(dim, chunks * dim, bias=True)


class AttentionZero(nn.Module):
    """`ocr_branch=True` gives AttentionZero_OCR2HQ2OCR (attention_ocrbranch_ocr2hq2ocr.py:471-514): the same module
    plus `to_{q,k,v,out}_ocr` for the third (OCR) stream."""

    def __init__(self, dim: int, heads: int, dim_head: int, context_pre_only: Optional[bool], qk_norm: Optional[str],
                 ocr_branch: bool = False):
        super().__init__()
        self.ocr_branch = ocr_branch
        self.heads = heads
        self.inner_dim = dim
        self.context_pre_only = context_pre_only
        if qk_norm == "rms_norm":
            self.norm_q = RMSNormWeight(dim_head)
            self.norm_k = RMSNormWeight(dim_head)
        elif qk_norm is None:
            self.norm_q = None
            self.norm_k = None
        else:
            raise ValueError(f"unknown qk_norm: {qk_norm}. Should be None or 'rms_norm'")
        self.to_q = nn.Linear(dim, dim, bias=True)
        self.to_k = nn.Linear(dim, dim, bias=True)
        self.to_v = nn.Linear(dim, dim, bias=True)
        self.add_k_proj = nn.Linear(dim, dim, bias=True)
        self.add_v_proj = nn.Linear(dim, dim, bias=True)
        self.add_q_proj = nn.Linear(dim, dim, bias=True)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim, bias=True), nn.Dropout(0.0)])
        if not context_pre_only:
            self.to_add_out = nn.Linear(dim, dim, bias=True)
        if qk_norm == "rms_norm":
            self.norm_added_q = RMSNormWeight(dim_head)
            self.norm_added_k = RMSNormWeight(dim_head)
        else:
            self.norm_added_q = None
            self.norm_added_k = None
        # attention.py:426-431: zero-initialised control projections; to_out_control starts as a copy of to_out[0]
        self.to_q_control = zero_module(nn.Linear(dim, dim, bias=True))
        self.to_k_control = zero_module(nn.Linear(dim, dim, bias=True))
        self.to_v_control = zero_module(nn.Linear(dim, dim, bias=True))
        self.to_out_control = nn.Linear(dim, dim, bias=True)
        self.to_out_control.weight.data.copy_(self.to_out[0].weight.data)
        self.to_out_control.bias.data.copy_(self.to_out[0].bias.data)
        if ocr_branch:
            self.to_q_ocr = zero_module(nn.Linear(dim, dim, bias=True))
            self.to_k_ocr = zero_module(nn.Linear(dim, dim, bias=True))
            self.to_v_ocr = zero_module(nn.Linear(dim, dim, bias=True))
            self.to_out_ocr = nn.Linear(dim, dim, bias=True)
            self.to_out_ocr.weight.data.copy_(self.to_out[0].weight.data)
            self.to_out_ocr.bias.data.copy_(self.to_out[0].bias.data)


def zero_module(module):
    for p in module.parameters():
        nn.init.zeros_(p)
    return module


class _GELUProj(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=True)


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        inner = dim * mult
        self.net = nn.ModuleList([_GELUProj(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim, bias=True)])


class FeedForwardControl(nn.Module):
    """`ocr_branch=True` gives FeedForwardControl_OCR2HQ2OCR (attention_ocrbranch_ocr2hq2ocr.py:102-211)."""

    def __init__(self, dim: int, mult: int = 4, ocr_branch: bool = False):
        super().__init__()
        inner = dim * mult
        self.net = nn.ModuleList([_GELUProj(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim, bias=True)])
        self.control_conv = zero_module(nn.Conv2d(inner, inner, 3, stride=1, padding=1, groups=inner))
        if ocr_branch:
            self.ocr_to_hidden_conv = zero_module(nn.Conv2d(inner, inner, 3, stride=1, padding=1, groups=inner))
            self.hidden_to_ocr_conv = zero_module(nn.Conv2d(inner, inner, 3, stride=1, padding=1, groups=inner))


class JointTransformerBlock(nn.Module):
    def __init__(self, dim, num_attention_heads, attention_head_dim, context_pre_only=False,
                 qk_norm: Optional[str] = None, use_dual_attention: bool = False, ocr_branch: bool = False):
        super().__init__()
        self.use_dual_attention = use_dual_attention
        self.context_pre_only = context_pre_only
        self.ocr_branch = ocr_branch
        self.norm1 = AdaNormLinear(dim, 9 if use_dual_attention else 6)
        self.norm1_context = AdaNormLinear(dim, 2 if context_pre_only else 6)
        self.attn = AttentionZero(dim, num_attention_heads, attention_head_dim, context_pre_only, qk_norm,
                                  ocr_branch=ocr_branch)
        # the OCR model keeps the plain 2-stream AttentionZero for attn2 (attention_ocrbranch_ocr2hq2ocr.py:279-293, Q6)
        self.attn2 = (AttentionZero(dim, num_attention_heads, attention_head_dim, context_pre_only, qk_norm)
                      if use_dual_attention else None)
        self.ff = FeedForwardControl(dim, ocr_branch=ocr_branch)
        self.ff_context = None if context_pre_only else FeedForward(dim)
        self._chunk_size = None
        self._chunk_dim = 0

    def set_chunk_feed_forward(self, chunk_size, dim: int = 0):
        self._chunk_size = chunk_size
        self._chunk_dim = dim

    def forward(self, *args, **kwargs):
        raise RuntimeError("dit4sr_b200 blocks are parameter holders; call SD3Transformer2DModel.forward "
                           "(the fused CUDA engine runs the whole block sequence)")

import torch
import torch.nn.functional as F
from torch import nn


class GELU(nn.Module):
    """Linear followed by GELU (optionally tanh-approximate)."""

    def __init__(self, dim_in, dim_out, approximate="none", bias=True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate

    def forward(self, hidden_states):
        return F.gelu(self.proj(hidden_states), approximate=self.approximate)


class _Unsupported(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("not on the DiT4SR hot path")


GEGLU = ApproximateGELU = SwiGLU = _Unsupported


class FP32SiLU(nn.Module):
    def forward(self, x):
        return F.silu(x.float(), inplace=False).to(x.dtype)

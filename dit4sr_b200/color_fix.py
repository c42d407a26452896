"""Colour fix of the SR result against the LQ input -- B200 mirror of the reference's
`utils/wavelet_color_fix.py`: same function names and argument meaning; tensors live on the GPU and every pixel
operation runs in the CUDA library (`csrc/colorfix.cu`, through the C ABI).  No CPU fallback.

Reference call site: `test/test.py:381-383` applies `wavelet_color_fix` / `adain_color_fix` to the decoded PIL image."""
import ctypes as C

import torch

from . import _lib
from .ops import _count, _ptr, _req, _stream, check


def _planes(t, name):
    if t.dim() != 4:
        raise _lib.Dit4srError(f"{name}: expected a 4-D [B, C, H, W] tensor")
    _req(t, torch.float32, name)
    if not t.is_contiguous():
        raise _lib.Dit4srError(f"{name}: tensor must be contiguous")
    return t.shape[0] * t.shape[1], t.shape[2], t.shape[3]


def wavelet_blur(image: torch.Tensor, radius: int, high_accum: torch.Tensor = None) -> torch.Tensor:
    """utils/wavelet_color_fix.py:76-94 (3x3 [1 2 1]x[1 2 1]/16 kernel, dilation = radius, replicate padding); with
    `high_accum` also adds `image - blur` into it (one level of `wavelet_decomposition`)."""
    planes, H, W = _planes(image, "image")
    low = torch.empty_like(image)
    if high_accum is not None:
        _planes(high_accum, "high_accum")
    check(_lib.load().dit4sr_wavelet_blur_f32(_ptr(image), _ptr(low), _ptr(high_accum), planes, H, W, int(radius),
                                              _stream()), "dit4sr_wavelet_blur_f32")
    _count()
    return low


def wavelet_decomposition(image: torch.Tensor, levels: int = 5):
    """utils/wavelet_color_fix.py:96-110: returns (high_freq, low_freq)."""
    high = torch.zeros_like(image)
    low = image
    for i in range(levels):
        low = wavelet_blur(image, 2 ** i, high_accum=high)
        image = low
    return high, low


def wavelet_reconstruction(content_feat: torch.Tensor, style_feat: torch.Tensor) -> torch.Tensor:
    """utils/wavelet_color_fix.py:112-124: high frequencies of the content + low frequencies of the style."""
    if content_feat.shape != style_feat.shape:
        raise _lib.Dit4srError("wavelet_reconstruction: content and style must have the same shape")
    content_high, _ = wavelet_decomposition(content_feat)
    _, style_low = wavelet_decomposition(style_feat)
    B, Cc, H, W = content_high.shape
    from . import ops
    ops.add_f32(content_high.view(B * Cc, H, W), style_low.view(B * Cc, H, W))
    return content_high


def calc_mean_std(feat: torch.Tensor, eps: float = 1e-5):
    """utils/wavelet_color_fix.py:47-60: per (batch, channel) mean and sqrt(unbiased variance + eps)."""
    planes, H, W = _planes(feat, "feat")
    mean = torch.empty(planes, device=feat.device, dtype=torch.float32)
    var = torch.empty_like(mean)
    check(_lib.load().dit4sr_plane_stats_f32(_ptr(feat), planes, H * W, _ptr(mean), _ptr(var), _stream()),
          "dit4sr_plane_stats_f32")
    _count()
    b, c = feat.shape[:2]
    return mean.view(b, c, 1, 1), (var + eps).sqrt().view(b, c, 1, 1)


def adaptive_instance_normalization(content_feat: torch.Tensor, style_feat: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """utils/wavelet_color_fix.py:62-74."""
    planes, H, W = _planes(content_feat, "content_feat")
    ps, Hs, Ws = _planes(style_feat, "style_feat")
    if ps != planes:
        raise _lib.Dit4srError("adaptive_instance_normalization: batch/channel mismatch")
    lib = _lib.load()
    st = torch.empty(4, planes, device=content_feat.device, dtype=torch.float32)
    check(lib.dit4sr_plane_stats_f32(_ptr(content_feat), planes, H * W, _ptr(st[0]), _ptr(st[1]), _stream()),
          "dit4sr_plane_stats_f32")
    check(lib.dit4sr_plane_stats_f32(_ptr(style_feat), planes, Hs * Ws, _ptr(st[2]), _ptr(st[3]), _stream()),
          "dit4sr_plane_stats_f32")
    out = torch.empty_like(content_feat)
    check(lib.dit4sr_adain_apply_f32(_ptr(content_feat), _ptr(out), planes, H * W, _ptr(st[0]), _ptr(st[1]), _ptr(st[2]),
                                     _ptr(st[3]), C.c_float(eps), _stream()), "dit4sr_adain_apply_f32")
    _count(3)
    return out


def _pil_to_tensor(img, device):
    import numpy as np
    a = np.asarray(img.convert("RGB"), dtype=np.uint8)
    return (torch.from_numpy(a.copy()).to(device).permute(2, 0, 1).float() / 255.0).unsqueeze(0).contiguous()


def _tensor_to_pil(t):
    from PIL import Image
    a = (t.squeeze(0).clamp_(0.0, 1.0) * 255.0).to(torch.uint8)      # ToPILImage: mul(255).byte()
    return Image.fromarray(a.permute(1, 2, 0).cpu().numpy())


def wavelet_color_fix(target, source, device="cuda"):
    """utils/wavelet_color_fix.py:30-44 on PIL images (same size)."""
    return _tensor_to_pil(wavelet_reconstruction(_pil_to_tensor(target, device), _pil_to_tensor(source, device)))


def adain_color_fix(target, source, device="cuda"):
    """utils/wavelet_color_fix.py:14-28 on PIL images."""
    return _tensor_to_pil(adaptive_instance_normalization(_pil_to_tensor(target, device), _pil_to_tensor(source, device)))

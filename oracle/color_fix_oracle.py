"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the reference's colour fix, `utils/wavelet_color_fix.py`, in numpy (float32 like the reference).
Pinned: `tests/golden/make_golden_colorfix.py` imports the reference module ITSELF (it only needs torch, torchvision
and PIL, all present) and records its outputs; `tests/test_colorfix.py` checks this restatement against them."""
import numpy as np

_K = np.array([[0.0625, 0.125, 0.0625], [0.125, 0.25, 0.125], [0.0625, 0.125, 0.0625]], dtype=np.float32)


def wavelet_blur(image: np.ndarray, radius: int) -> np.ndarray:
    """utils/wavelet_color_fix.py:76-94: replicate-pad by `radius`, depthwise 3x3 conv with dilation `radius`."""
    B, Cc, H, W = image.shape
    pad = np.pad(image, ((0, 0), (0, 0), (radius, radius), (radius, radius)), mode="edge")
    out = np.zeros_like(image, dtype=np.float32)
    for dy in range(3):
        for dx in range(3):
            out += _K[dy, dx] * pad[:, :, dy * radius: dy * radius + H, dx * radius: dx * radius + W]
    return out.astype(np.float32)


def wavelet_decomposition(image: np.ndarray, levels: int = 5):
    """utils/wavelet_color_fix.py:96-110."""
    high = np.zeros_like(image, dtype=np.float32)
    low = image
    for i in range(levels):
        low = wavelet_blur(image, 2 ** i)
        high += image - low
        image = low
    return high, low


def wavelet_reconstruction(content: np.ndarray, style: np.ndarray) -> np.ndarray:
    """utils/wavelet_color_fix.py:112-124."""
    ch, _ = wavelet_decomposition(content.astype(np.float32))
    _, sl = wavelet_decomposition(style.astype(np.float32))
    return ch + sl


def calc_mean_std(feat: np.ndarray, eps: float = 1e-5):
    """utils/wavelet_color_fix.py:47-60 (torch .var is the unbiased estimator)."""
    b, c = feat.shape[:2]
    f = feat.reshape(b, c, -1).astype(np.float64)
    var = f.var(axis=2, ddof=1) + eps
    return f.mean(axis=2).reshape(b, c, 1, 1).astype(np.float32), np.sqrt(var).reshape(b, c, 1, 1).astype(np.float32)


def adaptive_instance_normalization(content: np.ndarray, style: np.ndarray) -> np.ndarray:
    """utils/wavelet_color_fix.py:62-74."""
    sm, ss = calc_mean_std(style)
    cm, cs = calc_mean_std(content)
    return ((content - cm) / cs * ss + sm).astype(np.float32)

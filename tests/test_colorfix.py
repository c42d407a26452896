"""Colour fix (reference utils/wavelet_color_fix.py): the numpy oracle against golden vectors
recorded from the reference module itself (CPU), and the CUDA kernels against both (GPU, through the C ABI)."""
import os

import numpy as np
import pytest
import torch
from safetensors.torch import load_file

from oracle import color_fix_oracle as CO

GOLD = load_file(os.path.join(os.path.dirname(__file__), "golden", "reference_colorfix.safetensors"))
CASES = sorted({k.split(".")[0] for k in GOLD})
TOL = 2e-6     # fp32 images in [0, 1]; only the summation order of the 9 taps / of the statistics differs


def maxerr(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max())


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(case):
    c, s = GOLD[f"{case}.content"].numpy(), GOLD[f"{case}.style"].numpy()
    assert maxerr(CO.wavelet_blur(c, 4), GOLD[f"{case}.blur4"].numpy()) <= TOL
    assert maxerr(CO.wavelet_reconstruction(c, s), GOLD[f"{case}.wavelet"].numpy()) <= TOL
    assert maxerr(CO.adaptive_instance_normalization(c, s), GOLD[f"{case}.adain"].numpy()) <= 2e-5


def test_oracle_properties():
    rng = np.random.default_rng(0)
    x = rng.random((1, 3, 20, 31), dtype=np.float32)
    assert maxerr(CO.wavelet_reconstruction(x, x), x) <= 2e-6          # own high + own low frequencies
    const = np.full((1, 3, 9, 9), 0.3, dtype=np.float32)
    assert maxerr(CO.wavelet_blur(const, 8), const) <= 1e-7             # the kernel sums to one; replicate padding


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_matches_reference_golden(case):
    from dit4sr_b200 import color_fix as CF
    c, s = GOLD[f"{case}.content"].cuda(), GOLD[f"{case}.style"].cuda()
    assert maxerr(CF.wavelet_blur(c, 4).cpu(), GOLD[f"{case}.blur4"]) <= TOL
    assert maxerr(CF.wavelet_reconstruction(c, s).cpu(), GOLD[f"{case}.wavelet"]) <= TOL
    assert maxerr(CF.adaptive_instance_normalization(c, s).cpu(), GOLD[f"{case}.adain"]) <= 2e-5
    m, sd = CF.calc_mean_std(c)
    om, osd = CO.calc_mean_std(c.cpu().numpy())
    assert maxerr(m.cpu(), om) <= 1e-6 and maxerr(sd.cpu(), osd) <= 1e-6


@pytest.mark.gpu
def test_cuda_full_size_properties():
    """2048 x 2048 RGB (the C4 output size): size-independent properties instead of a CPU comparison."""
    from dit4sr_b200 import color_fix as CF
    torch.manual_seed(0)
    x = torch.rand(1, 3, 2048, 2048, device="cuda")
    y = torch.rand(1, 3, 2048, 2048, device="cuda") * 0.5 + 0.2
    assert (CF.wavelet_reconstruction(x, x) - x).abs().max().item() <= 4e-6
    hi, lo = CF.wavelet_decomposition(x)
    assert (hi + lo - x).abs().max().item() <= 4e-6                       # the decomposition telescopes
    out = CF.adaptive_instance_normalization(x, y)
    m, s = CF.calc_mean_std(out)
    ms, ss = CF.calc_mean_std(y)
    # takes the style's statistics (the eps inside both square roots shifts the std by ~eps / (2 std) = 4e-5 here)
    assert (m - ms).abs().max().item() <= 1e-5 and (s - ss).abs().max().item() <= 1e-4
    assert (CF.adaptive_instance_normalization(x, x) - x).abs().max().item() <= 2e-6


@pytest.mark.gpu
def test_cuda_pil_wrappers_match_reference_semantics():
    from PIL import Image
    from dit4sr_b200 import color_fix as CF
    rng = np.random.default_rng(1)
    a = Image.fromarray(rng.integers(0, 256, (37, 52, 3), dtype=np.uint8))
    b = Image.fromarray(rng.integers(0, 256, (37, 52, 3), dtype=np.uint8))
    ta = np.asarray(a, dtype=np.float32).transpose(2, 0, 1)[None] / 255.0
    tb = np.asarray(b, dtype=np.float32).transpose(2, 0, 1)[None] / 255.0
    for fn, ofn in ((CF.wavelet_color_fix, CO.wavelet_reconstruction), (CF.adain_color_fix, CO.adaptive_instance_normalization)):
        got = np.asarray(fn(a, b), dtype=np.int32)
        want = (np.clip(ofn(ta, tb), 0.0, 1.0) * 255.0).astype(np.uint8).transpose(0, 2, 3, 1)[0].astype(np.int32)
        assert np.abs(got - want).max() <= 1       # a float 1e-6 away from an integer boundary may truncate either way

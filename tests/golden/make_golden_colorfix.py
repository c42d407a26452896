"""Generates tests/golden/reference_colorfix.safetensors by running the REFERENCE's own, unmodified
utils/wavelet_color_fix.py (importable here: it needs only torch, torchvision and PIL) on seeded inputs.
Run in the build container:  python tests/golden/make_golden_colorfix.py"""
import importlib.util
import os

import torch
from safetensors.torch import save_file

REF = "/root/reference/utils/wavelet_color_fix.py"
spec = importlib.util.spec_from_file_location("ref_wavelet_color_fix", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

CASES = [("small", 1, 40, 56, 0), ("odd", 2, 33, 47, 1), ("tiny", 1, 5, 7, 2), ("mid", 1, 48, 64, 3)]
out = {}
for name, B, H, W, seed in CASES:
    g = torch.Generator().manual_seed(seed)
    content = torch.rand(B, 3, H, W, generator=g)
    style = (torch.rand(B, 3, H, W, generator=g) * 0.7 + 0.1)
    out[f"{name}.content"] = content
    out[f"{name}.style"] = style
    out[f"{name}.wavelet"] = ref.wavelet_reconstruction(content.clone(), style.clone())
    out[f"{name}.adain"] = ref.adaptive_instance_normalization(content.clone(), style.clone())
    out[f"{name}.blur4"] = ref.wavelet_blur(content.clone(), 4)
save_file(out, os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_colorfix.safetensors"))
print("wrote", len(out), "tensors")

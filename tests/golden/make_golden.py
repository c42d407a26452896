"""Generate golden vectors by running the REFERENCE's own first-party model code.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

`/root/reference/model_dit4sr/transformer_sd3.py` (+ attention.py) are imported unmodified; the absent
third-party dependency diffusers==0.31.0 is provided by `oracle/diffusers_shim` (a restatement, see its README).
Outputs (small .safetensors + json) are committed; tests compare the oracle and the CUDA path against them.
"""
import json
import os
import re
import sys
from types import SimpleNamespace as NS

import torch
from safetensors.torch import save_file

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "diffusers_shim"))
sys.path.insert(0, "/root/reference")

from oracle import dit4sr_oracle as O  # noqa: E402


def ref_cfg(feat_layers=(), feat="hq_feat", ts=False):
    return NS(train=NS(model=["ts_module"] if ts else [], transformer=NS(feat_extract_layer=list(feat_layers),
                                                                       feat_extract=feat)))


def run_reference(cfg, sd, inputs, ocr=False):
    if ocr:
        from model_dit4sr.transformer_sd3_ocrbranch_ocr2hq2ocr import SD3Transformer2DModel_OCRBranch_OCR2HQ2OCR as SD3Transformer2DModel
    else:
        from model_dit4sr.transformer_sd3 import SD3Transformer2DModel
    m = SD3Transformer2DModel(**vars(cfg)).eval()
    m.load_state_dict(sd, strict=True)
    hs, lr, enc, pooled, t = inputs
    with torch.no_grad():
        return m(hidden_states=hs, controlnet_image=lr, encoder_hidden_states=enc, pooled_projections=pooled,
                 timestep=t, return_dict=False, cfg=ref_cfg())[0]


def main():
    out = {}
    meta = {"cases": []}
    cases = [
        # name, cfg kwargs, batch, h, w, T, timestep
        ("tiny_qknorm_dual", dict(), 2, 16, 16, 13, 741.44),
        ("tiny_noqknorm", dict(qk_norm=None, dual_attention_layers=()), 1, 8, 8, 5, 1000.0),
        ("tiny_t9", dict(), 1, 16, 16, 13, 8.93),
        # 3-stream OCR-branch model (transformer_sd3_ocrbranch_ocr2hq2ocr.py)
        ("tiny_ocr_qknorm_dual", dict(), 2, 16, 16, 13, 741.44),
        ("tiny_ocr_noqknorm", dict(qk_norm=None, dual_attention_layers=()), 1, 8, 8, 5, 500.0),
    ]
    for name, kw, B, h, w, T, ts in cases:
        ocr = "_ocr_" in name
        cfg = O.tiny_config(**kw)
        sd = O.init_state_dict(cfg, seed=0, ocr=ocr)
        inputs = O.make_inputs(cfg, B, h, w, T, seed=1, timestep=ts)
        y = run_reference(cfg, sd, inputs, ocr=ocr)
        out[name + ".velocity"] = y.contiguous()
        meta["cases"].append(dict(name=name, ocr=ocr,
                                  cfg_kwargs={k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()},
                                  batch=B, height=h, width=w, text_tokens=T, timestep=ts, weight_seed=0, input_seed=1))
    save_file(out, os.path.join(HERE, "reference_tiny.safetensors"))

    # scheduler KAT: the 40 timesteps the reference logged
    log = ("/root/reference/result_val_dit4sr_3rd/few_eval/realtext/")
    sub = os.listdir(log)[0]
    f = os.path.join(log, sub, "final_pred_txt", "Canon_21_x4_5_crop_0.txt")
    ts = [float(x) for x in re.findall(r"timestep:\s+([0-9.]+)", open(f).read())]
    assert len(ts) == 40
    meta["logged_timesteps_40"] = ts
    meta["logged_timesteps_source"] = "result_val_dit4sr_3rd/few_eval/realtext/<run>/final_pred_txt/Canon_21_x4_5_crop_0.txt"

    # state_dict key manifest of the reference ctor at the full SD3.5-medium shape (names + shapes only)
    from model_dit4sr.transformer_sd3 import SD3Transformer2DModel
    with torch.device("meta"):
        full = SD3Transformer2DModel(**vars(O.sd35_medium_config(pos_embed_max_size=4, sample_size=8)))
    meta["full_manifest_note"] = "pos_embed_max_size shrunk to 4 to keep ctor cheap; other shapes are SD3.5-medium"
    meta["full_manifest"] = {k: list(v.shape) for k, v in full.state_dict().items()}
    from model_dit4sr.transformer_sd3_ocrbranch_ocr2hq2ocr import SD3Transformer2DModel_OCRBranch_OCR2HQ2OCR
    with torch.device("meta"):
        full = SD3Transformer2DModel_OCRBranch_OCR2HQ2OCR(**vars(O.sd35_medium_config(pos_embed_max_size=4, sample_size=8)))
    meta["full_manifest_ocr"] = {k: list(v.shape) for k, v in full.state_dict().items()}
    json.dump(meta, open(os.path.join(HERE, "reference_meta.json"), "w"), indent=1)
    print("wrote", len(out), "tensors;", len(meta["full_manifest"]), "manifest keys")


if __name__ == "__main__":
    main()

"""CPU tests: the oracle against the golden vectors made from the reference's own code, and its self-checks."""
import json
import os

import numpy as np
import pytest
import torch
from safetensors.torch import load_file

from oracle import dit4sr_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
META = json.load(open(os.path.join(GOLD, "reference_meta.json")))
REF = load_file(os.path.join(GOLD, "reference_tiny.safetensors"))


@pytest.mark.parametrize("case", META["cases"], ids=lambda c: c["name"])
@pytest.mark.parametrize("as_written", [True, False])
def test_oracle_matches_reference_golden(case, as_written):
    """fp32, tolerance 1e-6 abs (bit-exact in the generating container; slack for other BLAS builds)."""
    ocr = case.get("ocr", False)
    cfg = O.tiny_config(**case["cfg_kwargs"])
    sd = O.init_state_dict(cfg, seed=case["weight_seed"], ocr=ocr)
    inp = O.make_inputs(cfg, case["batch"], case["height"], case["width"], case["text_tokens"],
                        seed=case["input_seed"], timestep=case["timestep"])
    y = O.model_forward(sd, cfg, *inp, as_written=as_written, ocr=ocr)
    ref = REF[case["name"] + ".velocity"]
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 1e-6


def test_scheduler_known_answer():
    """The 40 timesteps the reference logged (result_val_dit4sr_3rd/...) == FlowMatchEuler(shift=3)."""
    sig = O.flow_match_sigmas(40)
    assert sig.shape == (41,) and sig[-1] == 0.0
    ts = 1000.0 * sig[:-1]
    logged = np.array(META["logged_timesteps_40"])
    assert np.all(np.abs(ts - logged) <= 0.005 + 1e-6)  # printed with 2 decimals


def test_manifest_matches_reference_ctor_full_shape():
    cfg = O.sd35_medium_config(pos_embed_max_size=4, sample_size=8)
    man = O.state_dict_manifest(cfg)
    ref = META["full_manifest"]
    assert list(man.keys()) == list(ref.keys())
    for k, shp in man.items():
        assert list(shp) == ref[k], k


def test_manifest_matches_reference_ctor_full_shape_ocr():
    cfg = O.sd35_medium_config(pos_embed_max_size=4, sample_size=8)
    man = O.state_dict_manifest(cfg, ocr=True)
    ref = META["full_manifest_ocr"]
    assert list(man.keys()) == list(ref.keys())
    for k, shp in man.items():
        assert list(shp) == ref[k], k


def test_ocr_attn2_splits_three_streams_in_half():
    """Q6: attn2 of the OCR model is the 2-stream processor fed 3N tokens -> halves of 1.5N. Feeding the 2-stream oracle
    attention a [1.5N | 1.5N] input must equal what the OCR block computes for attn2."""
    cfg = O.tiny_config()
    sd = O.init_state_dict(cfg, ocr=True)
    x = torch.randn(1, 3 * 16, 128)
    a = O.joint_attention(sd, cfg, "transformer_blocks.0.attn2", x, None)
    hq, lr = x[:, :24], x[:, 24:]
    b = O.joint_attention(sd, cfg, "transformer_blocks.0.attn2", torch.cat([hq, lr], 1), None)
    assert a.shape == (1, 48, 128) and torch.equal(a, b)


def test_dead_key_chain_elision_is_exact_fp64():
    """Q1: with qk_norm the to_k_control(to_k(lr)) chain is dead; eliding it changes nothing."""
    cfg = O.tiny_config()
    sd = {k: v.double() for k, v in O.init_state_dict(cfg).items()}
    inp = [t.double() if t.dtype.is_floating_point else t for t in O.make_inputs(cfg, 1, 8, 8, 7)]
    a = O.model_forward(sd, cfg, *inp, as_written=True)
    b = O.model_forward(sd, cfg, *inp, as_written=False)
    assert torch.equal(a, b)


def test_lr_path_is_exercised():
    """Q4: zeroing the control projections must change the output (i.e. random init exercises the LR path)."""
    cfg = O.tiny_config()
    sd = O.init_state_dict(cfg)
    inp = O.make_inputs(cfg, 1, 8, 8, 7)
    a = O.model_forward(sd, cfg, *inp)
    sd2 = dict(sd)
    for k in sd:
        if "control" in k:
            sd2[k] = torch.zeros_like(sd[k])
    b = O.model_forward(sd2, cfg, *inp)
    assert (a - b).abs().max() > 1e-4


def test_sample_loop_euler():
    cfg = O.tiny_config()
    sd = O.init_state_dict(cfg)
    hs, lr, enc, pooled, _ = O.make_inputs(cfg, 1, 8, 8, 5)
    x = O.sample_loop(sd, cfg, hs, lr, enc, pooled, num_steps=4)
    assert x.shape == hs.shape and torch.isfinite(x).all()
    # a constant-velocity field integrates to x0 - v (sum of dsigma = -1)
    v = torch.ones_like(hs)
    y = O.sample_loop(sd, cfg, hs, lr, enc, pooled, num_steps=7, model_fn=lambda x_, t_: v)
    assert torch.allclose(y, hs - v, atol=1e-5)


def test_cfg_loop_reduces_to_the_plain_loop_and_lre_blend():
    """oracle.sample_loop_cfg (pipeline_dit4sr.py:1085,1106-1108,1561-1567): with the negative prompt equal to the positive
    one the guided velocity equals the plain one whatever the scale; the LRE start point blends with control[:1]."""
    import torch
    from oracle import dit4sr_oracle as O
    cfg = O.tiny_config()
    sd = O.init_state_dict(cfg, seed=0)
    hs, lr, enc, pooled, _ = O.make_inputs(cfg, 2, 8, 8, 5, seed=1)
    with torch.no_grad():
        a = O.sample_loop(sd, cfg, hs, lr, enc, pooled, num_steps=3)
        b = O.sample_loop_cfg(sd, cfg, hs, lr, enc, pooled, enc, pooled, 8.0, num_steps=3)
    assert torch.allclose(a, b, atol=2e-4, rtol=1e-4)
    x0 = O.lre_start_point(hs, lr, 0.75)
    assert torch.allclose(x0[1], 0.25 * lr[0] + 0.75 * hs[1])

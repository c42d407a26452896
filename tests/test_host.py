"""CPU tests of the host side: C-ABI library loads and exports every declared symbol, drop-in module surface."""
import ctypes
import os
import re

import pytest
import torch

import dit4sr_b200
from dit4sr_b200 import _lib, build
from oracle import dit4sr_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return _lib.load()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "dit4sr_b200.h")).read()
    declared = set(re.findall(r"\b(dit4sr_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.dit4sr_abi_version() == 2


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    p = _lib.GemmParams()
    rc = lib.dit4sr_gemm_bf16(ctypes.byref(p), None)
    assert rc == -1 and b"null pointer" in lib.dit4sr_last_error()
    rc = lib.dit4sr_patchify(None, 0, None, 1, 16, 8, 8, None)
    assert rc == -1


def test_state_dict_layout_matches_reference_manifest():
    for cfg in (O.tiny_config(), O.tiny_config(qk_norm=None, dual_attention_layers=())):
        m = dit4sr_b200.SD3Transformer2DModel(**vars(cfg))
        man = O.state_dict_manifest(cfg)
        sd = m.state_dict()
        assert list(sd.keys()) == list(man.keys())
        assert all(tuple(sd[k].shape) == man[k] for k in sd)


def test_control_modules_zero_initialised_like_reference():
    m = dit4sr_b200.SD3Transformer2DModel(**vars(O.tiny_config()))
    a = m.transformer_blocks[0].attn
    assert a.to_q_control.weight.abs().max() == 0 and a.to_v_control.bias.abs().max() == 0
    assert torch.equal(a.to_out_control.weight, a.to_out[0].weight)
    assert m.transformer_blocks[0].ff.control_conv.weight.abs().max() == 0


def test_no_cpu_fallback():
    cfg = O.tiny_config()
    m = dit4sr_b200.SD3Transformer2DModel(**vars(cfg)).requires_grad_(False)
    hs, lr, enc, pooled, t = O.make_inputs(cfg, 1, 8, 8, 5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(hs, lr, enc, pooled, t)


def test_save_and_load_roundtrip(tmp_path):
    cfg = O.tiny_config()
    m = dit4sr_b200.SD3Transformer2DModel(**vars(cfg))
    m.load_state_dict(O.init_state_dict(cfg))
    m.save_pretrained(str(tmp_path / "transformer"))
    m2 = dit4sr_b200.SD3Transformer2DModel.from_pretrained_local(str(tmp_path), subfolder="transformer")
    assert m2.config.dual_attention_layers == (0,) and m2.config.qk_norm == "rms_norm"
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def test_sampler_sigmas_match_oracle_and_logged_timesteps():
    import json
    import numpy as np
    from dit4sr_b200.sampler import flow_match_sigmas
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_meta.json")))
    s = flow_match_sigmas(40)
    assert np.array_equal(s, O.flow_match_sigmas(40))
    assert np.all(np.abs(1000 * s[:-1] - np.array(meta["logged_timesteps_40"])) <= 0.005 + 1e-6)

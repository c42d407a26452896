"""Cross-checks of the diffusers-0.31.0 semantics the oracle's shim restates (diffusers itself is not
installable here, so that layer of the oracle is otherwise unpinned) against INDEPENDENT implementations of the same
building blocks that ship in this image's `transformers` package.  Skipped when those modules are not importable."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "diffusers_shim"))

from diffusers.models import embeddings as shim_emb        # noqa: E402  (the shim, not a real diffusers)
from diffusers.models import normalization as shim_norm    # noqa: E402


def _omni():
    return pytest.importorskip("transformers.models.qwen2_5_omni.modeling_qwen2_5_omni")


def test_adalayernorm_zero_matches_independent_implementation():
    """AdaLayerNormZero (attention.py:208-222 uses it for norm1 / norm1_context): Linear(dim, 6 dim) on silu(emb), chunk
    order shift/scale/gate/shift_mlp/scale_mlp/gate_mlp, LayerNorm(eps 1e-6, no affine) * (1 + scale) + shift."""
    m = _omni()
    torch.manual_seed(0)
    dim = 96
    ours, theirs = shim_norm.AdaLayerNormZero(dim), m.Qwen2_5_OmniAdaLayerNormZero(dim)
    theirs.linear.load_state_dict(ours.linear.state_dict())
    x, emb = torch.randn(3, 17, dim), torch.randn(3, dim)
    a, b = ours(x, emb=emb), theirs(x, emb=emb)
    assert len(a) == len(b) == 5
    for u, v in zip(a, b):
        assert torch.equal(u, v)


def test_adalayernorm_continuous_matches_independent_implementation():
    """AdaLayerNormContinuous as norm_out / the last block's norm1_context build it (transformer_sd3.py:116,
    attention.py:214: elementwise_affine=False, eps=1e-6): scale, shift = chunk(2) - in THAT order."""
    m = _omni()
    torch.manual_seed(1)
    dim = 64
    ours = shim_norm.AdaLayerNormContinuous(dim, dim, elementwise_affine=False, eps=1e-6)
    theirs = m.Qwen2_5_OmniAdaLayerNormZero_Final(dim)
    theirs.linear.load_state_dict(ours.linear.state_dict())
    x, emb = torch.randn(2, 9, dim), torch.randn(2, dim)
    assert torch.equal(ours(x, emb), theirs(x, emb))


def test_rms_norm_matches_llama_rms_norm():
    """diffusers RMSNorm(64, eps=1e-6) on q / k heads (attention.py:484-489,502-505): variance in fp32, x * rsqrt(var + eps)
    then * weight - the Llama formulation."""
    llama = pytest.importorskip("transformers.models.llama.modeling_llama")
    torch.manual_seed(2)
    ours, theirs = shim_norm.RMSNorm(64, eps=1e-6), llama.LlamaRMSNorm(64, eps=1e-6)
    w = 1 + 0.1 * torch.randn(64)
    with torch.no_grad():
        ours.weight.copy_(w)
        theirs.weight.copy_(w)
    x = torch.randn(2, 24, 50, 64) * 3
    assert torch.allclose(ours(x), theirs(x), rtol=1e-6, atol=1e-6)
    xb = x.bfloat16()
    assert torch.allclose(ours.bfloat16()(xb).float(), theirs.bfloat16()(xb).float(), rtol=2e-2, atol=2e-2)


def test_2d_sincos_pos_embed_matches_mae_when_unscaled():
    """PatchEmbed's pos_embed buffer: with base_size == grid_size and interpolation_scale 1 diffusers' formula reduces to
    the MAE one (w-first meshgrid, [emb_h | emb_w], [sin | cos] halves, omega = 1 / 10000^(k / (D/4)))."""
    mae = pytest.importorskip("transformers.models.vit_mae.modeling_vit_mae")
    for dim, g in ((64, 6), (1536, 12)):
        ours = shim_emb.get_2d_sincos_pos_embed(dim, g, base_size=g, interpolation_scale=1.0)
        theirs = mae.get_2d_sincos_pos_embed(dim, g, add_cls_token=False)
        assert ours.shape == theirs.shape == (g * g, dim)
        assert np.allclose(ours, theirs, rtol=0, atol=1e-12)


def test_oracle_functions_agree_with_the_shim_modules():
    """The functional oracle (oracle/dit4sr_oracle.py) and the shim modules are two restatements by construction pinned
    to each other through the golden vectors; spot-check the two pieces above directly."""
    from oracle import dit4sr_oracle as O
    torch.manual_seed(3)
    x = torch.randn(2, 5, 64)
    w = 1 + 0.1 * torch.randn(64)
    n = shim_norm.RMSNorm(64, eps=1e-6)
    with torch.no_grad():
        n.weight.copy_(w)
    assert torch.allclose(O.rms_norm(x, w), n(x), rtol=1e-6, atol=1e-6)
    pe = O.sincos_pos_embed_2d(64, 6, 6)
    ref = torch.from_numpy(shim_emb.get_2d_sincos_pos_embed(64, 6, base_size=6)).float()
    assert torch.allclose(pe.reshape(-1, 64).float(), ref, atol=1e-6)


def test_timesteps_match_the_flux_reference_embedding():
    """diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0) as CombinedTimestepTextProjEmbeddings builds it
    (transformer_sd3.py:95-97) against an independent implementation of the same rectified-flow-transformer embedding:
    torchtitan's copy of the Flux reference `timestep_embedding` ([cos | sin], freq_k = exp(-ln(1e4) k / 128), which takes
    t in [0, 1] and multiplies by 1000).  (TVM's port of diffusers' get_timestep_embedding, tilelang/3rdparty/tvm/python/
    tvm/relax/frontend/nn/op.py:1741-1800, documents the same flip: [sin | cos] -> [cos | sin].)"""
    layers = pytest.importorskip("torchtitan.experiments.flux.model.layers")
    t = torch.tensor([1000.0, 741.44, 8.93, 0.0, 333.3])
    ours = shim_emb.Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0)(t)
    theirs = layers.timestep_embedding(t / 1000.0, 256)
    assert ours.shape == theirs.shape == (5, 256)
    assert torch.allclose(ours, theirs, atol=2e-4, rtol=0)     # fp32 phase t * freq evaluated as (t/1000)*1000*freq there
    # and the oracle's own restatement used by model_forward
    from oracle import dit4sr_oracle as O
    assert torch.allclose(O.timestep_sincos(t, 256), ours, atol=1e-6, rtol=0)


def test_gelu_tanh_matches_the_independent_closed_form():
    """GELU(approximate='tanh') of FeedForward / FeedForwardControl (attention.py:75,135) against transformers' own
    closed-form implementation 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))."""
    act = pytest.importorskip("transformers.activations")
    x = torch.linspace(-6, 6, 4001)
    g = shim_emb.nn.functional.gelu(x, approximate="tanh")
    assert torch.allclose(g, act.NewGELUActivation()(x), atol=2e-6, rtol=0)


def test_pos_embed_crop_is_the_centred_window_of_the_independent_table():
    """PatchEmbed.cropped_pos_embed (transformer_sd3.py:326 -> pos_embed): the window of the pos_embed_max_size^2 table is
    centred, top = (max - h/2) // 2, left = (max - w/2) // 2.  The table itself is cross-checked against ViT-MAE's 2-D
    sincos; the centring rule is restated from memory of diffusers 0.31.0 (no independent copy ships in this image) - but
    real checkpoints carry the table as a buffer, so only the window arithmetic is at stake."""
    pe = shim_emb.PatchEmbed(height=16, width=16, patch_size=2, in_channels=4, embed_dim=32, pos_embed_max_size=12)
    table = pe.pos_embed.reshape(12, 12, 32)
    for h, w in ((16, 16), (8, 20), (24, 4)):
        got = pe.cropped_pos_embed(h, w).reshape(h // 2, w // 2, 32)
        top, left = (12 - h // 2) // 2, (12 - w // 2) // 2
        assert torch.equal(got, table[top:top + h // 2, left:left + w // 2])


# ---- VAE oracle (oracle/vae_oracle.py restates diffusers' AutoencoderKL from memory) against the LDM-lineage VQ-VAE encoder /
# decoder modules that ship in `transformers` (Chameleon's encoder, Janus' decoder): the architecture diffusers' AutoencoderKL
# checkpoints were converted from (convert_ldm_vae_checkpoint: down.{i}.block.{j} -> down_blocks.{i}.resnets.{j}, nin_shortcut ->
# conv_shortcut, mid.block_1 / attn_1 / block_2 -> mid_block.resnets.0 / attentions.0 / resnets.1, q / k / v / proj_out 1x1
# convolutions -> to_q / to_k / to_v / to_out.0 linears, norm_out -> conv_norm_out).
def _ldm_to_diffusers(sd_ldm, side, levels):
    """State dict of an LDM-style encoder / decoder (`side` = "encoder" | "decoder") under diffusers' AutoencoderKL key names."""
    out = {}
    blocks, res, samp, samp_new = (("down", "down_blocks", "downsample", "downsamplers.0") if side == "encoder"
                                   else ("up", "up_blocks", "upsample", "upsamplers.0"))
    for k, v in sd_ldm.items():
        p = k.split(".")
        if p[0] == blocks and p[2] == "block":
            name = f"{res}.{p[1]}.resnets.{p[3]}." + ".".join(p[4:]).replace("nin_shortcut", "conv_shortcut")
        elif p[0] == blocks and p[2] == samp:
            name = f"{res}.{p[1]}.{samp_new}." + ".".join(p[3:])
        elif p[0] == blocks and p[2] == "attn":
            continue                                        # (neutralised up-level attention of the Janus decoder, see below)
        elif p[0] == "mid" and p[1] in ("block_1", "block_2"):
            name = f"mid_block.resnets.{0 if p[1] == 'block_1' else 1}." + ".".join(p[2:])
        elif p[0] == "mid" and p[1] == "attn_1":
            sub = {"norm": "group_norm", "q": "to_q", "k": "to_k", "v": "to_v", "proj_out": "to_out.0"}[p[2]]
            name = f"mid_block.attentions.0.{sub}.{p[3]}"
            if p[2] != "norm" and p[3] == "weight":
                v = v[:, :, 0, 0]                           # 1x1 convolution -> linear
        elif p[0] == "norm_out":
            name = "conv_norm_out." + p[1]
        else:
            name = k                                        # conv_in, conv_out
        out[f"{side}.{name}"] = v.detach().clone()
    return out


def _ldm_config():
    from types import SimpleNamespace
    return SimpleNamespace(channel_multiplier=(1, 2, 4, 4), num_res_blocks=2, base_channels=32, resolution=64, in_channels=3,
                           out_channels=3, double_latent=True, latent_channels=16, attn_resolutions=None, attn_type="vanilla",
                           dropout=0.0)


def _randomise(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in module.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.dim() > 1 else 0.5) + (1.0 if p.dim() == 1 else 0.0))


def test_vae_oracle_encoder_matches_the_ldm_encoder_in_transformers():
    cham = pytest.importorskip("transformers.models.chameleon.modeling_chameleon")
    from oracle import vae_oracle as V
    torch.manual_seed(0)
    enc = cham.ChameleonVQVAEEncoder(_ldm_config()).eval()
    _randomise(enc, 1)
    cfg = V.sd3_vae_config(block_out_channels=(32, 64, 128, 128))
    sd = _ldm_to_diffusers(enc.state_dict(), "encoder", 4)
    want = {k: s for k, s in V.state_dict_shapes(cfg).items() if k.startswith("encoder.")}
    assert {k: tuple(v.shape) for k, v in sd.items()} == want          # same parameters, same shapes, nothing left over
    x = torch.randn(2, 3, 48, 80)
    with torch.no_grad():
        ref = enc(x.clone())
        got = V.encoder_forward(sd, cfg, x)
    assert got.shape == ref.shape == (2, 32, 6, 10)
    assert torch.allclose(got, ref, atol=2e-4, rtol=1e-4), (got - ref).abs().max()


def test_vae_oracle_decoder_matches_the_ldm_decoder_in_transformers():
    janus = pytest.importorskip("transformers.models.janus.modeling_janus")
    from oracle import vae_oracle as V
    torch.manual_seed(0)
    dec = janus.JanusVQVAEDecoder(_ldm_config()).eval()
    _randomise(dec, 2)
    # Janus adds attention blocks after the residual blocks of its lowest-resolution up level, which the SD VAE decoder does not
    # have: with a zero output projection each of them is `residual + 0`
    with torch.no_grad():
        for a in dec.up[0].attn:
            a.proj_out.weight.zero_()
            a.proj_out.bias.zero_()
    cfg = V.sd3_vae_config(block_out_channels=(32, 64, 128, 128))
    sd = _ldm_to_diffusers(dec.state_dict(), "decoder", 4)
    want = {k: s for k, s in V.state_dict_shapes(cfg).items() if k.startswith("decoder.")}
    assert {k: tuple(v.shape) for k, v in sd.items()} == want
    z = torch.randn(2, 16, 6, 10)
    with torch.no_grad():
        ref = dec(z.clone())
        got = V.decoder_forward(sd, cfg, z)
    assert got.shape == ref.shape == (2, 3, 48, 80)
    assert torch.allclose(got, ref, atol=2e-4, rtol=1e-4), (got - ref).abs().max()

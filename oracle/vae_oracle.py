"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch fp32) of the VAE the reference calls around its sampling loop.

The reference encodes the LR image with `self.vae.encode(control_image).latent_dist.sample()` and decodes the final latents
with `self.vae.decode(latents, return_dict=False)[0]` (/root/reference/pipelines/pipeline_dit4sr.py:1001-1002, 1602-1605);
`self.vae` is diffusers' `AutoencoderKL` (pipeline_dit4sr.py:154,193), a third-party class whose source is NOT under
/root/reference (diffusers==0.31.0, requirements.txt:3, not installed, no network).  The layer sequence is corroborated by
the reference's own tiled-VAE hook, which walks the same modules by name: /root/reference/utils/vaehook.py:279-358
(conv_in, down_blocks[i].resnets[j], downsamplers[0], mid_block.resnets[0] / attentions[0] / resnets[1], up_blocks[i]
.resnets[0..2], upsamplers[0], conv_norm_out, SiLU, conv_out), :247-276 (ResnetBlock: norm1, silu, conv1, norm2, silu, conv2,
conv_shortcut added to the input), :136-169 (attention: to_q / to_k / to_v, softmax(q k^T / sqrt(C)), to_out[0]) and
:370-412 (GroupNorm over (channels / groups) x H x W with eps 1e-6).  The rest (asymmetric (0,1,0,1) padding of the stride-2
down-sampling convolution, nearest 2x up-sampling, DiagonalGaussianDistribution with logvar clamped to [-30, 20], the SD3
VAE configuration) is restated from the published diffusers 0.31.0 sources from memory.

PINNING: the reference holds no golden vector for this path and diffusers cannot be imported here, so the oracle is pinned
against INDEPENDENT implementations of the same (LDM-lineage) architecture that ship in this image's `transformers` package:
`encoder_forward` equals Chameleon's `ChameleonVQVAEEncoder` and `decoder_forward` equals Janus' `JanusVQVAEDecoder` (its extra
up-level attention blocks neutralised) to 2e-4 on random weights mapped through the LDM -> diffusers key conversion, parameter
names and shapes included (tests/test_oracle_crosscheck.py::test_vae_oracle_*).  What stays unpinned w.r.t. a real diffusers
install: that AutoencoderKL 0.31.0 is this architecture under those key names (the reference's own vaehook walk above says so),
DiagonalGaussianDistribution's clamp, and the SD3 VAE configuration values.

The tiled variant (VAEHook.vae_tile_forward) is an out-of-memory work-around that reproduces the untiled network (GroupNorm
statistics are shared across tiles; its `fast` modes estimate them from a down-sampled copy); a 180 GB B200 decodes
2048 x 2048 images untiled, so the product implements the untiled network only.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F


def sd3_vae_config(**over):
    """AutoencoderKL config of the SD3 / SD3.5 VAE (vae/config.json of stabilityai/stable-diffusion-3.5-medium)."""
    cfg = dict(in_channels=3, out_channels=3, latent_channels=16, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
               norm_num_groups=32, scaling_factor=1.5305, shift_factor=0.0609, use_quant_conv=False,
               use_post_quant_conv=False, mid_block_add_attention=True, sample_size=1024)
    cfg.update(over)
    return SimpleNamespace(**cfg)


def _resnet_keys(prefix, cin, cout, shapes):
    shapes[prefix + "norm1.weight"] = (cin,)
    shapes[prefix + "norm1.bias"] = (cin,)
    shapes[prefix + "conv1.weight"] = (cout, cin, 3, 3)
    shapes[prefix + "conv1.bias"] = (cout,)
    shapes[prefix + "norm2.weight"] = (cout,)
    shapes[prefix + "norm2.bias"] = (cout,)
    shapes[prefix + "conv2.weight"] = (cout, cout, 3, 3)
    shapes[prefix + "conv2.bias"] = (cout,)
    if cin != cout:
        shapes[prefix + "conv_shortcut.weight"] = (cout, cin, 1, 1)
        shapes[prefix + "conv_shortcut.bias"] = (cout,)


def _mid_keys(prefix, c, shapes, attention=True):
    _resnet_keys(prefix + "resnets.0.", c, c, shapes)
    if attention:
        a = prefix + "attentions.0."
        shapes[a + "group_norm.weight"] = (c,)
        shapes[a + "group_norm.bias"] = (c,)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            shapes[a + n + ".weight"] = (c, c)
            shapes[a + n + ".bias"] = (c,)
    _resnet_keys(prefix + "resnets.1.", c, c, shapes)


def state_dict_shapes(cfg):
    """name -> shape of every AutoencoderKL parameter (diffusers key names)."""
    ch = list(cfg.block_out_channels)
    s = {}
    # encoder
    s["encoder.conv_in.weight"] = (ch[0], cfg.in_channels, 3, 3)
    s["encoder.conv_in.bias"] = (ch[0],)
    cin = ch[0]
    for i, cout in enumerate(ch):
        for j in range(cfg.layers_per_block):
            _resnet_keys(f"encoder.down_blocks.{i}.resnets.{j}.", cin if j == 0 else cout, cout, s)
        if i != len(ch) - 1:
            s[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            s[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"] = (cout,)
        cin = cout
    _mid_keys("encoder.mid_block.", ch[-1], s, cfg.mid_block_add_attention)
    s["encoder.conv_norm_out.weight"] = (ch[-1],)
    s["encoder.conv_norm_out.bias"] = (ch[-1],)
    s["encoder.conv_out.weight"] = (2 * cfg.latent_channels, ch[-1], 3, 3)
    s["encoder.conv_out.bias"] = (2 * cfg.latent_channels,)
    # decoder
    rev = ch[::-1]
    s["decoder.conv_in.weight"] = (rev[0], cfg.latent_channels, 3, 3)
    s["decoder.conv_in.bias"] = (rev[0],)
    _mid_keys("decoder.mid_block.", rev[0], s, cfg.mid_block_add_attention)
    cin = rev[0]
    for i, cout in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            _resnet_keys(f"decoder.up_blocks.{i}.resnets.{j}.", cin if j == 0 else cout, cout, s)
        if i != len(rev) - 1:
            s[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            s[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = (cout,)
        cin = cout
    s["decoder.conv_norm_out.weight"] = (rev[-1],)
    s["decoder.conv_norm_out.bias"] = (rev[-1],)
    s["decoder.conv_out.weight"] = (cfg.out_channels, rev[-1], 3, 3)
    s["decoder.conv_out.bias"] = (cfg.out_channels,)
    if cfg.use_quant_conv:
        s["quant_conv.weight"] = (2 * cfg.latent_channels, 2 * cfg.latent_channels, 1, 1)
        s["quant_conv.bias"] = (2 * cfg.latent_channels,)
    if cfg.use_post_quant_conv:
        s["post_quant_conv.weight"] = (cfg.latent_channels, cfg.latent_channels, 1, 1)
        s["post_quant_conv.bias"] = (cfg.latent_channels,)
    return s


def init_state_dict(cfg, seed=0):
    """Every parameter random (there are no weights offline): convolutions / linears ~ N(0, 1/fan_in) scaled so that the
    activations keep O(1) magnitude through the stack, norm weights 1 + 0.1 N, biases 0.02 N."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in state_dict_shapes(cfg).items():
        if "norm" in name and name.endswith("weight"):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            sd[name] = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_in = math.prod(shape[1:])
            sd[name] = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
    return sd


def _gn(x, sd, prefix, groups):
    return F.group_norm(x, groups, sd[prefix + ".weight"], sd[prefix + ".bias"], eps=1e-6)


def _resnet(x, sd, prefix, groups):
    h = F.silu(_gn(x, sd, prefix + "norm1", groups))
    h = F.conv2d(h, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], padding=1)
    h = F.silu(_gn(h, sd, prefix + "norm2", groups))
    h = F.conv2d(h, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"], padding=1)
    if prefix + "conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[prefix + "conv_shortcut.weight"], sd[prefix + "conv_shortcut.bias"])
    return x + h


def _attention(x, sd, prefix, groups):
    b, c, hh, ww = x.shape
    h = _gn(x.reshape(b, c, hh * ww), sd, prefix + "group_norm", groups).transpose(1, 2)      # [b, hw, c]
    q = F.linear(h, sd[prefix + "to_q.weight"], sd[prefix + "to_q.bias"])
    k = F.linear(h, sd[prefix + "to_k.weight"], sd[prefix + "to_k.bias"])
    v = F.linear(h, sd[prefix + "to_v.weight"], sd[prefix + "to_v.bias"])
    p = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(c), dim=-1)                            # one head, head_dim = c
    o = F.linear(p @ v, sd[prefix + "to_out.0.weight"], sd[prefix + "to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(b, c, hh, ww)


def _mid(x, sd, prefix, cfg):
    x = _resnet(x, sd, prefix + "resnets.0.", cfg.norm_num_groups)
    if cfg.mid_block_add_attention:
        x = _attention(x, sd, prefix + "attentions.0.", cfg.norm_num_groups)
    return _resnet(x, sd, prefix + "resnets.1.", cfg.norm_num_groups)


def encoder_forward(sd, cfg, x):
    """Encoder: image [B, 3, H, W] -> moments [B, 2 * latent_channels, H/8, W/8]."""
    n = len(cfg.block_out_channels)
    h = F.conv2d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    for i in range(n):
        for j in range(cfg.layers_per_block):
            h = _resnet(h, sd, f"encoder.down_blocks.{i}.resnets.{j}.", cfg.norm_num_groups)
        if i != n - 1:
            h = F.pad(h, (0, 1, 0, 1))                      # Downsample2D(padding=0): pad right / bottom only, stride 2
            h = F.conv2d(h, sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"],
                         sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"], stride=2)
    h = _mid(h, sd, "encoder.mid_block.", cfg)
    h = F.silu(_gn(h, sd, "encoder.conv_norm_out", cfg.norm_num_groups))
    h = F.conv2d(h, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)
    if cfg.use_quant_conv:
        h = F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])
    return h


def decoder_forward(sd, cfg, z):
    """Decoder: latents [B, latent_channels, h, w] -> image [B, 3, 8h, 8w]."""
    n = len(cfg.block_out_channels)
    if cfg.use_post_quant_conv:
        z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    h = F.conv2d(z, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    h = _mid(h, sd, "decoder.mid_block.", cfg)
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            h = _resnet(h, sd, f"decoder.up_blocks.{i}.resnets.{j}.", cfg.norm_num_groups)
        if i != n - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"],
                         sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
    h = F.silu(_gn(h, sd, "decoder.conv_norm_out", cfg.norm_num_groups))
    return F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)


def gaussian_sample(moments, noise=None):
    """DiagonalGaussianDistribution(moments).sample() (noise given) / .mode() (noise None)."""
    mean, logvar = moments.chunk(2, dim=1)
    if noise is None:
        return mean
    return mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise


def encode_latents(sd, cfg, image, noise=None):
    """pipeline_dit4sr.py:1001-1002: sample the posterior, then (z - shift_factor) * scaling_factor."""
    z = gaussian_sample(encoder_forward(sd, cfg, image), noise)
    return (z - cfg.shift_factor) * cfg.scaling_factor


def decode_latents(sd, cfg, latents):
    """pipeline_dit4sr.py:1602-1604: latents / scaling_factor + shift_factor, then decode."""
    return decoder_forward(sd, cfg, latents / cfg.scaling_factor + cfg.shift_factor)

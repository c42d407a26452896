"""Drop-in for the `AutoencoderKL` the reference's pipeline holds as `self.vae`.

Call sites mirrored (jinlovespho/DiT4SR): pipelines/pipeline_dit4sr.py:1001-1002
    control_image = self.vae.encode(control_image).latent_dist.sample()
    control_image = (control_image - self.vae.config.shift_factor) * self.vae.config.scaling_factor
and :1602-1605
    latents = (latents / self.vae.config.scaling_factor) + self.vae.config.shift_factor
    image = self.vae.decode(latents, return_dict=False)[0]
plus `.config.{block_out_channels, scaling_factor, shift_factor}` (:218-219), `.encoder` / `.decoder` with the module names
the reference's tiled-VAE hook walks (utils/vaehook.py:279-358), and the diffusers `state_dict` key names, so a real
`vae/diffusion_pytorch_model.safetensors` loads unchanged.

Every layer runs in the sm_100a library on channels-last bf16 activations: 3x3 convolutions as implicit GEMMs on tcgen05
(TMA box loads shifted per tap, out-of-bounds zero fill = padding; `ops.conv3x3`), skip connections fused into the GEMM
epilogue, Downsample2D as space-to-depth + a 2x2 convolution, GroupNorm / SiLU / up-sampling / softmax in `csrc/vae.cu`.
Any image size that is a multiple of 8 works: an (H x W) activation is stored with a row pitch that suits the convolution's
128-pixel tiles (ops.conv_pitch) and the tensor maps clip at W.  No tiling: vaehook's tiles exist to fit 8-24 GB GPUs; a
B200 holds the activations of a 2048 x 2048 decode (1 GB per 128-channel full-resolution tensor).  No PyTorch fallback.
"""
import inspect
import json
import math
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import nn

from . import ops
from .transformer_sd3 import CONFIG_NAME, SAFETENSORS_WEIGHTS_NAME, FrozenConfig


# ------------------------------------------------------------------------------------------------ parameter containers
class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6, affine=True)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6, affine=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None


class _Sampler(nn.Module):   # Downsample2D / Upsample2D: one 3x3 convolution named `conv`
    def __init__(self, c, stride):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=stride, padding=0 if stride == 2 else 1)


class Attention(nn.Module):   # single head, head_dim = channels (the VAE mid-block attention)
    def __init__(self, c, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6, affine=True)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])


class UNetMidBlock2D(nn.Module):
    def __init__(self, c, groups, attention):
        super().__init__()
        self.attentions = nn.ModuleList([Attention(c, groups)] if attention else [])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, groups), ResnetBlock2D(c, c, groups)])


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, groups) for j in range(layers)])
        self.downsamplers = nn.ModuleList([_Sampler(cout, 2)]) if add_downsample else None


class UpDecoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_upsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, groups) for j in range(layers)])
        self.upsamplers = nn.ModuleList([_Sampler(cout, 1)]) if add_upsample else None


class Encoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, groups, mid_attention):
        super().__init__()
        ch = list(block_out_channels)
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        blocks, cin = [], ch[0]
        for i, cout in enumerate(ch):
            blocks.append(DownEncoderBlock2D(cin, cout, layers_per_block, groups, i != len(ch) - 1))
            cin = cout
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = UNetMidBlock2D(ch[-1], groups, mid_attention)
        self.conv_norm_out = nn.GroupNorm(groups, ch[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[-1], 2 * out_channels, 3, padding=1)      # double_z: mean | logvar


class Decoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, groups, mid_attention):
        super().__init__()
        rev = list(block_out_channels)[::-1]
        self.conv_in = nn.Conv2d(in_channels, rev[0], 3, padding=1)
        self.mid_block = UNetMidBlock2D(rev[0], groups, mid_attention)
        blocks, cin = [], rev[0]
        for i, cout in enumerate(rev):
            blocks.append(UpDecoderBlock2D(cin, cout, layers_per_block + 1, groups, i != len(rev) - 1))
            cin = cout
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(groups, rev[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(rev[-1], out_channels, 3, padding=1)


# ------------------------------------------------------------------------------------------------ outputs
class DiagonalGaussianDistribution:
    """diffusers' posterior object: `.sample(generator)`, `.mode()`, `.mean`, `.logvar`, `.std`, `.var`."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device, dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self) -> torch.Tensor:
        return self.mean


@dataclass
class AutoencoderKLOutput:
    latent_dist: DiagonalGaussianDistribution


@dataclass
class DecoderOutput:
    sample: torch.Tensor


def _pad64(c):
    return (c + 63) // 64 * 64


def _pad32(c):
    return (c + 31) // 32 * 32


def pack_conv3x3(weight, bias):
    """Conv2d weight [N, C, 3, 3] -> implicit-GEMM matrix [pad32(N), 9 * pad64(C)] with K index = tap * C_pad + c,
    tap = dy * 3 + dx (dit4sr_gemm_params.conv_*: conv_kw 3, conv_taps 9, conv_org -1), zero rows / columns as padding."""
    n, c = weight.shape[:2]
    cp, npad = _pad64(c), _pad32(n)
    wk = weight.new_zeros(npad, 9, cp)
    wk[:n, :, :c] = weight.permute(0, 2, 3, 1).reshape(n, 9, c)
    b = bias.new_zeros(npad)
    b[:n] = bias
    return wk.reshape(npad, 9 * cp), b


def pack_conv_down(weight, bias):
    """Downsample2D: F.pad(x, (0, 1, 0, 1)) + Conv2d(3x3, stride 2) == a 2x2 stride-1 convolution (taps (oy, ox) at
    offsets 0 / +1, zero beyond the image) over the space-to-depth image with channel (py*2+px)*C + c, whose weight is
    w[:, :, 2*oy+py, 2*ox+px] (zero where that index would be 3).  Returns [N, 4 taps * 4C] (conv_kw 2, conv_taps 4,
    conv_org 0)."""
    n, c = weight.shape[:2]
    wk = weight.new_zeros(n, 2, 2, 2, 2, c)       # [n, oy, ox, py, px, c]
    for oy in range(2):
        for ox in range(2):
            for py in range(2):
                for px in range(2):
                    if 2 * oy + py <= 2 and 2 * ox + px <= 2:
                        wk[:, oy, ox, py, px] = weight[:, :, 2 * oy + py, 2 * ox + px]
    return wk.reshape(n, 16 * c), bias.clone()


# ------------------------------------------------------------------------------------------------ the model
class AutoencoderKL(nn.Module):
    config_name = CONFIG_NAME

    def __init__(self, in_channels: int = 3, out_channels: int = 3,
                 down_block_types: Tuple[str, ...] = ("DownEncoderBlock2D",) * 4,
                 up_block_types: Tuple[str, ...] = ("UpDecoderBlock2D",) * 4,
                 block_out_channels: Tuple[int, ...] = (128, 256, 512, 512), layers_per_block: int = 2,
                 act_fn: str = "silu", latent_channels: int = 16, norm_num_groups: int = 32, sample_size: int = 1024,
                 scaling_factor: float = 1.5305, shift_factor: Optional[float] = 0.0609, force_upcast: bool = True,
                 use_quant_conv: bool = False, use_post_quant_conv: bool = False, mid_block_add_attention: bool = True):
        super().__init__()
        if act_fn != "silu" or use_quant_conv or use_post_quant_conv:
            raise NotImplementedError("dit4sr_b200.AutoencoderKL implements the SD3 VAE flavour (silu, no quant convs)")
        if any(c % 64 for c in block_out_channels):
            raise NotImplementedError("block_out_channels must be multiples of 64 (tcgen05 K blocks)")
        cfg = dict(in_channels=in_channels, out_channels=out_channels, down_block_types=tuple(down_block_types),
                   up_block_types=tuple(up_block_types), block_out_channels=tuple(block_out_channels),
                   layers_per_block=layers_per_block, act_fn=act_fn, latent_channels=latent_channels,
                   norm_num_groups=norm_num_groups, sample_size=sample_size, scaling_factor=scaling_factor,
                   shift_factor=shift_factor, force_upcast=force_upcast, use_quant_conv=use_quant_conv,
                   use_post_quant_conv=use_post_quant_conv, mid_block_add_attention=mid_block_add_attention,
                   _class_name="AutoencoderKL")
        object.__setattr__(self, "_config", FrozenConfig(cfg))
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups,
                               mid_block_add_attention)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups,
                               mid_block_add_attention)
        self.quant_conv = None
        self.post_quant_conv = None
        self._packed = None
        # CUDA graphs of encode / decode per input shape (the ~140 launches of a decode are launch-paced at small sizes)
        self.use_cuda_graph = os.environ.get("DIT4SR_VAE_GRAPH", "1") != "0"
        self._graphs = {}

    # ---- config / io surface
    @property
    def config(self) -> FrozenConfig:
        return self._config

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__).parameters
        return cls(**{k: v for k, v in dict(config).items() if k in sig})

    @classmethod
    def from_pretrained(cls, path, subfolder=None, **kwargs):
        """Local directory with config.json + diffusion_pytorch_model.safetensors (diffusers layout; no hub access)."""
        import safetensors.torch
        if subfolder is not None:
            path = os.path.join(path, subfolder)
        with open(os.path.join(path, CONFIG_NAME)) as f:
            model = cls.from_config(json.load(f))
        model.load_state_dict(safetensors.torch.load_file(os.path.join(path, SAFETENSORS_WEIGHTS_NAME), device="cpu"))
        return model

    def save_pretrained(self, save_directory, **kwargs):
        from safetensors.torch import save_file
        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, CONFIG_NAME), "w") as f:
            json.dump({k: (list(v) if isinstance(v, tuple) else v) for k, v in self.config.items()}, f, indent=2)
        save_file({k: v.contiguous() for k, v in self.state_dict().items()}, os.path.join(save_directory, SAFETENSORS_WEIGHTS_NAME))

    def _apply(self, fn, *args, **kwargs):
        self._packed = None
        self._graphs = {}
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._packed = None
        self._graphs = {}
        return super().load_state_dict(*args, **kwargs)

    GRAPH_MAX_PIXELS = 1024 * 1024     # image pixels per batch entry times batch: beyond this the kernels are long enough
    GRAPH_MAX_CACHED = 4               # captured graphs keep their activations (about 2.9 GiB for a 1024px decode)

    @torch.no_grad()
    def _graphed(self, kind, fn, x, pixels):
        """fn(x) replayed from a CUDA graph captured per (kind, shape, dtype, device): same kernels, same results.  Large
        images (and anything that fails to capture) are launched directly."""
        if not self.use_cuda_graph or pixels > self.GRAPH_MAX_PIXELS or torch.cuda.is_current_stream_capturing():
            return fn(x)
        key = (kind, tuple(x.shape), x.dtype, x.device)
        ent = self._graphs.get(key)
        if ent is None:
            static_in = x.clone()
            fn(static_in)                                   # warm-up outside capture: packs weights, function attributes
            torch.cuda.synchronize(x.device)
            try:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    static_out = fn(static_in)
            except Exception as e:   # noqa: BLE001  (capture not possible here: keep launching directly)
                import warnings
                warnings.warn(f"dit4sr_b200.AutoencoderKL: CUDA graph capture failed ({e}); using direct launches")
                self.use_cuda_graph = False
                return fn(x)
            if len(self._graphs) >= self.GRAPH_MAX_CACHED:
                self._graphs.pop(next(iter(self._graphs)))
            ent = self._graphs[key] = (graph, static_in, static_out)
        graph, static_in, static_out = ent
        static_in.copy_(x)
        graph.replay()
        return static_out.clone()

    # tiling / slicing switches of diffusers' class: accepted, nothing to switch (the untiled network fits)
    def enable_tiling(self, *a, **k): pass
    def disable_tiling(self): pass
    def enable_slicing(self): pass
    def disable_slicing(self): pass

    # ---- weight packing (once per parameter state): bf16, GEMM layouts, channel padding
    def _pack(self):
        if self._packed is not None:
            return self._packed
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("dit4sr_b200.AutoencoderKL runs on a CUDA (sm_100a) device only: move the module with .to('cuda')")
        P = {}

        def conv3(name, conv):
            wk, bias = pack_conv3x3(conv.weight.detach().float().to(dev), conv.bias.detach().float().to(dev))
            P[name] = (wk.to(torch.bfloat16).contiguous(), bias)

        def conv_down(name, conv):
            wk, bias = pack_conv_down(conv.weight.detach().float().to(dev), conv.bias.detach().float().to(dev))
            P[name] = (wk.to(torch.bfloat16).contiguous(), bias)

        def lin(name, w, b):
            P[name] = (w.detach().to(torch.bfloat16).contiguous(), b.detach().float().contiguous())

        def norm(name, gn):
            P[name] = (gn.weight.detach().float().contiguous(), gn.bias.detach().float().contiguous())

        def resnet(name, r):
            norm(name + "norm1", r.norm1)
            conv3(name + "conv1", r.conv1)
            norm(name + "norm2", r.norm2)
            conv3(name + "conv2", r.conv2)
            if r.conv_shortcut is not None:
                lin(name + "conv_shortcut", r.conv_shortcut.weight.flatten(1), r.conv_shortcut.bias)

        def mid(name, m):
            resnet(name + "resnets.0.", m.resnets[0])
            resnet(name + "resnets.1.", m.resnets[1])
            if len(m.attentions):
                a = m.attentions[0]
                norm(name + "attn.norm", a.group_norm)
                lin(name + "attn.qk", torch.cat([a.to_q.weight, a.to_k.weight], 0), torch.cat([a.to_q.bias, a.to_k.bias], 0))
                P[name + "attn.v"] = a.to_v.weight.detach().to(torch.bfloat16).contiguous()
                # softmax rows sum to one, so P (V + 1 b_v^T) = P V + b_v: the value bias moves into the output bias
                wo = a.to_out[0].weight.detach().float()
                lin(name + "attn.out", wo, a.to_out[0].bias.detach().float() + wo @ a.to_v.bias.detach().float())

        e, d = self.encoder, self.decoder
        conv3("enc.conv_in", e.conv_in)
        for i, blk in enumerate(e.down_blocks):
            for j, r in enumerate(blk.resnets):
                resnet(f"enc.down.{i}.{j}.", r)
            if blk.downsamplers is not None:
                conv_down(f"enc.down.{i}.ds", blk.downsamplers[0].conv)
        mid("enc.mid.", e.mid_block)
        norm("enc.norm_out", e.conv_norm_out)
        conv3("enc.conv_out", e.conv_out)
        conv3("dec.conv_in", d.conv_in)
        mid("dec.mid.", d.mid_block)
        for i, blk in enumerate(d.up_blocks):
            for j, r in enumerate(blk.resnets):
                resnet(f"dec.up.{i}.{j}.", r)
            if blk.upsamplers is not None:
                conv3(f"dec.up.{i}.us", blk.upsamplers[0].conv)
        norm("dec.norm_out", d.conv_norm_out)
        conv3("dec.conv_out", d.conv_out)
        self._packed = P
        return P

    # ---- layers on channels-last activations [B, H*P, C]: the skip stream x is fp32, GEMM operands are bf16.  P is the row
    # pitch of the (H x W) image (ops.conv_pitch: divides 128 or a multiple of it; pad pixels are never read as data: the
    # convolutions' tensor maps clip at W, GroupNorm skips them, attention runs on the compacted tokens)
    def _gn(self, x, key, silu, W, P):
        g, b = self._packed[key]
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
        ops.groupnorm_nhwc(x, g, b, out, groups=self.config.norm_num_groups, eps=1e-6, silu=silu, width=W, pitch=P)
        return out

    def _conv(self, x, H, W, P, key, dtype=torch.float32):
        w, b = self._packed[key]
        out = torch.empty(x.shape[0], H * P, w.shape[0], device=x.device, dtype=dtype)
        ops.conv3x3(x, H, P, w, b, out, width=W)
        return out

    def _resnet(self, x, H, W, P, key):
        """x (fp32) -> shortcut(x) + conv2(silu(norm2(conv1(silu(norm1(x)))))); the sum is formed in place in the fp32
        stream by the conv2 GEMM's epilogue (TMA reduce-adds)."""
        Pk = self._packed
        h = self._gn(x, key + "norm1", True, W, P)
        h = self._conv(h, H, W, P, key + "conv1", torch.bfloat16)
        h = self._gn(h, key + "norm2", True, W, P)
        if key + "conv_shortcut" in Pk:
            w, b = Pk[key + "conv_shortcut"]
            xb = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
            ops.cast(x, xb)
            x = torch.empty(x.shape[0], H * P, w.shape[0], device=x.device, dtype=torch.float32)
            ops.gemm(xb, w, b, x)
        w, b = Pk[key + "conv2"]
        ops.conv3x3(h, H, P, w, b, x, resid=x, width=W)
        return x

    def _attention(self, x, H, W, P, key):
        """x += to_out(softmax(q k^T / sqrt(C)) v) with one head of head_dim C: scores and P V are plain tcgen05 GEMMs,
        the fp32 score rows (chunked to <= 1 GB) are normalised into bf16 probabilities by a row kernel.  A pitched image
        is compacted first (pad pixels must not act as keys) and written back afterwards (strided device copies)."""
        Pk = self._packed
        B, _, Cc = x.shape
        xp = None
        if P != W:
            xp = x
            x = x.view(B, H, P, Cc)[:, :, :W].reshape(B, H * W, Cc)
        S = H * W
        Sp = (S + 31) // 32 * 32           # key / token padding to the GEMM's granularity (pad keys get probability 0)
        g, bt = Pk[key + "attn.norm"]
        h = torch.zeros(B, Sp, Cc, device=x.device, dtype=torch.bfloat16) if Sp != S else \
            torch.empty(B, S, Cc, device=x.device, dtype=torch.bfloat16)
        ops.groupnorm_nhwc(x, g, bt, h[:, :S], groups=self.config.norm_num_groups, eps=1e-6, silu=False)
        wqk, bqk = Pk[key + "attn.qk"]
        qk = torch.empty(B, Sp, 2 * Cc, device=x.device, dtype=torch.bfloat16)
        ops.gemm(h, wqk, bqk, qk)
        wo, bo = Pk[key + "attn.out"]
        rows = max(128, min(S, (1 << 28) // Sp // 128 * 128))
        scores = torch.empty(rows, Sp, device=x.device, dtype=torch.float32)
        probs = torch.empty(rows, Sp, device=x.device, dtype=torch.bfloat16)
        vt = torch.empty(Cc, Sp, device=x.device, dtype=torch.bfloat16)
        o = torch.empty(S, Cc, device=x.device, dtype=torch.bfloat16)
        for b in range(B):
            ops.gemm(Pk[key + "attn.v"], h[b], None, vt)                 # V^T [C, S] = W_v h^T (bias folded into to_out)
            for r0 in range(0, S, rows):
                r1 = min(S, r0 + rows)
                ops.gemm(qk[b, r0:r1, :Cc], qk[b, :, Cc:], None, scores[:r1 - r0])
                ops.softmax_rows_to_bf16(scores[:r1 - r0], probs[:r1 - r0], 1.0 / math.sqrt(Cc), valid_cols=S)
                ops.gemm(probs[:r1 - r0], vt, None, o[r0:r1])
            ops.gemm(o, wo, bo, x[b], ops.EPI_GATE_RESID)                # in place: x[b] += o W_o^T + bias
        if xp is not None:
            xp.view(B, H, P, Cc)[:, :, :W].copy_(x.view(B, H, W, Cc))
            x = xp
        return x

    def _mid(self, x, H, W, P, key, attention):
        x = self._resnet(x, H, W, P, key + "resnets.0.")
        if attention:
            x = self._attention(x, H, W, P, key)
        return self._resnet(x, H, W, P, key + "resnets.1.")

    # ---- public API
    def _check_input(self, t, channels):
        if not t.is_cuda:
            raise RuntimeError("dit4sr_b200.AutoencoderKL has no CPU path: pass CUDA tensors")
        if t.dim() != 4 or t.shape[1] != channels:
            raise ValueError(f"expected [B, {channels}, H, W], got {tuple(t.shape)}")

    @torch.no_grad()
    def _encode_moments(self, x):
        cfg = self.config
        self._check_input(x, cfg.in_channels)
        self._pack()
        B, _, H, W = x.shape
        n = len(cfg.block_out_channels)
        if H % (1 << (n - 1)) or W % (1 << (n - 1)):
            raise ValueError("image sides must be multiples of the VAE scale factor")
        P = ops.conv_pitch(W)
        h = ops.nchw_to_nhwc_bf16(x, _pad64(cfg.in_channels), pitch=P)
        h = self._conv(h, H, W, P, "enc.conv_in")
        for i, blk in enumerate(self.encoder.down_blocks):
            for j in range(len(blk.resnets)):
                h = self._resnet(h, H, W, P, f"enc.down.{i}.{j}.")
            if blk.downsamplers is not None:
                Cc = h.shape[2]
                P2 = ops.conv_pitch(W // 2)
                s2d = torch.empty(B, (H // 2) * P2, 4 * Cc, device=h.device, dtype=torch.bfloat16)
                ops.space_to_depth2_nhwc(h, H, W, s2d, pitch_in=P, pitch_out=P2)
                H, W, P = H // 2, W // 2, P2
                w, b = self._packed[f"enc.down.{i}.ds"]
                h = torch.empty(B, H * P, w.shape[0], device=h.device, dtype=torch.float32)
                ops.conv2x2_s2d(s2d, H, P, w, b, h, width=W)
        h = self._mid(h, H, W, P, "enc.mid.", cfg.mid_block_add_attention)
        h = self._gn(h, "enc.norm_out", True, W, P)
        m = self._conv(h, H, W, P, "enc.conv_out")
        return ops.nhwc_to_nchw_f32(m, 2 * cfg.latent_channels, H, W, pitch=P)

    def encode(self, x: torch.Tensor, return_dict: bool = True):
        self._check_input(x, self.config.in_channels)
        moments = self._graphed("enc", self._encode_moments, x, x.shape[0] * x.shape[2] * x.shape[3])
        posterior = DiagonalGaussianDistribution(moments.to(x.dtype if x.is_floating_point() else torch.float32))
        return AutoencoderKLOutput(latent_dist=posterior) if return_dict else (posterior,)

    @torch.no_grad()
    def _decode(self, z):
        cfg = self.config
        self._check_input(z, cfg.latent_channels)
        self._pack()
        B, _, H, W = z.shape
        P = ops.conv_pitch(W)
        h = ops.nchw_to_nhwc_bf16(z, _pad64(cfg.latent_channels), pitch=P)
        h = self._conv(h, H, W, P, "dec.conv_in")
        h = self._mid(h, H, W, P, "dec.mid.", cfg.mid_block_add_attention)
        for i, blk in enumerate(self.decoder.up_blocks):
            for j in range(len(blk.resnets)):
                h = self._resnet(h, H, W, P, f"dec.up.{i}.{j}.")
            if blk.upsamplers is not None:
                P2 = ops.conv_pitch(2 * W)
                up = torch.empty(B, 2 * H * P2, h.shape[2], device=h.device, dtype=torch.bfloat16)
                ops.upsample2x_nhwc(h, H, W, up, pitch_in=P, pitch_out=P2)
                H, W, P = 2 * H, 2 * W, P2
                h = self._conv(up, H, W, P, f"dec.up.{i}.us")
        h = self._gn(h, "dec.norm_out", True, W, P)
        img = self._conv(h, H, W, P, "dec.conv_out")
        return ops.nhwc_to_nchw_f32(img, cfg.out_channels, H, W, pitch=P)

    def decode(self, z: torch.Tensor, return_dict: bool = True, generator=None):
        self._check_input(z, self.config.latent_channels)
        f = 1 << (len(self.config.block_out_channels) - 1)
        dec = self._graphed("dec", self._decode, z, z.shape[0] * z.shape[2] * z.shape[3] * f * f).to(z.dtype)
        return DecoderOutput(sample=dec) if return_dict else (dec,)

    def forward(self, sample: torch.Tensor, sample_posterior: bool = False, return_dict: bool = True,
                generator: Optional[torch.Generator] = None):
        posterior = self.encode(sample).latent_dist
        z = posterior.sample(generator=generator) if sample_posterior else posterior.mode()
        dec = self.decode(z).sample
        return DecoderOutput(sample=dec) if return_dict else (dec,)

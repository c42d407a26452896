"""VAE: the oracle restatement, the host-side weight packing (CPU), and the CUDA path against the oracle (GPU).

Parity is UNPINNED for this row (oracle/vae_oracle.py header): diffusers' AutoencoderKL is absent, so the oracle restates its
layer semantics; the layer ORDER and names are those the reference's own tiled-VAE hook walks (utils/vaehook.py:279-358)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import vae_oracle as VO

SMALL = dict(block_out_channels=(64, 128, 128), layers_per_block=1)


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_state_dict_layout_matches_the_oracle_manifest():
    import dit4sr_b200
    for over in ({}, SMALL):
        cfg = VO.sd3_vae_config(**over)
        m = dit4sr_b200.AutoencoderKL(block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block)
        mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert mine == VO.state_dict_shapes(cfg)
    # the SD3 VAE: 83.8 M parameters (34.2 M encoder + 49.5 M decoder), no quant convs
    n = sum(math.prod(s) for s in VO.state_dict_shapes(VO.sd3_vae_config()).values())
    assert 83_000_000 < n < 84_500_000, n
    assert m.config.scaling_factor == 1.5305 and m.config.shift_factor == 0.0609
    # the names the reference's tiled-VAE hook walks (utils/vaehook.py:279-358)
    m = dit4sr_b200.AutoencoderKL()
    assert len(m.decoder.up_blocks) == 4 and len(m.decoder.up_blocks[0].resnets) == 3 and m.decoder.up_blocks[3].upsamplers is None
    assert len(m.encoder.down_blocks[0].resnets) == 2 and m.encoder.down_blocks[3].downsamplers is None
    assert hasattr(m.decoder.mid_block.attentions[0], "to_q") and hasattr(m.decoder, "conv_norm_out")
    with pytest.raises(RuntimeError, match="CUDA"):
        m.decode(torch.zeros(1, 16, 8, 8))


def _implicit_gemm_cpu(x_nhwc, wk, kw, taps, org):
    """The arithmetic of the conv mode of dit4sr_gemm_bf16 on the CPU: K-block tap t reads the image shifted by
    (t // kw + org, t % kw + org), zero outside."""
    B, H, W, C = x_nhwc.shape
    cols = []
    for t in range(taps):
        dy, dx = t // kw + org, t % kw + org
        sh = torch.zeros_like(x_nhwc)
        ys, ye = max(0, -dy), min(H, H - dy)
        xs, xe = max(0, -dx), min(W, W - dx)
        sh[:, ys:ye, xs:xe] = x_nhwc[:, ys + dy:ye + dy, xs + dx:xe + dx]
        cols.append(sh)
    a = torch.cat(cols, dim=-1).reshape(B, H * W, taps * C)
    return a @ wk.t()


def test_conv_weight_packing_is_the_convolution():
    from dit4sr_b200.vae import pack_conv3x3, pack_conv_down
    torch.manual_seed(0)
    x = torch.randn(2, 5, 6, 8)                                  # NCHW, 5 channels (padded to 64 by the packer)
    w, b = torch.randn(7, 5, 3, 3), torch.randn(7)
    wk, bk = pack_conv3x3(w, b)
    assert wk.shape == (32, 9 * 64) and bk.shape == (32,)
    xp = torch.zeros(2, 6, 8, 64)
    xp[..., :5] = x.permute(0, 2, 3, 1)
    got = (_implicit_gemm_cpu(xp, wk, 3, 9, -1) + bk)[..., :7].reshape(2, 6, 8, 7).permute(0, 3, 1, 2)
    assert rel(got, F.conv2d(x, w, b, padding=1)) < 1e-6
    # Downsample2D = pad (0,1,0,1) + 3x3 stride 2  ==  space-to-depth + 2x2 taps at offsets 0 / +1
    x = torch.randn(2, 4, 6, 8)
    w, b = torch.randn(3, 4, 3, 3), torch.randn(3)
    wk, bk = pack_conv_down(w, b)
    assert wk.shape == (3, 16 * 4)
    nhwc = x.permute(0, 2, 3, 1)
    s2d = torch.stack([nhwc[:, py::2, px::2] for py in range(2) for px in range(2)], dim=3).reshape(2, 3, 4, 16)
    got = (_implicit_gemm_cpu(s2d, wk, 2, 4, 0) + bk).reshape(2, 3, 4, 3).permute(0, 3, 1, 2)
    assert rel(got, F.conv2d(F.pad(x, (0, 1, 0, 1)), w, b, stride=2)) < 1e-6


def test_oracle_against_independent_torch_modules():
    """The functional oracle against nn.Module compositions written separately (GroupNorm / Conv2d / SDPA modules)."""
    torch.manual_seed(0)
    cfg = VO.sd3_vae_config(**SMALL)
    sd = VO.init_state_dict(cfg, seed=3)
    x = torch.randn(1, 64, 8, 8)
    pre = "decoder.mid_block.resnets.0."
    n1, n2 = torch.nn.GroupNorm(32, 64, eps=1e-6), torch.nn.GroupNorm(32, 64, eps=1e-6)
    c1, c2 = torch.nn.Conv2d(64, 64, 3, padding=1), torch.nn.Conv2d(64, 64, 3, padding=1)
    cfg2 = VO.sd3_vae_config(block_out_channels=(64, 64), layers_per_block=1)
    sd2 = VO.init_state_dict(cfg2, seed=4)
    for mod, k in ((n1, "norm1"), (c1, "conv1"), (n2, "norm2"), (c2, "conv2")):
        mod.load_state_dict({"weight": sd2[pre + k + ".weight"], "bias": sd2[pre + k + ".bias"]})
    with torch.no_grad():
        want = x + c2(F.silu(n2(c1(F.silu(n1(x))))))
        assert rel(VO._resnet(x, sd2, pre, 32), want) < 1e-6
        # attention: one head over H*W tokens with head_dim = C, through torch's SDPA
        a = "decoder.mid_block.attentions.0."
        h = F.group_norm(x, 32, sd2[a + "group_norm.weight"], sd2[a + "group_norm.bias"], eps=1e-6).flatten(2).transpose(1, 2)
        q, k, v = (F.linear(h, sd2[a + n + ".weight"], sd2[a + n + ".bias"]) for n in ("to_q", "to_k", "to_v"))
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = F.linear(o, sd2[a + "to_out.0.weight"], sd2[a + "to_out.0.bias"]).transpose(1, 2).reshape(1, 64, 8, 8)
        assert rel(VO._attention(x, sd2, a, 32), x + o) < 1e-6
        # shapes through the whole network and the posterior
        img = torch.randn(1, 3, 32, 32)
        mom = VO.encoder_forward(sd, cfg, img)
        assert mom.shape == (1, 32, 8, 8)
        assert VO.decoder_forward(sd, cfg, mom[:, :16]).shape == (1, 3, 32, 32)
        assert torch.equal(VO.gaussian_sample(mom), mom[:, :16])
        noise = torch.randn(1, 16, 8, 8)
        assert rel(VO.gaussian_sample(mom, noise), mom[:, :16] + torch.exp(0.5 * mom[:, 16:]) * noise) < 1e-6


# ----------------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def ops():
    from dit4sr_b200 import ops as o
    return o


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,C,N", [(1, 8, 8, 64, 64), (2, 16, 16, 128, 256), (1, 64, 64, 512, 512), (2, 128, 128, 64, 128),
                                       (1, 256, 256, 128, 32), (1, 32, 64, 256, 256), (3, 4, 32, 64, 96)])
def test_conv3x3_implicit_gemm(ops, B, H, W, C, N):
    """3x3 / pad 1 convolution on tcgen05 (TMA boxes shifted per tap, out-of-bounds zero fill) against F.conv2d on the same
    bf16-rounded operands, with and without the fused skip connection; single-CTA and CTA-pair schedules."""
    from dit4sr_b200.vae import pack_conv3x3
    torch.manual_seed(0)
    x = torch.randn(B, C, H, W, device="cuda").bfloat16()
    w = (torch.randn(N, C, 3, 3, device="cuda") / math.sqrt(9 * C)).bfloat16()
    b = torch.randn(N, device="cuda")
    res = torch.randn(B, H * W, N, device="cuda").bfloat16()
    wk, bk = pack_conv3x3(w.float(), b)
    xn = x.permute(0, 2, 3, 1).reshape(B, H * W, C).contiguous()
    ref = F.conv2d(x.float(), w.float(), b, padding=1).permute(0, 2, 3, 1).reshape(B, H * W, N)
    out = torch.empty(B, H * W, wk.shape[0], device="cuda", dtype=torch.bfloat16)
    ops.conv3x3(xn, H, W, wk.bfloat16(), bk, out)
    assert rel(out[..., :N], ref) < 4e-3
    ops.conv3x3(xn, H, W, wk.bfloat16(), bk, out, resid_bf16=torch.nn.functional.pad(res, (0, wk.shape[0] - N)).contiguous())
    assert rel(out[..., :N], ref + res.float()) < 4e-3
    o32 = torch.empty(B, H * W, wk.shape[0], device="cuda", dtype=torch.float32)
    ops.conv3x3(xn, H, W, wk.bfloat16(), bk, o32)
    assert rel(o32[..., :N], ref) < 1e-4      # fp32 accumulate and output of bf16 products
    # the fp32 skip stream of the residual blocks: out-of-place, and in place (TMA reduce-adds into the stream)
    skip = torch.randn(B, H * W, wk.shape[0], device="cuda")
    ops.conv3x3(xn, H, W, wk.bfloat16(), bk, o32, resid=skip)
    assert rel(o32[..., :N], ref + skip[..., :N]) < 1e-4
    stream = skip.clone()
    ops.conv3x3(xn, H, W, wk.bfloat16(), bk, stream, resid=stream)
    assert rel(stream[..., :N], ref + skip[..., :N]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,C", [(1, 16, 16, 64), (2, 64, 64, 128), (1, 256, 256, 128)])
def test_downsample_conv_as_space_to_depth(ops, B, H, W, C):
    from dit4sr_b200.vae import pack_conv_down
    torch.manual_seed(0)
    x = torch.randn(B, C, H, W, device="cuda").bfloat16()
    w = (torch.randn(C, C, 3, 3, device="cuda") / math.sqrt(9 * C)).bfloat16()
    b = torch.randn(C, device="cuda")
    wk, bk = pack_conv_down(w.float(), b)
    xn = x.permute(0, 2, 3, 1).reshape(B, H * W, C).contiguous()
    s2d = torch.empty(B, H * W // 4, 4 * C, device="cuda", dtype=torch.bfloat16)
    ops.space_to_depth2_nhwc(xn, H, W, s2d)
    s2d_f = torch.empty_like(s2d)
    ops.space_to_depth2_nhwc(xn.float(), H, W, s2d_f)       # the fp32 skip stream as input
    assert torch.equal(s2d, s2d_f)
    out = torch.empty(B, H * W // 4, C, device="cuda", dtype=torch.bfloat16)
    ops.conv2x2_s2d(s2d, H // 2, W // 2, wk.bfloat16(), bk, out)
    ref = F.conv2d(F.pad(x.float(), (0, 1, 0, 1)), w.float(), b, stride=2).permute(0, 2, 3, 1).reshape(B, -1, C)
    assert rel(out, ref) < 4e-3


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("B,P,C,silu", [(1, 64, 64, False), (2, 4096, 512, True), (1, 65536, 128, True), (3, 1000, 256, False)])
def test_groupnorm_silu_nhwc(ops, B, P, C, silu, dtype):
    torch.manual_seed(0)
    x = (torch.randn(B, P, C, device="cuda") * 2.0 + 0.7).to(dtype)
    g, b = 1 + 0.1 * torch.randn(C, device="cuda"), 0.1 * torch.randn(C, device="cuda")
    out = torch.empty(B, P, C, device="cuda", dtype=torch.bfloat16)
    ops.groupnorm_nhwc(x, g, b, out, groups=32, silu=silu)
    again = torch.empty_like(out)
    ops.groupnorm_nhwc(x, g, b, again, groups=32, silu=silu)
    assert torch.equal(out, again)      # fixed summation order: bit-reproducible
    ref = F.group_norm(x.float().transpose(1, 2), 32, g, b, eps=1e-6)
    ref = (F.silu(ref) if silu else ref).transpose(1, 2)
    assert rel(out, ref) < 4e-3


@pytest.mark.gpu
def test_upsample_softmax_and_layout_kernels(ops):
    torch.manual_seed(0)
    x = torch.randn(2, 6 * 10, 64, device="cuda").bfloat16()
    up = torch.empty(2, 4 * 60, 64, device="cuda", dtype=torch.bfloat16)
    ops.upsample2x_nhwc(x, 6, 10, up)
    ref = F.interpolate(x.float().reshape(2, 6, 10, 64).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    assert torch.equal(up.float().reshape(2, 12, 20, 64).permute(0, 3, 1, 2), ref)
    up_f = torch.empty_like(up)
    ops.upsample2x_nhwc(x.float(), 6, 10, up_f)
    assert torch.equal(up, up_f)
    s = torch.randn(300, 4096, device="cuda") * 30
    p = torch.empty(300, 4096, device="cuda", dtype=torch.bfloat16)
    ops.softmax_rows_to_bf16(s, p, 0.05)
    assert rel(p, torch.softmax(s * 0.05, dim=-1)) < 4e-3
    ops.softmax_rows_to_bf16(s, p, 0.05, valid_cols=4001)          # key padding: columns >= 4001 get probability 0
    assert rel(p[:, :4001], torch.softmax(s[:, :4001] * 0.05, dim=-1)) < 4e-3 and float(p[:, 4001:].abs().max()) == 0.0
    img = torch.randn(2, 3, 8, 12, device="cuda")
    n = ops.nchw_to_nhwc_bf16(img, 64, a=0.5, b=0.25)
    assert n.shape == (2, 96, 64) and float(n[..., 3:].abs().max()) == 0.0
    assert rel(n[..., :3], (img * 0.5 + 0.25).permute(0, 2, 3, 1).reshape(2, 96, 3)) < 4e-3
    f = torch.randn(2, 96, 32, device="cuda")
    assert torch.equal(ops.nhwc_to_nchw_f32(f, 3, 8, 12), f[..., :3].reshape(2, 8, 12, 3).permute(0, 3, 1, 2))


def _build(cfg, seed):
    import dit4sr_b200
    sd = VO.init_state_dict(cfg, seed=seed)
    # fp32 master weights cast to bf16 for BOTH sides (the product keeps GEMM operands in bf16)
    sd = {k: (v.bfloat16().float() if v.dim() > 1 else v) for k, v in sd.items()}
    m = dit4sr_b200.AutoencoderKL(block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block)
    m.load_state_dict(sd)
    return m.to("cuda").eval(), sd


@pytest.mark.gpu
@pytest.mark.parametrize("over,size,batch", [(SMALL, 64, 2), ({}, 128, 1), ({}, 256, 2)])
def test_vae_decode_and_encode_against_the_oracle(over, size, batch):
    """Whole decoder and encoder (bf16 activations, fp32 accumulation) against the fp32 CPU oracle on the same random
    weights: decode of pipeline-scaled latents (pipeline_dit4sr.py:1602-1604) and the posterior moments / sample of an
    image (:1001-1002).  Tolerance 1e-2 relative L2 (the bf16 gate of the denoiser)."""
    torch.manual_seed(1)
    cfg = VO.sd3_vae_config(**over)
    m, sd = _build(cfg, seed=0)
    f = 2 ** (len(cfg.block_out_channels) - 1)
    lat = torch.randn(batch, 16, size // f, size // f)
    with torch.no_grad():
        z = lat / cfg.scaling_factor + cfg.shift_factor
        got = m.decode(z.cuda(), return_dict=False)[0].float().cpu()
        ref = VO.decoder_forward(sd, cfg, z)
    assert got.shape == ref.shape == (batch, 3, size, size)
    e = rel(got, ref)
    print(f"decode rel-L2 {e:.3e}")
    assert e < 1e-2
    img = torch.rand(batch, 3, size, size) * 2 - 1
    with torch.no_grad():
        post = m.encode(img.cuda()).latent_dist
        mom = VO.encoder_forward(sd, cfg, img)
    e = rel(post.parameters.cpu(), mom)
    print(f"encode moments rel-L2 {e:.3e}")
    assert e < 1e-2
    g = torch.Generator(device="cuda").manual_seed(5)
    s1 = post.sample(generator=g)
    g.manual_seed(5)
    noise = torch.randn(post.mean.shape, generator=g, device="cuda")
    assert torch.allclose(s1, post.mean + post.std * noise)
    assert rel(post.mode().cpu(), mom[:, :16]) < 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("over,hh,ww", [(SMALL, 96, 160), (SMALL, 40, 24), (SMALL, 12, 20), ({}, 192, 320), ({}, 136, 1032)])
def test_vae_arbitrary_image_sizes_pitched_rows(over, hh, ww):
    """Image sides that are multiples of 8 but not powers of two: every level is stored with a row pitch that suits the
    convolution's 128-pixel tiles (ops.conv_pitch), the tensor maps clip at the real width, GroupNorm skips the pad pixels
    and the mid-block attention runs on the compacted tokens.  Same gates as the square cases."""
    from dit4sr_b200 import ops
    torch.manual_seed(2)
    cfg = VO.sd3_vae_config(**over)
    m, sd = _build(cfg, seed=1)
    f = 2 ** (len(cfg.block_out_channels) - 1)
    assert ops.conv_pitch(ww // f) != ww // f or ops.conv_pitch(ww) != ww      # at least one level is really pitched
    z = torch.randn(1, 16, hh // f, ww // f) / cfg.scaling_factor + cfg.shift_factor
    img = torch.rand(1, 3, hh, ww) * 2 - 1
    with torch.no_grad():
        got = m.decode(z.cuda(), return_dict=False)[0].float().cpu()
        ref = VO.decoder_forward(sd, cfg, z)
        assert got.shape == ref.shape == (1, 3, hh, ww)
        e = rel(got, ref)
        print(f"decode {hh}x{ww} rel-L2 {e:.3e}")
        assert e < 1e-2
        mom = m.encode(img.cuda()).latent_dist.parameters.cpu()
        ref = VO.encoder_forward(sd, cfg, img)
        assert mom.shape == ref.shape
        e = rel(mom, ref)
        print(f"encode {hh}x{ww} rel-L2 {e:.3e}")
        assert e < 1e-2


@pytest.mark.gpu
def test_pitched_layer_kernels(ops):
    """GroupNorm / upsample / space-to-depth / layout kernels and the 3x3 convolution on a pitched (H x W, pitch P) image
    against the dense ones: pad pixels never influence valid pixels."""
    from dit4sr_b200.vae import pack_conv3x3
    torch.manual_seed(0)
    B, H, W, C, N = 2, 6, 20, 64, 64
    P = ops.conv_pitch(W)
    assert P == 32
    dense = torch.randn(B, H, W, C, device="cuda")
    pitched = torch.full((B, H, P, C), 7.0, device="cuda")                   # pad pixels hold garbage on purpose
    pitched[:, :, :W] = dense
    g, b = 1 + 0.1 * torch.randn(C, device="cuda"), 0.1 * torch.randn(C, device="cuda")
    for dt in (torch.float32, torch.bfloat16):
        a = torch.empty(B, H * W, C, device="cuda", dtype=torch.bfloat16)
        c = torch.empty(B, H * P, C, device="cuda", dtype=torch.bfloat16)
        ops.groupnorm_nhwc(dense.to(dt).reshape(B, H * W, C), g, b, a, silu=True)
        ops.groupnorm_nhwc(pitched.to(dt).reshape(B, H * P, C), g, b, c, silu=True, width=W, pitch=P)
        assert rel(c.view(B, H, P, C)[:, :, :W], a.view(B, H, W, C)) < 2e-3
        up = torch.empty(B, 2 * H * 2 * W, C, device="cuda", dtype=torch.bfloat16)
        upp = torch.empty(B, 2 * H * 64, C, device="cuda", dtype=torch.bfloat16)
        ops.upsample2x_nhwc(dense.to(dt).reshape(B, H * W, C), H, W, up)
        ops.upsample2x_nhwc(pitched.to(dt).reshape(B, H * P, C), H, W, upp, pitch_in=P, pitch_out=64)
        assert torch.equal(upp.view(B, 2 * H, 64, C)[:, :, :2 * W], up.view(B, 2 * H, 2 * W, C))
        assert float(upp.view(B, 2 * H, 64, C)[:, :, 2 * W:].abs().max()) == 0.0
        sd_ = torch.empty(B, (H // 2) * (W // 2), 4 * C, device="cuda", dtype=torch.bfloat16)
        sdp = torch.empty(B, (H // 2) * 16, 4 * C, device="cuda", dtype=torch.bfloat16)
        ops.space_to_depth2_nhwc(dense.to(dt).reshape(B, H * W, C), H, W, sd_)
        ops.space_to_depth2_nhwc(pitched.to(dt).reshape(B, H * P, C), H, W, sdp, pitch_in=P, pitch_out=16)
        assert torch.equal(sdp.view(B, H // 2, 16, 4 * C)[:, :, :W // 2], sd_.view(B, H // 2, W // 2, 4 * C))
    w = (torch.randn(N, C, 3, 3, device="cuda") / 24).bfloat16()
    bias = torch.randn(N, device="cuda")
    wk, bk = pack_conv3x3(w.float(), bias)
    out = torch.empty(B, H * P, N, device="cuda", dtype=torch.float32)
    ops.conv3x3(pitched.bfloat16().reshape(B, H * P, C), H, P, wk.bfloat16(), bk, out, width=W)
    ref = F.conv2d(dense.bfloat16().float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1)
    assert rel(out.view(B, H, P, N)[:, :, :W], ref) < 1e-4
    img = torch.randn(B, 3, H, W, device="cuda")
    n = ops.nchw_to_nhwc_bf16(img, 64, pitch=P).view(B, H, P, 64)
    assert float(n[:, :, W:].abs().max()) == 0.0 and rel(n[:, :, :W, :3], img.permute(0, 2, 3, 1)) < 4e-3
    f32 = torch.randn(B, H * P, 32, device="cuda")
    assert torch.equal(ops.nhwc_to_nchw_f32(f32, 3, H, W, pitch=P), f32.view(B, H, P, 32)[:, :, :W, :3].permute(0, 3, 1, 2))


@pytest.mark.gpu
def test_vae_round_trip_is_deterministic_and_roundtrips_through_save(tmp_path):
    import dit4sr_b200
    cfg = VO.sd3_vae_config(**SMALL)
    m, _ = _build(cfg, seed=2)
    m.save_pretrained(tmp_path)
    m2 = dit4sr_b200.AutoencoderKL.from_pretrained(str(tmp_path)).to("cuda")
    assert m2.config.block_out_channels == (64, 128, 128)
    z = torch.randn(1, 16, 16, 16, device="cuda")
    a = m.decode(z).sample
    assert torch.equal(a, m.decode(z).sample) and torch.equal(a, m2.decode(z).sample)
    assert m(torch.rand(1, 3, 64, 64, device="cuda")).sample.shape == (1, 3, 64, 64)


@pytest.mark.gpu
def test_super_resolve_image_path_against_the_oracle():
    """LR image -> VAE encode -> 8-step loop of a 4-block full-width model -> VAE decode -> [0, 1] image, every stage on
    the CUDA library, against the same chain of CPU oracles (posterior mode instead of a sample so that both sides see the
    same control latent).  PSNR of the final image >= 35 dB (bf16 through two VAE passes and 8 denoiser calls)."""
    import dit4sr_b200
    from dit4sr_b200.sampler import super_resolve
    from oracle import dit4sr_oracle as O
    torch.manual_seed(0)
    vcfg = VO.sd3_vae_config()
    vae, vsd = _build(vcfg, seed=0)
    cfg = O.sd35_medium_config(num_layers=4, dual_attention_layers=(0, 1), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    m = dit4sr_b200.SD3Transformer2DModel(**vars(cfg))
    m.load_state_dict(sd)
    m = m.to("cuda", torch.bfloat16).eval()
    lr_img = torch.rand(1, 3, 128, 128) * 2 - 1
    noise, _, enc, pooled, _ = O.make_inputs(cfg, 1, 16, 16, 77, seed=1)
    bf = torch.bfloat16
    img = super_resolve(m, vae, lr_img.cuda(), enc.cuda().to(bf), pooled.cuda().to(bf), num_inference_steps=8,
                        latents=noise.cuda(), sample_posterior=False).float().cpu()
    with torch.no_grad():
        control = VO.encode_latents(vsd, vcfg, lr_img)
        final = O.sample_loop(sd, cfg, noise, control, enc, pooled, num_steps=8)
        ref = (VO.decode_latents(vsd, vcfg, final) / 2 + 0.5).clamp(0, 1)
    assert img.shape == ref.shape == (1, 3, 128, 128)
    mse = ((img - ref) ** 2).mean().item()
    psnr = 10 * math.log10(1.0 / mse)
    print(f"super_resolve image PSNR vs the oracle chain: {psnr:.1f} dB")
    assert psnr >= 35.0


@pytest.mark.gpu
def test_super_resolve_non_square_image_through_latent_tiles():
    """A 256x384 image (latent 32x48): VAE encode on pitched rows, two Euler steps whose velocities are blended from
    overlapping 24x24 latent tiles (pipeline_dit4sr.py:1088,1449-1558), VAE decode - against the same chain of CPU oracles."""
    import dit4sr_b200
    from dit4sr_b200.sampler import super_resolve
    from oracle import dit4sr_oracle as O
    torch.manual_seed(0)
    vcfg = VO.sd3_vae_config(**SMALL)
    vae, vsd = _build(vcfg, seed=0)
    f = 4
    cfg = O.sd35_medium_config(num_layers=2, dual_attention_layers=(0,), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    m = dit4sr_b200.SD3Transformer2DModel(**vars(cfg))
    m.load_state_dict(sd)
    m = m.to("cuda", torch.bfloat16).eval()
    lr_img = torch.rand(1, 3, 32 * f, 48 * f) * 2 - 1
    _, _, enc, pooled, _ = O.make_inputs(cfg, 1, 8, 8, 20, seed=1)
    noise = torch.randn(1, 16, 32, 48)
    bf = torch.bfloat16
    img = super_resolve(m, vae, lr_img.cuda(), enc.cuda().to(bf), pooled.cuda().to(bf), num_inference_steps=2,
                        latents=noise.cuda(), sample_posterior=False, latent_tiled_size=24,
                        latent_tiled_overlap=8).float().cpu()
    with torch.no_grad():
        control = VO.encode_latents(vsd, vcfg, lr_img)
        sig = O.flow_match_sigmas(2)
        x = noise.clone()
        for i in range(2):
            t = torch.full((1,), float(sig[i]) * 1000.0)

            def cpu_fn(xt, ct):
                n = xt.shape[0]
                return O.model_forward(sd, cfg, xt, ct, enc.expand(n, -1, -1), pooled.expand(n, -1), t.expand(n), as_written=False)

            x = x + (float(sig[i + 1]) - float(sig[i])) * O.tiled_velocity_as_written(cpu_fn, x, control, 24, 8)
        ref = (VO.decode_latents(vsd, vcfg, x) / 2 + 0.5).clamp(0, 1)
    assert img.shape == ref.shape == (1, 3, 32 * f, 48 * f)
    psnr = 10 * math.log10(1.0 / ((img - ref) ** 2).mean().item())
    print(f"tiled non-square super_resolve PSNR vs the oracle chain: {psnr:.1f} dB")
    assert psnr >= 35.0

"""GPU parity of the whole denoiser (drop-in module -> engine -> C-ABI kernels) against the oracle and the golden
vectors recorded from the reference's own code.

Gates (BASELINE.json north_star): per-step velocity rel-L2 <= 1e-2 for the bf16 product; final latents PSNR >= 40 dB
against the reference bf16 sampling loop (no VAE weights offline, so PSNR is taken on the final latents with
peak = max - min of the reference latents)."""
import json
import os

import pytest
import torch
from safetensors.torch import load_file

from oracle import dit4sr_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
VEL_TOL = 1e-2


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def build(cfg, sd, dtype=torch.bfloat16, ocr=False):
    import dit4sr_b200
    cls = dit4sr_b200.SD3Transformer2DModel_OCRBranch_OCR2HQ2OCR if ocr else dit4sr_b200.SD3Transformer2DModel
    m = cls(**vars(cfg))
    m.load_state_dict(sd)
    return m.to("cuda", dtype).requires_grad_(False).eval()


def run(m, inputs, dtype=torch.bfloat16):
    hs, lr, enc, pooled, t = inputs
    with torch.no_grad():
        y = m(hidden_states=hs.cuda().to(dtype), controlnet_image=lr.cuda().to(dtype),
              encoder_hidden_states=enc.cuda().to(dtype), pooled_projections=pooled.cuda().to(dtype),
              timestep=t.cuda(), return_dict=False)[0]
    return y.float().cpu()


META = json.load(open(os.path.join(GOLD, "reference_meta.json")))


@pytest.mark.parametrize("case", META["cases"], ids=lambda c: c["name"])
def test_tiny_model_matches_reference_golden(case):
    """The committed outputs of the reference's own model code (fp32) vs the bf16 CUDA path."""
    ref = load_file(os.path.join(GOLD, "reference_tiny.safetensors"))[case["name"] + ".velocity"]
    ocr = case.get("ocr", False)
    cfg = O.tiny_config(**case["cfg_kwargs"])
    sd = O.init_state_dict(cfg, seed=case["weight_seed"], ocr=ocr)
    inputs = O.make_inputs(cfg, case["batch"], case["height"], case["width"], case["text_tokens"],
                           seed=case["input_seed"], timestep=case["timestep"])
    y = run(build(cfg, sd, ocr=ocr), inputs)
    assert y.shape == ref.shape
    assert rel(y, ref) <= VEL_TOL


def _oracle_gpu(sd, cfg, inputs, dtype, ocr=False):
    sdg = {k: v.cuda().to(dtype) for k, v in sd.items()}
    hs, lr, enc, pooled, t = inputs
    with torch.no_grad():
        return O.model_forward(sdg, cfg, hs.cuda().to(dtype), lr.cuda().to(dtype), enc.cuda().to(dtype),
                               pooled.cuda().to(dtype), t.cuda(), as_written=False, ocr=ocr).float().cpu()


def test_ocr_branch_model_full_width():
    """BASELINE config #5 shape family: 3-stream OCR2HQ2OCR model, full width, 4 blocks (2 dual -> exercises the Q6
    half-split of attn2), 512px, B=1; returns (output, trans_blk_outs) like the reference."""
    cfg = O.sd35_medium_config(num_layers=4, dual_attention_layers=(0, 1), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0, ocr=True)
    inputs = O.make_inputs(cfg, 1, 64, 64, 333, seed=1, timestep=600.0)
    ref32 = _oracle_gpu(sd, cfg, inputs, torch.float32, ocr=True)
    m = build(cfg, sd, ocr=True)
    hs, lr, enc, pooled, t = inputs
    bf = torch.bfloat16
    with torch.no_grad():
        out = m(hs.cuda().to(bf), lr.cuda().to(bf), enc.cuda().to(bf), pooled.cuda().to(bf), t.cuda(), return_dict=False)
    assert isinstance(out, tuple) and len(out) == 2 and out[1] == []
    e = rel(out[0].float().cpu(), ref32)
    print(f"ocr model: cuda-vs-fp32 {e:.3e}")
    assert e <= VEL_TOL


@pytest.mark.parametrize("timestep", [1000.0, 741.44, 8.93])
def test_full_width_model_velocity(timestep):
    """SD3.5-medium width (D=1536, 24 heads, F=6144), 4 blocks (2 dual, 1 plain, 1 context_pre_only), 512px, B=2."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = O.sd35_medium_config(num_layers=4, dual_attention_layers=(0, 1), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    inputs = O.make_inputs(cfg, 2, 64, 64, 333, seed=1, timestep=timestep)
    ref32 = _oracle_gpu(sd, cfg, inputs, torch.float32)
    refbf = _oracle_gpu(sd, cfg, inputs, torch.bfloat16)
    y = run(build(cfg, sd), inputs)
    e32, ebf, floor = rel(y, ref32), rel(y, refbf), rel(refbf, ref32)
    print(f"t={timestep}: cuda-vs-fp32 {e32:.3e}  cuda-vs-refbf16 {ebf:.3e}  refbf16-vs-fp32 (noise floor) {floor:.3e}")
    assert e32 <= VEL_TOL
    assert ebf <= VEL_TOL + floor


def test_full_model_24_blocks_velocity():
    """The BASELINE config C1 shape on the GPU: 24 blocks, 512px, B=1, T=333."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = O.sd35_medium_config(pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    inputs = O.make_inputs(cfg, 1, 64, 64, 333, seed=1, timestep=741.44)
    ref32 = _oracle_gpu(sd, cfg, inputs, torch.float32)
    refbf = _oracle_gpu(sd, cfg, inputs, torch.bfloat16)
    y = run(build(cfg, sd), inputs)
    e32, ebf, floor = rel(y, ref32), rel(y, refbf), rel(refbf, ref32)
    print(f"24 blocks: cuda-vs-fp32 {e32:.3e}  cuda-vs-refbf16 {ebf:.3e}  refbf16-vs-fp32 {floor:.3e}")
    assert e32 <= VEL_TOL
    assert ebf <= VEL_TOL + floor


def test_c3_shape_1024px_velocity():
    """BASELINE config C3 / C4 shape family: 1024px (4096 HQ + 4096 LR + 333 text tokens, S = 8525: 67 KV tiles, the
    non-persistent attention path), 2 blocks (1 dual), B=1."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = O.sd35_medium_config(num_layers=2, dual_attention_layers=(0,), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    inputs = O.make_inputs(cfg, 1, 128, 128, 333, seed=1, timestep=500.0)
    ref32 = _oracle_gpu(sd, cfg, inputs, torch.float32)
    e = rel(run(build(cfg, sd), inputs), ref32)
    print(f"1024px, 2 blocks: cuda-vs-fp32 {e:.3e}")
    assert e <= VEL_TOL


def test_unfolded_lr_chains_as_written():
    """fold_lr_chains=False runs the chained LR projections with a bf16 intermediate, as the reference is written."""
    cfg = O.sd35_medium_config(num_layers=3, dual_attention_layers=(0,), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    inputs = O.make_inputs(cfg, 1, 32, 32, 77, seed=1, timestep=300.0)
    ref32 = _oracle_gpu(sd, cfg, inputs, torch.float32)
    m = build(cfg, sd)
    m.engine(fold_lr_chains=False)
    e_unfolded = rel(run(m, inputs), ref32)
    m.engine(fold_lr_chains=True)
    e_folded = rel(run(m, inputs), ref32)
    print(f"unfolded {e_unfolded:.3e}  folded {e_folded:.3e}")
    assert e_unfolded <= VEL_TOL and e_folded <= VEL_TOL
    # and the no-qk_norm flavour unfolded (three live control chains)
    cfg2 = O.sd35_medium_config(num_layers=2, dual_attention_layers=(), qk_norm=None, pos_embed_max_size=96)
    sd2 = O.init_state_dict(cfg2, seed=0)
    m2 = build(cfg2, sd2)
    m2.engine(fold_lr_chains=False)
    assert rel(run(m2, inputs), _oracle_gpu(sd2, cfg2, inputs, torch.float32)) <= VEL_TOL


def test_sd3_medium_flavour_without_qk_norm():
    """qk_norm=None, no dual attention: the to_k_control(to_k(lr)) chain is live (attention.py:476)."""
    cfg = O.sd35_medium_config(num_layers=3, dual_attention_layers=(), qk_norm=None, pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    inputs = O.make_inputs(cfg, 1, 32, 32, 154, seed=1, timestep=500.0)
    ref32 = _oracle_gpu(sd, cfg, inputs, torch.float32)
    assert rel(run(build(cfg, sd), inputs), ref32) <= VEL_TOL


def test_sampling_loop_psnr():
    """10-step flow-match Euler loop (guidance 1): final latents vs the oracle's bf16 loop, PSNR >= 40 dB."""
    from dit4sr_b200.sampler import denoise
    cfg = O.sd35_medium_config(num_layers=4, dual_attention_layers=(0, 1), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    hs, lr, enc, pooled, _ = O.make_inputs(cfg, 1, 32, 32, 77, seed=1)
    m = build(cfg, sd)
    bf = torch.bfloat16
    out = denoise(m, hs.cuda().to(bf), lr.cuda().to(bf), enc.cuda().to(bf), pooled.cuda().to(bf),
                  num_inference_steps=10).float().cpu()
    sdg = {k: v.cuda().to(bf) for k, v in sd.items()}
    sig = O.flow_match_sigmas(10)
    x = hs.cuda().to(bf)
    with torch.no_grad():
        for i in range(10):
            t = torch.full((1,), float(sig[i]) * 1000.0, device="cuda")
            v = O.model_forward(sdg, cfg, x, lr.cuda().to(bf), enc.cuda().to(bf), pooled.cuda().to(bf), t, as_written=False)
            x = (x.float() + (float(sig[i + 1]) - float(sig[i])) * v.float()).to(v.dtype)
    ref = x.float().cpu()
    mse = ((out - ref) ** 2).mean().item()
    peak = (ref.max() - ref.min()).item()
    psnr = 10 * torch.log10(torch.tensor(peak ** 2 / mse)).item()
    print(f"PSNR of final latents vs reference bf16 loop: {psnr:.1f} dB")
    assert psnr >= 40.0


def test_feature_extraction_outputs():
    """'ts_module' path: per-block hidden-state slices are returned like transformer_sd3.py:361-368,404-405."""
    from types import SimpleNamespace as NS
    cfg = O.sd35_medium_config(num_layers=2, dual_attention_layers=(0,), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    hs, lr, enc, pooled, t = O.make_inputs(cfg, 1, 64, 64, 20, seed=1)
    m = build(cfg, sd)
    rcfg = NS(train=NS(model=["ts_module"], transformer=NS(feat_extract_layer=[0, 1], feat_extract="lq_feat")))
    bf = torch.bfloat16
    with torch.no_grad():
        out, feats = m(hs.cuda().to(bf), lr.cuda().to(bf), enc.cuda().to(bf), pooled.cuda().to(bf), t.cuda(),
                       return_dict=False, cfg=rcfg)
    assert len(feats) == 2 and feats[0]["extract_feat"].shape == (1, 1024, 1536)
    sdg = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        _, rf = O.model_forward(sdg, cfg, hs.cuda(), lr.cuda(), enc.cuda(), pooled.cuda(), t.cuda(), as_written=False,
                                feat_extract_layers=(0, 1), feat_extract="lq_feat")
    assert rel(feats[1]["extract_feat"].float().cpu(), rf[1].cpu()) <= VEL_TOL


# ----------------------------------------------------------------------------------------------------------------------
# Round 2: the configurations round 1 left untested (VERDICT "what's weak" 1)
def test_c4_shape_2048px_velocity():
    """BASELINE config C4 shape on one GPU: 2048px (16384 HQ + 16384 LR + 333 text tokens, S = 33101), 2 blocks (1 dual),
    B=1, against the fp32 oracle."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = O.sd35_medium_config(num_layers=2, dual_attention_layers=(0,), pos_embed_max_size=128)
    sd = O.init_state_dict(cfg, seed=0)
    inputs = O.make_inputs(cfg, 1, 256, 256, 333, seed=1, timestep=500.0)
    ref32 = _oracle_gpu(sd, cfg, inputs, torch.float32)
    e = rel(run(build(cfg, sd), inputs), ref32)
    print(f"2048px, 2 blocks: cuda-vs-fp32 {e:.3e}")
    assert e <= VEL_TOL


def test_c5_shape_ocr_model_1024px_velocity():
    """BASELINE config C5 shape: the OCR-branch model at 1024px (3 x 4096 + 333 tokens), 2 blocks (1 dual), B=1."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = O.sd35_medium_config(num_layers=2, dual_attention_layers=(0,), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0, ocr=True)
    inputs = O.make_inputs(cfg, 1, 128, 128, 333, seed=1, timestep=500.0)
    ref32 = _oracle_gpu(sd, cfg, inputs, torch.float32, ocr=True)
    e = rel(run(build(cfg, sd, ocr=True), inputs), ref32)
    print(f"OCR model 1024px, 2 blocks: cuda-vs-fp32 {e:.3e}")
    assert e <= VEL_TOL


def test_ocr_model_unfolded_lr_chains():
    """fold_lr_chains=False on the 3-stream model (the LR scratch buffer is sized for attn2's 1.5 N rows)."""
    cfg = O.sd35_medium_config(num_layers=2, dual_attention_layers=(0,), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0, ocr=True)
    inputs = O.make_inputs(cfg, 1, 32, 32, 77, seed=1, timestep=300.0)
    ref32 = _oracle_gpu(sd, cfg, inputs, torch.float32, ocr=True)
    m = build(cfg, sd, ocr=True)
    m.engine(fold_lr_chains=False)
    assert rel(run(m, inputs), ref32) <= VEL_TOL


def test_c2_batch8_24_blocks_velocity_online_and_prescaled():
    """The benchmarked configuration itself (C2: B=8, 24 blocks, 512px): every image against the fp32 oracle, with the
    pre-scaled bounded-score attention (default for this checkpoint) and with the online softmax."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = O.sd35_medium_config(pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    inputs = O.make_inputs(cfg, 8, 64, 64, 333, seed=1, timestep=741.44)
    ref32 = _oracle_gpu(sd, cfg, inputs, torch.float32)
    from dit4sr_b200.engine import Engine
    for prescale in (True, False):
        m = build(cfg, sd)
        m._engine = Engine(m, prescale_attention=prescale)
        assert (m._engine.blocks[0].attn.score_bound is not None) == prescale
        y = run(m, inputs)
        errs = [rel(y[i], ref32[i]) for i in range(8)]
        print(f"C2 B=8 24 blocks prescale={prescale}: max per-image rel-L2 {max(errs):.3e}")
        assert max(errs) <= VEL_TOL
        del m


def test_sampling_loop_psnr_40_steps_full_model():
    """The north star's PSNR gate at full scale: 40-step loop of the 24-block model at 512px (B=1) against the oracle's
    bf16 loop on the GPU; PSNR of the final latents >= 40 dB."""
    from dit4sr_b200.sampler import denoise
    cfg = O.sd35_medium_config(pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    hs, lr, enc, pooled, _ = O.make_inputs(cfg, 1, 64, 64, 333, seed=1)
    m = build(cfg, sd)
    bf = torch.bfloat16
    out = denoise(m, hs.cuda().to(bf), lr.cuda().to(bf), enc.cuda().to(bf), pooled.cuda().to(bf),
                  num_inference_steps=40).float().cpu()
    sdg = {k: v.cuda().to(bf) for k, v in sd.items()}
    sig = O.flow_match_sigmas(40)
    x = hs.cuda().to(bf)
    with torch.no_grad():
        for i in range(40):
            t = torch.full((1,), float(sig[i]) * 1000.0, device="cuda")
            v = O.model_forward(sdg, cfg, x, lr.cuda().to(bf), enc.cuda().to(bf), pooled.cuda().to(bf), t, as_written=False)
            x = (x.float() + (float(sig[i + 1]) - float(sig[i])) * v.float()).to(v.dtype)
    ref = x.float().cpu()
    mse = ((out - ref) ** 2).mean().item()
    peak = (ref.max() - ref.min()).item()
    psnr = 10 * torch.log10(torch.tensor(peak ** 2 / mse)).item()
    print(f"40 steps, 24 blocks: PSNR of final latents vs reference bf16 loop: {psnr:.1f} dB")
    assert psnr >= 40.0


def test_fp32_module_feature_outputs_do_not_alias_the_workspace():
    """with an fp32 module `.to(fp32)` of the fp32 residual workspace is a view; the hook must copy, otherwise
    an early block's feature ends up holding the last block's state."""
    from types import SimpleNamespace as NS
    cfg = O.sd35_medium_config(num_layers=3, dual_attention_layers=(0,), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    hs, lr, enc, pooled, t = O.make_inputs(cfg, 1, 64, 64, 20, seed=1)
    m = build(cfg, sd, dtype=torch.float32)
    rcfg = NS(train=NS(model=["ts_module"], transformer=NS(feat_extract_layer=[0], feat_extract="hq_feat")))
    with torch.no_grad():
        out, feats = m(hs.cuda(), lr.cuda(), enc.cuda(), pooled.cuda(), t.cuda(), return_dict=False, cfg=rcfg)
    sdg = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        _, rf = O.model_forward(sdg, cfg, hs.cuda(), lr.cuda(), enc.cuda(), pooled.cuda(), t.cuda(), as_written=False,
                                feat_extract_layers=(0,), feat_extract="hq_feat")
    assert rel(feats[0]["extract_feat"].float().cpu(), rf[0].cpu()) <= 1e-3


# ----------------------------------------------------------------------------------------------------------------------
# N1 / N2 of SURVEY 8f: classifier-free guidance + LRE start point, feature outputs for the text-spotting module
def _psnr(out, ref):
    mse = ((out - ref) ** 2).mean().item()
    peak = (ref.max() - ref.min()).item()
    return 10 * torch.log10(torch.tensor(peak ** 2 / mse)).item()


@pytest.mark.parametrize("start_point", ["noise", "lr"])
def test_sampling_loop_with_guidance_and_lre_start(start_point):
    """guidance_scale 8.0 (9 of the 36 shipped configs) with [negative | positive] prompt embeddings and the LRE start point,
    against the oracle's restatement of pipelines/pipeline_dit4sr.py:1025-1034,1085,1106-1108,1561-1567 in bf16."""
    from dit4sr_b200.sampler import denoise
    cfg = O.sd35_medium_config(num_layers=4, dual_attention_layers=(0, 1), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    hs, lr, enc, pooled, _ = O.make_inputs(cfg, 2, 32, 32, 77, seed=1)
    g = torch.Generator().manual_seed(5)
    neg, negp = 0.5 * torch.randn(enc.shape, generator=g), torch.randn(pooled.shape, generator=g)
    bf = torch.bfloat16
    c = lambda t: t.cuda().to(bf)
    m = build(cfg, sd)
    out = denoise(m, c(hs), c(lr), c(enc), c(pooled), num_inference_steps=10, guidance_scale=8.0,
                  negative_prompt_embeds=c(neg), negative_pooled_prompt_embeds=c(negp), start_point=start_point).float().cpu()
    sdg = {k: v.cuda().to(bf) for k, v in sd.items()}
    with torch.no_grad():
        ref = O.sample_loop_cfg(sdg, cfg, c(hs), c(lr), c(enc), c(pooled), c(neg), c(negp), 8.0, num_steps=10,
                                start_point=start_point).float().cpu()
    # the guided loop amplifies bf16 differences by the guidance scale: the two bf16 runs are compared as images
    psnr = _psnr(out, ref)
    print(f"guidance 8.0, start point {start_point}: PSNR vs the oracle's bf16 loop {psnr:.1f} dB")
    assert psnr >= 40.0
    # and one guided step against fp32 (velocity tolerance scaled by the guidance amplification)
    from dit4sr_b200 import ops as _ops  # noqa: F401
    eng = m.engine()
    x2 = torch.cat([c(hs)] * 2)
    t = torch.full((4,), 500.0, device="cuda")
    v = eng.step(x2, torch.cat([c(lr)] * 2), torch.cat([c(neg), c(enc)]), torch.cat([c(negp), c(pooled)]), t,
                 guidance=8.0).float().cpu()
    sd32 = {k: v_.cuda() for k, v_ in sd.items()}
    with torch.no_grad():
        vv = O.model_forward(sd32, cfg, x2.float(), torch.cat([lr.cuda()] * 2), torch.cat([neg, enc]).cuda(),
                             torch.cat([negp, pooled]).cuda(), t, as_written=False).cpu()
    vu, vt = vv.chunk(2)
    assert v.shape == vu.shape
    assert rel(v, vu + 8.0 * (vt - vu)) <= 8 * VEL_TOL


def test_feature_outputs_nchw_fp32_inside_the_graph():
    """pipelines/pipeline_dit4sr.py:1214-1245: 'b (N H W) (pH pW d) -> b (N d) (H pH) (W pW)' to fp32, produced by the engine
    while the graph replays (two calls: capture, then replay with other inputs), for the hq+lq concatenation."""
    from types import SimpleNamespace as NS
    from einops import rearrange
    import dit4sr_b200
    cfg = O.sd35_medium_config(num_layers=3, dual_attention_layers=(0,), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    m = build(cfg, sd)
    rcfg = NS(train=NS(model=["ts_module"], transformer=NS(feat_extract_layer=[0, 2], feat_extract="hqlq_feat")))
    bf = torch.bfloat16
    sdg = {k: v.cuda() for k, v in sd.items()}
    for seed in (1, 2):      # first call captures the graph, second replays it
        hs, lr, enc, pooled, t = O.make_inputs(cfg, 1, 64, 64, 20, seed=seed)
        with torch.no_grad():
            out, feats = m(hs.cuda().to(bf), lr.cuda().to(bf), enc.cuda().to(bf), pooled.cuda().to(bf), t.cuda(),
                           return_dict=False, cfg=rcfg)
            _, rf = O.model_forward(sdg, cfg, hs.cuda(), lr.cuda(), enc.cuda(), pooled.cuda(), t.cuda(), as_written=False,
                                    feat_extract_layers=(0, 2), feat_extract="hqlq_feat")
        assert len(feats) == 2 and feats[0]["extract_feat"].shape == (1, 2048, 1536)
        nchw = dit4sr_b200.unpatchify_features(feats, "hqlq_feat")
        for k in range(2):
            want = rearrange(rf[k], "b (N H W) (pH pW d) -> b (N d) (H pH) (W pW)", N=2, H=32, W=32, pH=2, pW=2).float()
            assert nchw[k].dtype == torch.float32 and nchw[k].shape == (1, 768, 64, 64)
            assert rel(nchw[k].cpu(), want.cpu()) <= VEL_TOL
            assert rel(feats[k]["extract_feat"].float().cpu(), rf[k].cpu()) <= VEL_TOL
            # the layout itself is exact: the engine's two outputs describe the same numbers up to the bf16 cast of one
            again = rearrange(feats[k]["extract_feat"].float(), "b (N H W) (pH pW d) -> b (N d) (H pH) (W pW)", N=2, H=32,
                              W=32, pH=2, pW=2)
            assert rel(again, nchw[k]) <= 4e-3
    # a bare extract_feat (e.g. from a saved tensor) goes through the same kernel: exact rearrange
    bare = [{"extract_feat": feats[0]["extract_feat"]}]
    got = dit4sr_b200.unpatchify_features(bare, "hqlq_feat")[0]
    assert torch.equal(got, rearrange(feats[0]["extract_feat"].float(), "b (N H W) (pH pW d) -> b (N d) (H pH) (W pW)",
                                      N=2, H=32, W=32, pH=2, pW=2))


def test_fp16_module_runs_in_bf16_within_the_stated_tolerance():
    """Every shipped reference config runs `mixed_precision: fp16` (run_configs/val/*.yaml).  This product has no fp16
    arithmetic: an fp16 module is computed with bf16 tensor-core operands and fp32 accumulation / residual streams, fp16
    inputs are converted at the boundary and the output is returned in fp16.  Stated tolerance: the velocity stays within
    the north star's bf16 gate (1e-2) of the fp32 oracle, and is no further from it than the reference's own fp16 run
    (the oracle evaluated in fp16 on the GPU) plus that gate."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = O.sd35_medium_config(num_layers=4, dual_attention_layers=(0, 1), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    inputs = O.make_inputs(cfg, 1, 64, 64, 333, seed=1, timestep=600.0)
    ref32 = _oracle_gpu(sd, cfg, inputs, torch.float32)
    ref16 = _oracle_gpu(sd, cfg, inputs, torch.float16)
    m = build(cfg, sd, dtype=torch.float16)
    y = run(m, inputs, dtype=torch.float16)
    e_ours, e_ref16 = rel(y, ref32), rel(ref16, ref32)
    print(f"fp16 module: cuda(bf16 arithmetic) vs fp32 {e_ours:.3e}; reference fp16 run vs fp32 {e_ref16:.3e}")
    assert e_ours <= VEL_TOL
    assert rel(y, ref16) <= VEL_TOL + e_ref16


def test_sampling_loop_hoisted_invariants_match_per_step_head():
    """denoise() forms the adaLN vectors of all steps and the embedded text tokens once (Engine.prepare_loop) - the same
    kernels on the same operands as the per-step head, so the loop must end where the per-step loop ends (with and without
    guidance; graph replay and direct launches)."""
    from dit4sr_b200.sampler import denoise
    cfg = O.sd35_medium_config(num_layers=3, dual_attention_layers=(0,), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    hs, lr, enc, pooled, _ = O.make_inputs(cfg, 2, 32, 32, 77, seed=1)
    bf = torch.bfloat16
    m = build(cfg, sd)
    a = [t.cuda().to(bf) for t in (hs, lr, enc, pooled)]
    neg = dict(negative_prompt_embeds=torch.zeros_like(a[2]), negative_pooled_prompt_embeds=torch.zeros_like(a[3]))
    for kw in (dict(), dict(guidance_scale=4.0, **neg)):
        for graph in (True, False):
            m.engine().use_cuda_graph = graph
            y1 = denoise(m, *a, num_inference_steps=6, hoist_loop_invariants=True, **kw).float()
            y0 = denoise(m, *a, num_inference_steps=6, hoist_loop_invariants=False, **kw).float()
            err = rel(y1, y0)
            print(f"hoisted vs per-step head (guidance {kw.get('guidance_scale', 1.0)}, graph {graph}): rel-L2 {err:.2e}")
            assert err < 1e-5

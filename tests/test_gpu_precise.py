"""GPU parity of the fp32-accurate mode (BASELINE north_star: per-step velocity rel-L2 <= 1e-4 against the fp32
reference).  In this mode activations stay fp32 and every tensor-core product runs on bf16 hi/lo splits (K' = 3K);
it is selected automatically for fp32 modules.  All calls go through the C ABI."""
import json
import os

import pytest
import torch
from safetensors.torch import load_file

from oracle import dit4sr_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FP32_TOL = 1e-4   # BASELINE.json north_star, fp32 gate


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def test_split3_reconstructs_fp32_to_16_bits():
    from dit4sr_b200 import ops
    x = torch.randn(3, 50, 64, device="cuda") * 3
    out = torch.empty(3 * 56, 192, device="cuda", dtype=torch.bfloat16)
    ops.split3(x, out, pattern=0, rows_pad=56)
    o = out.view(3, 56, 3, 64).float()
    assert torch.equal(o[:, :50, 0], o[:, :50, 1])                       # [hi | hi | lo]
    assert torch.equal(o[:, 50:], torch.zeros_like(o[:, 50:]))           # padded rows are zero
    assert rel(o[:, :50, 0] + o[:, :50, 2], x) < 2 ** -16
    ops.split3(x, out, pattern=1, rows_pad=56)
    o = out.view(3, 56, 3, 64).float()
    assert torch.equal(o[:, :50, 0], o[:, :50, 2])                       # [hi | lo | hi]
    assert rel(o[:, :50, 0] + o[:, :50, 1], x) < 2 ** -16
    v = torch.randn(70, 64, device="cuda")
    vt = torch.empty(64, 3 * 128, device="cuda", dtype=torch.bfloat16)
    ops.split3_transposed(v, vt, 128)
    t = vt.view(64, 3, 128).float()
    assert torch.equal(t[:, 0], t[:, 2]) and torch.equal(t[:, :, 70:], torch.zeros_like(t[:, :, 70:]))
    assert rel((t[:, 0] + t[:, 1])[:, :70].t(), v) < 2 ** -16


@pytest.mark.parametrize("epi", ["bias", "gelu", "silu", "gate_resid"])
def test_split_gemm_matches_fp64(epi):
    """A W^T on bf16x3 splits vs an fp64 matmul: <= 2e-5 (the bf16 single-product path is at ~3e-3)."""
    from dit4sr_b200 import ops
    torch.manual_seed(0)
    B, R, K, N = 2, 333, 1536, 768
    A = torch.randn(B, R, K, device="cuda")
    W = torch.randn(N, K, device="cuda") * 0.02
    bias = torch.randn(N, device="cuda") * 0.02
    A3 = torch.empty(B, R, 3 * K, device="cuda", dtype=torch.bfloat16)
    W3 = torch.empty(N, 3 * K, device="cuda", dtype=torch.bfloat16)
    ops.split3(A, A3.view(-1), 0)
    ops.split3(W, W3.view(-1), 1)
    y = (A.double() @ W.double().t() + bias.double())
    out = torch.randn(B, R, N, device="cuda")
    if epi == "bias":
        ops.gemm(A3, W3, bias, out, ops.EPI_BIAS)
        ref = y
    elif epi == "gelu":
        ops.gemm(A3, W3, bias, out, ops.EPI_BIAS_GELU)
        ref = torch.nn.functional.gelu(y, approximate="tanh")
    elif epi == "silu":
        ops.gemm(A3, W3, bias, out, ops.EPI_BIAS_SILU)
        ref = torch.nn.functional.silu(y)
    else:
        gate = torch.randn(B, N, device="cuda")
        ref = out.double() + gate.double()[:, None] * y
        ops.gemm(A3, W3, bias, out, ops.EPI_GATE_RESID, gate=gate)
    assert rel(out, ref) <= 2e-5


def test_split_gemm_qkv_norm_fp32_out():
    from dit4sr_b200 import ops
    torch.manual_seed(0)
    B, R, K, D = 1, 200, 256, 128
    A = torch.randn(B, R, K, device="cuda")
    W = torch.randn(3 * D, K, device="cuda") * 0.05
    bias = torch.randn(3 * D, device="cuda") * 0.02
    wq, wk = 1 + 0.1 * torch.randn(64, device="cuda"), 1 + 0.1 * torch.randn(64, device="cuda")
    A3 = torch.empty(B, R, 3 * K, device="cuda", dtype=torch.bfloat16)
    W3 = torch.empty(3 * D, 3 * K, device="cuda", dtype=torch.bfloat16)
    ops.split3(A, A3.view(-1), 0)
    ops.split3(W, W3.view(-1), 1)
    out = torch.zeros(B, 2 * R, 3 * D, device="cuda")
    ops.gemm(A3, W3, bias, out[:, :R], ops.EPI_QKV_NORM, norm_w0=wq, norm_w1=wk, sec_cols=D, dbl_row_offset=R)
    y = (A.double() @ W.double().t() + bias.double()).view(B, R, 3, 2, 64)

    def rms(t, w):
        return t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-6) * w.double()
    q, k, v = rms(y[:, :, 0], wq), rms(y[:, :, 1], wk), y[:, :, 2]
    assert rel(out[:, :R, :D], q.reshape(B, R, D)) <= 2e-5
    assert rel(out[:, :R, D:2 * D], k.reshape(B, R, D)) <= 2e-5
    assert rel(out[:, :R, 2 * D:], v.reshape(B, R, D)) <= 2e-5
    assert rel(out[:, R:, D:2 * D], rms(k, wk).reshape(B, R, D)) <= 2e-5     # quirk Q1: double-normed LR keys


def test_fp32_helpers_match_torch():
    from dit4sr_b200 import ops
    torch.manual_seed(0)
    s = torch.randn(37, 256, device="cuda") * 4
    ref = torch.softmax(s[:, :201].double() * 0.125, -1)
    ops.softmax_rows(s, 201, 0.125)
    assert rel(s[:, :201], ref) < 1e-6 and s[:, 201:].abs().max().item() == 0.0
    B, G, Cc = 2, 12, 64
    u = torch.randn(B, 2 * G * G, Cc, device="cuda")
    w = torch.randn(Cc, 1, 3, 3, device="cuda")
    b = torch.randn(Cc, device="cuda")
    lr = u[:, G * G:].reshape(B, G, G, Cc).permute(0, 3, 1, 2).double()
    ref = u[:, :G * G].double() + torch.nn.functional.conv2d(lr, w.double(), b.double(), padding=1, groups=Cc) \
        .permute(0, 2, 3, 1).reshape(B, G * G, Cc)
    ops.dwconv3x3_add_f32(u[:, :G * G], u[:, G * G:], w.reshape(Cc, 9).t().contiguous(), b, G)
    assert rel(u[:, :G * G], ref) < 1e-6
    x = torch.randn(2, 16, 8, 8, device="cuda")
    p = torch.empty(2, 16, 64, device="cuda")
    ops.patchify(x, p)
    assert torch.equal(p, x.reshape(2, 16, 4, 2, 4, 2).permute(0, 2, 4, 1, 3, 5).reshape(2, 16, 64))
    xx = torch.randn(5, 40, 256, device="cuda") * 2 + 0.3
    sh, sc = torch.randn(5, 256, device="cuda"), torch.randn(5, 256, device="cuda")
    o = torch.empty(5, 40, 256, device="cuda")
    ops.ln_modulate(xx, sh, sc, o)
    ref = torch.nn.functional.layer_norm(xx.double(), (256,), eps=1e-6) * (1 + sc.double()[:, None]) + sh.double()[:, None]
    assert rel(o, ref) < 1e-6


def build(cfg, sd, ocr=False):
    import dit4sr_b200
    cls = dit4sr_b200.SD3Transformer2DModel_OCRBranch_OCR2HQ2OCR if ocr else dit4sr_b200.SD3Transformer2DModel
    m = cls(**vars(cfg))
    m.load_state_dict(sd)
    m = m.to("cuda", torch.float32).requires_grad_(False).eval()
    assert m.engine().precise            # fp32 module -> fp32-accurate engine
    return m


def run(m, inputs):
    hs, lr, enc, pooled, t = inputs
    with torch.no_grad():
        y = m(hidden_states=hs.cuda(), controlnet_image=lr.cuda(), encoder_hidden_states=enc.cuda(),
              pooled_projections=pooled.cuda(), timestep=t.cuda(), return_dict=False)[0]
    assert y.dtype == torch.float32
    return y.cpu()


def _oracle_gpu(sd, cfg, inputs, ocr=False, dtype=torch.float32):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sdg = {k: v.cuda().to(dtype) for k, v in sd.items()}
    hs, lr, enc, pooled, t = inputs
    with torch.no_grad():
        return O.model_forward(sdg, cfg, hs.cuda().to(dtype), lr.cuda().to(dtype), enc.cuda().to(dtype),
                               pooled.cuda().to(dtype), t.cuda(), as_written=False, ocr=ocr).float().cpu()


META = json.load(open(os.path.join(GOLD, "reference_meta.json")))


@pytest.mark.parametrize("case", META["cases"], ids=lambda c: c["name"])
def test_tiny_model_fp32_matches_reference_golden(case):
    """The committed fp32 outputs of the reference's own model code vs the fp32-accurate CUDA path."""
    ref = load_file(os.path.join(GOLD, "reference_tiny.safetensors"))[case["name"] + ".velocity"]
    ocr = case.get("ocr", False)
    cfg = O.tiny_config(**case["cfg_kwargs"])
    sd = O.init_state_dict(cfg, seed=case["weight_seed"], ocr=ocr)
    inputs = O.make_inputs(cfg, case["batch"], case["height"], case["width"], case["text_tokens"],
                           seed=case["input_seed"], timestep=case["timestep"])
    e = rel(run(build(cfg, sd, ocr=ocr), inputs), ref)
    print(f"{case['name']}: fp32-accurate mode vs reference golden {e:.3e}")
    assert e <= FP32_TOL


@pytest.mark.parametrize("fold", [True, False])
def test_full_width_model_fp32(fold):
    """SD3.5-medium width, 4 blocks (2 dual, 1 plain, 1 context_pre_only), 512px, B=2, vs the fp32 and fp64 oracle."""
    cfg = O.sd35_medium_config(num_layers=4, dual_attention_layers=(0, 1), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    inputs = O.make_inputs(cfg, 2, 32, 32, 77, seed=1, timestep=741.44)
    ref32 = _oracle_gpu(sd, cfg, inputs)
    ref64 = _oracle_gpu(sd, cfg, inputs, dtype=torch.float64)
    m = build(cfg, sd)
    m.engine(fold_lr_chains=fold)
    y = run(m, inputs)
    e32, e64, floor = rel(y, ref32), rel(y, ref64), rel(ref32, ref64)
    print(f"fold={fold}: cuda-vs-fp32 {e32:.3e}  cuda-vs-fp64 {e64:.3e}  fp32-vs-fp64 (oracle's own floor) {floor:.3e}")
    assert e32 <= FP32_TOL and e64 <= FP32_TOL


def test_c1_24_blocks_fp32():
    """BASELINE config C1 on the GPU: 24 blocks, 512px, B=1, T=333, fp32 gate 1e-4."""
    cfg = O.sd35_medium_config(pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    inputs = O.make_inputs(cfg, 1, 64, 64, 333, seed=1, timestep=741.44)
    ref32 = _oracle_gpu(sd, cfg, inputs)
    y = run(build(cfg, sd), inputs)
    e = rel(y, ref32)
    print(f"C1 24 blocks: fp32-accurate mode vs fp32 oracle {e:.3e}")
    assert e <= FP32_TOL


def test_ocr_model_fp32():
    cfg = O.sd35_medium_config(num_layers=3, dual_attention_layers=(0,), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0, ocr=True)
    inputs = O.make_inputs(cfg, 1, 32, 32, 77, seed=1, timestep=600.0)
    ref32 = _oracle_gpu(sd, cfg, inputs, ocr=True)
    m = build(cfg, sd, ocr=True)
    with torch.no_grad():
        out = m(inputs[0].cuda(), inputs[1].cuda(), inputs[2].cuda(), inputs[3].cuda(), inputs[4].cuda(),
                return_dict=False)
    e = rel(out[0].cpu(), ref32)
    print(f"ocr model fp32-accurate: {e:.3e}")
    assert e <= FP32_TOL

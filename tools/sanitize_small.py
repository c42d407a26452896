"""Small-shape invocation of every kernel family through the C ABI with a parity check against torch and SENTINEL GUARD
REGIONS around every output (a write outside the tensor is reported).  It can also run under compute-sanitizer
(`compute-sanitizer --tool memcheck python tools/sanitize_small.py`, one tool per gpurun call) where that is open; on this
pool it is closed (profiles/r02_sanitizer_closed.log), so the guards are the bounds check."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dit4sr_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
bf = torch.bfloat16
rel = lambda a, b: ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()
ok = True


GUARD = 4096          # elements of sentinel on either side of every guarded output
_guards = []


def guarded(shape, dtype, fill=None):
    """An output tensor carved out of a larger sentinel-filled allocation: `check_guards` verifies after the kernels that
    nothing outside the tensor was written (compute-sanitizer is closed on this pool; this is the bounds check of our own
    the pool's notice asks for)."""
    n = 1
    for d in shape:
        n *= d
    sentinel = 165 if dtype == torch.uint8 else 1234.5
    big = torch.full((n + 2 * GUARD,), sentinel, device=dev, dtype=dtype)
    t = big[GUARD:GUARD + n].view(*shape)
    if fill is not None:
        t.fill_(fill)
    _guards.append((big, n, sentinel))
    return t


def check_guards(name):
    global ok
    bad = 0
    for big, n, sentinel in _guards:
        bad += int((big[:GUARD] != sentinel).sum()) + int((big[GUARD + n:] != sentinel).sum())
    print(f"{name:34s} guard regions {'intact' if bad == 0 else 'OVERWRITTEN: %d elements' % bad}", flush=True)
    ok = ok and bad == 0
    _guards.clear()


def report(name, err, tol):
    global ok
    print(f"{name:34s} rel {err:.3e} {'ok' if err <= tol else 'BAD'}", flush=True)
    ok = ok and err <= tol


# ---- attention: both kernels, both softmax forms, ragged tails, residual rows
for variant in (2, 3, 4):
    for prescaled in (False, True):
        B, H, S, Sq = 2, 2, 333, 300
        D = H * 64
        qkv = torch.randn(B, S, 3 * D, device=dev)
        if prescaled:
            qkv[..., :D] *= 1.4426950408889634 / 8.0
        qkv = qkv.to(bf)
        q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
        res = torch.randn(B, S, D, device=dev).to(bf)
        o = guarded((B, Sq, D), bf, 0.0)
        ops.attention(q, k, v, o, H, seq_q=Sq, lr_res=res, lr_rows=(100, 200), score_bound=30.0 if prescaled else None,
                      variant=variant)
        hd = lambda t: t.float().reshape(B, S, H, 64).transpose(1, 2)
        ref = torch.nn.functional.scaled_dot_product_attention(hd(q), hd(k), hd(v), scale=math.log(2.0) if prescaled else None)
        ref = ref.transpose(1, 2).reshape(B, S, D)[:, :Sq].clone()
        ref[:, 100:200] += res[:, 100:200].float()
        report(f"attention v{variant} prescaled={prescaled}", rel(o, ref), 4e-3)
        if variant >= 3:      # split-KV tail: slices + merge kernel through the scratch buffer
            ws = guarded((ops.attention_workspace_bytes(),), torch.uint8)
            o2 = guarded((B, Sq, D), bf, 0.0)
            ops.attention(q, k, v, o2, H, seq_q=Sq, lr_res=res, lr_rows=(100, 200), score_bound=30.0 if prescaled else None,
                          variant=variant, workspace=ws, split_kv=2)
            report(f"attention v{variant} prescaled={prescaled} split 2", rel(o2, ref), 6e-3)
check_guards("attention")

# ---- GEMM epilogues (grouped, tail-split path included by the row count)
M, K, N = 300, 128, 192
A = torch.randn(2, M, K, device=dev).to(bf)
W = (torch.randn(N, K, device=dev) * 0.1).to(bf)
bias = torch.randn(N, device=dev)
y = A.float() @ W.float().t() + bias
out = guarded((2, M, N), bf)
ops.gemm(A, W, bias, out, ops.EPI_BIAS)
report("gemm BIAS", rel(out, y), 5e-3)
ops.gemm(A, W, bias, out, ops.EPI_BIAS_GELU)
report("gemm BIAS_GELU", rel(out, torch.nn.functional.gelu(y, approximate="tanh")), 8e-3)
ops.gemm(A, W, bias, out, ops.EPI_BIAS_SILU)
report("gemm BIAS_SILU", rel(out, torch.nn.functional.silu(y)), 8e-3)
w0, w1 = torch.rand(64, device=dev) + 0.5, torch.rand(64, device=dev) + 0.5
out2 = guarded((2, 2 * M, N), bf, 0.0)
ops.gemm(A, W, bias, out2[:, :M], ops.EPI_QKV_NORM, norm_w0=w0, norm_w1=w1, sec_cols=64, dbl_row_offset=M)


def rms(t, w):
    return t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-6) * w


refq = torch.cat([rms(y[..., :64], w0), rms(y[..., 64:128], w1), y[..., 128:]], -1)
report("gemm QKV_NORM", rel(out2[:, :M], refq), 6e-3)
report("gemm QKV_NORM double-normed keys", rel(out2[:, M:, 64:128], rms(rms(y[..., 64:128], w1), w1)), 6e-3)
x = torch.randn(2, M, N, device=dev)
gate = torch.randn(2, N, device=dev)
ref = x + gate[:, None] * y
xs = guarded((2, M, N), torch.float32)
xs.copy_(x)
ops.gemm(A, W, bias, xs, ops.EPI_GATE_RESID, gate=gate)
report("gemm GATE_RESID (reduce-add)", rel(xs, ref), 5e-3)
probs = [dict(A=A, W=W, bias=bias, out=torch.empty(2, M, N, device=dev, dtype=bf), epilogue=ops.EPI_BIAS) for _ in range(3)]
ops.gemm_grouped(probs)
report("gemm grouped x3", max(rel(p["out"], y) for p in probs), 5e-3)
check_guards("gemm")

# ---- LayerNorm + modulation (single / dual), depthwise stencil, small movers
Dm = 256
x = torch.randn(2, 70, Dm, device=dev)
mod = torch.randn(2, 4 * Dm, device=dev)
o1, o2 = guarded((2, 70, Dm), bf), guarded((2, 70, Dm), bf)
ops.ln_modulate(x, mod[:, :Dm], mod[:, Dm:2 * Dm], o1, mod[:, 2 * Dm:3 * Dm], mod[:, 3 * Dm:], o2)
ln = torch.nn.functional.layer_norm(x, (Dm,), eps=1e-6)
report("ln_modulate dual", max(rel(o1, ln * (1 + mod[:, None, Dm:2 * Dm]) + mod[:, None, :Dm]),
                               rel(o2, ln * (1 + mod[:, None, 3 * Dm:]) + mod[:, None, 2 * Dm:3 * Dm])), 5e-3)
g, F = 6, 64
u = torch.randn(2, 2 * g * g, F, device=dev).to(bf)
w9 = torch.randn(9, F, device=dev) * 0.1
b9 = torch.randn(F, device=dev) * 0.1
refc = u[:, :g * g].float() + torch.nn.functional.conv2d(
    u[:, g * g:].float().reshape(2, g, g, F).permute(0, 3, 1, 2), w9.t().reshape(F, 1, 3, 3), b9, padding=1,
    groups=F).permute(0, 2, 3, 1).reshape(2, g * g, F)
ops.dwconv3x3_add(u[:, :g * g], u[:, g * g:], w9, b9, g)
report("dwconv3x3_add", rel(u[:, :g * g], refc), 8e-3)
img = torch.randn(2, 16, 8, 8, device=dev).to(bf)
pt = guarded((2, 16, 64), bf)
ops.patchify(img, pt)
refp = img.reshape(2, 16, 4, 2, 4, 2).permute(0, 2, 4, 1, 3, 5).reshape(2, 16, 64)
report("patchify", rel(pt, refp), 0.0)
tok = torch.randn(2, 16, 64, device=dev)
im2 = guarded((2, 16, 8, 8), bf)
ops.unpatchify(tok, im2)
refu = tok.reshape(2, 4, 4, 2, 2, 16).permute(0, 5, 1, 3, 2, 4).reshape(2, 16, 8, 8)
report("unpatchify", rel(im2, refu), 4e-3)
a_, b_ = torch.randn(2, 10, 64, device=dev).to(bf), torch.randn(2, 10, 64, device=dev).to(bf)
refa = a_.float() + b_.float()
ops.add_bf16(a_, b_)
report("add_bf16", rel(a_, refa), 4e-3)
# feature outputs and the guided unpatchify
xf = torch.randn(2, 2 * 16, 128, device=dev)
nchw, tokf = guarded((2, 2 * 32, 8, 8), torch.float32), guarded((2, 2 * 16, 128), bf)
ops.feature_extract(xf, 0, 2, 4, 4, nchw=nchw, tok=tokf)
want = xf.reshape(2, 2, 4, 4, 2, 2, 32).permute(0, 1, 6, 2, 4, 3, 5).reshape(2, 64, 8, 8)
report("feature_extract nchw", rel(nchw, want), 0.0)
report("feature_extract tok", rel(tokf, xf), 4e-3)
tok2 = torch.randn(4, 16, 64, device=dev)
lat = torch.randn(2, 16, 8, 8, device=dev)
outg = guarded((2, 16, 8, 8), torch.float32)
ops.unpatchify(tok2, outg, latents=lat, dsigma=-0.05, guidance=8.0)
vv = tok2.reshape(4, 4, 4, 2, 2, 16).permute(0, 5, 1, 3, 2, 4).reshape(4, 16, 8, 8)
report("unpatchify_cfg", rel(outg, lat - 0.05 * (vv[:2] + 8.0 * (vv[2:] - vv[:2]))), 1e-6)
check_guards("ln / stencil / movers")
torch.cuda.synchronize()
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)

"""Developer harness: runs each CUDA kernel against a plain torch fp32 reference in its own subprocess (a faulting
kernel cannot take the others down) and writes gpurun_out/check.json.  Usage: python tools/gpu_check.py [case ...]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item(), (a - b).abs().max().item()


def case_ln():
    import torch
    from dit4sr_b200 import ops
    torch.manual_seed(0)
    out = {}
    for (B, R, D) in [(2, 333, 1536), (8, 2048, 1536), (1, 7, 128)]:
        x = torch.randn(B, R, D, device="cuda") * 2 + 0.5
        mod = torch.randn(B, 4 * D, device="cuda") * 0.5
        sh, sc, sh2, sc2 = mod[:, :D], mod[:, D:2 * D], mod[:, 2 * D:3 * D], mod[:, 3 * D:]
        o1 = torch.empty(B, R, D, device="cuda", dtype=torch.bfloat16)
        o2 = torch.empty_like(o1)
        ops.ln_modulate(x, sh, sc, o1, sh2, sc2, o2)
        ln = torch.nn.functional.layer_norm(x, (D,), eps=1e-6)
        r1 = ln * (1 + sc[:, None]) + sh[:, None]
        r2 = ln * (1 + sc2[:, None]) + sh2[:, None]
        o3 = torch.empty_like(o1)
        ops.ln_modulate(x, sh, sc, o3)
        out[f"{B}x{R}x{D}"] = dict(out1=rel(o1.float(), r1), out2=rel(o2.float(), r2), single=rel(o3.float(), r1))
    return out


def case_dwconv():
    import torch
    from dit4sr_b200 import ops
    torch.manual_seed(0)
    out = {}
    for (B, G, C) in [(2, 32, 6144), (1, 5, 512), (1, 64, 256)]:
        N = G * G
        u = (torch.randn(B, 2 * N, C, device="cuda")).bfloat16()
        w = torch.randn(C, 1, 3, 3, device="cuda") * 0.1
        bias = torch.randn(C, device="cuda") * 0.1
        ref_lr = u[:, N:].float().reshape(B, G, G, C).permute(0, 3, 1, 2)
        conv = torch.nn.functional.conv2d(ref_lr, w, bias, padding=1, groups=C).reshape(B, C, N).permute(0, 2, 1)
        ref = u[:, :N].float() + conv
        w9 = w.reshape(C, 9).t().contiguous()
        u2 = u.clone()
        ops.dwconv3x3_add(u2[:, :N], u2[:, N:], w9, bias, G)
        out[f"{B}x{G}x{C}"] = dict(hq=rel(u2[:, :N].float(), ref), lr_untouched=bool(torch.equal(u2[:, N:], u[:, N:])))
    return out


def case_gemm():
    import torch
    from dit4sr_b200 import ops
    torch.manual_seed(0)
    out = {}
    dev = "cuda"

    def mk(B, R, K, N):
        A = (torch.randn(B, R, K, device=dev)).bfloat16()
        W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        bias = torch.randn(N, device=dev) * 0.1
        ref = A.float() @ W.float().t() + bias
        return A, W, bias, ref

    for (B, R, K, N) in [(1, 128, 64, 256), (1, 256, 1536, 1536), (2, 333, 1536, 4608), (1, 8, 1536, 512),
                         (1, 1024, 1536, 64), (1, 4096, 6144, 1536)]:
        A, W, bias, ref = mk(B, R, K, N)
        o = torch.empty(B, R, N, device=dev, dtype=torch.bfloat16)
        ops.gemm(A, W, bias, o, ops.EPI_BIAS)
        o32 = torch.empty(B, R, N, device=dev, dtype=torch.float32)
        ops.gemm(A, W, bias, o32, ops.EPI_BIAS)
        out[f"bias_{B}x{R}x{K}x{N}"] = dict(bf16=rel(o.float(), ref), f32=rel(o32, ref))
    # strided A (column slice) + strided C (row segment of a bigger buffer)
    A, W, bias, ref = mk(2, 200, 512, 256)
    big = torch.zeros(2, 200, 1024, device=dev, dtype=torch.bfloat16)
    big[:, :, 512:] = A
    cbuf = torch.zeros(2, 500, 768, device=dev, dtype=torch.bfloat16)
    ops.gemm(big[:, :, 512:], W, bias, cbuf[:, 100:300, 256:512], ops.EPI_BIAS)
    untouched = cbuf.clone()
    untouched[:, 100:300, 256:512] = 0
    out["strided"] = dict(val=rel(cbuf[:, 100:300, 256:512].float(), ref), clean=float(untouched.abs().max()))
    # activations
    A, W, bias, ref = mk(1, 512, 1536, 6144)
    o = torch.empty(1, 512, 6144, device=dev, dtype=torch.bfloat16)
    ops.gemm(A, W, bias, o, ops.EPI_BIAS_GELU)
    out["gelu"] = rel(o.float(), torch.nn.functional.gelu(ref, approximate="tanh"))
    ops.gemm(A, W, bias, o, ops.EPI_BIAS_SILU)
    out["silu"] = rel(o.float(), torch.nn.functional.silu(ref))
    # gate + residual (in place and with a broadcast residual)
    A, W, bias, ref = mk(2, 333, 1536, 1536)
    gate = torch.randn(2, 1536, device=dev)
    x = torch.randn(2, 333, 1536, device=dev)
    x0 = x.clone()
    ops.gemm(A, W, bias, x, ops.EPI_GATE_RESID, gate=gate)
    out["gate_resid_inplace"] = rel(x, x0 + gate[:, None] * ref)
    pos = torch.randn(1, 333, 1536, device=dev)
    y = torch.empty(2, 333, 1536, device=dev)
    ops.gemm(A, W, bias, y, ops.EPI_GATE_RESID, resid=pos)
    out["resid_broadcast"] = rel(y, pos + ref)
    # qkv + rmsnorm (+ double-normed k written N rows further down)
    Bq, Nq, D = 2, 256, 1536
    S = 2 * Nq + 77
    A, W, bias, ref = mk(Bq, Nq, 1536, 3 * D)
    wq = 1 + 0.1 * torch.randn(64, device=dev)
    wk = 1 + 0.1 * torch.randn(64, device=dev)
    qkv = torch.zeros(Bq, S, 3 * D, device=dev, dtype=torch.bfloat16)
    ops.gemm(A, W, bias, qkv[:, :Nq], ops.EPI_QKV_NORM, norm_w0=wq, norm_w1=wk, sec_cols=D, dbl_row_offset=Nq)

    def rms(t, w):
        t = t.reshape(*t.shape[:-1], -1, 64)
        t = t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-6) * w
        return t.reshape(*t.shape[:-2], -1)

    rq, rk, rv = rms(ref[..., :D], wq), rms(ref[..., D:2 * D], wk), ref[..., 2 * D:]
    rk2 = rms(rk, wk)
    out["qkv_norm"] = dict(q=rel(qkv[:, :Nq, :D].float(), rq), k=rel(qkv[:, :Nq, D:2 * D].float(), rk),
                           v=rel(qkv[:, :Nq, 2 * D:].float(), rv), k_lr=rel(qkv[:, Nq:2 * Nq, D:2 * D].float(), rk2),
                           rest=float(qkv[:, 2 * Nq:].abs().max()) + float(qkv[:, Nq:2 * Nq, :D].abs().max()))
    return out


def case_attn():
    import torch
    from dit4sr_b200 import ops
    torch.manual_seed(0)
    out = {}
    dev = "cuda"
    for (B, H, S, Sq, lr) in [(1, 2, 256, 256, None), (2, 3, 333, 333, None), (1, 24, 2381, 2381, (1024, 2048)),
                              (2, 4, 2381, 2048, (1024, 2048)), (1, 2, 100, 100, None)]:
        D = H * 64
        qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16()
        q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
        o = torch.zeros(B, Sq, D, device=dev, dtype=torch.bfloat16)
        res = torch.randn(B, S, D, device=dev).bfloat16() if lr else None
        ops.attention(q, k, v, o, H, seq_q=Sq, lr_res=res, lr_rows=lr or (0, 0))

        def hd(t):
            return t.float().reshape(B, S, H, 64).transpose(1, 2)

        ref = torch.nn.functional.scaled_dot_product_attention(hd(q), hd(k), hd(v)).transpose(1, 2).reshape(B, S, D)
        ref = ref[:, :Sq].clone()
        if lr:
            ref[:, lr[0]:lr[1]] += res[:, lr[0]:lr[1]].float()
        out[f"{B}x{H}x{S}q{Sq}"] = rel(o.float(), ref)
    return out


def case_misc():
    import torch
    from dit4sr_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    out = {}
    x = torch.randn(2, 16, 8, 12, device=dev)
    p = torch.empty(2, 24, 64, device=dev, dtype=torch.bfloat16)
    ops.patchify(x, p)
    ref = x.reshape(2, 16, 4, 2, 6, 2).permute(0, 2, 4, 1, 3, 5).reshape(2, 24, 64)
    out["patchify"] = rel(p.float(), ref.bfloat16().float())
    tok = torch.randn(2, 24, 64, device=dev)
    o = torch.empty(2, 16, 8, 12, device=dev)
    ops.unpatchify(tok, o)
    ref = tok.reshape(2, 4, 6, 2, 2, 16).permute(0, 5, 1, 3, 2, 4).reshape(2, 16, 8, 12)
    out["unpatchify"] = rel(o, ref)
    lat = torch.randn(2, 16, 8, 12, device=dev)
    ops.unpatchify(tok, o, latents=lat, dsigma=-0.025)
    out["euler"] = rel(o, lat - 0.025 * ref)
    t = torch.tensor([1000.0, 741.44, 8.93], device=dev)
    f = torch.empty(3, 256, device=dev, dtype=torch.bfloat16)
    ops.timestep_features(t, f)
    import math
    ex = torch.exp(-math.log(10000) * torch.arange(128, device=dev, dtype=torch.float32) / 128)
    e = t[:, None] * ex[None]
    out["timestep"] = rel(f.float(), torch.cat([torch.cos(e), torch.sin(e)], -1))
    a = torch.randn(1000, device=dev)
    b = torch.empty(1000, device=dev, dtype=torch.bfloat16)
    ops.silu_cast(a, b)
    out["silu"] = rel(b.float(), torch.nn.functional.silu(a))
    return out


CASES = dict(ln=case_ln, dwconv=case_dwconv, gemm=case_gemm, attn=case_attn, misc=case_misc)


def child(name):
    import torch
    res = CASES[name]()
    torch.cuda.synchronize()
    print("RESULT " + json.dumps(res))


def main():
    names = sys.argv[1:] or list(CASES)
    summary = {}
    for n in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--child", n], capture_output=True, text=True, timeout=300)
            lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if p.returncode == 0 and lines:
                summary[n] = json.loads(lines[-1][7:])
            else:
                summary[n] = dict(error=f"rc={p.returncode}", stderr=p.stderr[-3000:], stdout=p.stdout[-1000:])
        except subprocess.TimeoutExpired:
            summary[n] = dict(error="timeout")
        summary[n + "_secs"] = round(time.time() - t0, 1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "check.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        main()

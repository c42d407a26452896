"""Developer micro-benchmarks of single kernels at the C2 / 1024px shapes (CUDA-event timing, L2 flushed by size).
Usage: python tools/kbench.py attn|gemm|dwconv|ln [--reps N] [--shape c2|px1024]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dit4sr_b200 import ops  # noqa: E402


GRAPH = False


def timeit(fn, reps, warm=3):
    """Average device time of fn() over `reps` back-to-back launches; with --graph the launches are replayed from a
    CUDA graph so that host-side launch cost (python + ctypes, ~10-20 us per call) cannot bound short kernels."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if GRAPH:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
    else:
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--shape", default="c2")
    ap.add_argument("--split", type=int, default=-1, help="attn: -1 no workspace, 0 automatic split-KV tail, 2..4 forced")
    ap.add_argument("--variant", type=int, default=0, help="attn: 0 default, 2 rotating buffers, 3 double-buffered 96-key tiles, 4 one tile per CTA, 5 = 4 with split rows")
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--sustain", type=float, default=0.0, help="attn: also run back to back for this many seconds, sampling clocks and power")
    ap.add_argument("--prescaled", action="store_true", help="attn: DIT4SR_ATTN_PRESCALED mode (q pre-multiplied)")
    ap.add_argument("--exp-offload", type=int, default=0, help="attn: 0 default, 1/2/3 = 0/8, 2/8, 4/8 polynomial")
    a = ap.parse_args()
    global GRAPH
    GRAPH = a.graph
    B, N, T = (8, 1024, 333) if a.shape == "c2" else (2, 4096, 333)
    D, H, F = 1536, 24, 6144
    if a.shape == "c4rank":      # what one of 8 ranks runs under Ulysses at 2048px: 3 heads over the whole sequence
        B, N, T, D, H = 1, 16384, 333, 192, 3
    S = 2 * N + T
    if a.shape == "attn2":       # the image-only second attention of the dual blocks at C2
        S, T = 2 * N, 0
    dev = "cuda"
    torch.manual_seed(0)
    if a.what == "attn":
        qkv = torch.randn(B, S, 3 * D, device=dev)
        if a.prescaled:
            qkv[..., :D] *= 1.4426950408889634 / 8.0
        qkv = qkv.bfloat16()
        o = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
        res = torch.randn(B, 2 * N, D, device=dev).bfloat16()
        last = ops.attention_last_plan
        wsp = torch.empty(ops.attention_workspace_bytes(), device=dev, dtype=torch.uint8) if a.split >= 0 else None
        fn = lambda: ops.attention(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], o, H, lr_res=res, lr_rows=(N, 2 * N),
                                   score_bound=30.0 if a.prescaled else None, exp_offload=a.exp_offload, variant=a.variant,
                                   workspace=wsp, split_kv=max(a.split, 0))
        ms = timeit(fn, a.reps)
        fl = 4.0 * S * S * D * B
        if a.sustain > 0:   # same kernel back to back for `sustain` seconds with nvidia-smi sampling clocks / power
            import subprocess, tempfile, time
            f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            pr = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                                   "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=f)
            n = max(int(a.sustain * 1e3 / ms), 10)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            pr.terminate(); pr.wait()
            rows = [l.strip().split(", ") for l in open(f.name) if l.strip()]
            rows = rows[len(rows) // 3:]
            mhz = sorted(float(r[0]) for r in rows); watts = sorted(float(r[1]) for r in rows)
            cap = sum(r[2].lower().startswith("active") for r in rows)
            print(f"  sustained {n} launches: {e0.elapsed_time(e1) / n:.4f} ms  {fl * n / e0.elapsed_time(e1) / 1e9:.1f} TFLOP/s; "
                  f"median {mhz[len(mhz) // 2]:.0f} MHz, {watts[len(watts) // 2]:.0f} W, power-cap active in {cap}/{len(rows)} samples")
        print(f"attn B={B} S={S} prescaled={a.prescaled} exp_offload={a.exp_offload} variant={a.variant} split={a.split}: {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s  plan {last()}")
    elif a.what == "gemm":
        for name, M, K, Nn, epi in [("ff1", B * 2 * N, D, F, ops.EPI_BIAS_GELU), ("ff2", B * 2 * N, F, D, ops.EPI_GATE_RESID),
                                    ("qkv", B * N, D, 3 * D, ops.EPI_QKV_NORM), ("out", B * N, D, D, ops.EPI_GATE_RESID)]:
            A = torch.randn(M, K, device=dev).bfloat16()
            W = (torch.randn(Nn, K, device=dev) * 0.02).bfloat16()
            bias = torch.randn(Nn, device=dev)
            out = torch.empty(M, Nn, device=dev, dtype=torch.float32 if epi == ops.EPI_GATE_RESID else torch.bfloat16)
            if epi == ops.EPI_GATE_RESID:
                out.zero_()
            kw = {}
            if epi == ops.EPI_QKV_NORM:
                w = torch.ones(64, device=dev)
                kw = dict(norm_w0=w, norm_w1=w, sec_cols=D)
            ms = timeit(lambda: ops.gemm(A, W, bias, out, epi, **kw), a.reps)
            print(f"gemm {name} M={M} K={K} N={Nn}: {ms:.4f} ms  {2.0 * M * K * Nn / ms / 1e9:.1f} TFLOP/s")
    elif a.what == "dwconv":
        g = int(N ** 0.5)
        u = torch.randn(B, 2 * N, F, device=dev).bfloat16()
        w9 = torch.randn(9, F, device=dev) * 0.1
        bias = torch.randn(F, device=dev) * 0.1
        ms = timeit(lambda: u[:, :N].add_(u[:, N:]), a.reps)
        print(f"torch u_hq += u_lr (same traffic): {ms:.4f} ms  {3.0 * B * N * F * 2 / ms / 1e6:.0f} GB/s")
        ms = timeit(lambda: ops.dwconv3x3_add(u[:, :N], u[:, N:], w9, bias, g), a.reps)
        print(f"dwconv B={B} grid={g}: {ms:.4f} ms  {3.0 * B * N * F * 2 / ms / 1e6:.0f} GB/s (algorithmic)")
    elif a.what == "ln":
        x = torch.randn(B, 2 * N, D, device=dev)
        mod = torch.randn(B, 4 * D, device=dev)
        o1 = torch.empty(B, 2 * N, D, device=dev, dtype=torch.bfloat16)
        o2 = torch.empty_like(o1)
        ms = timeit(lambda: o1.copy_(x), a.reps)
        print(f"torch fp32->bf16 copy of the same tensors (same traffic as ln single): {ms:.4f} ms  {B * 2 * N * D * 6 / ms / 1e6:.0f} GB/s")
        ms = timeit(lambda: ops.ln_modulate(x, mod[:, :D], mod[:, D:2 * D], o1), a.reps)
        print(f"ln single: {ms:.4f} ms  {B * 2 * N * D * 6 / ms / 1e6:.0f} GB/s")
        ms = timeit(lambda: ops.ln_modulate(x, mod[:, :D], mod[:, D:2 * D], o1, mod[:, 2 * D:3 * D], mod[:, 3 * D:], o2), a.reps)
        print(f"ln dual: {ms:.4f} ms  {B * 2 * N * D * 8 / ms / 1e6:.0f} GB/s")
        # text stream: its own launch against riding in the image stream's launch (companion problem, ABI v5)
        ctx = torch.randn(B, T, D, device=dev)
        oc = torch.empty(B, T, D, device=dev, dtype=torch.bfloat16)
        ms_t = timeit(lambda: ops.ln_modulate(ctx, mod[:, :D], mod[:, D:2 * D], oc), a.reps)
        ms_c = timeit(lambda: ops.ln_modulate(x, mod[:, :D], mod[:, D:2 * D], o1, companion=(ctx, mod[:, 2 * D:3 * D], mod[:, 3 * D:], oc)), a.reps)
        print(f"ln text alone: {ms_t:.4f} ms; image + text in one launch: {ms_c:.4f} ms  {B * (2 * N + T) * D * 6 / ms_c / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main()

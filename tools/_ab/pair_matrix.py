import os, sys, torch
os.environ["DIT4SR_GEMM_PAIR"] = "1"
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dit4sr_b200 import ops
torch.manual_seed(0)
dev = "cuda"
def run(name, M, K, N, epi, reps=3):
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    bias = torch.randn(N, device=dev)
    f32 = epi == ops.EPI_GATE_RESID
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    ref = A.float() @ W.float().t() + bias
    if epi == ops.EPI_BIAS_GELU:
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    try:
        ops.gemm(A, W, bias, out, epi)
        torch.cuda.synchronize()
        err = ((out.float() - ref).norm() / ref.norm()).item()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if f32: out.zero_()
        e0.record()
        for _ in range(reps): ops.gemm(A, W, bias, out, epi)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{name:28s} M={M} K={K} N={N}: rel {err:.2e}  {ms:.4f} ms {2.0*M*K*N/ms/1e9:.0f} TFLOP/s", flush=True)
    except Exception as e:
        print(f"{name:28s} M={M} K={K} N={N}: FAILED {str(e)[:80]}", flush=True)
        sys.exit(1)
run("bias small", 512, 256, 256, ops.EPI_BIAS)
run("bias K=6144 small M", 512, 6144, 256, ops.EPI_BIAS)
run("gate_resid small", 512, 256, 256, ops.EPI_GATE_RESID)
run("gate_resid K=6144", 512, 6144, 512, ops.EPI_GATE_RESID)
run("gate_resid M=16384 K=1536", 16384, 1536, 1536, ops.EPI_GATE_RESID)
run("ff1", 16384, 1536, 6144, ops.EPI_BIAS_GELU, 10)
run("ff2", 16384, 6144, 1536, ops.EPI_GATE_RESID, 10)

"""B=1 512px step latency with and without CUDA graph replay (full 24-block model)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

m, cfg = bench.build_model_gpu(torch.device("cuda"), 0)
eng = m.engine()
bf = torch.bfloat16
for B in (1, 8):
    x = torch.randn(B, 16, 64, 64, device="cuda")
    c = torch.randn(B, 16, 64, 64, device="cuda").to(bf)
    e = torch.randn(B, 333, 4096, device="cuda").to(bf)
    p = torch.randn(B, 2048, device="cuda").to(bf)
    t = torch.full((B,), 500.0, device="cuda")
    for graph in (False, True):
        eng.use_cuda_graph = graph
        for _ in range(3):
            y = eng.step(x, c, e, p, t)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            y = eng.step(x, c, e, p, t)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        if graph:
            eng.use_cuda_graph = False
            ref = eng.forward(x, c, e, p, t)
            print(f"B={B} graph={graph}: {dt * 1e3:.2f} ms/step  max|graph-eager| = {(y - ref).abs().max().item():.3e}")
        else:
            print(f"B={B} graph={graph}: {dt * 1e3:.2f} ms/step")

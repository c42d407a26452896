"""Developer bench of the whole sampling loop (sampler.denoise, 40 steps, 512px, SD3.5-medium shape, random weights):
ms per step with the loop-invariant work hoisted (Engine.prepare_loop) and with the per-step head.
Usage (GPU box): python tools/loop_bench.py [--batch 1] [--steps 40] [--guidance 1.0]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--side", type=int, default=64)
    ap.add_argument("--guidance", type=float, default=1.0)
    a = ap.parse_args()
    import bench
    from dit4sr_b200.sampler import denoise
    torch.manual_seed(0)
    bf, dev = torch.bfloat16, "cuda"
    m, _ = bench.build_model_gpu(dev)
    B, s = a.batch, a.side
    x = torch.randn(B, 16, s, s, device=dev).to(bf)
    c = torch.randn(B, 16, s, s, device=dev).to(bf)
    enc = (torch.randn(B, 333, 4096, device=dev) * 0.5).to(bf)
    pooled = torch.randn(B, 2048, device=dev).to(bf)
    kw = {}
    if a.guidance > 1.0:
        kw = dict(guidance_scale=a.guidance, negative_prompt_embeds=torch.zeros_like(enc),
                  negative_pooled_prompt_embeds=torch.zeros_like(pooled))
    res = {}
    import time
    eng = m.engine()
    for _ in range(2):      # what the table itself costs (host wall clock around a synchronised call)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cache = eng.prepare_loop(enc, pooled, [500.0] * a.steps)
        torch.cuda.synchronize()
        print(f"prepare_loop({a.steps} steps, B={B}): {(time.perf_counter() - t0) * 1e3:.2f} ms, table "
              f"{cache['mod'].numel() * 4 / 2 ** 20:.0f} MiB")
        del cache
    for hoist in (True, False, True, False, False, True):
        denoise(m, x, c, enc, pooled, num_inference_steps=3, hoist_loop_invariants=hoist, **kw)      # graphs, attributes
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = denoise(m, x, c, enc, pooled, num_inference_steps=a.steps, hoist_loop_invariants=hoist, **kw)
        e1.record()
        torch.cuda.synchronize()
        res.setdefault(hoist, []).append(e0.elapsed_time(e1))
        assert torch.isfinite(y.float()).all()
    for hoist, v in res.items():
        print(f"denoise B={B} {8 * s}px {a.steps} steps guidance {a.guidance} hoisted={hoist}: "
              + ", ".join(f"{t:.1f} ms ({t / a.steps:.3f} ms/step)" for t in v))


if __name__ == "__main__":
    main()

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
    ap.add_argument("--tile", type=int, default=0, help="latent tile size: time the tiled loop (overlap 24) against the per-step composition it replaced")
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
    if a.tile:
        import numpy as np
        from dit4sr_b200 import ops
        from dit4sr_b200.sampler import flow_match_sigmas
        from dit4sr_b200.tiling import tiled_velocity

        def legacy(steps):     # what sampler._denoise_tiled did before: everything re-formed per step, model called as a module
            sig = flow_match_sigmas(steps)
            xx = x.float().contiguous()
            nxt = torch.empty_like(xx)
            for i in range(steps):
                t_i = float(sig[i]) * 1000.0

                def model_fn(lat, ctl):
                    n = lat.shape[0] // B
                    t = torch.full((lat.shape[0],), t_i, device=lat.device, dtype=torch.float32)
                    return m(lat.to(bf), ctl, enc.repeat(n, 1, 1), pooled.repeat(n, 1), t, return_dict=False)[0]

                v = tiled_velocity(model_fn, xx, c, a.tile, 24).contiguous()
                ops.euler_update(xx, v, nxt, float(sig[i + 1]) - float(sig[i]))
                xx, nxt = nxt, xx
            return xx.to(bf)

        fns = {"tiled loop (plan + loop table)": lambda n: denoise(m, x, c, enc, pooled, num_inference_steps=n, latent_tiled_size=a.tile,
                                                                   latent_tiled_overlap=24),
               "per-step composition": legacy}
        outs = {}
        for _ in range(2):
            for name, fn in fns.items():
                fn(2)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                outs[name] = fn(a.steps)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                print(f"{name}: B={B} latent {s}x{s} tile {a.tile} {a.steps} steps: {ms:.1f} ms ({ms / a.steps:.3f} ms/step)")
        d = (outs["tiled loop (plan + loop table)"].float() - outs["per-step composition"].float()).abs().max().item()
        print(f"max abs difference of the final latents: {d:.3e}")
        return
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

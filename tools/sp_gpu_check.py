"""Multi-GPU check of Ulysses sequence parallelism:  torchrun --nproc-per-node P tools/sp_gpu_check.py [--side 64]
[--layers 4] [--steps 3].  Every rank runs the same image through (a) the single-GPU engine, (b) the sequence-parallel
engine over torch.distributed (NCCL) and (c) the sequence-parallel engine over NVLink peer memory (captured in a CUDA
graph); all must agree to bf16 noise and match the oracle (<= 1e-2, the bf16 tolerance of BASELINE.json), and a short
sharded sampling loop must match the single-GPU loop.  Exit code 0 iff every rank passes."""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dit4sr_b200  # noqa: E402
from dit4sr_b200.parallel import PeerSequenceParallel, SequenceParallel  # noqa: E402
from dit4sr_b200.sampler import denoise  # noqa: E402
from oracle import dit4sr_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=64)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--no-oracle", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    side = a.side
    cfg = O.sd35_medium_config(num_layers=a.layers, dual_attention_layers=tuple(range(min(2, a.layers))),
                               pos_embed_max_size=max(96, side // 2))
    sd = O.init_state_dict(cfg, seed=0)
    m = dit4sr_b200.SD3Transformer2DModel(**vars(cfg))
    m.load_state_dict(sd)
    m = m.to("cuda", torch.bfloat16).requires_grad_(False).eval()
    hs, lr, enc, pooled, t = O.make_inputs(cfg, 1, side, side, 333, seed=1, timestep=600.0)
    bf = torch.bfloat16
    args = (hs.cuda().to(bf), lr.cuda().to(bf), enc.cuda().to(bf), pooled.cuda().to(bf), t.cuda())
    rel = lambda x, y: ((x - y).norm() / y.norm()).item()

    def timed(fn, n=5):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    ok = True
    with torch.no_grad():
        y1 = m(*args, return_dict=False)[0].float()
        x1 = denoise(m, args[0], args[1], args[2], args[3], num_inference_steps=a.steps).float()
        ms1 = timed(lambda: m(*args, return_dict=False))
        ref = None
        if not a.no_oracle:
            sdg = {k: v.cuda() for k, v in sd.items()}
            ref = O.model_forward(sdg, cfg, hs.cuda(), lr.cuda(), enc.cuda(), pooled.cuda(), t.cuda(), as_written=False)
            del sdg
        for name, sp in (("nccl", SequenceParallel()), ("peer", PeerSequenceParallel())):
            m.engine(sequence_parallel=sp)
            y2 = m(*args, return_dict=False)[0].float()
            x2 = denoise(m, args[0], args[1], args[2], args[3], num_inference_steps=a.steps).float()
            ms2 = timed(lambda: m(*args, return_dict=False))
            line = (f"rank {rank} [{name}]: sp-vs-single {rel(y2, y1):.3e}  loop({a.steps} steps) sp-vs-single "
                    f"{rel(x2, x1):.3e}")
            if ref is not None:
                line += f"  single-vs-oracle {rel(y1, ref):.3e}  sp-vs-oracle {rel(y2, ref):.3e}"
                ok = ok and rel(y2, ref) <= 1e-2
            print(line + f"  step single {ms1:.2f} ms / sp {ms2:.2f} ms", flush=True)
            ok = ok and rel(y2, y1) <= 1e-2 and rel(x2, x1) <= 2e-2
            m.engine(sequence_parallel=None)
            if hasattr(sp, "close"):
                dist.barrier()
                sp.close()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()

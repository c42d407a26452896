"""Multi-GPU check of Ulysses sequence parallelism: torchrun --nproc-per-node P tools/sp_gpu_check.py
Each rank runs the same image through (a) the single-GPU engine and (b) the sequence-parallel engine; outputs must agree
(bf16 noise only) and both must match the oracle."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dit4sr_b200  # noqa: E402
from dit4sr_b200.parallel import SequenceParallel  # noqa: E402
from oracle import dit4sr_oracle as O  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ocr = "--ocr" in sys.argv
    side = 64
    cfg = O.sd35_medium_config(num_layers=4, dual_attention_layers=(0, 1), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0, ocr=ocr)
    cls = dit4sr_b200.SD3Transformer2DModel_OCRBranch_OCR2HQ2OCR if ocr else dit4sr_b200.SD3Transformer2DModel
    m = cls(**vars(cfg))
    m.load_state_dict(sd)
    m = m.to("cuda", torch.bfloat16).requires_grad_(False).eval()
    hs, lr, enc, pooled, t = O.make_inputs(cfg, 1, side, side, 333, seed=1, timestep=600.0)
    bf = torch.bfloat16
    args = (hs.cuda().to(bf), lr.cuda().to(bf), enc.cuda().to(bf), pooled.cuda().to(bf), t.cuda())
    with torch.no_grad():
        y1 = m(*args, return_dict=False)[0].float()
        m.engine(sequence_parallel=SequenceParallel())
        y2 = m(*args, return_dict=False)[0].float()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            y2 = m(*args, return_dict=False)[0].float()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        sdg = {k: v.cuda() for k, v in sd.items()}
        ref = O.model_forward(sdg, cfg, hs.cuda(), lr.cuda(), enc.cuda(), pooled.cuda(), t.cuda(), as_written=False, ocr=ocr)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    print(f"rank {rank}: sp-vs-single {rel(y2, y1):.3e}  single-vs-oracle {rel(y1, ref):.3e}  sp-vs-oracle {rel(y2, ref):.3e}"
          f"  sp step {dt * 1e3:.2f} ms", flush=True)
    ok = rel(y2, ref) <= 1e-2 and rel(y2, y1) <= 1e-2
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

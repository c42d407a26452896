#!/usr/bin/env python
"""Benchmark of the DiT4SR denoising hot path (BASELINE.json metric: DiT denoise steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|px1024|px2048]

Workload at N=1 (default): BASELINE config C2 -- SD3.5-medium-shaped DiT4SR transformer (24 joint blocks, hidden 1536,
24 heads, dual attention in blocks 0-12, QK-RMSNorm, random init), bf16, batch 8, 512px (64x64 latents -> 1024 HQ +
1024 LR + 333 text tokens), stepping through the 40-step flow-match schedule.  One "step" = one denoise step of the
whole batch (model forward + fused Euler update).  N>1: one process per GPU (torchrun), every rank runs its own batch
(data-parallel sampling, no data-path collective) -> weak scaling; time = max over ranks.

Prints ONE JSON line (see README / DESIGN.md for the keys).  `--impl reference` times the CPU restatement of the
reference (oracle/, plain PyTorch fp32, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (latent side, batch, text tokens, mode, description)
    #   mode dp_weak  : `batch` images PER GPU, data-parallel sampling, no data-path collective (weak scaling)
    #   mode dp_strong: `batch` images IN TOTAL, split over the GPUs (strong scaling)
    #   mode sp       : `batch` images IN TOTAL, every image sharded over all GPUs with Ulysses sequence parallelism
    "c2": (64, 8, 333, "dp_weak", "C2: SD3.5-medium-shaped DiT4SR, bf16, 512px (1024 HQ + 1024 LR + 333 text tokens), "
                                  "batch 8 per GPU, 40-step flow-match schedule"),
    "c3": (128, 32, 333, "dp_strong", "C3: 1024px (4096 HQ + 4096 LR + 333 text tokens), batch 32 in total, data-parallel"),
    "c4": (256, 1, 333, "sp", "C4: 2048px (16384 HQ + 16384 LR + 333 text tokens), batch 1, Ulysses sequence parallelism "
                              "(head<->sequence all-to-all over NCCL)"),
    "c5": (128, 8, 333, "dp_strong", "C5: OCR-branch model (transformer_sd3_ocrbranch_ocr2hq2ocr: HQ | LR | OCR | text joint "
                                     "attention, 3 x 4096 + 333 tokens), 1024px, batch 8 in total, data-parallel"),
    "px1024": (128, 1, 333, "dp_weak", "1024px batch 1 per GPU"),
}


UNIT = "image-steps/s"   # one image advanced one denoise step; the SAME string in both arms (the driver divides them)


def workload_config(workload, world):
    """`config` of the JSON line: identical for the product arm and the reference arm of one workload / world size."""
    side, B, T, mode, desc = WORKLOADS[workload]
    per_gpu = B // world if mode == "dp_strong" else B
    total = B if mode in ("dp_strong", "sp") else B * world
    N = (side // 2) ** 2
    return dict(workload=desc, per_gpu_batch=per_gpu, global_batch=total, latent=f"{side}x{side}", tokens_hq=N,
                tokens_lr=N, tokens_text=T, parallelism=(f"ulysses{world}" if mode == "sp" else f"dp{world}"),
                l2="working set (5.4 GB bf16 weights + activations) >> 126 MB L2; no explicit flush")


def flops_per_image_step_ocr(N, T, L=24, dual=13, D=1536):
    """Same for the 3-stream OCR-branch model (LR and OCR projection chains folded, as the engine runs them):
    SURVEY.md 8d quotes 54.90 TFLOP at 1024px with folded chains."""
    F = 4 * D
    gemm = attn = 0.0
    for i in range(L):
        last = i == L - 1
        # q|k|v: HQ 3N, LR q,v 2N, OCR q,k,v 3N, text 3T; out: 3N (+T)
        gemm += 2 * D * D * (3 * N + 2 * N + 3 * N + 3 * T + 3 * N + (0 if last else T))
        attn += 4 * (3 * N + T) ** 2 * D
        if i < dual:   # attn2 is the 2-stream processor on the 3N image tokens split at 1.5N
            h = 1.5 * N
            gemm += 2 * D * D * (3 * h + 2 * h + 2 * h)
            attn += 4 * (3 * N) ** 2 * D
        gemm += 4 * (3 * N) * D * F + (0 if last else 4 * T * D * F)
    return gemm, attn


def flops_per_image_step(N, T, L=24, dual=13, D=1536, folded=True):
    """Algorithmic FLOPs of one denoise step of one image (dead LR-key GEMMs removed).
    folded=True counts the LR projection chains as ONE GEMM each (what the engine runs by default; BASELINE.md's
    bracketed totals: 4.724 TFLOP at 512px); folded=False counts them as two (5.081 TFLOP)."""
    F = 4 * D
    gemm = 0.0
    attn = 0.0
    lr = 2 * N if folded else 4 * N
    for i in range(L):
        last = i == L - 1
        gemm += 2 * D * D * (3 * N + lr + 3 * T + 2 * N + (0 if last else T))
        attn += 4 * (2 * N + T) ** 2 * D
        if i < dual:
            gemm += 2 * D * D * (3 * N + lr + 2 * N)
            attn += 4 * (2 * N) ** 2 * D
        gemm += 4 * (2 * N) * D * F + (0 if last else 4 * T * D * F)
    return gemm, attn


# The SD3.5-medium shape the reference loads (SURVEY.md 8d "Model config"); the product arm builds its model from this
# dict - nothing under oracle/ is imported outside the cpu_baseline / reference legs.
SD35_MEDIUM_CONFIG = dict(num_layers=24, num_attention_heads=24, attention_head_dim=64, caption_projection_dim=1536,
                          joint_attention_dim=4096, pooled_projection_dim=2048, patch_size=2, in_channels=16,
                          out_channels=16, sample_size=128, pos_embed_max_size=384, qk_norm="rms_norm",
                          dual_attention_layers=tuple(range(13)))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="MEASURED_PEAKS.json (measured)")
    return dict(tf_burst=1590.0, tf_sustained=1400.0, hbm=6650.0, source="B200_PROFILING.md fallback")


def hbm_rooflines(breakdown, B, N, T, n_layers, n_dual, peaks, D=1536, F=6144):
    """`roofline_hbm`: the two HBM-bound kernel classes of a 2-stream step against the measured copy bandwidth, from the
    instrumented step's CUDA-event times (in situ: cold L2, clocks of the power-capped step).  Algorithmic bytes (SURVEY
    8d with the fp32 residual stream this build keeps): LayerNorm+modulate reads 4 B and writes 2 B per element and output
    (two outputs on the dual blocks; the text rows ride in the same launches); the stencil reads u_lr, reads and writes u_hq.
    In situ these kernels also carry the write-back of what the kernel before them left dirty in the 126 MB L2 (the fp32
    residual stream after a reduce-add epilogue, the MLP hidden state after FF1), which is not in their algorithmic bytes:
    alone, on warm buffers, tools/kbench.py measures 0.84 (LayerNorm) and 0.91 (stencil) of the same peak."""
    if not breakdown or "ln_mod" not in breakdown or "dwconv" not in breakdown:
        return None
    img, txt = 2 * N * B, T * B
    ln = 0
    for i in range(n_layers):
        ln += img * D * (4 + 2 * (2 if i < n_dual else 1)) + txt * D * 6      # norm1 (+ second modulation), norm1_context
        ln += img * D * 6 + (txt * D * 6 if i < n_layers - 1 else 0)          # norm2, text norm2 (absent in the last block)
    conv = n_layers * 3 * N * B * F * 2
    out = {}
    for name, key, nbytes in (("ln_modulate", "ln_mod", ln), ("dwconv3x3", "dwconv", conv)):
        ms = breakdown[key]["total_ms"]
        ach = nbytes / (ms / 1e3) / 1e9
        out[name] = dict(bound="hbm", achieved=round(ach, 1), peak=peaks["hbm"], unit="GB/s", frac=round(ach / peaks["hbm"], 4),
                         launches=breakdown[key]["launches"], total_ms=ms, algorithmic_bytes_per_step=nbytes)
    return out


def attention_roofline(launch_ms, flops_per_step, launches_per_step, peaks, root=ROOT):
    """`roofline` entry of the joint-attention kernel: algorithmic FLOPs per launch (average over the joint and the
    image-only calls of a step) / average launch duration measured with CUDA events; traffic from the committed ncu
    --set full capture of the same kernel."""
    if not launch_ms or not launches_per_step:
        return None
    avg_ms = sum(launch_ms) / len(launch_ms)
    flops = flops_per_step / launches_per_step
    ach = flops / (avg_ms / 1e3) / 1e12
    traffic = src = None
    try:
        prof = json.load(open(os.path.join(root, "profiles", "r02f_ncu_full.json")))
        k = next(k for k in prof if "attn_fwd5_kernel" in k["Kernel Name"])
        traffic = round((float(k["dram__bytes_read.sum [Mbyte]"]) + float(k["dram__bytes_write.sum [Mbyte]"])) * 1e6)
        src = ("from committed capture profiles/r02f_ncu_full.json of the same kernel (attn_fwd5_kernel; not measured in this run): "
               "dram__bytes_read.sum + dram__bytes_write.sum of one joint-attention "
               "launch (S=2381, B=8; algorithmic 263 MB q|k|v read + 59 MB o written + 50 MB LR residual read)")
    except (OSError, KeyError, ValueError, StopIteration):
        pass
    return dict(bound="tensor", kernel="attn_fwd5_kernel (joint attention over [HQ|LR|text] and the image-only attn2, "
                "head_dim 64; the largest share of the step: MUFU.EX2 / issue limited, DESIGN.md 3.2)", achieved=round(ach, 1),
                peak=peaks["tf_sustained"], unit="TFLOP/s", frac=round(ach / peaks["tf_sustained"], 4),
                peak_source=peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                frac_of_burst=round(ach / peaks["tf_burst"], 4), frac_of_spec_2250=round(ach / 2250.0, 4),
                launch_ms=round(avg_ms, 4), flops_per_launch=flops, launches_timed=len(launch_ms), traffic=traffic,
                traffic_source=src)


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            hi = [s for s in sm if s >= 0.5 * sm[-1]]   # samples under load
            out.update(sm_mhz=hi[len(hi) // 2], sm_max_mhz=mx, samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def build_model_gpu(device, seed=0, ocr=False):
    """SD3.5-medium-shaped model with EVERY parameter random (SURVEY 8d scales), created directly on the GPU."""
    import torch
    import dit4sr_b200
    cfg = dict(SD35_MEDIUM_CONFIG)
    cls = dit4sr_b200.SD3Transformer2DModel_OCRBranch_OCR2HQ2OCR if ocr else dit4sr_b200.SD3Transformer2DModel
    with torch.device(device):
        m = cls(**cfg)
    g = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if ".norm_" in name and p.dim() == 1 and p.numel() == 64:
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, device=device, generator=g))
            elif name.endswith("conv.weight") and p.dim() == 4:      # control_conv (+ the OCR model's two convs)
                p.copy_(0.1 * torch.randn(p.shape, device=device, generator=g))
            else:
                p.copy_(0.02 * torch.randn(p.shape, device=device, generator=g))
    return m.to(torch.bfloat16).requires_grad_(False).eval(), cfg


class Stepper:
    """Synthetic inputs of one workload on this rank + the closure that advances them one denoise step (model forward +
    fused Euler update) through the engine.  Under sequence parallelism the image tensors are this rank's row strips."""

    def __init__(self, eng, B, side, T, dev, seed, sp=None):
        import torch
        from dit4sr_b200.sampler import flow_match_sigmas
        bf = torch.bfloat16
        g = torch.Generator(device=dev).manual_seed(seed)
        x = torch.randn(B, 16, side, side, device=dev, generator=g)
        ctrl = torch.randn(B, 16, side, side, device=dev, generator=g).to(bf)
        self.full = (x, ctrl)
        if sp is not None:
            g0, g1 = sp.row_slice(side // 2)
            x, ctrl = x[:, :, 2 * g0:2 * g1].contiguous(), ctrl[:, :, 2 * g0:2 * g1].contiguous()
        self.eng, self.sp = eng, sp
        self.ctrl = ctrl
        self.enc = (0.5 * torch.randn(B, T, 4096, device=dev, generator=g)).to(bf)
        self.pooled = torch.randn(B, 2048, device=dev, generator=g).to(bf)
        self.sig = flow_match_sigmas(40)
        self.t = torch.empty(B, device=dev, dtype=torch.float32)
        self.cur, self.oth = x, torch.empty_like(x)
        self.i = 0

    def step(self):
        k = self.i % 40
        self.t.fill_(float(self.sig[k]) * 1000.0)
        self.eng.step(self.cur, self.ctrl, self.enc, self.pooled, self.t, out=self.oth,
                      euler=(self.cur, float(self.sig[k + 1]) - float(self.sig[k])), sharded_io=self.sp is not None)
        self.cur, self.oth = self.oth, self.cur
        self.i += 1


def timed_steps(stepper, steps, warmup, barrier, world, dev):
    """`warmup` untimed steps, then exactly `steps` steps between CUDA events, barrier + synchronize on both sides; returns
    the MAX over ranks of the elapsed milliseconds."""
    import torch
    import torch.distributed as dist
    for _ in range(warmup):
        stepper.step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        stepper.step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt.item()
    return ms


def instrumented_step(eng, stepper, rank):
    """One extra direct-launch step with CUDA events around every kernel class (every rank runs it: under sequence
    parallelism the ranks must issue the same exchanges); rank 0's totals are reported."""
    import torch
    eng.profile_kernel("*")
    stepper.step()
    torch.cuda.synchronize()
    bd = eng.profile_result()
    return {k: dict(launches=len(v), total_ms=round(sum(v), 4)) for k, v in bd.items()} if rank == 0 else None


def release(eng):
    """Drop an engine's workspaces and captured graphs (the next workload needs the memory)."""
    import torch
    eng._graphs.clear()
    eng._ws.clear()
    torch.cuda.empty_cache()


def north_star_scaling(model, world, rank, dev, barrier, steps=5, warmup=3):
    """The two partitioned workloads BASELINE.json's north star names, at THIS world size (the driver runs 1, 2, 4, 8):
      c4_ulysses    one 2048px image, Ulysses sequence parallelism over all ranks (plain engine at 1 rank)
      c3_dp_strong  1024px, batch 32 IN TOTAL, data-parallel: this rank's share is 32 / world images, no collective
    Device-timed (CUDA events, max over ranks), `steps` timed steps after `warmup`; each entry names its limiter split."""
    import torch
    out = {}
    # ---- C4
    side, _, T, _, _ = WORKLOADS["c4"]
    N = (side // 2) ** 2
    try:
        sp = None
        if world > 1:
            from dit4sr_b200.parallel import PeerSequenceParallel
            sp = PeerSequenceParallel()
        eng = model.engine(sequence_parallel=sp)
        st = Stepper(eng, 1, side, T, dev, seed=1, sp=sp)
        ms = timed_steps(st, steps, warmup, barrier, world, dev)
        bd = instrumented_step(eng, st, rank)
        gemm_f, attn_f = flops_per_image_step(N, T, folded=eng.fold_lr_chains)
        entry = dict(ranks=world, path="nvlink-peer-memory ulysses, cuda graph" if sp else "single gpu, cuda graph",
                     ms_per_step=round(ms / steps, 3), steps=steps, image_steps_per_sec=round(steps / (ms / 1e3), 4),
                     step_tflop=round((gemm_f + attn_f) / 1e12, 2),
                     tflops_per_gpu=round((gemm_f + attn_f) / world / (ms / steps / 1e3) / 1e12, 1))
        from dit4sr_b200 import ops as _ops
        entry["attention_plan_last_call"] = dict(zip(("variant", "kv_slices", "whole_items", "tail_items"),
                                                     _ops.attention_last_plan()))
        if bd:
            keep = ("attn", "a2a", "halo", "qkv_group", "out_group", "ff1", "ff2", "dwconv", "ln_mod")
            entry["breakdown_ms"] = {k: bd[k]["total_ms"] for k in keep if k in bd}
            entry["breakdown_note"] = ("direct-launch instrumented step on rank 0; a2a = scatter kernels + arrival "
                                       "barriers (includes waiting for the slowest rank), attn = attention kernel with "
                                       "the way back fused into its epilogue")
        if sp is not None:   # parity of the sharded path against the single-GPU engine on the same image (not timed)
            x_full, c_full = st.full
            t = torch.full((1,), 741.44, device=dev)
            v_sp = eng.step(x_full.to(torch.bfloat16), c_full, st.enc, st.pooled, t).float()
            release(eng)
            eng1 = model.engine(sequence_parallel=None)
            v_1 = eng1.step(x_full.to(torch.bfloat16), c_full, st.enc, st.pooled, t).float()
            entry["parity_rel_l2_vs_single_gpu"] = float(((v_sp - v_1).norm() / v_1.norm()).item())
            release(eng1)
            barrier()
            sp.close()
        else:
            release(eng)
        out["c4_ulysses"] = entry
    except Exception as e:   # noqa: BLE001  (never lose the headline line to the extra measurements)
        out["c4_ulysses"] = dict(error=f"{type(e).__name__}: {e}"[:300])
    # ---- C3 (this rank's share of the batch)
    side, Btot, T, _, _ = WORKLOADS["c3"]
    N = (side // 2) ** 2
    try:
        eng = model.engine(sequence_parallel=None)
        Bl = Btot // world
        st = Stepper(eng, Bl, side, T, dev, seed=1 + rank)
        ms = timed_steps(st, steps, warmup, barrier, world, dev)
        gemm_f, attn_f = flops_per_image_step(N, T, folded=eng.fold_lr_chains)
        out["c3_dp_strong"] = dict(ranks=world, global_batch=Btot, per_gpu_batch=Bl, ms_per_step=round(ms / steps, 3),
                                   steps=steps, image_steps_per_sec=round(Btot * steps / (ms / 1e3), 3),
                                   tflops_per_gpu=round((gemm_f + attn_f) * Bl / (ms / steps / 1e3) / 1e12, 1),
                                   collectives_in_step=0)
        release(eng)
    except Exception as e:   # noqa: BLE001
        out["c3_dp_strong"] = dict(error=f"{type(e).__name__}: {e}"[:300])
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from dit4sr_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    side, B, T, mode, desc = WORKLOADS[args.workload]
    if mode == "dp_strong":
        if B % world:
            raise SystemExit(f"workload {args.workload}: batch {B} does not divide over {world} GPUs")
        B //= world
    N = (side // 2) ** 2
    bf = torch.bfloat16

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    model, cfg = build_model_gpu(dev, seed=0, ocr=args.workload == "c5")
    sp = None
    if mode == "sp" and world > 1:
        from dit4sr_b200.parallel import PeerSequenceParallel, SequenceParallel
        sp = SequenceParallel() if os.environ.get("DIT4SR_SP_BACKEND") == "nccl" else PeerSequenceParallel()
    eng = model.engine(sequence_parallel=sp)
    images_total = B * (1 if mode == "sp" else world)       # images advanced per step by the whole job
    rank_seed = 1 if mode == "sp" else 1 + rank              # sequence parallelism: every rank holds the same image
    st = Stepper(eng, B, side, T, dev, rank_seed, sp=sp)

    # ---- device-resident timing (`value`)
    for _ in range(args.warmup):
        st.step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_range = os.environ.get("DIT4SR_PROFILE_RANGE") == "1"   # ncu --profile-from-start off
    if prof_range:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        st.step()
    e1.record()
    barrier()
    if prof_range:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count() - l0
    # dominant kernels: CUDA events around every FF1-group / attention launch of the same steps, run again right after the
    # timed region (same stream, same shapes, same clocks; direct launches because events cannot be recorded inside a
    # graph replay).  Every rank runs them (sequence parallelism: same exchanges everywhere).
    ff1_ms = attn_ms = None
    try:
        eng.profile_kernel("ff1")
        for _ in range(args.steps):
            st.step()
        ff1_ms = eng.profile_result()
        eng.profile_kernel("attn")
        for _ in range(args.steps):
            st.step()
        attn_ms = (eng.profile_result() or {}).get("attn")
    except Exception:   # noqa: BLE001  (never allowed to break the bench line)
        if sp is not None:
            raise
    clocks = sampler.stop() if sampler else None
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt.item()

    # ---- end-to-end through the public module API with host buffers (`e2e`)
    hx = torch.randn(B, 16, side, side).to(bf).pin_memory()
    hc, he, hp = st.full[1].cpu().pin_memory(), st.enc.cpu().pin_memory(), st.pooled.cpu().pin_memory()
    ht = torch.full((B,), 741.44).pin_memory()
    hout = torch.empty(B, 16, side, side, dtype=bf).pin_memory()
    e2e_velocity = {}

    # Input pipeline of the e2e loop: every step's inputs are copied from pinned host memory inside the timed region, on a
    # copy stream into one of two device slots, so that step i+1's 24 MB ride under step i's kernels (what a serving loop
    # does); the forward waits for its slot's copy, the copy waits for the forward that last read the slot.
    hosts = (hx, hc, he, hp, ht)
    dslots = [tuple(torch.empty(h.shape, device=dev, dtype=h.dtype) for h in hosts) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def e2e_prefetch(i):
        sl = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[sl])
            for d, h in zip(dslots[sl], hosts):
                d.copy_(h, non_blocking=True)
            ready[sl].record(copy_stream)

    def e2e_step(i, last):
        sl = i & 1
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[sl])
        if not last:
            e2e_prefetch(i + 1)
        dx, dc, de, dp, dt = dslots[sl]
        with torch.no_grad():
            v = model(hidden_states=dx, controlnet_image=dc, encoder_hidden_states=de, pooled_projections=dp, timestep=dt,
                      return_dict=False)[0]
        consumed[sl].record(cur)
        hout.copy_(v, non_blocking=True)
        e2e_velocity["v"] = v

    nw = max(1, args.warmup)
    e2e_prefetch(0)
    for i in range(nw):
        e2e_step(i, i + 1 == nw)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    copy_stream.wait_event(f0)          # the first copy of the timed region starts inside it
    e2e_prefetch(nw)
    for i in range(nw, nw + args.steps):
        e2e_step(i, i + 1 == nw + args.steps)
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    if world > 1:
        tt = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = tt.item()
    h2d = sum(a.numel() * a.element_size() for a in (hx, hc, he, hp, ht))
    d2h = hout.numel() * hout.element_size()

    # ---- per-kernel-class breakdown of one extra, instrumented step (not part of any timed number)
    breakdown = None
    if not args.no_breakdown:
        breakdown = instrumented_step(eng, st, rank)

    # ---- CPU baseline on rank 0 (bounded sample); on C2 the same timed CPU forward is also the parity reference of
    #      image 0 of the e2e call above (same bf16-rounded weights, same inputs)
    cpu = None
    parity = None
    if rank == 0 and not args.no_cpu:
        torch.cuda.synchronize()
        cpu, parity = cpu_baseline_leg(model if args.workload == "c2" else None, (hx, hc, he, hp, ht), hout)

    # ---- the partitioned north-star workloads at this world size (C2 headline only; placed last in the JSON line)
    ns = None
    if args.workload == "c2" and not args.no_north_star:
        release(eng)
        del st
        ns = north_star_scaling(model, world, rank, dev, barrier)

    if rank == 0:
        peaks = measured_peaks()
        if args.workload == "c5":
            gemm_f, attn_f = flops_per_image_step_ocr(N, T)
        else:
            gemm_f, attn_f = flops_per_image_step(N, T, folded=eng.fold_lr_chains)
        step_flops = (gemm_f + attn_f) * images_total / world      # per-GPU share of one step
        img_steps = images_total * args.steps
        value = img_steps / (ms / 1e3)
        n_layers, n_dual = cfg["num_layers"], len(cfg["dual_attention_layers"])
        # FF1 launch = image FF1 [B*2N,1536]x[1536,6144] grouped with the text FF1 [B*T,...] (absent in the last block)
        streams = 3 if args.workload == "c5" else 2
        rows_local = N // (world if mode == "sp" else 1)
        ff1_flops = 2.0 * (B * streams * rows_local + B * T * (n_layers - 1.0) / n_layers) * 1536 * 6144
        ff1_avg_ms = (sum(ff1_ms["ff1"]) / len(ff1_ms["ff1"])) if ff1_ms and ff1_ms.get("ff1") else None
        # DRAM traffic of the same kernel: NOT measured in this run - read from the committed `ncu --set full` capture
        traffic = None
        traffic_source = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r01d_ncu_full_kernels.json")))
            vals = [float(k["dram__bytes_read.sum [Mbyte]"]) + float(k["dram__bytes_write.sum [Mbyte]"])
                    for k in prof if "gemm_kernel<1>" in k["Kernel Name"]]
            if vals:
                traffic = round(sum(vals) / len(vals) * 1e6)
                traffic_source = ("from committed capture profiles/r01d_ncu_full_kernels.json (not measured in this run): "
                                  "dram__bytes_read.sum + dram__bytes_write.sum per launch of the FF1 image problem alone "
                                  "(M=16384; algorithmic 50.3 MB A + 18.9 MB W read + 201.3 MB written)")
        except (OSError, KeyError, ValueError):
            pass
        roof = None
        if ff1_avg_ms:
            ach = ff1_flops / (ff1_avg_ms / 1e3) / 1e12
            roof = dict(bound="tensor", kernel="gemm_kernel<BIAS_GELU> (FF1 group: image [B*2N,1536]x[1536,6144] + text [B*T,..])",
                        achieved=round(ach, 1),
                        peak=peaks["tf_sustained"], unit="TFLOP/s", frac=round(ach / peaks["tf_sustained"], 4),
                        peak_source=peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                        frac_of_burst=round(ach / peaks["tf_burst"], 4), frac_of_spec_2250=round(ach / 2250.0, 4),
                        launch_ms=round(ff1_avg_ms, 4), flops_per_launch=ff1_flops, launches_timed=len(ff1_ms["ff1"]),
                        traffic=traffic, traffic_source=traffic_source)
        # the dominant kernel by time is the joint attention: it is the `roofline` entry; the largest GEMM follows
        roof_attn = None
        try:
            n_attn = n_layers + n_dual                                    # joint + image-only (dual) calls per step
            if mode == "sp":
                pass   # the 'attn' class also holds the residual-add kernels there: GEMM entry only
            elif attn_ms:
                roof_attn = attention_roofline(attn_ms, attn_f * images_total / world, n_attn, peaks)
            elif breakdown and breakdown.get("attn"):
                per = breakdown["attn"]["total_ms"] / breakdown["attn"]["launches"]
                roof_attn = attention_roofline([per] * breakdown["attn"]["launches"], attn_f * images_total / world,
                                               breakdown["attn"]["launches"], peaks)
        except Exception:   # noqa: BLE001
            roof_attn = None
        if roof_attn is None:
            roof_attn, roof = roof, None                                   # keep a `roofline` entry whatever happens
        line = dict(
            metric="denoise_steps_per_sec", value=round(value, 3), unit=UNIT,
            n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=round(ms / args.steps, 3),
            higher_is_better=True, scaling="weak" if mode == "dp_weak" else "strong", vs_baseline=None, dtype="bf16",
            data="synthetic",
            config=workload_config(args.workload, world),
            batch_steps_per_sec=round(args.steps * (world if mode == "dp_weak" else 1) / (ms / 1e3), 3),
            step_tflop=round(step_flops / 1e12, 3), lr_chains_folded=bool(eng.fold_lr_chains),
            attention_softmax=("pre-scaled bounded scores (QK-RMSNorm bound checked per attention module)"
                               if any(b.attn.score_bound is not None for b in eng.blocks) else "online"),
            step_tflops_achieved=round(step_flops * world / (ms / args.steps / 1e3) / 1e12, 1),   # whole job
            step_frac_of_sustained_peak=round(step_flops / (ms / args.steps / 1e3) / 1e12 / peaks["tf_sustained"], 4),
            e2e=dict(value=round(img_steps / (e2e_ms / 1e3), 3), unit=UNIT, h2d_bytes_per_step=h2d,
                     d2h_bytes_per_step=d2h, ms_per_step=round(e2e_ms / args.steps, 3),
                     api="dit4sr_b200.SD3Transformer2DModel.forward; every step copies its inputs from pinned host memory (double-buffered on a copy stream, one step ahead) and reads the velocity back"),
            gpu_launches=launches, clocks=clocks, roofline=roof_attn, roofline_gemm=roof,
            roofline_hbm=(hbm_rooflines(breakdown, B, N, T, n_layers, n_dual, peaks)
                          if mode != "sp" and args.workload != "c5" else None),
            cpu_baseline=cpu,
            parity=parity, kernel_breakdown_ms=breakdown, north_star_scaling=ns,
        )
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_reference_sample(max_seconds=40.0, steps=3, warmup=1, state_dict=None, inputs=None):
    """Times the oracle (plain-PyTorch fp32 restatement of the reference) on the host cores: B=1, 512px, 24 blocks.
    With `state_dict` / `inputs` (the benchmarked model's weights and image 0 of its batch) the same forward is also the
    parity reference; returns (cpu_baseline entry, velocity)."""
    import torch
    from oracle import dit4sr_oracle as O
    ncpu = os.cpu_count() or 1
    torch.set_num_threads(ncpu)
    if state_dict is None:
        cfg = O.sd35_medium_config(pos_embed_max_size=96)   # the cropped 32x32 window is what the step reads
        sd = {k: torch.empty(shp).normal_(0, 0.02) for k, shp in O.state_dict_manifest(cfg).items()}
        for k in sd:
            if ".norm_" in k and sd[k].dim() == 1 and sd[k].numel() == 64:
                sd[k] = 1 + 0.1 * torch.randn(64)
        inp = O.make_inputs(cfg, 1, 64, 64, 333, seed=1, timestep=741.44)
    else:
        cfg = O.sd35_medium_config(**{k: v for k, v in SD35_MEDIUM_CONFIG.items()})
        sd, inp = state_dict, inputs
    times = []
    t_begin = time.perf_counter()
    v = None
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            v = O.model_forward(sd, cfg, *inp, as_written=True)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            if time.perf_counter() - t_begin > max_seconds and times:
                break
    if not times:
        times = [dt]
    sec = sum(times) / len(times)
    return dict(value=round(1.0 / sec, 4), unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=f"oracle (plain PyTorch fp32 restatement of the reference), B=1 512px 24 blocks, "
                       f"{len(times)} timed step(s) after {warmup} warm-up, {sec:.2f} s/step, host cpu_count={ncpu}"), v


def cpu_baseline_leg(model, host_inputs, host_velocity):
    """The `cpu_baseline` entry and, when `model` is given, the `parity` entry: the CPU restatement of the reference runs
    image 0 of the benchmarked call on the benchmarked model's own (bf16-rounded) weights, and its velocity is compared
    with what the GPU path returned for that image.  The only place of the product arm that executes oracle/."""
    if model is None:
        cpu, _ = cpu_reference_sample()
        return cpu, None
    sd_cpu = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    inp = tuple(t[:1].float() for t in host_inputs)
    cpu, v_ref = cpu_reference_sample(state_dict=sd_cpu, inputs=inp)
    v_gpu = host_velocity[:1].float()
    parity = dict(rel_l2=float(((v_gpu - v_ref).norm() / v_ref.norm()).item()), tolerance=1e-2,
                  what="velocity of image 0 of the benchmarked C2 call (B=8, 24 blocks, bf16, through the module API) "
                       "against the fp32 CPU restatement of the reference on the same weights and inputs")
    return cpu, parity


def run_reference(args):
    """Reference arm: the reference's algorithm on the host CPU (oracle port; diffusers is not installable offline so
    the reference module itself cannot run), all host threads, same metric/unit/config, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import dit4sr_oracle as O
    ncpu = os.cpu_count() or 1
    torch.set_num_threads(ncpu)
    side, B, T, mode, desc = WORKLOADS[args.workload]
    cfg = O.sd35_medium_config(pos_embed_max_size=max(96, side // 2))
    sd = {k: torch.empty(shp).normal_(0, 0.02) for k, shp in O.state_dict_manifest(cfg).items()}
    inp = O.make_inputs(cfg, 1, side, side, T, seed=1, timestep=741.44)
    budget = float(os.environ.get("DIT4SR_REF_BUDGET_S", "240"))
    t_begin = time.perf_counter()
    times = []
    done_w = 0
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            O.model_forward(sd, cfg, *inp, as_written=True)
            dt = time.perf_counter() - t0
            if i < args.warmup and time.perf_counter() - t_begin < budget / 3:
                done_w += 1
                continue
            times.append(dt)
            if time.perf_counter() - t_begin > budget:
                break
    sec = sum(times) / len(times)
    val = round(1.0 / sec, 4)
    unit = UNIT
    world = int(os.environ.get("WORLD_SIZE", "1"))
    print(json.dumps(dict(
        impl="reference", metric="denoise_steps_per_sec", value=val, unit=unit, n_gpus=int(os.environ.get("WORLD_SIZE", "1")),
        steps=len(times), warmup=done_w, ms_per_step=round(sec * 1e3, 1), higher_is_better=True, scaling="weak",
        vs_baseline=None, dtype="f32", data="synthetic",
        config=workload_config(args.workload, world),
        cpu_baseline=dict(value=val, unit=unit, cores=torch.get_num_threads(), kind="port",
                          sample=f"oracle port; each step = ONE image of the batch (B=1) at {side * 8}px through all 24 blocks "
                                 f"on the CPU, {len(times)} timed steps, wall budget {budget:.0f}s"),
        e2e=dict(value=val, unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-breakdown", action="store_true", help="skip the per-kernel breakdown step")
    ap.add_argument("--no-north-star", action="store_true",
                    help="skip the extra C4 (Ulysses) / C3 (strong data-parallel) measurements of the default run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

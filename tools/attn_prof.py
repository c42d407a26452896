"""Developer tool: clock64 phase timers of the joint-attention kernel (one CTA of a late wave), printed per KV step.

  python tools/attn_prof.py --build          (here, no GPU: compiles tools/_timing/libdit4sr_b200.so with -DDIT4SR_ATTN_TIMING)
  python tools/attn_prof.py [--prescaled] [--exp-offload N] [--shape c2|px1024]      (on the GPU box)
"""
import argparse
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_timing")
LIB = os.path.join(OUT, "libdit4sr_b200.so")


def build():
    from dit4sr_b200 import build as b
    os.makedirs(OUT, exist_ok=True)
    objs = []
    procs = []
    for src in b.sources():
        obj = os.path.join(OUT, os.path.basename(src) + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen([b._nvcc()] + b.NVCC_FLAGS + ["-DDIT4SR_ATTN_TIMING", "-c", src, "-o", obj],
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate()
        if p.returncode:
            sys.stderr.write(out)
            raise SystemExit("nvcc failed")
    subprocess.check_call([b._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs)
    print(LIB)


SOFTMAX = {1: "wait S", 2: "TMEM load", 3: "max/exp/sum/pack", 4: "wait PV(j-1)", 5: "O rescale", 6: "P store", 7: "arrive",
           9: "loop"}
ISSUER = {1: "wait V", 2: "wait P", 3: "issue PV+commits", 4: "wait K", 5: "issue QK+commits", 9: "loop"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--prescaled", action="store_true")
    ap.add_argument("--exp-offload", type=int, default=0)
    ap.add_argument("--shape", default="c2")
    ap.add_argument("--variant", type=int, default=0, help="attn: 0 default, 2 rotating buffers, 3 double-buffered 96-key tiles")
    a = ap.parse_args()
    if a.build:
        return build()
    import torch
    from dit4sr_b200 import _lib
    _lib.LIB_PATH = LIB
    from dit4sr_b200 import ops
    B, N, T = (8, 1024, 333) if a.shape == "c2" else (2, 4096, 333)
    D, H = 1536, 24
    S = 2 * N + T
    torch.manual_seed(0)
    qkv = torch.randn(B, S, 3 * D, device="cuda")
    if a.prescaled:
        qkv[..., :D] *= 1.4426950408889634 / 8.0
    qkv = qkv.bfloat16()
    o = torch.empty(B, S, D, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.attention(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], o, H,
                               score_bound=30.0 if a.prescaled else None, exp_offload=a.exp_offload, variant=a.variant)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    buf = (C.c_longlong * (12 * 16))()
    lib = _lib.load()
    lib.dit4sr_attn_timing_read.restype = C.c_int
    rc = lib.dit4sr_attn_timing_read(buf)
    assert rc == 0
    steps = (S + 127) // 128 if a.variant == 2 else (S + 95) // 96
    print(f"shape B={B} S={S} prescaled={a.prescaled} exp_offload={a.exp_offload} variant={a.variant}: {e0.elapsed_time(e1) / 10:.4f} ms "
          f"(instrumented build), {steps} KV steps; cycles per KV step of one CTA:")
    for w in range(12):
        vals = [buf[w * 16 + i] for i in range(10)]
        if not any(vals):
            continue
        names = ISSUER if w in (1, 3) else SOFTMAX
        per = {names[i]: vals[i] / steps for i in names}
        tot = sum(per.values())
        print(f"  warp {w:2d} ({'issuer' if w in (1, 3) else 'softmax tile %d' % ((w - 4) // 4)}): total {tot:7.0f} | " +
              " | ".join(f"{k} {v:.0f}" for k, v in per.items()))


if __name__ == "__main__":
    main()

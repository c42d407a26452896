#!/bin/bash
# Gated-residual epilogue A/B: TMA reduce-adds (L2 atomics) against load / add / store, per GEMM class (out-projection K = 1536,
# FF2 K = 6144), alone (kbench) and inside the C2 step (kernel_breakdown_ms).
cd "$(dirname "$0")/.."
L=gpurun_out/r02f_reduce_ab.log
: > $L
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run DIT4SR_GEMM_REDUCE=1 timeout 60 python tools/kbench.py gemm --reps 20
run DIT4SR_GEMM_REDUCE=0 timeout 60 python tools/kbench.py gemm --reps 20
run DIT4SR_GEMM_REDUCE=1 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-north-star
run DIT4SR_GEMM_REDUCE=0 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-north-star
python - <<'P' >> $L
import json
for l in open("gpurun_out/r02f_reduce_ab.log"):
    if l.startswith('{"metric"'):
        d = json.loads(l)
        print("ms_per_step", d["ms_per_step"], {k: v["total_ms"] for k, v in d["kernel_breakdown_ms"].items()})
P
grep -E "^== |^gemm|^ms_per_step" $L | cut -c1-400

#!/bin/bash
# One GPU call: parity of the companion LayerNorm / skinny GEMM / forked head, their micro-benchmarks, and the C2 step A/B.
cd "$(dirname "$0")/.."
L=gpurun_out/r02e_head_ab.log
: > $L
run() { echo "== $*" >> $L; "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "ln_modulate or gemm or head_tail"
run timeout 400 python -m pytest tests/test_gpu_model.py -q -x
run timeout 60 python tools/kbench.py ln --reps 30 --graph
run timeout 60 python tools/kbench.py skinny --reps 10
run timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-north-star
DIT4SR_HEAD_FORK=0 run timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-north-star --no-breakdown
DIT4SR_HEAD_FORK=0 DIT4SR_GEMM_SKINNY=0 run timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-north-star --no-breakdown
tail -c 3000 $L

#!/bin/bash
# One GPU call: the two attentions of a dual block on two streams (DIT4SR_DUAL_OVERLAP=1) - parity, then the C2 step A/B.
cd "$(dirname "$0")/.."
L=gpurun_out/r02e_dual_ab.log
: > $L
run() { echo "== $*" >> $L; "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
export DIT4SR_DUAL_OVERLAP=1
run timeout 300 python -m pytest tests/test_gpu_model.py -q -x -k "tiny or full_width_model or 24_blocks_velocity or feature_outputs_nchw"
run timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-north-star --no-breakdown
export DIT4SR_DUAL_OVERLAP=0
run timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-north-star --no-breakdown
export DIT4SR_DUAL_OVERLAP=1
run timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-north-star --no-breakdown
export DIT4SR_DUAL_OVERLAP=0
run timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-north-star --no-breakdown
grep -o 'rc=[0-9]*\|"ms_per_step": [0-9.]*, "higher\|passed.*\|failed.*' $L

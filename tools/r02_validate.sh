#!/bin/bash
# Round 2: parity + stress + timing of the default attention instances (model-checked barrier scheme, online and
# pre-scaled softmax selected per call through ABI v3), then the model suite and a C2 bench line.
# Usage: gpurun --timeout 900 -- 'bash tools/r02_validate.sh'   -> gpurun_out/r02_validate.log
mkdir -p gpurun_out
L=gpurun_out/r02_validate.log
: > $L
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x
run timeout 60 python tools/kbench.py attn --reps 30
run timeout 60 python tools/kbench.py attn --reps 30 --prescaled
run timeout 60 python tools/kbench.py attn --reps 30 --prescaled --exp-offload 3
run timeout 60 python tools/kbench.py attn --reps 10 --shape px1024 --prescaled
run timeout 300 python tools/stress_attn.py
run timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_precise.py tests/test_tiling.py tests/test_colorfix.py -q -x
run timeout 200 python bench.py --steps 10 --no-cpu
run DIT4SR_ATTN_PRESCALE=0 timeout 200 python bench.py --steps 10 --no-cpu
grep -E "^== |^rc=|passed|failed|TFLOP/s|ms_per_step|bad" $L | cut -c1-400

#!/bin/bash
# First GPU call of the next round: parity + timing of the opt-in attention instances prepared (statically ranked,
# model-checked, CPU-emulated, but not yet run through the GPU suite) at the end of round 1.
#   XREG=1 setmaxnreg 40/232      XREG=2 + six-barrier scheme with per-step o_done wait
#   XREG=3 + fixed softmax reference (needs DIT4SR_ATTN_SCORE_BOUND)      XREG=4 + pre-scaled scores, p = 2^s (engine folds the scale)
# Usage: gpurun --timeout 900 -- 'bash tools/round2_attention_ab.sh'   -> gpurun_out/r02_attn_ab.log
mkdir -p gpurun_out
L=gpurun_out/r02_attn_ab.log
: > $L
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for x in 0 1 2; do
  run DIT4SR_ATTN_XREG=$x timeout 120 python -m pytest tests/test_gpu_kernels.py -q -x -k attention
  run DIT4SR_ATTN_XREG=$x timeout 60 python tools/kbench.py attn --reps 30
  run DIT4SR_ATTN_XREG=$x timeout 60 python tools/kbench.py attn --reps 10 --shape px1024
  run DIT4SR_ATTN_XREG=$x timeout 200 python tools/stress_attn.py
done
run DIT4SR_ATTN_XREG=3 DIT4SR_ATTN_SCORE_BOUND=40 timeout 120 python -m pytest tests/test_gpu_kernels.py -q -x -k attention
run DIT4SR_ATTN_XREG=3 DIT4SR_ATTN_SCORE_BOUND=40 timeout 60 python tools/kbench.py attn --reps 30
# XREG=4 changes the contract between the projection epilogue and the kernel (q arrives pre-scaled): model-level parity only
for p in 2 4; do run DIT4SR_ATTN_XREG=4 DIT4SR_ATTN_POLY=$p timeout 60 python tools/kbench.py attn --reps 30; done
run DIT4SR_ATTN_XREG=4 timeout 400 python -m pytest tests/test_gpu_model.py -q -x
run DIT4SR_ATTN_XREG=2 timeout 400 python -m pytest tests/test_gpu_model.py -q -x
for x in 0 2 4; do run DIT4SR_ATTN_XREG=$x timeout 200 python bench.py --steps 10 --no-cpu; done
grep -E "^== |^rc=|passed|failed|TFLOP/s|ms_per_step|bad" $L | cut -c1-220

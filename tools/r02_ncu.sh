#!/bin/bash
# Round 2 ncu evidence (one gpurun call, ncu only): launch list of the bench timed region, then a full capture of the
# attention kernel (pre-scaled v3 and v2, and the online form) with source-level stall sampling.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-north-star --no-breakdown"
$B > gpurun_out/r02_ncu_plain_bench.log 2>&1 &&
DIT4SR_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_ncu_launches.csv $B > gpurun_out/r02_ncu_bench.log 2>&1
echo "launch list rc=$?"
K="python tools/kbench.py attn --reps 2 --prescaled"
$K > gpurun_out/r02_ncu_plain_attn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -f -o gpurun_out/r02_attn_v3_prescaled $K \
    > gpurun_out/r02_ncu_attn.log 2>&1
echo "attn v3 rc=$?"
K="python tools/kbench.py attn --reps 2"
$K >> gpurun_out/r02_ncu_plain_attn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -f -o gpurun_out/r02_attn_v3_online $K \
    >> gpurun_out/r02_ncu_attn.log 2>&1
echo "attn online rc=$?"
K="python tools/kbench.py gemm --reps 1"
$K >> gpurun_out/r02_ncu_plain_attn.log 2>&1 &&
ncu --set full --clock-control none -k regex:gemm_kernel -s 4 -c 4 -f -o gpurun_out/r02_gemm_full $K >> gpurun_out/r02_ncu_attn.log 2>&1
echo "gemm rc=$?"
ls -la gpurun_out | tail -12

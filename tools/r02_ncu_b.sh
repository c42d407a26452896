#!/bin/bash
# Round 2 (second half) ncu evidence: launch list of the bench timed region with attention v5, full captures of attention v5
# (pre-scaled, C2 shape), of the LN / stencil kernels and of one 3x3 convolution GEMM of the VAE.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-north-star --no-breakdown"
$B > gpurun_out/r02b_ncu_plain_bench.log 2>&1 &&
DIT4SR_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02b_ncu_launches.csv $B > gpurun_out/r02b_ncu_bench.log 2>&1
echo "launch list rc=$?"
K="python tools/kbench.py attn --reps 2 --prescaled"
$K > gpurun_out/r02b_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -f -o gpurun_out/r02b_attn_v5 $K \
    > gpurun_out/r02b_ncu.log 2>&1
echo "attn v5 rc=$?"
K="python tools/kbench.py ln --reps 1"
$K >> gpurun_out/r02b_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:ln_modulate -s 2 -c 2 -f -o gpurun_out/r02b_ln $K >> gpurun_out/r02b_ncu.log 2>&1
echo "ln rc=$?"
K="python tools/kbench.py dwconv --reps 1"
$K >> gpurun_out/r02b_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:dwconv3x3 -s 2 -c 1 -f -o gpurun_out/r02b_dwconv $K >> gpurun_out/r02b_ncu.log 2>&1
echo "dwconv rc=$?"
K="python tools/vae_bench.py --size 512 --batch 1 --reps 1"
$K >> gpurun_out/r02b_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_vae_launches.csv $K >> gpurun_out/r02b_ncu.log 2>&1
echo "vae rc=$?"
ls -la gpurun_out | tail -8

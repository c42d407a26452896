#!/bin/bash
# Round 2, final code: full GPU suite, smoke(), the default bench line, then ncu (launch list of the bench timed region;
# full captures of attention v5, LayerNorm with its companion problem, the stencil, the GEMM epilogues, one VAE decode list).
# Usage: gpurun --timeout 1500 -- 'bash tools/r02_final.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r02f
timeout 600 python -m pytest tests -m gpu -q -x --durations=8 > ${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py > ${P}_bench_n1.log 2>&1; echo "bench rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-north-star --no-breakdown"
DIT4SR_PROFILE_RANGE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file ${P}_ncu_launches.csv $B > ${P}_ncu_bench.log 2>&1
echo "launch list rc=$?"
K="python tools/kbench.py attn --reps 2 --prescaled"
$K > ${P}_ncu_plain.log 2>&1 &&
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -f -o ${P}_attn_v5 $K > ${P}_ncu.log 2>&1
echo "attn rc=$?"
K="python tools/kbench.py ln --reps 1"
$K >> ${P}_ncu_plain.log 2>&1 &&
timeout 200 ncu --set full --clock-control none -k regex:ln_modulate -s 3 -c 13 -f -o ${P}_ln $K >> ${P}_ncu.log 2>&1
echo "ln rc=$?"
K="python tools/kbench.py dwconv --reps 1"
$K >> ${P}_ncu_plain.log 2>&1 &&
timeout 200 ncu --set full --clock-control none -k regex:dwconv3x3 -s 2 -c 1 -f -o ${P}_dwconv $K >> ${P}_ncu.log 2>&1
echo "dwconv rc=$?"
K="python tools/kbench.py gemm --reps 1"
$K >> ${P}_ncu_plain.log 2>&1 &&
timeout 200 ncu --set full --clock-control none -k regex:gemm_kernel -s 7 -c 9 -f -o ${P}_gemm $K >> ${P}_ncu.log 2>&1
echo "gemm rc=$?"
K="python tools/vae_bench.py --size 512 --batch 1 --reps 1"
$K >> ${P}_ncu_plain.log 2>&1 &&
DIT4SR_VAE_GRAPH=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${P}_vae_launches.csv $K >> ${P}_ncu.log 2>&1
echo "vae rc=$?"
tail -3 ${P}_pytest_gpu.log; tail -2 ${P}_smoke.log; tail -c 1500 ${P}_bench_n1.log

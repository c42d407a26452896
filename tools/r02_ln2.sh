mkdir -p gpurun_out; L=gpurun_out/r02_ln2.log; : > $L
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "ln or layer or modulate or attention_split or norm"
run DIT4SR_LN_VARIANT=8 timeout 60 python tools/kbench.py ln --reps 50
run DIT4SR_LN_VARIANT=9 timeout 60 python tools/kbench.py ln --reps 50
run DIT4SR_LN_VARIANT=8 timeout 60 python tools/kbench.py ln --reps 50 --graph
run DIT4SR_LN_VARIANT=9 timeout 60 python tools/kbench.py ln --reps 50 --graph
run timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-north-star
grep -E "^== |rc=[1-9]|passed|failed|GB/s" $L | cut -c1-300; grep -o '"ms_per_step": [0-9.]*' $L | head -2; grep -o '"kernel_breakdown_ms.*' $L | cut -c1-700

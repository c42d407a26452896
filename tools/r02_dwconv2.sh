mkdir -p gpurun_out; L=gpurun_out/r02_dwconv2.log; : > $L
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "conv"
run DIT4SR_DWCONV_VARIANT=0 timeout 60 python tools/kbench.py dwconv --reps 50 --graph
run DIT4SR_DWCONV_VARIANT=1 timeout 60 python tools/kbench.py dwconv --reps 50 --graph
run DIT4SR_DWCONV_VARIANT=0 timeout 60 python tools/kbench.py dwconv --reps 50 --graph --shape px1024
run DIT4SR_DWCONV_VARIANT=1 timeout 60 python tools/kbench.py dwconv --reps 50 --graph --shape px1024
run timeout 600 python -m pytest tests/test_gpu_model.py -q -x
run timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-north-star
grep -E "^== |rc=[1-9]|passed|failed|GB/s" $L | cut -c1-300; grep -o '"ms_per_step": [0-9.]*' $L | head -2; grep -o '"kernel_breakdown_ms.*' $L | cut -c1-700

mkdir -p gpurun_out; L=gpurun_out/r02_attn_v3.log; : > $L
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k attention
for v in 2 3; do
run timeout 60 python tools/kbench.py attn --reps 30 --variant $v
run timeout 60 python tools/kbench.py attn --reps 30 --prescaled --variant $v
run timeout 60 python tools/kbench.py attn --reps 30 --prescaled --exp-offload 3 --variant $v
run timeout 60 python tools/kbench.py attn --reps 30 --prescaled --exp-offload 1 --variant $v
run timeout 60 python tools/kbench.py attn --reps 10 --shape px1024 --prescaled --variant $v
done
run timeout 60 python tools/attn_prof.py --prescaled --variant 3
run timeout 60 python tools/attn_prof.py --prescaled --variant 3 --exp-offload 3
run timeout 60 python tools/attn_prof.py --variant 3
run timeout 300 python tools/stress_attn.py
grep -E "^== |^rc=|passed|failed|TFLOP/s|bad|warp" $L | cut -c1-330

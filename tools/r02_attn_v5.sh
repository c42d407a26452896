mkdir -p gpurun_out; L=gpurun_out/r02_attn_v5.log; : > $L
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run timeout 400 python -m pytest tests/test_gpu_kernels.py -q -x -k "prescaled or split_kv"
for v in 4 5; do
for e in 1 2 3; do
run timeout 60 python tools/kbench.py attn --reps 30 --prescaled --exp-offload $e --variant $v
done
run timeout 60 python tools/kbench.py attn --reps 10 --shape px1024 --prescaled --variant $v
run timeout 60 python tools/kbench.py attn --reps 10 --shape c4rank --prescaled --variant $v --split 0
done
run timeout 300 python tools/stress_attn.py 5
grep -E "^== |^rc=|passed|failed|TFLOP/s|bad|warp" $L | cut -c1-330

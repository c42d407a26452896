mkdir -p gpurun_out; L=gpurun_out/r02_attn_v5b.log; : > $L
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for v in 4 5; do
run timeout 60 python tools/kbench.py attn --reps 30 --prescaled --variant $v --sustain 3
done
run timeout 60 python tools/kbench.py attn --reps 30 --prescaled --variant 5 --exp-offload 1 --sustain 3
for e in 1 2 3; do
run timeout 60 python tools/attn_prof.py --prescaled --variant 5 --exp-offload $e
done
run timeout 60 python tools/kbench.py gemm --reps 30
cat $L | cut -c1-400

mkdir -p gpurun_out; L=gpurun_out/r02_attn_split_c2.log; : > $L
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for s in -1 0 2 3 4; do
run timeout 60 python tools/kbench.py attn --reps 50 --prescaled --split $s --graph
done
for s in -1 0; do
run timeout 60 python tools/kbench.py attn --reps 50 --prescaled --split $s --graph --shape attn2
run timeout 60 python tools/kbench.py attn --reps 20 --prescaled --split $s --graph --shape px1024
run timeout 60 python tools/kbench.py attn --reps 50 --split $s --graph
done
grep -E "TFLOP/s|rc=[1-9]" $L

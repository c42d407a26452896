cd /root/repo
L=gpurun_out/r02f_pdl_recheck.log
: > $L
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run DIT4SR_PDL=1 timeout 200 python -m pytest tests/test_gpu_model.py -q -s -k "c2_batch8_24_blocks"
run DIT4SR_PDL=0 timeout 200 python -m pytest tests/test_gpu_model.py -q -s -k "c2_batch8_24_blocks"
run DIT4SR_PDL=1 timeout 100 python tools/loop_bench.py --batch 2
run DIT4SR_PDL=1 timeout 100 python tools/loop_bench.py --batch 1
grep -E "^== |rc=|rel-L2|passed|failed|ms/step" $L | cut -c1-330

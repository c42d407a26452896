#!/bin/bash
# Loop-invariant hoisting (Engine.prepare_loop): loop timings at B = 1 and 8 (parity: tests/test_gpu_model.py -k hoisted).
cd "$(dirname "$0")/.."
L=gpurun_out/r02e_loop_b.log
: > $L
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run timeout 200 python tools/loop_bench.py --batch 1
run timeout 200 python tools/loop_bench.py --batch 2
grep -E "^== |^rc=|ms" $L | cut -c1-400

#!/bin/bash
# VAE: parity suite, then decode / encode timings (graphs on / off); also the GEMM suite (ragged MMA N) .
cd "$(dirname "$0")/.."
L=gpurun_out/r02e_vae_gn.log
: > $L
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run timeout 300 python -m pytest tests/test_vae.py -q -x -m gpu
run timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm or head_tail"
for s in 512 1024; do
  run DIT4SR_VAE_GRAPH=1 timeout 100 python tools/vae_bench.py --size $s --batch 1
done
run DIT4SR_VAE_GRAPH=0 timeout 100 python tools/vae_bench.py --size 512 --batch 1
run timeout 100 python tools/vae_bench.py --size 2048 --batch 1
grep -E "^== |^rc=|passed|failed|ms" $L | cut -c1-300

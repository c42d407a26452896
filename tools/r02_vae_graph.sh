#!/bin/bash
# VAE encode / decode replayed from CUDA graphs: parity suite, then decode / encode timings with and without the graphs.
cd "$(dirname "$0")/.."
L=gpurun_out/r02e_vae_graph.log
: > $L
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run timeout 300 python -m pytest tests/test_vae.py -q -x -m gpu
for s in 512 1024; do
  run DIT4SR_VAE_GRAPH=1 timeout 100 python tools/vae_bench.py --size $s --batch 1
  run DIT4SR_VAE_GRAPH=0 timeout 100 python tools/vae_bench.py --size $s --batch 1
done
grep -E "^== |^rc=|passed|failed|ms" $L | cut -c1-300

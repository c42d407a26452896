#!/bin/bash
cd "$(dirname "$0")/.."
L=gpurun_out/r02g_tiled.log
: > $L
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run timeout 200 python -m pytest tests/test_tiling.py -q -x -s -m gpu
run timeout 200 python -m pytest tests/test_vae.py -q -x -s -m gpu -k "tiled"
run timeout 200 python tools/loop_bench.py --batch 1 --side 128 --tile 64 --steps 10
grep -E "^== |^rc=|passed|failed|ms/step|difference|PSNR" $L | cut -c1-300

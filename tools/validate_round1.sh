#!/bin/bash
# Round-end style validation on one box: GPU test suite, bench line, ncu capture of the reduce-add GEMM epilogue.
mkdir -p gpurun_out
timeout 420 python -m pytest tests -q -m gpu > gpurun_out/val_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/val_pytest.log
timeout 170 python bench.py --steps 10 > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err
echo "bench rc=$?"; cut -c1-2500 gpurun_out/val_bench.json; tail -3 gpurun_out/val_bench.err
timeout 110 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel --launch-skip 4 --launch-count 12 \
  -o gpurun_out/r01e_gemm_full -f python tools/kbench.py gemm --reps 1 > gpurun_out/val_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/val_ncu.log

#!/bin/bash
# Sequence-parallel A/B on N GPUs (gpurun --gpus N): parity check, then C4 with / without the fused scatter epilogue and
# over NCCL.   Usage: gpurun --gpus 2 --timeout 900 -- 'bash tools/r02_sp_ab.sh 2'
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/r02_sp_ab_n$N.log
: > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run timeout 300 $TR --master-port 29701 tools/sp_gpu_check.py --side 64 --layers 4
run timeout 300 $TR --master-port 29702 tools/sp_gpu_check.py --side 128 --layers 2
run timeout 300 $TR --master-port 29703 bench.py --gpus $N --workload c4 --steps 10 --no-cpu
run DIT4SR_SP_FUSED_SCATTER=0 timeout 300 $TR --master-port 29704 bench.py --gpus $N --workload c4 --steps 10 --no-cpu
run DIT4SR_SP_BACKEND=nccl timeout 300 $TR --master-port 29705 bench.py --gpus $N --workload c4 --steps 5 --no-cpu
grep -E "^== |^rc=|rank 0|ms_per_step" $L | sed -E 's/.*("ms_per_step": [0-9.]+).*("kernel_breakdown_ms".*)/\1 \2/' | cut -c1-900

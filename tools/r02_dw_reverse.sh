#!/bin/bash
# Stencil walking its tiles in descending order (FF1 wrote the hidden state ascending): parity, then C2 step A/B.
cd "$(dirname "$0")/.."
L=gpurun_out/r02f_dw_reverse.log
: > $L
run() { echo "== $*" >> $L; env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run DIT4SR_DWCONV_REVERSE=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x -k "dwconv"
for r in 1 0 1 0; do
  run DIT4SR_DWCONV_REVERSE=$r timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-north-star
done
python - <<'P' >> $L
import json
for l in open("gpurun_out/r02f_dw_reverse.log"):
    if l.startswith('{"metric"'):
        d = json.loads(l)
        print("ms_per_step", d["ms_per_step"], {k: v["total_ms"] for k, v in d["kernel_breakdown_ms"].items()})
P
grep -E "^== |passed|failed|^ms_per_step" $L | cut -c1-400

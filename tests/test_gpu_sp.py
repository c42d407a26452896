"""Ulysses sequence parallelism on real GPUs (needs >= 2 of them; skipped on a single-GPU box): tools/sp_gpu_check.py
under torchrun compares the torch.distributed path and the NVLink peer-memory path (CUDA-graph captured, attention
epilogue scattering straight into the peers) with the single-GPU engine and with the oracle, for one step and a short
sharded sampling loop.  BASELINE config C4 is this path at 2048px."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world,side,layers", [(2, 64, 4), (4, 128, 2), (8, 128, 2)])
def test_sequence_parallel_matches_single_gpu_and_oracle(world, side, layers):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29610 + world), os.path.join(ROOT, "tools", "sp_gpu_check.py"),
           "--side", str(side), "--layers", str(layers)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]

"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding and the final gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dit4sr_b200.parallel import dp_denoise, shard_range
from oracle import dit4sr_oracle as O


def test_shard_range_partitions_exactly():
    for total in (0, 1, 5, 8, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = O.tiny_config()
    sd = O.init_state_dict(cfg, seed=0)
    hs, lr, enc, pooled, _ = O.make_inputs(cfg, batch, 8, 8, 5, seed=1)

    def denoise_fn(x, c, e, p, steps):   # the oracle stands in for the CUDA sampler on CPU
        return O.sample_loop(sd, cfg, x, c, e, p, num_steps=steps)

    out = dp_denoise(denoise_fn, hs, lr, enc, pooled, num_inference_steps=2)
    if rank == 0:
        ref = O.sample_loop(sd, cfg, hs, lr, enc, pooled, num_steps=2)
        q.put((tuple(out.shape), float((out - ref).abs().max())))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [4, 3])
def test_dp_denoise_matches_single_process(batch):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    shape, err = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert shape == (batch, 16, 8, 8)
    assert err <= 1e-5   # per-image results do not depend on the batch split (fp32 oracle; BLAS batching noise only)

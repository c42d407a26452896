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


# ------------------------------------------------------------------------------------------------------------------
# Ulysses sequence parallelism: the exchange logic, checked on CPU/gloo against single-process math
# ------------------------------------------------------------------------------------------------------------------
def _sp_worker(rank, world, port, q):
    import torch.nn.functional as F
    from dit4sr_b200.parallel import SequenceParallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    sp = SequenceParallel()
    g = torch.Generator().manual_seed(0)
    B, G, H, dh, T, Fc = 2, 8, 4, 64, 5, 16
    N = G * G
    D = H * dh
    # full joint sequence [HQ | LR | text] as the single-process reference sees it
    qkv_img = torch.randn(B, 2, N, 3, H, dh, generator=g)          # stream, token
    qkv_txt = torch.randn(B, T, 3, H, dh, generator=g)
    full = torch.cat([qkv_img[:, 0], qkv_img[:, 1], qkv_txt], dim=1)            # [B, S, 3, H, dh]
    qf, kf, vf = (full[:, :, i].transpose(1, 2) for i in range(3))
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B, 2 * N + T, D)
    # this rank's tokens: contiguous grid rows of both streams
    r0, r1 = sp.row_slice(G)
    loc = torch.cat([qkv_img[:, 0, r0 * G:r1 * G], qkv_img[:, 1, r0 * G:r1 * G]], dim=1)   # [B, 2*Nl, 3, H, dh]
    Nl = (r1 - r0) * G
    xs = sp.seq_to_heads(loc)                                                     # [B, P*2Nl, 3, Hl, dh]
    h0, h1 = sp.head_slice(H)
    joint = torch.cat([xs, qkv_txt[:, :, :, h0:h1]], dim=1)
    q_, k_, v_ = (joint[:, :, i].transpose(1, 2) for i in range(3))
    o = F.scaled_dot_product_attention(q_, k_, v_).transpose(1, 2).reshape(B, joint.shape[1], (h1 - h0) * dh)
    o_img = sp.heads_to_seq(o[:, : world * 2 * Nl].contiguous())                  # [B, 2Nl, D]
    o_txt = sp.gather_heads(o[:, world * 2 * Nl:])                                # [B, T, D]
    ref_img = torch.cat([ref[:, r0 * G:r1 * G], ref[:, N + r0 * G:N + r1 * G]], dim=1)
    e_attn = max(float((o_img - ref_img).abs().max()), float((o_txt - ref[:, 2 * N:]).abs().max()))
    # halo exchange + row-sharded depthwise conv == conv on the whole image
    u_lr = torch.randn(B, N, Fc, generator=g)
    w = torch.randn(Fc, 1, 3, 3, generator=g)
    conv_full = F.conv2d(u_lr.reshape(B, G, G, Fc).permute(0, 3, 1, 2), w, padding=1, groups=Fc)
    padded = sp.halo_rows(u_lr[:, r0 * G:r1 * G].contiguous(), G)                 # [B, (rows+2)*G, F]
    rows = r1 - r0
    img = padded.reshape(B, rows + 2, G, Fc).permute(0, 3, 1, 2)
    conv_loc = F.conv2d(img, w, padding=(0, 1), groups=Fc)                        # valid in y (halo), padded in x
    e_conv = float((conv_loc - conv_full[:, :, r0:r1]).abs().max())
    q.put((rank, e_attn, e_conv))
    dist.destroy_process_group()


def test_sequence_parallel_exchanges_match_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e_attn, e_conv in res:
        assert e_attn <= 1e-5, (rank, e_attn)
        assert e_conv <= 1e-5, (rank, e_conv)

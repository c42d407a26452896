"""Multi-GPU partitioning of the denoise path: one process per GPU (torchrun), `torch.distributed` for the plumbing.

Batch / tile data-parallel sampling (BASELINE configs C3, C5): images are independent, weights are replicated, so the
batch is split into contiguous per-rank slices and every rank runs its own sampling loop; there is NO collective on the
data path -- only an optional all-gather of the final latents.
"""
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of `total` items for `rank` (first `total % world` ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def dp_denoise(denoise_fn: Callable, latents, controlnet_image, encoder_hidden_states, pooled_projections,
               num_inference_steps: int = 40, group: Optional[dist.ProcessGroup] = None, gather: bool = True):
    """Data-parallel sampling. `denoise_fn(latents, control, enc, pooled, steps)` runs the loop on this rank's slice
    (on the GPU: `functools.partial(dit4sr_b200.sampler.denoise, model)`). Inputs are the FULL batch on every rank
    (same tensors / same seed); returns the full batch of final latents if `gather`, else this rank's slice."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = latents.shape[0]
    b0, b1 = shard_range(B, world, rank)
    if b1 > b0:
        local = denoise_fn(latents[b0:b1], controlnet_image[b0:b1], encoder_hidden_states[b0:b1],
                           pooled_projections[b0:b1], num_inference_steps)
    else:
        local = latents[:0].clone()
    if not gather or world == 1:
        return local
    # ragged all-gather: pad every slice to the largest one
    per = (B + world - 1) // world
    pad = torch.zeros((per,) + tuple(latents.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: b1 - b0] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = []
    for r, p in enumerate(parts):
        r0, r1 = shard_range(B, world, r)
        out.append(p[: r1 - r0])
    return torch.cat(out, dim=0)


# ======================================================================================================================
# Ulysses-style sequence parallelism (BASELINE config C4: one 2048px image over P GPUs)
# ======================================================================================================================
class SequenceParallel:
    """Head <-> sequence exchange for joint attention plus the 1-row halo exchange of the depthwise stencil.

    Partitioning: every rank owns `grid/P` contiguous rows of the sqrt(N) x sqrt(N) token grid of BOTH image streams
    (so LayerNorm, GEMMs and the HQ/LR pairing of the conv stay local); the T text tokens are replicated.  Per attention:
      all-to-all #1   [my tokens, all heads]  ->  [all tokens, my H/P heads]      (q, k, v)
      local attention over the full joint sequence on H/P heads                   (text heads: local slice, no comm)
      all-to-all #2   [all tokens, my heads]  ->  [my tokens, all heads]          (o)
      all-gather      text rows of o over the head dimension (before to_add_out)
    Token order on the attention side is rank-major ([r0 tokens | r1 tokens | ... | text]): softmax attention is
    invariant to a permutation of the keys, and the query permutation is undone by all-to-all #2.
    The collectives go through torch.distributed (NCCL over NVLink on the GPUs; gloo in the CPU tests).
    """

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    # ---- attention exchange -------------------------------------------------------------------------------------
    def seq_to_heads(self, t: torch.Tensor) -> torch.Tensor:
        """t [B, Sl, C, H, dh] (my tokens, all heads) -> [B, P*Sl, C, H/P, dh] (all tokens rank-major, my heads)."""
        P = self.world
        B, Sl, C, H, dh = t.shape
        if H % P:
            raise ValueError(f"{H} heads do not divide over {P} ranks")
        Hl = H // P
        send = t.reshape(B, Sl, C, P, Hl, dh).permute(3, 0, 1, 2, 4, 5).contiguous()      # [P, B, Sl, C, Hl, dh]
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        return recv.permute(1, 0, 2, 3, 4, 5).reshape(B, P * Sl, C, Hl, dh)

    def heads_to_seq(self, o: torch.Tensor) -> torch.Tensor:
        """o [B, P*Sl, Hl*dh] (all tokens rank-major, my heads) -> [B, Sl, P*Hl*dh] (my tokens, all heads)."""
        P = self.world
        B, S, Dl = o.shape
        Sl = S // P
        send = o.reshape(B, P, Sl, Dl).permute(1, 0, 2, 3).contiguous()                    # [P, B, Sl, Dl]
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        return recv.permute(1, 2, 0, 3).reshape(B, Sl, P * Dl)

    def gather_heads(self, t: torch.Tensor) -> torch.Tensor:
        """t [B, T, Dl] (replicated tokens, my heads) -> [B, T, P*Dl] (all heads)."""
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t.contiguous(), group=self.group)
        return torch.cat(parts, dim=-1)

    def head_slice(self, H: int) -> Tuple[int, int]:
        Hl = H // self.world
        return self.rank * Hl, (self.rank + 1) * Hl

    # ---- stencil halo -------------------------------------------------------------------------------------------
    def halo_rows(self, u_lr: torch.Tensor, grid_w: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """u_lr [B, rows*grid_w, F]: my rows of the LR hidden state. Returns [B, (rows+2)*grid_w, F] with the neighbours'
        adjacent row above and below (zeros at the image border)."""
        B, n, F = u_lr.shape
        rows = n // grid_w
        if out is None:
            out = torch.empty(B, (rows + 2) * grid_w, F, dtype=u_lr.dtype, device=u_lr.device)
        out[:, grid_w:(rows + 1) * grid_w] = u_lr
        top, bot = out[:, :grid_w], out[:, (rows + 1) * grid_w:]
        ops_, bufs = [], []
        up, down = self.rank - 1, self.rank + 1

        def peer(r):
            return r if self.group is None else dist.get_global_rank(self.group, r)

        if up >= 0:
            s = u_lr[:, :grid_w].contiguous()
            r_ = torch.empty_like(s)
            ops_ += [dist.P2POp(dist.isend, s, peer(up), self.group), dist.P2POp(dist.irecv, r_, peer(up), self.group)]
            bufs.append((top, r_))
        else:
            top.zero_()
        if down < self.world:
            s = u_lr[:, (rows - 1) * grid_w:].contiguous()
            r_ = torch.empty_like(s)
            ops_ += [dist.P2POp(dist.isend, s, peer(down), self.group), dist.P2POp(dist.irecv, r_, peer(down), self.group)]
            bufs.append((bot, r_))
        else:
            bot.zero_()
        if ops_:
            for w in dist.batch_isend_irecv(ops_):
                w.wait()
        for dst, src in bufs:
            dst.copy_(src)
        return out

    # ---- row sharding of images -------------------------------------------------------------------------------------
    def row_slice(self, rows: int) -> Tuple[int, int]:
        if rows % self.world:
            raise ValueError(f"{rows} rows do not divide over {self.world} ranks")
        per = rows // self.world
        return self.rank * per, (self.rank + 1) * per

    def gather_rows(self, t: torch.Tensor, dim: int) -> torch.Tensor:
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t.contiguous(), group=self.group)
        return torch.cat(parts, dim=dim)

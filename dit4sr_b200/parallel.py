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

"""Latent tiling around the denoiser (row N3 of SURVEY 8f; reference: pipelines/pipeline_dit4sr.py:746-761,1449-1558).

Large latents are cut into overlapping square tiles, every tile goes through the model (here: ALL tiles of a step as
one batch -- tiles are independent, so this is the same batch data-parallelism as `parallel.dp_denoise`), and the
per-tile velocities are blended with a gaussian weight window.  Tile placement and weights follow the reference
exactly (including its `cfg`-less call, which crashes there, SURVEY Q7, and works here)."""
from typing import Callable, List, Tuple

import numpy as np
import torch


def tile_offsets(size: int, tile: int, overlap: int) -> List[int]:
    """Start offsets along one axis (pipeline_dit4sr.py:1454-1486): stride tile-overlap, last tile flush with the end."""
    n, cur = 0, 0
    while cur < size:
        cur = max(n * tile - overlap * n, 0) + tile
        n += 1
    return [size - tile if i == n - 1 else max(i * tile - overlap * i, 0) for i in range(n)]


def gaussian_weights(tile_w: int, tile_h: int, device=None, dtype=torch.float64) -> torch.Tensor:
    """pipeline_dit4sr.py:746-761 (`_gaussian_weights`, var 0.01; note the asymmetric midpoints of the original)."""
    var = 0.01
    mid = (tile_w - 1) / 2
    xs = [np.exp(-(x - mid) * (x - mid) / (tile_w * tile_w) / (2 * var)) / np.sqrt(2 * np.pi * var) for x in range(tile_w)]
    mid = tile_h / 2
    ys = [np.exp(-(y - mid) * (y - mid) / (tile_h * tile_h) / (2 * var)) / np.sqrt(2 * np.pi * var) for y in range(tile_h)]
    return torch.tensor(np.outer(ys, xs), device=device, dtype=dtype)


class TilePlan:
    """Tile placement, gaussian window and weight sum of one latent size (everything about the tiling that does not change
    from step to step): `extract` cuts a [B, C, h, w] tensor into the tile batch (batch index = tile * B + image, the
    reference's loop order: rows over x first), `blend` folds per-tile results back with the gaussian weights
    (pipeline_dit4sr.py:1449-1558)."""

    def __init__(self, h: int, w: int, tile: int, overlap: int, device=None):
        self.h, self.w = h, w
        self.tile = tile = min(tile, min(h, w))
        xs, ys = tile_offsets(w, tile, overlap), tile_offsets(h, tile, overlap)
        self.coords: List[Tuple[int, int]] = [(y, x) for x in xs for y in ys]
        self.weights = gaussian_weights(tile, tile, device=device, dtype=torch.float32)
        cnt = torch.zeros(h, w, device=device, dtype=torch.float32)
        for y, x in self.coords:
            cnt[y:y + tile, x:x + tile] += self.weights
        self.weight_sum = cnt

    def __len__(self):
        return len(self.coords)

    def extract(self, t: torch.Tensor) -> torch.Tensor:
        T = self.tile
        return torch.cat([t[:, :, y:y + T, x:x + T] for y, x in self.coords], dim=0)

    def blend(self, vel: torch.Tensor) -> torch.Tensor:
        n, T = len(self.coords), self.tile
        B = vel.shape[0] // n
        acc = torch.zeros(B, vel.shape[1], self.h, self.w, device=vel.device, dtype=torch.float32)
        for i, (y, x) in enumerate(self.coords):
            acc[:, :, y:y + T, x:x + T] += vel[i * B:(i + 1) * B] * self.weights
        return acc / self.weight_sum


def tiled_velocity(model_fn: Callable, latents: torch.Tensor, controlnet_image: torch.Tensor, tile: int, overlap: int,
                   tile_batch: int = 0) -> torch.Tensor:
    """Velocity of a large latent from overlapping tiles. model_fn(latent_tiles, control_tiles) -> velocity tiles
    (batch = tiles x images). tile_batch > 0 caps the number of tiles per model call."""
    B, C, h, w = latents.shape
    plan = TilePlan(h, w, tile, overlap, device=latents.device)
    lat, ctl = plan.extract(latents), plan.extract(controlnet_image)
    step = tile_batch * B if tile_batch > 0 else lat.shape[0]
    outs = [model_fn(lat[i:i + step].contiguous(), ctl[i:i + step].contiguous()) for i in range(0, lat.shape[0], step)]
    return plan.blend(torch.cat(outs, dim=0).float())

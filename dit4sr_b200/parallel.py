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

    peer = False   # True in PeerSequenceParallel: exchanges are CUDA kernels over NVLink peer memory

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


# ======================================================================================================================
# The same partitioning with the exchanges over NVLink peer memory (csrc/sp_comm.cu) instead of torch.distributed
# ======================================================================================================================
class _RawCuda:
    """Minimal __cuda_array_interface__ holder: lets torch alias a raw device allocation without copying."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerSequenceParallel(SequenceParallel):
    """SequenceParallel whose data path never enters a collective library: every exchange buffer is allocated with
    dit4sr_peer_alloc, exported over CUDA IPC and mapped by the peers (torch.distributed only carries the 64-byte handles
    at set-up), and the all-to-all / halo kernels store straight into the destination GPU's memory (include/dit4sr_b200.h,
    "Ulysses sequence parallelism over NVLink peer memory").  The engine can therefore capture the whole sequence-parallel
    step in ONE CUDA graph.  One process per GPU of one NVLink/NVSwitch box, at most 8 ranks."""

    peer = True

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        super().__init__(group)
        import ctypes as C

        from . import _lib
        if self.world > _lib.SP_MAX_RANKS:
            raise ValueError(f"PeerSequenceParallel supports at most {_lib.SP_MAX_RANKS} ranks")
        self._C, self._lib = C, _lib
        self._allocs = []      # local allocations (freed by close())
        self._opened = []      # peer mappings (closed by close())
        self.device = torch.device("cuda", torch.cuda.current_device())
        flags_local, flags_peers = self._alloc_shared_raw(_lib.SP_CHANNELS * _lib.SP_MAX_RANKS * 4)
        epoch = C.c_void_p()
        _lib.check(_lib.load().dit4sr_peer_alloc(_lib.SP_CHANNELS * 4, C.byref(epoch)), "dit4sr_peer_alloc")
        self._allocs.append(epoch.value)
        self.comm = _lib.SpComm()
        self.comm.world, self.comm.rank = self.world, self.rank
        for r, p in enumerate(flags_peers):
            self.comm.flags[r] = p
        self.comm.epoch = epoch.value

    def _alloc_shared_raw(self, nbytes: int):
        """Collective: every rank allocates `nbytes`, the handles go round, everybody maps everybody.  Returns
        (local pointer, [pointer of rank r's allocation as mapped here for r in range(world)])."""
        C, _lib = self._C, self._lib
        lib = _lib.load()
        nbytes = (int(nbytes) + 255) // 256 * 256
        ptr = C.c_void_p()
        _lib.check(lib.dit4sr_peer_alloc(nbytes, C.byref(ptr)), "dit4sr_peer_alloc")
        self._allocs.append(ptr.value)
        handle = C.create_string_buffer(64)
        _lib.check(lib.dit4sr_peer_export(ptr, handle), "dit4sr_peer_export")
        handles = [None] * self.world
        dist.all_gather_object(handles, (bytes(handle.raw), nbytes), group=self.group)
        peers = []
        for r, (h, n) in enumerate(handles):
            if n != nbytes:
                raise RuntimeError(f"PeerSequenceParallel: rank {r} allocated {n} bytes, this rank {nbytes}: the ranks "
                                   "must request the same buffers in the same order")
            if r == self.rank:
                peers.append(ptr.value)
                continue
            q = C.c_void_p()
            _lib.check(lib.dit4sr_peer_open(C.create_string_buffer(h, 64), C.byref(q)), "dit4sr_peer_open")
            self._opened.append(q.value)
            peers.append(q.value)
        return ptr.value, peers

    def shared_empty(self, shape, dtype=torch.bfloat16):
        """Collective. A zero-filled tensor of `shape` on this rank that every peer has mapped: returns (tensor, peers)
        with peers[r] = device pointer of rank r's tensor in THIS process."""
        numel = 1
        for s in shape:
            numel *= int(s)
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        ptr, peers = self._alloc_shared_raw(nbytes)
        flat = torch.as_tensor(_RawCuda(ptr, nbytes), device=self.device)
        if flat.data_ptr() != ptr:
            raise RuntimeError("PeerSequenceParallel: torch copied the peer allocation instead of aliasing it")
        return flat.view(dtype).view(*shape), peers

    def close(self):
        torch.cuda.synchronize()
        lib = self._lib.load()
        for q in self._opened:
            lib.dit4sr_peer_close(q)
        for p in self._allocs:
            lib.dit4sr_peer_free(p)
        self._opened, self._allocs = [], []

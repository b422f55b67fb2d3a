"""Single-node data parallelism for the two places the hot path shards:

* independent units (files for create_dataset, clips for generate): contiguous shards, NO communication --
  `shard_bounds`, `dataset_shard_plan`;
* training: batch-sharded replicas whose gradients are averaged with ONE all-reduce per optimiser step over a flat
  contiguous bucket (NCCL over NVLink on the GPU box; gloo in the CPU tests) -- `FlatGradBucket`.

One process per GPU, `torch.distributed` already initialised by the launcher (torchrun).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch as th
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of `n_items` units for `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def chunks_written(n_samples: int, hop: int = 256, nb_vec: int = 512) -> int:
    """How many `magn_phase_<idx>.pt` files the reference writes for a clip of `n_samples` samples
    (create_dataset.py:41-64): none when T < nb_vec; ONE (empty) chunk when T == nb_vec; else (T-1)//nb_vec."""
    t = 1 + n_samples // hop
    if t < nb_vec:
        return 0
    return max((t - 1) // nb_vec, 1)


def dataset_shard_plan(sample_counts: Sequence[int], rank: int, world_size: int, hop: int = 256, nb_vec: int = 512):
    """Files [begin, end) of this rank and the global index of its first output chunk, so that the union of all
    ranks' outputs carries exactly the idx numbering of the sequential reference loop (an exclusive prefix sum of
    the per-file chunk counts, computed on the host from the sample counts alone)."""
    begin, end = shard_bounds(len(sample_counts), rank, world_size)
    first_idx = sum(chunks_written(n, hop, nb_vec) for n in sample_counts[:begin])
    return begin, end, first_idx


class FlatGradBucket:
    """Gradients of a parameter list viewed as one flat fp32 buffer: `sync()` copies the existing .grad tensors in,
    all-reduces once, scales by 1/world and copies back.  Parameters whose .grad is None (blocks not yet reached by
    the progressive growing) are skipped on every rank alike, so they stay None and Adam keeps ignoring them -- the
    set of active parameters is a function of the growth stage only, identical on all ranks."""

    def __init__(self, params: Iterable[th.nn.Parameter]):
        self.params: List[th.nn.Parameter] = [p for p in params]
        self._flat = None

    def rebuild(self, params: Iterable[th.nn.Parameter]) -> None:       # after next_layer() / add_param_group
        self.params = [p for p in params]
        self._flat = None

    def sync(self) -> int:
        rank, ws = world()
        active = [p for p in self.params if p.grad is not None]
        if ws == 1 or not active:
            return 0
        n = sum(p.grad.numel() for p in active)
        if self._flat is None or self._flat.numel() < n or self._flat.device != active[0].grad.device:
            self._flat = th.empty(n, dtype=th.float32, device=active[0].grad.device)
        flat = self._flat[:n]
        off = 0
        for p in active:
            k = p.grad.numel()
            flat[off:off + k].copy_(p.grad.reshape(-1))
            off += k
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.mul_(1.0 / ws)
        off = 0
        for p in active:
            k = p.grad.numel()
            p.grad.copy_(flat[off:off + k].view_as(p.grad))
            off += k
        return n

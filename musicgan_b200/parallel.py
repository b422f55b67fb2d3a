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
    """Gradients of a parameter list living in ONE flat fp32 buffer: `adopt(grads)` gathers freshly computed gradient
    tensors into the buffer with one multi-tensor copy, all-reduces the buffer once (averaging inside NCCL), and installs
    VIEWS of the buffer as the parameters' .grad -- nothing is copied back.  Parameters whose gradient is None (blocks not
    yet reached by the progressive growing) are skipped on every rank alike, so they stay None and Adam keeps ignoring them
    -- the set of active parameters is a function of the growth stage only, identical on all ranks.

    `sync()` is the same for gradients that already sit in .grad (eager `backward()`)."""

    def __init__(self, params: Iterable[th.nn.Parameter]):
        self.params: List[th.nn.Parameter] = [p for p in params]
        self._flat = None
        self._views = None
        self._key = None

    def rebuild(self, params: Iterable[th.nn.Parameter]) -> None:       # after next_layer() / add_param_group
        self.params = [p for p in params]
        self._flat = self._views = self._key = None

    def _layout(self, active: List[th.nn.Parameter]):
        key = tuple(id(p) for p in active)
        if key != self._key:
            n = sum(p.numel() for p in active)
            self._flat = th.empty(n, dtype=th.float32, device=active[0].device)
            self._views, off = [], 0
            for p in active:
                self._views.append(self._flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            self._key = key
        return self._flat, self._views

    def adopt(self, grads: Sequence) -> int:
        """`grads[i]` is the new gradient of `params[i]` or None.  Returns the number of all-reduced elements."""
        rank, ws = world()
        pairs = [(p, g) for p, g in zip(self.params, grads) if g is not None]
        for p, g in zip(self.params, grads):
            if g is None:
                p.grad = None
        if not pairs:
            return 0
        if ws == 1:
            for p, g in pairs:
                p.grad = g
            return 0
        flat, views = self._layout([p for p, _ in pairs])
        th._foreach_copy_(views, [g.to(th.float32) for _, g in pairs])
        _all_reduce_mean(flat, ws)
        for (p, _), v in zip(pairs, views):
            p.grad = v
        return flat.numel()

    def sync(self) -> int:
        return self.adopt([p.grad for p in self.params])


def _all_reduce_mean(flat: th.Tensor, ws: int) -> None:
    if flat.is_cuda:
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)          # NCCL averages in the collective: no scaling kernel
    else:                                                   # gloo (CPU tests) has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.mul_(1.0 / ws)


def ensure_dir(path: str, not_a_directory_message: str) -> None:
    """Create `path` like the reference's entry points do (one `mkdir`, the parent must exist; a non-directory in the way
    raises NotADirectoryError with the reference's message) -- but safely when EVERY rank of a torchrun launch calls it:
    the reference's test-then-create (create_dataset.py:27-31, generate.py:25-29, train.py:48-52) races between ranks."""
    import os
    try:
        os.mkdir(path)
    except FileExistsError:
        if not os.path.isdir(path):
            raise NotADirectoryError(not_a_directory_message) from None

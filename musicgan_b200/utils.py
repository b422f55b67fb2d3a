"""Host-side orchestration helpers of the training loop (reference utils.py:14-242): the progressive-growing
schedule and the periodic checkpoint writer.  The schedule arithmetic (sample counters, growth thresholds, fade-in
alpha) is integer / scalar host logic and matches the reference exactly; the real-batch transform runs on the GPU."""
from __future__ import annotations

from os.path import join
from typing import List

import torch as th
import torch.nn.functional as F

from . import audio


def antialias_bilinear_matrix(in_size: int, out_size: int, device=None) -> th.Tensor:
    """(out_size, in_size) 1-D resampling matrix of torch's antialiased bilinear interpolation (align_corners=False),
    i.e. of torchvision's tensor Resize that the reference applies to every real batch (utils.py:70-82): a triangle
    filter whose support is the scale factor, weights normalised to sum 1.  The 2-D resize is R_h @ x @ R_w^T -- two
    small GEMMs instead of ATen's gather kernel (which cannot run the 512 -> 4 case on CUDA: shared-memory limit)."""
    scale = in_size / out_size
    support = scale if scale >= 1.0 else 1.0
    inv = 1.0 / scale if scale >= 1.0 else 1.0
    m = th.zeros(out_size, in_size, dtype=th.float64)
    for i in range(out_size):
        center = scale * (i + 0.5)
        lo = max(0, int(center - support + 0.5))
        hi = min(in_size, int(center + support + 0.5))
        j = th.arange(lo, hi, dtype=th.float64)
        w = (1.0 - ((j - center + 0.5) * inv).abs()).clamp_min(0.0)
        m[i, lo:hi] = w / w.sum()
    return m.to(dtype=th.float32, device=device)


class DevicePrefetcher:
    """Iterate over (pinned) host batches and yield device tensors whose host->device copy was issued `depth` steps
    ahead on a separate stream, so the upload of batch t+1 overlaps the compute of batch t (the reference uploads
    synchronously inside the step, train.py:139-140).

    The device buffers are a fixed ring (no allocation per batch: an allocator round trip in the middle of a step costs
    far more than the copy).  A yielded tensor stays valid until the NEXT batch is requested; work enqueued on the
    current stream before that request is ordered before the buffer's reuse."""

    def __init__(self, iterable, device, depth: int = 2):
        import collections
        self.it, self.dev = iter(iterable), th.device(device)
        self.stream = th.cuda.Stream(self.dev)
        self.queue = collections.deque()
        self.depth = max(1, depth)
        self.bufs, self.free_ev = [], []
        self.issued, self.last = 0, None
        for _ in range(self.depth):
            self._issue()

    def _issue(self):
        try:
            host = next(self.it)
        except StopIteration:
            return
        k = self.issued % (self.depth + 2)
        if len(self.bufs) <= k:
            self.bufs.append(None)
            self.free_ev.append(None)
        buf = self.bufs[k]
        if buf is None or buf.shape != host.shape or buf.dtype != host.dtype:
            buf = self.bufs[k] = th.empty(host.shape, dtype=host.dtype, device=self.dev)
        with th.cuda.stream(self.stream):
            if self.free_ev[k] is not None:
                self.stream.wait_event(self.free_ev[k])      # the consumer is done with the previous content
            buf.copy_(host, non_blocking=True)
            ev = th.cuda.Event()
            ev.record(self.stream)
        self.queue.append((k, ev, host))              # `host` kept alive until its copy has been consumed
        self.issued += 1

    def __iter__(self):
        return self

    def __next__(self):
        if not self.queue:
            raise StopIteration
        cur = th.cuda.current_stream(self.dev)
        if self.last is not None:                     # everything enqueued so far used the previous batch at most
            ev = th.cuda.Event()
            ev.record(cur)
            self.free_ev[self.last] = ev
        k, ev, _ = self.queue.popleft()
        cur.wait_event(ev)
        self.last = k
        self._issue()
        return self.bufs[k]

    def __len__(self):
        return len(self.it) if hasattr(self.it, "__len__") else NotImplemented


class Grower:
    """Sample-count driven growth schedule (utils.py:14-86)."""

    def __init__(self, n_grow: int, fadein_lengths: List[int], train_lengths: List[int]):
        assert len(fadein_lengths) == n_grow + 1        # +1: the last layer also fades in
        assert len(train_lengths) == n_grow
        self._n_grow = n_grow
        self._curr_grow = 0
        self._seen = 0                 # samples seen since the start
        self._seen_in_stage = 0        # samples seen since the last growth
        self._downscale = 7
        self._fadein = list(fadein_lengths)
        acc, self._thresholds = 0, []
        for n in train_lengths:        # cumulative sums (utils.py:41-45)
            acc += n
            self._thresholds.append(acc)
        self._resize = {}
        self._norm = audio.ChannelMinMaxNorm()
        self._range = audio.ChangeRange(-1., 1.)

    def grow(self, viewed_samples: int) -> bool:
        self._seen += viewed_samples
        self._seen_in_stage += viewed_samples
        if self._curr_grow >= self._n_grow:
            return False
        if self._thresholds[self._curr_grow] < self._seen:
            self._seen_in_stage = 0
            self._curr_grow += 1
            self._downscale -= 1
            return True
        return False

    @property
    def alpha(self) -> float:
        return min(1., (1. + self._seen_in_stage) / self._fadein[self._curr_grow])

    @property
    def curr_grow(self) -> int:
        return self._curr_grow

    @property
    def target_size(self) -> int:
        return 512 // 2 ** self._downscale

    def scale_transform(self, x: th.Tensor) -> th.Tensor:
        """ChannelMinMaxNorm -> ChangeRange(-1, 1) -> Resize(512 / 2^k) (utils.py:70-82); torchvision's Resize on a
        tensor is bilinear interpolation with antialiasing, i.e. exactly this F.interpolate call."""
        x = self._range(self._norm(x))
        size = self.target_size
        if x.size(-1) != size or x.size(-2) != size:
            key = (x.size(-2), x.size(-1), size, str(x.device))
            if key not in self._resize:
                self._resize[key] = (antialias_bilinear_matrix(x.size(-2), size, x.device),
                                     antialias_bilinear_matrix(x.size(-1), size, x.device))
            r_h, r_w = self._resize[key]
            x = th.matmul(th.matmul(r_h, x), r_w.t())
        return x


class Saver:
    """Every `save_every` calls: the four state dicts under the reference's file names (utils.py:118-145, 209-233).
    The preview PNGs / sounds of the reference (:147-207) need matplotlib and are not part of the hot path."""

    def __init__(self, output_dir: str, save_every: int, rand_channels: int, rand_height: int = 2, rand_width: int = 2):
        self._dir, self._every = output_dir, save_every
        self._counter, self._curr_save = 0, 0

    @property
    def curr_save(self) -> int:
        return self._curr_save - 1          # utils.py:236-239: the last SAVED index

    @property
    def save_counter(self) -> int:
        return self._counter % self._every  # utils.py:241-242

    def request_save(self, gen, disc, optim_gen, optim_disc, alpha: float) -> bool:
        # utils.py:209-233: count first, save when the count is a multiple of save_every (first files at call #save_every)
        self._counter += 1
        if self._counter % self._every == 0:
            k = self._curr_save
            th.save(disc.state_dict(), join(self._dir, f"disc_{k}.pt"))
            th.save(optim_disc.state_dict(), join(self._dir, f"optim_disc_{k}.pt"))
            th.save(gen.state_dict(), join(self._dir, f"gen_{k}.pt"))
            th.save(optim_gen.state_dict(), join(self._dir, f"optim_gen_{k}.pt"))
            self._curr_save += 1
            return True
        return False

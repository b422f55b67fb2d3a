"""Real-batch normalisation helpers with the reference's names and semantics (audio/transforms.py:4-40);
plain tensor expressions that run on whatever device the batch lives on (the trainer keeps it on the GPU)."""
import torch as th


class ChannelMinMaxNorm:
    """Per sample and per channel: (x - min) / (max - min + eps) over the (H, W) plane."""

    def __init__(self, epsilon: float = 1e-8):
        self.epsilon = epsilon

    def __call__(self, x: th.Tensor) -> th.Tensor:
        assert len(x.size()) == 4
        assert x.size()[1] == 2
        flat = x.flatten(2)
        hi = flat.amax(dim=-1)[:, :, None, None]
        lo = flat.amin(dim=-1)[:, :, None, None]
        return (x - lo) / (hi - lo + self.epsilon)


class ChangeRange:
    """x * (upper - lower) + lower."""

    def __init__(self, lower_bond: float, upper_bound: float):
        self.span = upper_bound - lower_bond
        self.start = lower_bond

    def __call__(self, x: th.Tensor) -> th.Tensor:
        return x * self.span + self.start

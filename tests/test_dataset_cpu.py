"""CPU: the packed float32 dataset format beside the reference's one-.pt-per-chunk format
(create_dataset.py:51-64, audio/dataset.py:15-44): same numbering, same values, DataLoader workers, export back."""
import os

import numpy as np
import torch
from torch.utils.data import DataLoader

from musicgan_b200.audio import dataset as ds


def _chunks(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 2, 512, 512, generator=g) * 2 - 1


def test_packed_equals_reference_format(tmp_path):
    a, b = _chunks(3, 1), _chunks(2, 2)
    packed, legacy = tmp_path / "packed", tmp_path / "legacy"
    packed.mkdir(); legacy.mkdir()
    # file 1 -> chunks 0..2, a skipped file (no shard), file 3 -> chunks 3..4: the reference's running idx
    ds.write_packed_shard(str(packed), 3, b)
    ds.write_packed_shard(str(packed), 0, a)
    assert ds.write_packed_index(str(packed)) == 5
    for i, c in enumerate(torch.cat([a, b])):
        torch.save(c.to(torch.float64), str(legacy / f"magn_phase_{i}.pt"))
    dp, dl = ds.AudioDataset(str(packed)), ds.AudioDataset(str(legacy))
    assert len(dp) == len(dl) == 5
    # the reference sorts file names lexicographically; with < 10 chunks that is the numeric order
    for i in range(5):
        x, y = dp[i], dl[i]
        assert x.dtype == torch.float32 and y.dtype == torch.float64 and tuple(x.shape) == (2, 512, 512)
        assert torch.equal(x.double(), y)
    assert torch.equal(dp[-1], dp[4])
    # half the bytes of the reference's float64 pickles
    size = lambda d: sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
    assert size(str(packed)) < 0.51 * size(str(legacy))


def test_packed_dataset_through_dataloader_workers(tmp_path):
    a = _chunks(6, 3)
    ds.write_packed_shard(str(tmp_path), 0, a[:4])
    ds.write_packed_shard(str(tmp_path), 4, a[4:])
    ds.write_packed_index(str(tmp_path))
    loader = DataLoader(ds.AudioDataset(str(tmp_path)), batch_size=3, shuffle=False, num_workers=2, drop_last=True)
    got = torch.cat(list(loader))
    assert torch.equal(got, a)


def test_export_pt_round_trip(tmp_path):
    a = _chunks(2, 4)
    src, out = tmp_path / "p", tmp_path / "pt"
    src.mkdir()
    ds.write_packed_shard(str(src), 7, a)
    ds.write_packed_index(str(src))
    assert ds.export_pt(str(src), str(out)) == 2
    assert sorted(os.listdir(out)) == ["magn_phase_7.pt", "magn_phase_8.pt"]
    t = torch.load(str(out / "magn_phase_8.pt"))
    assert t.dtype == torch.float64 and torch.equal(t, a[1].double())

"""Dataset of the chunks written by create_dataset.

Two on-disk formats, told apart by the presence of `index.json`:

* the reference's (audio/dataset.py:15-44, create_dataset.py:51-64): one float64 (2, 512, 512) `magn_phase_<n>.pt` per
  chunk, read with `th.load` (4 MiB of pickle per sample);
* packed: per input file ONE raw little-endian float32 shard `shard_<first idx>.f32` holding its chunks
  back to back as (n, 2, 512, 512), listed in `index.json` with the global index of their first chunk -- the numbering
  is the reference's.  Items are read through `numpy.memmap` (no pickle, no float64: 2 MiB per sample, page-cache
  friendly, safe with DataLoader worker processes) and come back as float32 tensors.

`export_pt(dataset_path, out_dir)` turns a packed dataset into the reference's `.pt` files for tools that expect them.
"""
import json
import re
from os import listdir, makedirs
from os.path import exists, getsize, isdir, isfile, join

import numpy as np
import torch as th
from torch.utils.data import Dataset

_PATTERN = re.compile(r"^magn_phase_\d+\.pt$")
INDEX = "index.json"
CHUNK_SHAPE = (2, 512, 512)
_CHUNK_ELEMS = 2 * 512 * 512


def write_packed_shard(dataset_path: str, first_idx: int, chunks) -> str:
    """Write the chunks of one input file -- (n, 2, 512, 512), any float dtype -- as one float32 shard.  Returns the file
    name.  The index is written separately (`write_packed_index`) once all shards exist."""
    a = np.ascontiguousarray(chunks.detach().cpu().numpy() if th.is_tensor(chunks) else chunks, dtype="<f4")
    assert a.ndim == 4 and tuple(a.shape[1:]) == CHUNK_SHAPE, a.shape
    name = f"shard_{first_idx:08d}.f32"
    a.tofile(join(dataset_path, name))
    return name


def write_packed_index(dataset_path: str) -> int:
    """(Re)build `index.json` from the shards present (their first index is in the name, their count is their size);
    every rank of a sharded create_dataset run may call it after its own shards are written -- last writer wins with
    the complete list.  Returns the number of chunks."""
    shards = []
    for f in sorted(listdir(dataset_path)):
        m = re.match(r"^shard_(\d+)\.f32$", f)
        if m:
            n_bytes = getsize(join(dataset_path, f))
            assert n_bytes % (_CHUNK_ELEMS * 4) == 0, f
            shards.append({"file": f, "first_idx": int(m.group(1)), "count": n_bytes // (_CHUNK_ELEMS * 4)})
    shards.sort(key=lambda s: s["first_idx"])
    with open(join(dataset_path, INDEX), "w") as fh:
        json.dump({"dtype": "float32", "chunk_shape": list(CHUNK_SHAPE), "shards": shards}, fh)
    return sum(s["count"] for s in shards)


class AudioDataset(Dataset):
    def __init__(self, dataset_path: str) -> None:
        super().__init__()
        assert isdir(dataset_path)
        self._root = dataset_path
        self._packed = exists(join(dataset_path, INDEX))
        if self._packed:
            with open(join(dataset_path, INDEX)) as fh:
                meta = json.load(fh)
            assert meta["dtype"] == "float32" and tuple(meta["chunk_shape"]) == CHUNK_SHAPE
            self._shards = [s for s in meta["shards"] if s["count"] > 0]
            self._starts = np.cumsum([0] + [s["count"] for s in self._shards])
            self._maps = {}
        else:
            names = [f for f in listdir(dataset_path) if isfile(join(dataset_path, f)) and _PATTERN.match(f)]
            self._files = np.array(sorted(names))          # same (lexicographic) order as the reference

    def _map(self, k: int):
        m = self._maps.get(k)              # opened lazily, per process (DataLoader workers re-open their own)
        if m is None:
            s = self._shards[k]
            m = self._maps[k] = np.memmap(join(self._root, s["file"]), dtype="<f4", mode="r", shape=(s["count"],) + CHUNK_SHAPE)
        return m

    def __getitem__(self, index: int):
        if not self._packed:
            return th.load(join(self._root, self._files[index]))
        if index < 0:
            index += len(self)
        k = int(np.searchsorted(self._starts, index, side="right")) - 1
        return th.from_numpy(np.array(self._map(k)[index - self._starts[k]]))      # a private float32 copy

    def __len__(self):
        return int(self._starts[-1]) if self._packed else len(self._files)

    def __getstate__(self):                # memmaps are not pickled into DataLoader workers
        d = dict(self.__dict__)
        if d.get("_packed"):
            d["_maps"] = {}
        return d


def export_pt(dataset_path: str, out_dir: str) -> int:
    """Packed dataset -> the reference's float64 `magn_phase_<idx>.pt` files (create_dataset.py:51-64)."""
    with open(join(dataset_path, INDEX)) as fh:
        meta = json.load(fh)
    makedirs(out_dir, exist_ok=True)
    n = 0
    for s in meta["shards"]:
        if s["count"] == 0:
            continue
        m = np.memmap(join(dataset_path, s["file"]), dtype="<f4", mode="r", shape=(s["count"],) + CHUNK_SHAPE)
        for i in range(s["count"]):
            th.save(th.from_numpy(np.array(m[i])).to(th.float64), join(out_dir, f"magn_phase_{s['first_idx'] + i}.pt"))
            n += 1
    return n

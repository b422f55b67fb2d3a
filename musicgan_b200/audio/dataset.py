"""Dataset of `magn_phase_<n>.pt` chunks written by create_dataset (reference audio/dataset.py:15-44)."""
import re
from os import listdir
from os.path import isdir, isfile, join

import numpy as np
import torch as th
from torch.utils.data import Dataset

_PATTERN = re.compile(r"^magn_phase_\d+\.pt$")


class AudioDataset(Dataset):
    def __init__(self, dataset_path: str) -> None:
        super().__init__()
        assert isdir(dataset_path)
        names = [f for f in listdir(dataset_path) if isfile(join(dataset_path, f)) and _PATTERN.match(f)]
        self._files = np.array(sorted(names))          # same (lexicographic) order as the reference
        self._root = dataset_path

    def __getitem__(self, index: int):
        return th.load(join(self._root, self._files[index]))

    def __len__(self):
        return len(self._files)

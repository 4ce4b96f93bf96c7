"""DAPS noisy/clean enhancement loader -- drop-in for /root/reference/onssen/data/daps_enhance.py:30-126.

Same factory signature `daps_enhance_dataloader(num_batch, feature_options, partition, device=None)` and the same
stateful semantics: an epoch is `num_batch * batch_size` items (:125-126); an item is the next non-overlapping
`frame_length`-frame segment of the CURRENT recording (:117-122); when fewer than `frame_length` frames remain a new
file is popped from the shuffled list at `index % len(list)` (:91-95), the list being re-read when empty (:92-93),
and its clean twin is `<data_path>/clean/<a>_<b>_clean.wav` (:96-97).  Items are `[feature, mag_noisy]`,
`[mag_clean, cos_diff]` (:101-108), batched by a shuffling DataLoader with default collate (:31-35).

Where the reference runs three numpy STFTs per file on the host, a whole recording is featurized by ONE pair of
kernel launches on the device (stft_features with frame_length = all frames) and the segments are slices of that
result; batches are stacked on the device.  As in the reference, a new file's first segment can be shorter than
`frame_length` only if the recording itself is (the reference would then fail in default_collate; here it raises)."""
import os
import random

import numpy as np
import torch

from .. import _lib
from .wsj0_2mix import _opt, _read_wav


class SegmentCursor:
    """The stateful part of daps_dataset.__getitem__ (:90-122), independent of how features are computed:
    `featurize(path) -> (list_of_tensors_input, list_of_tensors_label)` with time on dim 0."""

    def __init__(self, list_path, frame_length, featurize):
        self.list_path, self.frame_length, self.featurize = list_path, frame_length, featurize
        self.length_remaining = 0
        self.input = self.label = None
        self.get_item_list()

    def get_item_list(self):
        with open(self.list_path) as f:
            self.file_list = [line.replace("\n", "") for line in f if line.strip()]
        random.shuffle(self.file_list)

    def next_item(self, index):
        if self.length_remaining < self.frame_length:
            if len(self.file_list) == 0:
                self.get_item_list()
            f_noisy = self.file_list.pop(index % len(self.file_list))
            self.input, self.label = self.featurize(f_noisy)
        T = self.frame_length
        inp, lab = [e[0:T] for e in self.input], [e[0:T] for e in self.label]
        self.input, self.label = [e[T:] for e in self.input], [e[T:] for e in self.label]
        self.length_remaining = self.input[0].shape[0]
        if inp[0].shape[0] != T:
            raise ValueError(f"recording shorter than frame_length={T} frames")
        return inp, lab


class _DapsLoader:
    def __init__(self, num_batch, feature_options, partition, device):
        self.fo = feature_options
        self.num_batch = num_batch
        self.batch_size = _opt(feature_options, "batch_size")
        self.base_path = _opt(feature_options, "data_path")
        self.device = torch.device("cpu") if device is None else torch.device(device)
        self.cursor = SegmentCursor(self.base_path + "/" + partition, _opt(feature_options, "frame_length"),
                                    self._featurize_file)

    def clean_path(self, f_noisy):
        a = os.path.basename(f_noisy).split("_")
        return self.base_path + "/clean/" + a[0] + "_" + a[1] + "_clean.wav"

    def _featurize_file(self, f_noisy):
        sr, n_fft, hop = (_opt(self.fo, k) for k in ("sampling_rate", "window_size", "hop_size"))
        noisy, clean = _read_wav(f_noisy, sr), _read_wav(self.clean_path(f_noisy), sr)
        n = min(len(noisy), len(clean))
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a[:n])[None]).to(self.device)
        frames = 1 + n // hop
        o = _lib.stft_features(to(noisy), to(clean), to(noisy - clean), n_fft, hop,
                               torch.zeros(1, dtype=torch.int32), frames, ("feature", "mag_mix", "mag_s1", "cos_s1"))
        return [o["feature"][0], o["mag_mix"][0]], [o["mag_s1"][0], o["cos_s1"][0]]

    def __len__(self):
        return self.num_batch

    def __iter__(self):
        order = list(range(self.num_batch * self.batch_size))
        random.shuffle(order)                                  # DataLoader(shuffle=True) over __len__ items
        for i in range(0, len(order), self.batch_size):
            items = [self.cursor.next_item(j) for j in order[i:i + self.batch_size]]
            yield ([torch.stack([it[0][k] for it in items]) for k in range(2)],
                   [torch.stack([it[1][k] for it in items]) for k in range(2)])


def daps_enhance_dataloader(num_batch, feature_options, partition, device=None):
    return _DapsLoader(num_batch, feature_options, partition, device)

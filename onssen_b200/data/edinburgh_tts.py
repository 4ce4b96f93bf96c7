"""Edinburgh-TTS noisy/clean loader -- drop-in for /root/reference/onssen/data/edinburgh_tts.py:10-120.

Same factory signature, file-list convention (`<data_path>/<partition>` lists file names under
`noisy_trainset_28spk_wav/` with the clean twin under `clean_trainset_28spk_wav/`, :58-63,68-70), fixed crop
`[:frame_length]` after tiling (:73-82) and per-model label layouts (:91-112); "speaker 2" is the noise
mix - clean (get_stft_from_subtraction, feature_utils.py:24-33).  The batch is featurized on the device by the same
two kernel launches as the wsj0-2mix loader (crop_start = 0)."""
import random

import numpy as np
import torch

from . import feature_utils, wavio
from .wsj0_2mix import _opt, _read_wav


class _EdinburghLoader:
    def __init__(self, model_name, feature_options, partition, device):
        self.model_name = model_name
        self.fo = feature_options
        self.device = torch.device("cpu") if device is None else torch.device(device)
        self.batch_size = _opt(feature_options, "batch_size")
        root = _opt(feature_options, "data_path")
        with open(root + "/" + partition, "r") as f:
            self.file_list = [root + "/noisy_trainset_28spk_wav/" + ln.replace("\n", "") for ln in f if ln.strip()]
        random.shuffle(self.file_list)            # edinburgh_tts.py:65

    def __len__(self):
        return (len(self.file_list) + self.batch_size - 1) // self.batch_size

    def _load(self, names):
        sr = _opt(self.fo, "sampling_rate")
        mix = [_read_wav(fn, sr) for fn in names]
        clean = [_read_wav(fn.replace("/noisy_trainset_28spk_wav", "/clean_trainset_28spk_wav"), sr) for fn in names]
        lengths = np.array([min(len(a), len(b)) for a, b in zip(mix, clean)], dtype=np.int32)
        pitch = int(lengths.max())
        bufs = [np.zeros((len(names), pitch), dtype=np.float32) for _ in range(3)]
        for i, (m, c) in enumerate(zip(mix, clean)):
            n = lengths[i]
            bufs[0][i, :n] = m[:n]
            bufs[1][i, :n] = c[:n]
            bufs[2][i, :n] = m[:n] - c[:n]        # noise as "speaker 2"
        to = (lambda a: torch.from_numpy(a).pin_memory().to(self.device, non_blocking=True)) \
            if self.device.type == "cuda" else torch.from_numpy
        return [to(b) for b in bufs], torch.from_numpy(lengths)

    def _waveform_batches(self, batches):
        """-> ((mix, clean, noise) device waveforms, lengths) per batch; on CUDA the raw PCM is staged by a thread pool
        ahead of the consumer and decoded / resampled on the device (data/wavio.py)"""
        if self.device.type != "cuda":
            for names in batches:
                yield self._load(names)
            return
        pairs = [[(fn, fn.replace("/noisy_trainset_28spk_wav", "/clean_trainset_28spk_wav")) for fn in names]
                 for names in batches]
        sr = _opt(self.fo, "sampling_rate")
        for staged in wavio.PcmStager(pairs):
            (mix, clean), lengths = wavio.device_waveforms(staged, sr, self.device)
            yield (mix, clean, mix - clean), lengths          # noise as "speaker 2" (get_stft_from_subtraction)

    def __iter__(self):
        order = list(range(len(self.file_list)))
        random.shuffle(order)
        fo = self.fo
        batches = [[self.file_list[j] for j in order[i:i + self.batch_size]] for i in range(0, len(order), self.batch_size)]
        for names, ((mix, s1, s2), lengths) in zip(batches, self._waveform_batches(batches)):
            inp, lab = feature_utils.featurize_batch(mix, s1, s2, "dc" if self.model_name == "dc" else self.model_name,
                                                     _opt(fo, "window_size"), _opt(fo, "hop_size"),
                                                     _opt(fo, "frame_length"), _opt(fo, "db_threshold"),
                                                     crop_start=torch.zeros(len(names), dtype=torch.int32),
                                                     lengths=lengths)
            if self.model_name == "dc":
                lab = lab[:1]                     # the reference yields label=[one_hot_label] only (:92)
            yield inp, lab


def edinburgh_tts_dataloader(model_name, feature_options, partition, device=None):
    return _EdinburghLoader(model_name, feature_options, partition, device)

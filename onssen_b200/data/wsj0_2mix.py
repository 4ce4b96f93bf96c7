"""wsj0-2mix loaders -- drop-in for /root/reference/onssen/data/wsj0_2mix.py:26-254 (STFT models only; the
TasNet raw-waveform branch :86-112,216-234 is out of scope).

Same factory signature and directory convention (`<data_path>/wav8k/min/<partition>/{mix,s1,s2}/*.wav`,
:78,89-90) and the same per-model (input_list, label_list) layouts (:137-152).  Differences by design:
  * a whole batch of waveforms is featurized on the DEVICE by two kernel launches (feature_utils.featurize_batch)
    instead of 3 librosa STFTs per item on the host; wav decoding uses scipy.io.wavfile (int16 -> /32768), the
    soundfile/librosa convention;
  * the random crop start is drawn per utterance with numpy's global RNG exactly like :125 (exclusive upper
    bound frames - frame_length after tiling), so a seeded run selects the same frames as the reference;
  * the `tt` loader repairs the reference's undefined `get_ref_sig` (:240): labels are
    [stft_r_mix, stft_i_mix, sig_ref(2, nsample)].
Optional sharding (rank, world_size) splits the file list for one-process-per-GPU runs.
"""
import glob
import random

import numpy as np
import torch

from . import feature_utils, wavio


def _read_wav(fn, sampling_rate):
    """PCM -> mono float32 at `sampling_rate` (feature_utils.py:15-19: librosa.load(sr=None) then
    librosa.core.resample when the file's rate differs).  Resampling stays on the host like in the reference, as a
    polyphase filter (scipy.signal.resample_poly); librosa's default kernel (resampy kaiser_best, version unpinned)
    is a different low-pass, so resampled audio is NOT sample-identical to the reference's: parity unpinned for files
    whose rate differs from the configured one (wsj0-2mix 8 kHz needs no resampling)."""
    from scipy.io import wavfile
    rate, x = wavfile.read(fn)
    if x.dtype == np.int16:
        x = x.astype(np.float32) / 32768.0
    elif x.dtype == np.int32:
        x = x.astype(np.float32) / 2147483648.0
    else:
        x = x.astype(np.float32)
    if x.ndim > 1:
        x = x.mean(axis=1)
    if rate != sampling_rate:
        from math import gcd
        from scipy.signal import resample_poly
        g = gcd(int(rate), int(sampling_rate))
        x = resample_poly(x, int(sampling_rate) // g, int(rate) // g).astype(np.float32)
    return x


def shard_files(files, rank, world_size, batch_size):
    """Rank's share of a sorted file list for one-process-per-GPU runs.  With world_size > 1 the list is first
    wrap-padded (files from the start are repeated) to a multiple of world_size * batch_size, so that EVERY rank runs
    the same number of steps with the same batch size: a rank with one step more would issue gradient all-reduces
    nobody answers (a silent NCCL hang), and cross-rank BatchNorm statistics assume equal per-rank row counts."""
    if world_size <= 1:
        return list(files)
    if not files:
        return []
    unit = world_size * batch_size
    padded = list(files) + [files[i % len(files)] for i in range((-len(files)) % unit)]
    return padded[rank::world_size]


def _opt(feature_options, key):
    return feature_options[key] if isinstance(feature_options, dict) else getattr(feature_options, key)


class _WavBatchLoader:
    """Iterable with __len__, yielding (input_list, label_list) on `device` (contract of :26-37)."""

    def __init__(self, model_name, feature_options, partition, device, batch_size, shuffle, rank=0, world_size=1):
        assert model_name in ("dc", "chimera", "chimera++", "phase"), model_name
        self.model_name = model_name
        self.fo = feature_options
        self.partition = partition
        self.device = torch.device("cpu") if device is None else torch.device(device)
        self.batch_size = batch_size
        self.shuffle = shuffle
        full_path = _opt(feature_options, "data_path") + "/wav8k/min/" + partition + "/mix/*.wav"
        files = sorted(glob.glob(full_path))
        self.file_list = shard_files(files, rank, world_size, batch_size)
        opt = lambda k, d: (feature_options.get(k, d) if isinstance(feature_options, dict)
                            else getattr(feature_options, k, d))
        self.workers = int(opt("num_workers", 8))       # decode threads (extra keys; the reference has no workers)
        self.prefetch = int(opt("prefetch_batches", 2))  # batches staged ahead of the training step

    def __len__(self):
        return (len(self.file_list) + self.batch_size - 1) // self.batch_size

    def _load(self, names):
        sr = _opt(self.fo, "sampling_rate")
        trip = [[_read_wav(fn.replace("/mix", "/" + k) if k != "mix" else fn, sr) for fn in names]
                for k in ("mix", "s1", "s2")]
        lengths = np.array([len(x) for x in trip[0]], dtype=np.int32)
        pitch = int(lengths.max())
        out = []
        for sigs in trip:
            buf = np.zeros((len(names), pitch), dtype=np.float32)
            for i, x in enumerate(sigs):
                buf[i, :len(x)] = x[:lengths[i]]
            out.append(torch.from_numpy(buf).pin_memory().to(self.device, non_blocking=True)
                       if self.device.type == "cuda" else torch.from_numpy(buf))
        return out, torch.from_numpy(lengths)

    def __iter__(self):
        order = list(range(len(self.file_list)))
        if self.shuffle:
            random.shuffle(order)
        batches = [[self.file_list[j] for j in order[i:i + self.batch_size]] for i in range(0, len(order), self.batch_size)]
        if self.device.type != "cuda":
            for names in batches:                      # host decode (the featurizer itself needs a CUDA device)
                (mix, s1, s2), lengths = self._load(names)
                yield self._featurize(mix, s1, s2, lengths)
            return
        # CUDA: raw PCM is staged by a thread pool `prefetch` batches ahead (data/wavio.py); int16 -> float, channel
        # mean and (if the rate differs) resampling run on the device; nothing is decoded on this thread
        triples = [[(fn, fn.replace("/mix", "/s1"), fn.replace("/mix", "/s2")) for fn in names] for names in batches]
        sr = _opt(self.fo, "sampling_rate")
        for staged in wavio.PcmStager(triples, workers=self.workers, depth=self.prefetch):
            (mix, s1, s2), lengths = wavio.device_waveforms(staged, sr, self.device)
            yield self._featurize(mix, s1, s2, lengths)

    def _featurize(self, mix, s1, s2, lengths):
        fo = self.fo
        return feature_utils.featurize_batch(mix, s1, s2, self.model_name, _opt(fo, "window_size"), _opt(fo, "hop_size"),
                                             _opt(fo, "frame_length"), _opt(fo, "db_threshold"), lengths=lengths,
                                             label_dtype=torch.float32)


class _EvalLoader(_WavBatchLoader):
    """full-length features + mixture STFT + reference signals (:170-254, repaired).  Batch 1 like the reference by
    default; `feature_options.eval_batch_size` > 1 (an extra key) zero-pads B utterances of different lengths into one
    batch and appends their sample counts as a 4th label element -- `utils.tester` then runs the model once per batch
    with per-utterance lengths in the recurrence and post-processes every utterance at its own length."""

    def _featurize(self, mix, s1, s2, lengths):
        from .. import _lib
        fo = self.fo
        n_fft, hop = _opt(fo, "window_size"), _opt(fo, "hop_size")
        B = mix.shape[0]
        ns = int(lengths.max())
        frames = 1 + ns // hop
        o = _lib.stft_features(mix, None, None, n_fft, hop, torch.zeros(B, dtype=torch.int32), frames,
                               ["feature", "ph_mix"], lengths=lengths)
        ph = o["ph_mix"]
        sig_ref = torch.stack([s1[:, :ns], s2[:, :ns]], 1)                       # (B, 2, nsample of the longest)
        label = [ph[..., 0].contiguous(), ph[..., 1].contiguous(), sig_ref]
        if B > 1:
            label.append(lengths.to(torch.int32))                                 # host tensor: samples per utterance
        return [o["feature"]], label


def wsj0_2mix_dataloader(model_name, feature_options, partition, device=None, rank=0, world_size=1):
    if partition in ("tr", "cv"):
        return _WavBatchLoader(model_name, feature_options, partition, device, _opt(feature_options, "batch_size"), True,
                               rank, world_size)
    if partition == "tt":
        eb = (feature_options.get("eval_batch_size", 1) if isinstance(feature_options, dict)
              else getattr(feature_options, "eval_batch_size", 1))
        return _EvalLoader(model_name, feature_options, partition, device, int(eb), False, rank, world_size)
    raise ValueError(partition)

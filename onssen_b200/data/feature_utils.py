"""Device featurizer -- batch replacement of /root/reference/onssen/data/feature_utils.py:5-95 and the
crop/label logic of /root/reference/onssen/data/wsj0_2mix.py:114-152.

The reference featurizes one wav at a time on the host (librosa).  Here a whole batch of waveforms already
on the device is turned into the model's (input_list, label_list) by two kernel launches.
"""
import numpy as np
import torch

from .. import _lib

_WANT = {
    "dc": ["feature", "mag_mix", "mag_s1", "mag_s2", "feat_max"],
    "chimera": ["feature", "mag_mix", "mag_s1", "mag_s2", "feat_max"],
    "chimera++": ["feature", "mag_mix", "mag_s1", "mag_s2", "cos_s1", "cos_s2", "feat_max"],
    "phase": ["feature", "mag_mix", "mag_s1", "mag_s2", "ph_mix", "ph_s1", "ph_s2", "feat_max"],
}


def num_crop_starts(nsample, hop_size, frame_length):
    """Exclusive upper bound of the crop start = argument of np.random.randint at wsj0_2mix.py:125
    (after the tiling of wsj0_2mix.py:118-123)."""
    frames = 1 + nsample // hop_size
    if frames <= frame_length:
        frames *= frame_length // frames + 1
    return frames - frame_length


def featurize_batch(wav_mix, wav_s1, wav_s2, model_name, window_size, hop_size, frame_length, db_threshold,
                    crop_start=None, label_dtype=torch.float32, lengths=None):
    """wav_* (B, nsample) fp32 CUDA tensors -> (input_list, label_list) with the reference's per-model layout
    (wsj0_2mix.py:137-152). crop_start: int32 (B,) or None (drawn with numpy's global RNG like the reference)."""
    B, ns = wav_mix.shape
    if crop_start is None:
        lens = [ns] * B if lengths is None else [int(v) for v in lengths]
        crop_start = torch.from_numpy(np.array(
            [np.random.randint(num_crop_starts(n, hop_size, frame_length)) for n in lens], dtype=np.int32))
    o = _lib.stft_features(wav_mix, wav_s1, wav_s2, window_size, hop_size, crop_start, frame_length,
                           _WANT[model_name], lengths=lengths)
    one_hot = _lib.one_hot_vad(o["feature"], o["mag_s1"], o["mag_s2"], o["feat_max"], db_threshold, label_dtype)
    if model_name == "dc":
        return [o["feature"]], [one_hot, o["mag_mix"]]
    if model_name == "chimera":
        return [o["feature"]], [one_hot, o["mag_mix"], o["mag_s1"], o["mag_s2"]]
    if model_name == "chimera++":
        return [o["feature"]], [one_hot, o["mag_mix"], o["mag_s1"], o["mag_s2"], o["cos_s1"], o["cos_s2"]]
    if model_name == "phase":
        return [o["feature"], o["ph_mix"]], [one_hot, o["mag_mix"], o["mag_s1"], o["mag_s2"], o["ph_s1"], o["ph_s2"]]
    raise ValueError(f"unknown model_name {model_name!r}")

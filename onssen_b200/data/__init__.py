"""Data plugins (contract of /root/reference/onssen/data/__init__.py:1-3). The Edinburgh-TTS / DAPS loaders are
"next" rows of SURVEY.md section 8(f); their featurisation is the same featurize_batch call."""
from .feature_utils import featurize_batch, num_crop_starts
from .prefetch import DevicePrefetcher
from .wsj0_2mix import wsj0_2mix_dataloader

__all__ = ["featurize_batch", "num_crop_starts", "wsj0_2mix_dataloader", "DevicePrefetcher"]

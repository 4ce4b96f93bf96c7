"""Data plugins (contract of /root/reference/onssen/data/__init__.py:1-3). The DAPS loader (stateful sequential
chunking, DataLoader-hostile) is out of scope."""
from .feature_utils import featurize_batch, num_crop_starts
from .edinburgh_tts import edinburgh_tts_dataloader
from .prefetch import DevicePrefetcher
from .wsj0_2mix import wsj0_2mix_dataloader

__all__ = ["featurize_batch", "num_crop_starts", "wsj0_2mix_dataloader", "edinburgh_tts_dataloader", "DevicePrefetcher"]

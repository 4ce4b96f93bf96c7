"""Data plugins (contract of /root/reference/onssen/data/__init__.py:1-3)."""
from .feature_utils import featurize_batch, num_crop_starts
from .daps_enhance import daps_enhance_dataloader
from .edinburgh_tts import edinburgh_tts_dataloader
from . import wavio
from .wsj0_2mix import wsj0_2mix_dataloader

__all__ = ["featurize_batch", "num_crop_starts", "wsj0_2mix_dataloader", "edinburgh_tts_dataloader",
           "daps_enhance_dataloader", "wavio"]

"""Data plugins (contract of /root/reference/onssen/data/__init__.py:1-3)."""
from .feature_utils import featurize_batch, num_crop_starts

__all__ = ["featurize_batch", "num_crop_starts"]

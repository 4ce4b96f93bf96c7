"""onssen_b200 -- B200-native (sm_100a) replacement of onssen's STFT-mask separation hot path.

Same plugin surface as the reference package (`onssen.nn`, `onssen.loss`, `onssen.data`, `onssen.utils`,
/root/reference/onssen/__init__.py:1-5) so that `egs/*/run.py` only changes its import line.  All arithmetic
on the hot path runs in hand-written CUDA (libonssen_b200.so, C ABI in include/onssen_b200.h); there is no
CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401
from . import data, loss, nn, utils  # noqa: F401

__all__ = ["data", "loss", "nn", "utils"]

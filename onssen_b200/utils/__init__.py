"""Runtime helpers mirroring /root/reference/onssen/utils/__init__.py:1-9."""
from .basic import AttrDict, AverageMeter, build_optimizer

__all__ = ["AttrDict", "AverageMeter", "build_optimizer"]

"""Runtime helpers mirroring /root/reference/onssen/utils/__init__.py:1-9."""
from .basic import AttrDict, AverageMeter, build_optimizer
from .test import tester, tester_chimera, tester_dc
from .train import trainer
from .optim import Adam, clip_grad_norm_
from . import ddp, dist, experiment, optim

__all__ = ["AttrDict", "AverageMeter", "build_optimizer", "trainer", "tester", "tester_dc", "tester_chimera",
           "dist", "ddp", "optim", "experiment", "Adam", "clip_grad_norm_"]

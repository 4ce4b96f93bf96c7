"""build_optimizer / AverageMeter (behaviour of /root/reference/onssen/utils/basic.py:5-27) and an AttrDict
shim: the reference imports the unmaintained `attrdict` package (egs/*/run.py), which is not installable
here; the scripts only need nested dict + attribute access + item assignment."""
import torch


class AttrDict(dict):
    """dict with attribute access; nested dicts are wrapped on read (run.py:19-29 usage pattern)."""

    def __getattr__(self, name):
        try:
            v = self[name]
        except KeyError:
            raise AttributeError(name)
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
            self[name] = v
        return v

    def __setattr__(self, name, value):
        self[name] = value


def build_optimizer(params, optimizer_options):
    name = optimizer_options["name"]
    lr = optimizer_options["lr"]
    if name == "adam":
        params = list(params)
        if params and all(p.is_cuda for p in params):
            from .optim import Adam          # multi-tensor kernel of libonssen_b200.so (csrc/optim.cu)
            return Adam(params, lr=lr)
        return torch.optim.Adam(params, lr=lr)
    if name == "sgd":
        return torch.optim.SGD(params, lr=lr, momentum=0.9)
    if name == "rmsprop":
        return torch.optim.RMSprop(params, lr=lr)
    raise ValueError(f"unknown optimizer {name!r}")


class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

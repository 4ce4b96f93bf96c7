"""Optimiser step of the training loop on the device, without host synchronisation.

`clip_grad_norm_` and `Adam` are drop-ins for `torch.nn.utils.clip_grad_norm_(params, 5)` and the
`torch.optim.Adam(params, lr)` of /root/reference/onssen/utils/train.py:83-84 and onssen/utils/basic.py:6-7:
same arguments, same state layout (`state[p] = {step, exp_avg, exp_avg_sq}`), same arithmetic (amsgrad off),
but each is a handful of multi-tensor launches of libonssen_b200.so (csrc/optim.cu) instead of ATen's foreach
kernels, and the clip coefficient never leaves the device."""
import torch

from .. import _lib


def _rows(params, with_state=None):
    rows = []
    for p in params:
        g = p.grad
        if not g.is_contiguous():
            p.grad = g = g.contiguous()
        if not (p.is_cuda and g.is_cuda):
            raise _lib.OnssenB200Error("optimiser kernels run on CUDA tensors only (no CPU fallback by design)")
        if g.dtype != torch.float32 or p.dtype != torch.float32 or not p.is_contiguous():
            raise _lib.OnssenB200Error("optimiser kernels need contiguous fp32 parameters and gradients")
        st = (0, 0) if with_state is None else (with_state[p]["exp_avg"].data_ptr(), with_state[p]["exp_avg_sq"].data_ptr())
        rows.append((p.data_ptr(), g.data_ptr(), st[0], st[1]))
    return rows


_clip_tables = {}


def clip_grad_norm_(parameters, max_norm):
    """-> total gradient norm as a 0-dim CUDA tensor (torch returns the same); gradients are scaled in place."""
    params = [p for p in ([parameters] if isinstance(parameters, torch.Tensor) else list(parameters))
              if p.grad is not None]
    if not params:
        return torch.tensor(0.0)
    key = tuple((p.data_ptr(), p.numel()) for p in params)
    tb = _clip_tables.get(key)
    if tb is None:
        _clip_tables.clear()               # one live parameter set at a time is the normal case
        tb = _clip_tables[key] = _lib.OptTables([p.numel() for p in params], params[0].device)
    out2 = torch.empty(2, device=params[0].device, dtype=torch.float32)
    _lib.clip_grad_norm(tb, _rows(params), max_norm, out2)
    return out2[0]


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}

    def __getstate__(self):
        # the device pointer tables are a cache keyed on live tensors: never pickled (checkpoints pickle the optimizer)
        st = super().__getstate__() if hasattr(super(), "__getstate__") else dict(self.__dict__)
        st = dict(st)
        st.pop("_tables", None)
        return st

    def __setstate__(self, state):
        super().__setstate__(state)
        self._tables = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            for p in params:
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            steps = {int(self.state[p]["step"]) for p in params}
            if len(steps) != 1:
                raise _lib.OnssenB200Error("parameters of one group must share the step count")
            step = steps.pop() + 1
            key = (gi, tuple((p.data_ptr(), p.numel()) for p in params))
            tb = self._tables.get(key)
            if tb is None:
                tb = self._tables[key] = _lib.OptTables([p.numel() for p in params], params[0].device)
            b1, b2 = group["betas"]
            _lib.adam_step(tb, _rows(params, self.state), group["lr"], b1, b2, group["eps"], group["weight_decay"], step)
            for p in params:
                self.state[p]["step"] = step
                torch.autograd.graph.increment_version(p)   # packed fp16 weight caches key on _version
        return loss

"""Enhancement mask losses -- drop-in for /root/reference/onssen/loss/loss_mask.py:6-40."""
import torch

from .. import _lib


class _MSE(torch.autograd.Function):
    """nn.MSELoss()(a, b) -> scalar with a hand-written backward w.r.t. a."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return _lib.loss_mse_fwd(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return _lib.loss_mse_bwd(a, b, g), None


def loss_mask_msa(output, label):
    [clean_est] = output
    [mag_clean, cos_diff] = label
    return _MSE.apply(clean_est.float().contiguous(), mag_clean.detach().float().contiguous())


class _L1Psa(torch.autograd.Function):
    """per-utterance L1 of mask*noisy - min(noisy, relu(clean*cos)) with a hand-written backward w.r.t. mask."""

    @staticmethod
    def forward(ctx, mask, noisy, clean, cosd):
        ctx.save_for_backward(mask, noisy, clean, cosd)
        return _lib.loss_l1_psa_fwd(mask, noisy, clean, cosd)

    @staticmethod
    def backward(ctx, g):
        mask, noisy, clean, cosd = ctx.saved_tensors
        return _lib.loss_l1_psa_bwd(mask, noisy, clean, cosd, g), None, None, None


def loss_mask_psa(output, label):
    [mask] = output
    [mag_noisy, mag_clean, cos_diff] = label
    c = lambda t: t.detach().float().contiguous()
    return _L1Psa.apply(mask.float().contiguous(), c(mag_noisy), c(mag_clean), c(cos_diff))

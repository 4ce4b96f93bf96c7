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


def loss_mask_psa(output, label):
    [mask] = output
    [mag_noisy, mag_clean, cos_diff] = label
    c = lambda t: t.detach().float().contiguous()
    if torch.is_grad_enabled() and mask.requires_grad:
        raise NotImplementedError("loss_mask_psa backward is not part of this build (use loss_mask_msa for training)")
    return _lib.loss_l1_psa_fwd(c(mask), c(mag_noisy), c(mag_clean), c(cos_diff))

"""chimera / chimera++ losses -- drop-in for /root/reference/onssen/loss/loss_chimera.py:6-59:
0.975 * loss_dc (B,B) + 0.025 * PIT-L1 mask loss (B,) -> broadcast (B,B)."""
import torch

from .. import _lib
from .loss_dc import loss_dc


def _mask_args(mask_A, mask_B):
    """the reference's masks are strided views masks[..., k] of one (B,T,F,S) tensor; accept both layouts"""
    if mask_A.is_contiguous() and mask_B.is_contiguous():
        return mask_A, mask_B, 1
    stride = mask_A.stride(-1)
    exp = (mask_A.shape[1] * mask_A.shape[2] * stride, mask_A.shape[2] * stride, stride)
    if tuple(mask_A.stride()) != exp or tuple(mask_B.stride()) != exp:
        return mask_A.contiguous(), mask_B.contiguous(), 1
    return mask_A, mask_B, stride


class _PitL1(torch.autograd.Function):
    """PIT L1 mask loss with a hand-written backward (sign(mask*mix - target) * mix under the chosen permutation)."""

    @staticmethod
    def forward(ctx, mask_A, mask_B, mag_mix, mag_s1, mag_s2, cos_s1, cos_s2):
        ma, mb, stride = _mask_args(mask_A, mask_B)
        out, perm = _lib.loss_pit_l1_fwd(ma, mb, stride, mag_mix, mag_s1, mag_s2, cos_s1, cos_s2)
        ctx.save_for_backward(ma, mb, mag_mix, mag_s1, mag_s2, cos_s1, cos_s2, perm)
        ctx.stride = stride
        ctx.mark_non_differentiable(perm)
        return out, perm

    @staticmethod
    def backward(ctx, g, _g_perm):
        ma, mb, mix, s1, s2, c1, c2, perm = ctx.saved_tensors
        da, db = _lib.loss_pit_l1_bwd(ma, mb, ctx.stride, mix, s1, s2, c1, c2, perm, g)
        return da, db, None, None, None, None, None


def _pit(mask_A, mask_B, mag_mix, mag_s1, mag_s2, cos_s1=None, cos_s2=None, want_perm=False):
    c = lambda t: None if t is None else t.detach().float().contiguous()
    out, perm = _PitL1.apply(mask_A, mask_B, c(mag_mix), c(mag_s1), c(mag_s2), c(cos_s1), c(cos_s2))
    return (out, perm) if want_perm else out


def loss_chimera_msa(output, label):
    [embedding, mask_A, mask_B] = output
    [one_hot_label, mag_mix, mag_s1, mag_s2] = label
    loss_embedding = loss_dc([embedding], [one_hot_label, mag_mix])
    loss_mask = _pit(mask_A, mask_B, mag_mix, mag_s1, mag_s2)
    return loss_embedding * 0.975 + loss_mask * 0.025


def loss_chimera_psa(output, label):
    [embedding, mask_A, mask_B] = output
    [one_hot_label, mag_mix, mag_s1, mag_s2, cos_s1, cos_s2] = label
    loss_embedding = loss_dc([embedding], [one_hot_label, mag_mix])
    loss_mask = _pit(mask_A, mask_B, mag_mix, mag_s1, mag_s2, cos_s1, cos_s2)
    return loss_embedding * 0.975 + loss_mask * 0.025

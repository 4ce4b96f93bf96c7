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


def _pit(mask_A, mask_B, mag_mix, mag_s1, mag_s2, cos_s1=None, cos_s2=None):
    mask_A, mask_B, stride = _mask_args(mask_A, mask_B)
    c = lambda t: None if t is None else t.float().contiguous()
    out, _ = _lib.loss_pit_l1_fwd(mask_A, mask_B, stride, c(mag_mix), c(mag_s1), c(mag_s2), c(cos_s1), c(cos_s2))
    return out


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

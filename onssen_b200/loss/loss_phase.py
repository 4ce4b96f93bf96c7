"""loss_phase -- REPAIRED restatement of /root/reference/onssen/loss/loss_phase.py:6-37 (the reference raises:
it asserts 6 outputs then unpacks 5 at :7,9 and calls loss_dc with mag_mix on the wrong side at :13).
Repairs: 5 outputs; loss_dc([embedding], [one_hot_label, mag_mix]).  The phase term uses the
permutation chosen by the mask loss (:21-24,33-35)."""
import torch

from .. import _lib
from .loss_dc import loss_dc
from .loss_chimera import _pit


class _PhaseCos(torch.autograd.Function):
    """-sum_n mag_mix * cos_sim(phase_est, phase_ref) under the permutation chosen by the mask loss, (B,)."""

    @staticmethod
    def forward(ctx, phase_A, phase_B, phase_s1, phase_s2, mag_mix, perm):
        ctx.save_for_backward(phase_A, phase_B, phase_s1, phase_s2, mag_mix, perm)
        return _lib.loss_phase_cos_fwd(phase_A, phase_B, phase_s1, phase_s2, mag_mix, perm)

    @staticmethod
    def backward(ctx, g):
        pa, pb, s1, s2, mix, perm = ctx.saved_tensors
        d_pa, d_pb = _lib.loss_phase_cos_bwd(pa, pb, s1, s2, mix, perm, g)
        return d_pa, d_pb, None, None, None, None


def loss_phase(output, label):
    assert len(output) == 5, "There must be 5 tensors in the output"
    assert len(label) == 6, "There must be 6 tensors in the label"
    [embedding, mask_A, mask_B, phase_A, phase_B] = output
    [one_hot_label, mag_mix, mag_s1, mag_s2, phase_s1, phase_s2] = label
    loss_embedding = loss_dc([embedding], [one_hot_label, mag_mix])
    c = lambda t: t.detach().float().contiguous()
    loss_mask, perm = _pit(mask_A, mask_B, mag_mix, mag_s1, mag_s2, want_perm=True)
    loss_ph = _PhaseCos.apply(phase_A.float().contiguous(), phase_B.float().contiguous(), c(phase_s1), c(phase_s2),
                              c(mag_mix), perm)
    return loss_embedding * 0.975 + loss_mask * 0.025 + loss_ph * 0.025

"""loss_phase -- REPAIRED restatement of /root/reference/onssen/loss/loss_phase.py:6-37 (the reference raises:
it asserts 6 outputs then unpacks 5 at :7,9 and calls loss_dc with mag_mix on the wrong side at :13).
Repairs: 5 outputs; loss_dc([embedding], [one_hot_label, mag_mix]).  The phase term uses the
permutation chosen by the mask loss (:21-24,33-35)."""
from .. import _lib
from .loss_dc import loss_dc
from .loss_chimera import _mask_args


def loss_phase(output, label):
    assert len(output) == 5, "There must be 5 tensors in the output"
    assert len(label) == 6, "There must be 6 tensors in the label"
    [embedding, mask_A, mask_B, phase_A, phase_B] = output
    [one_hot_label, mag_mix, mag_s1, mag_s2, phase_s1, phase_s2] = label
    loss_embedding = loss_dc([embedding], [one_hot_label, mag_mix])
    c = lambda t: t.float().contiguous()
    ma, mb, stride = _mask_args(mask_A, mask_B)
    loss_mask, perm = _lib.loss_pit_l1_fwd(ma, mb, stride, c(mag_mix), c(mag_s1), c(mag_s2))
    loss_ph = _lib.loss_phase_cos_fwd(c(phase_A), c(phase_B), c(phase_s1), c(phase_s2), c(mag_mix), perm)
    return loss_embedding * 0.975 + loss_mask * 0.025 + loss_ph * 0.025

"""Deep-clustering affinity loss on the device -- drop-in for /root/reference/onssen/loss/loss_dc.py:6-44
(same asserts, same un-squared Frobenius norms, same (B,B) return shape)."""
import torch

from .. import _lib


def loss_dc(output, label):
    assert len(output) == 1, "Number of output must be 1 for Deep Clustering"
    assert len(label) == 2, "Number of label must be 2 for Deep Clustering"
    embedding, = output
    label, mag_mix = label
    if torch.is_grad_enabled() and embedding.requires_grad:
        raise NotImplementedError("loss_dc backward kernel is not part of this build yet (no autograd fallback)")
    B, T, F, S = label.shape
    D = embedding.shape[-1]
    emb = embedding.contiguous().view(B, T * F, D)
    lab = label.contiguous().view(B, T * F, S)
    mag = mag_mix.detach().float().contiguous().view(B, T * F)
    loss_bb, _, _ = _lib.loss_dc_fwd(emb, lab, mag)
    return loss_bb

"""Deep-clustering affinity loss on the device -- drop-in for /root/reference/onssen/loss/loss_dc.py:6-44
(same asserts, same un-squared Frobenius norms, same (B,B) return shape), differentiable w.r.t. the embedding
through a hand-written backward kernel (no autograd graph of torch ops)."""
import torch

from .. import _lib


class _LossDC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, lab, mag):
        loss_bb, _, _, rec = _lib.loss_dc_fwd(emb, lab, mag, return_record=True)
        ctx.save_for_backward(emb, lab, mag, rec)
        return loss_bb

    @staticmethod
    def backward(ctx, g_bb):
        emb, lab, mag, rec = ctx.saved_tensors
        return _lib.loss_dc_bwd(emb, lab, mag, rec, g_bb), None, None


def loss_dc(output, label):
    assert len(output) == 1, "Number of output must be 1 for Deep Clustering"
    assert len(label) == 2, "Number of label must be 2 for Deep Clustering"
    embedding, = output
    label, mag_mix = label
    B, T, F, S = label.shape
    D = embedding.shape[-1]
    if label.dtype not in (torch.float32, torch.float64, torch.uint8):
        label = label.float()
    emb = embedding.contiguous().view(B, T * F, D)
    lab = label.contiguous().view(B, T * F, S)
    mag = mag_mix.detach().float().contiguous().view(B, T * F)
    return _LossDC.apply(emb, lab, mag)

"""Deep-clustering affinity loss on the device -- drop-in for /root/reference/onssen/loss/loss_dc.py:6-44
(same asserts, same un-squared Frobenius norms, same (B,B) return shape), differentiable w.r.t. the embedding
through a hand-written backward kernel (no autograd graph of torch ops)."""
import torch

from .. import _lib


GLOBAL_MEAN = [False]      # utils.ddp.enable_global_loss_mean


def _global_first_factor(msum):
    """(B,) per-utterance sum of mag_mix -> the mean over the WHOLE data-parallel batch (one scalar all-reduce), or
    None when not distributed."""
    import torch.distributed as dist
    if not (GLOBAL_MEAN[0] and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return None
    # (the count is produced by a fill kernel: torch.tensor(x, device=...) is a pageable host-to-device copy, which
    # makes the host wait for the stream in the middle of the step and leaves the GPU idle while the backward's launches
    # are enqueued -- +0.5 ms per training step at cfg2)
    t = torch.stack([msum.double().sum(), torch.full((), float(msum.numel()), device=msum.device, dtype=torch.float64)])
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return (t[0] / t[1]).float()


class _LossDC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, lab, mag):
        loss_bb, l, msum, rec = _lib.loss_dc_fwd(emb, lab, mag, return_record=True)
        mbar = _global_first_factor(msum)
        if mbar is None:
            ctx.save_for_backward(emb, lab, mag, rec)
            ctx.coupled = False
            return loss_bb
        # element [i,j] = (global mean of sum m) * l_j : same (B,B) shape, and its mean averaged over ranks is the
        # single-device mean_i(sum m_i) * mean_j(l_j) of the whole batch
        ctx.save_for_backward(emb, lab, mag, rec, mbar / msum)
        ctx.coupled = True
        return (mbar * l).unsqueeze(0).expand(l.numel(), l.numel()).contiguous()

    @staticmethod
    def backward(ctx, g_bb):
        if ctx.coupled:
            emb, lab, mag, rec, ratio = ctx.saved_tensors
            g_bb = g_bb * ratio.unsqueeze(1)          # the kernel forms sum_i g[i][j] * msum_i
        else:
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

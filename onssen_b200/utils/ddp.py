"""Gradient synchronisation for one-process-per-GPU data parallel training (the reference has none; SURVEY.md
section 8e): bucketed all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests) launched as soon as a
group of gradients is final -- head/BN first, then each BLSTM layer top-down -- (overlap mode) or as one grouped
launch after the backward (the NCCL default, see GradSync); results are averaged over ranks.

    sync = GradSync()                 # after dist.init_process_group
    model.grad_sync = sync            # DCFunction.backward feeds it bucket by bucket
    loss.backward(); sync.wait()      # gradients in .grad are now the rank average
"""
import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, group=None, overlap=None):
        """overlap: launch each bucket's all-reduce as soon as it is final (beside the remaining BPTT launches) instead
        of one grouped all-reduce after the backward.  Default: off for NCCL.  Measured on 2 x B200 at cfg2
        (scripts/ddp_train_steps.py): the persistent recurrence kernels are cooperative launches that need 114-120 free
        SMs, so an all-reduce kernel that is resident when one of them is launched (or the reverse) makes one wait for
        the other to drain, and each rank's NCCL kernel spins on its peer meanwhile: overlapped 12.85 ms/step vs 11.77
        on one GPU, while the whole 108 MB all-reduce costs ~0.3 ms when it runs alone.  ONSSEN_DDP_OVERLAP=1 turns the
        overlap back on (it pays when the collective is slow: gloo, PCIe)."""
        import os
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.pending = []
        self.deferred = []
        self.bytes_reduced = 0
        self.nccl = dist.is_initialized() and dist.get_backend(group) == "nccl"
        if overlap is None:
            overlap = (not self.nccl) or os.environ.get("ONSSEN_DDP_OVERLAP", "0") == "1"
        self.overlap = bool(overlap)
        if self.nccl and self.world > 1 and self.overlap:
            # leave the collective's share of the SMs free (two batch slices instead of three in the BPTT kernel)
            from .. import _lib
            _lib.load().onssen_blstm_rec_bwd_set_sm_reserve(int(os.environ.get("ONSSEN_DDP_SM_RESERVE", "48")))

    def _launch(self, names, grads):
        """one collective over a bucket -> an entry of self.pending"""
        tensors = [grads[n] for n in names]
        if self.nccl:
            # ONE grouped launch of in-place AVG all-reduces (ncclGroupStart/End through torch's coalescing manager): no
            # flatten copy, no copy-back, no divide pass
            with dist._coalescing_manager(group=self.group, device=tensors[0].device, async_ops=True) as cm:
                for t in tensors:
                    dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
            return (cm, None, names, grads)
        # other backends (gloo in the CPU tests): flatten, SUM, and write the average back in wait()
        flat = torch.cat([t.reshape(-1) for t in tensors])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return (work, flat, names, grads)

    def reduce_bucket(self, grads):
        """grads: dict name -> tensor (final).  Launched now (overlap) or held back until flush() / wait()."""
        if self.world == 1 or not grads:
            return
        names = sorted(grads)
        self.bytes_reduced += sum(grads[n].numel() * grads[n].element_size() for n in names)
        if self.overlap:
            self.pending.append(self._launch(names, grads))
        else:
            self.deferred.append((names, grads))

    def flush(self):
        """Launch the all-reduce of the buckets held back so far as ONE collective (called by the backward once its
        last cooperative kernel is enqueued: the collective then runs beside the remaining weight-gradient GEMMs)."""
        if self.deferred:
            merged = {}
            for names, grads in self.deferred:
                for n in names:
                    merged[n] = grads[n]
            self.deferred = []
            self.pending.append(self._launch(sorted(merged), merged))

    def wait(self):
        """Blocks (the stream, for NCCL) until every bucket is reduced; afterwards the tensors hold the rank average."""
        self.flush()
        for work, flat, names, grads in self.pending:
            work.wait()
            if flat is None:
                continue
            off = 0
            for n in names:
                g = grads[n]
                k = g.numel()
                g.copy_(flat[off:off + k].view_as(g) / self.world)
                off += k
        self.pending = []


def broadcast_parameters(model, src=0):
    """Start every rank from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.detach(), src)
            torch.autograd.graph.increment_version(t)     # packed fp16 weight caches key on (data_ptr, _version)


def enable_sync_batchnorm(model, on=True):
    """Train-mode BatchNorm statistics (and their backward) over the WHOLE data-parallel batch: one all-reduce of
    2 x 2Hp doubles per BatchNorm call, forward and backward.  Without it each rank normalises with its own shard's
    statistics, which differs from the single-device arithmetic of deep_clustering.py:36-38.
    Equal per-rank batch sizes are assumed."""
    for m in model.modules():
        if hasattr(m, "bn"):
            m.sync_bn = bool(on)
    return model


def enable_global_loss_mean(on=True):
    """Make `torch.mean(loss_dc(...))` (and the chimera / phase losses built on it) on every rank contribute to the
    GLOBAL-batch scalar of the single-device arithmetic: the reference's (B,B) mean is mean_i(sum m_i) * mean_j(l_j)
    (loss_dc.py:44 + train.py:79), a product of two batch means, so the average of per-rank products is not the product
    of global means.  With this on, `loss_dc` all-reduces one scalar (sum_i sum m_i, plus the batch size) in its
    forward and uses the global mean for the first factor: the rank-average of the per-rank scalars, and of the
    gradients GradSync averages, is then exactly the full-batch value.  The trainer turns it on under torch.distributed."""
    import importlib
    # (`onssen_b200.loss.loss_dc` the attribute is the function, like in the reference package; get the module)
    importlib.import_module("onssen_b200.loss.loss_dc").GLOBAL_MEAN[0] = bool(on)

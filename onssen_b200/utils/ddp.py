"""Gradient synchronisation for one-process-per-GPU data parallel training (the reference has none; SURVEY.md
section 8e): bucketed all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests) launched as soon as a
group of gradients is final -- head/BN first, then each BLSTM layer top-down -- so the collective overlaps the
remaining BPTT launches; results are averaged over ranks.

    sync = GradSync()                 # after dist.init_process_group
    model.grad_sync = sync            # DCFunction.backward feeds it bucket by bucket
    loss.backward(); sync.wait()      # gradients in .grad are now the rank average
"""
import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.pending = []
        self.bytes_reduced = 0

    def reduce_bucket(self, grads):
        """grads: dict name -> tensor (final). Launches one async all-reduce over the flattened bucket."""
        if self.world == 1 or not grads:
            return
        names = sorted(grads)
        flat = torch.cat([grads[n].reshape(-1) for n in names])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending.append((work, flat, names, grads))
        self.bytes_reduced += flat.numel() * flat.element_size()

    def wait(self):
        """Blocks until every bucket is reduced and writes the rank-averaged values back in place."""
        for work, flat, names, grads in self.pending:
            work.wait()
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
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src)


def enable_sync_batchnorm(model, on=True):
    """Train-mode BatchNorm statistics (and their backward) over the WHOLE data-parallel batch: one all-reduce of
    2 x 2Hp doubles per BatchNorm call, forward and backward.  Without it each rank normalises with its own shard's
    statistics, which differs from the single-device arithmetic of deep_clustering.py:36-38.
    Equal per-rank batch sizes are assumed."""
    for m in model.modules():
        if hasattr(m, "bn"):
            m.sync_bn = bool(on)
    return model

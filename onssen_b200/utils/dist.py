"""One-process-per-GPU helpers (the reference has no distributed code at all).

Utterances are independent through featurizer, BLSTM, heads and losses, so the batch is sharded across ranks
with no data-path collective for forward + loss; ranks only synchronise for timing / logging."""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced [lo, hi) slice of n_items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, device=None):
    """MAX all-reduce of a python float (step time): every rank gets the slowest rank's value."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def global_loss_mean(l, mag_sum, device=None):
    """The reference's `torch.mean(loss_dc)` over a (B,B) outer product equals mean_i(mag_sum_i)*mean_j(l_j)
    (loss_dc.py:44 + train.py:79).  Under batch sharding the GLOBAL value is obtained from four scalar sums
    (SURVEY.md section 8e), not from the mean of per-shard means."""
    n = sum_over_ranks(l.numel(), device)
    s_l = sum_over_ranks(float(l.double().sum().item()), device)
    s_m = sum_over_ranks(float(mag_sum.double().sum().item()), device)
    return (s_m / n) * (s_l / n)

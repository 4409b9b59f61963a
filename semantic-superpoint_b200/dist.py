"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

The loss step shards by batch; its only exchange is the global-batch normalisers (both reference losses
divide by whole-batch quantities: Train_model_heatmap_all.py:178, utils/utils.py:886-887).  Homography
adaptation shards by source image with no data-path collective.
"""
import os

import torch
import torch.distributed as tdist


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not tdist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        tdist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def _group(group):
    return None if group is True else group


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class DeferredExchange(object):
    """Launches the tiny all-reduces asynchronously and applies their fix-ups later, so that the exchange of one loss
    overlaps the kernels of the next (the reductions only gate the final loss values and the backward scales).
    Pass an instance as `dist_group`; call finish() before the losses are used."""

    def __init__(self, group=True):
        self.group = group
        self.pending = []

    def all_reduce(self, sums, fixup):
        if os.environ.get("SSP_DIST_SYNC") == "1":  # debugging aid: no NCCL kernel ever overlaps the loss kernels
            tdist.all_reduce(sums, op=tdist.ReduceOp.SUM, group=_group(self.group))
            fixup()
            return
        work = tdist.all_reduce(sums, op=tdist.ReduceOp.SUM, group=_group(self.group), async_op=True)
        self.pending.append((work, fixup))

    def finish(self):
        for work, fixup in self.pending:
            work.wait()
            fixup()
        self.pending = []


def _reduce(sums, group, fixup):
    if isinstance(group, DeferredExchange):
        group.all_reduce(sums, fixup)
    else:
        tdist.all_reduce(sums, op=tdist.ReduceOp.SUM, group=_group(group))
        fixup()


def _world(group):
    g = group.group if isinstance(group, DeferredExchange) else group
    return tdist.get_world_size(_group(g))


def globalize_detector(out3, group, after=None):
    """out3 = [loss, numerator, sum(mask) + 1e-5] of the local shard -> same triple for the global batch.
    `after` (optional callable) runs once the global values are in place (immediately, or at DeferredExchange.finish)."""
    sums = torch.stack((out3[1], out3[2] - 1e-5))

    def fixup():
        with torch.no_grad():
            out3[1] = sums[0]
            out3[2] = sums[1] + 1e-5
            out3[0] = out3[1] / out3[2]
            if after is not None:
                after()

    _reduce(sums, group, fixup)
    return out3


def globalize_semantic(out3, group, after=None):
    """out3 = [loss, sum, count] of the local shard -> global batch (mean over every counted pixel of every rank)."""
    sums = out3[1:3].clone()

    def fixup():
        with torch.no_grad():
            out3[1:3] = sums
            out3[0] = sums[0] / sums[1]
            if after is not None:
                after()

    _reduce(sums, group, fixup)
    return out3


def globalize_descriptor(out8, B_local, Hc, Wc, group, after=None):
    """out8 = [loss, pos, neg, norm, num_loss, num_pos, num_neg, sum(mask_valid)] of the local shard -> global batch.
    norm_global = B_global * (sum_global(mask_valid) + 1) * Hc * Wc  (utils/utils.py:886-887)."""
    sums = out8[4:8].clone()
    world = _world(group)

    def fixup():
        with torch.no_grad():
            norm = float(B_local * world) * (sums[3] + 1.0) * float(Hc * Wc)
            out8[3] = norm
            out8[0:3] = sums[0:3] / norm
            out8[4:8] = sums
            if after is not None:
                after()

    _reduce(sums, group, fixup)
    return out8


def max_over_ranks(value, device):
    """Max of a python float over all ranks (timing is reported as the slowest rank)."""
    if not tdist.is_initialized():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    return float(t.item())

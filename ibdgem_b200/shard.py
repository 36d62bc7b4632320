"""Multi-GPU plumbing for the scoring path: targets are independent (the outer loop of
compare_impute, src/ibdgem.c:522), so they are partitioned contiguously across ranks, the packed
panel is replicated, every rank scores its own targets, and the per-window scores are brought
together with ONE all_gather (NCCL over NVLink on the GPU box, gloo in the CPU tests).  There is
no other exchange step on this path."""
from __future__ import annotations

import numpy as np


def shard_bounds(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced partition: the first n % world ranks hold one extra item."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_targets(targets, world: int, rank: int) -> np.ndarray:
    """This rank's slice of the target list, in output order."""
    targets = np.asarray(targets)
    lo, hi = shard_bounds(len(targets), world, rank)
    return targets[lo:hi]


def gather_window_scores(local, n_targets_total: int, group=None):
    """all_gather of per-window scores.

    local: torch tensor [T_local, maxW, 3] (fp64) of this rank's targets — on the GPU for the NCCL
    backend, on the host for gloo.  Returns [n_targets_total, maxW, 3] in target-list order on
    every rank.  Shards may differ by one target: the collective runs on equally sized, padded
    blocks (one all_gather_into_tensor), the padding is dropped afterwards."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if world == 1:
        return local
    per = -(-n_targets_total // world)
    block = local
    if local.shape[0] != per:
        block = torch.full((per,) + tuple(local.shape[1:]), float("nan"), dtype=local.dtype, device=local.device)
        block[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), block.contiguous().view(-1), group=group)
    pieces = []
    for r in range(world):
        lo, hi = shard_bounds(n_targets_total, world, r)
        pieces.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(pieces, dim=0)

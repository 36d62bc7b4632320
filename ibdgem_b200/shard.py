"""Multi-GPU plumbing for the scoring path: targets are independent (the outer loop of
compare_impute, src/ibdgem.c:522), so they are partitioned contiguously across ranks, the packed
panel is replicated, every rank scores its own targets, and the per-window scores are brought
together with ONE all_gather (NCCL over NVLink on the GPU box, gloo in the CPU tests).  There is
no other exchange step on the scoring path.  The replication of the panel itself can also go over
NVLink (replicate_panel): every rank copies 1/N of the packed rows over PCIe and the ranks
all_gather the pieces, instead of N full uploads through the host bridges."""
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


def panel_pieces(n_sites: int, world: int, pieces: int):
    """Row layout of replicate_panel: piece c covers rows [c*per*world, (c+1)*per*world) of the
    (padded) device buffer and rank r contributes rows [c*per*world + r*per, +per) of it.
    Returns (per, padded_rows)."""
    per = -(-int(n_sites) // (int(pieces) * int(world)))
    return per, per * world * pieces


def replicate_panel(engine, h_bits, d_panel, n_indiv: int, pieces: int = 1, stream=None, group=None):
    """Fills `d_panel` (torch int32 [padded_rows, Wh], padded_rows from panel_pieces) with the packed
    panel `h_bits` (torch int32 [S, Wh], identical on every rank; pinned for an asynchronous copy)
    and hands it to `engine` (ibdgem_engine_set_panel_device / _panel_rows_ready).

    Every rank copies only its 1/world of each piece from the host; one in-place
    all_gather_into_tensor per piece (NCCL over NVLink) completes it on all ranks.  With a side
    `stream` (torch.cuda.Stream) the copies and collectives are enqueued there, so the engine scores
    the windows of the pieces that have arrived while later ones are still moving."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    S, Wh = int(h_bits.shape[0]), int(h_bits.shape[1])
    per, padded = panel_pieces(S, world, pieces)
    if d_panel.shape[0] < padded or d_panel.shape[1] != Wh:
        raise ValueError(f"d_panel must be [{padded}, {Wh}], got {tuple(d_panel.shape)}")
    on_gpu = d_panel.is_cuda
    engine.set_panel_device(d_panel.data_ptr(), S, n_indiv, Wh)
    ctx = torch.cuda.stream(stream) if (on_gpu and stream is not None) else _null_ctx()
    with ctx:
        for c in range(pieces):
            c0 = c * per * world
            lo = c0 + rank * per
            hi = min(lo + per, S)
            if hi > lo:
                d_panel[lo:hi].copy_(h_bits[lo:hi], non_blocking=True)
            if world > 1:
                mine = d_panel[lo:lo + per]
                if not on_gpu:
                    mine = mine.clone()  # gloo: no aliasing of input and output
                dist.all_gather_into_tensor(d_panel[c0:c0 + per * world].view(-1), mine.reshape(-1), group=group)
            ready = min(c0 + per * world, S)
            if ready > 0:
                engine.panel_rows_ready(ready, stream.cuda_stream if (on_gpu and stream is not None) else 0)
            if ready >= S:
                break


class _null_ctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

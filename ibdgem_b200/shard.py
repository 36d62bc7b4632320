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


def window_shard_bounds(n_windows: int, world: int, rank: int) -> tuple[int, int]:
    """Windows [lo, hi) of rank `rank` under the partition by windows — the same arithmetic as
    ibdgem_engine_set_window_shard (engine.cu window_shard_bounds)."""
    return int(n_windows) * rank // world, int(n_windows) * (rank + 1) // world


def gather_window_columns(local, group=None):
    """Partition by windows, collective form (gloo / NCCL fallback of the peer-store gather): every rank
    holds a full-size table [T, maxW, 3] in which only its own window columns are set (NaN elsewhere);
    the ranks' tables are combined column-wise on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if world == 1:
        return local
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous(), group=group)
    out = parts[0].clone()
    for p in parts[1:]:
        out = torch.where(torch.isnan(out), p, out)
    return out


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


class _DevArray:
    """__cuda_array_interface__ view of raw device memory (for torch.as_tensor)."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3, "strides": None}


class PeerTable:
    """The gathered window-score table [rows, maxW, 3] fp64 in the ROOT rank's HBM, mapped into every rank
    of the node through CUDA IPC (ibdgem_peer_alloc / ibdgem_peer_open).  A rank hands `block_ptr(row0)`
    (targets partition: its own block of rows) or `ptr` (window partition: every row, its own columns) to
    the engine as ibdgem_scores.w_loglik_device; the engine's copies then ARE the gather — device to device
    over NVLink, range by range, with no rendezvous between the ranks inside the scoring loop.

    `ok` is False on every rank when any rank could not map the buffer (IPC unavailable in this
    container): callers then fall back to gather_window_scores (one NCCL collective)."""

    def __init__(self, rows: int, max_windows: int, device_index: int, group=None, root: int = 0):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from ._lib import load_library
        self._lib = load_library()
        self.shape = (int(rows), int(max_windows), 3)
        self.nbytes = int(rows) * int(max_windows) * 24
        self.device_index = device_index
        self.root = root
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.is_root = self.rank == root
        self.ptr = 0
        handle = (C.c_ubyte * 64)()
        good = 1
        if self.is_root:
            p = C.c_void_p()
            if self._lib.ibdgem_peer_alloc(C.c_int32(device_index), C.c_int64(self.nbytes), C.byref(p), handle) != 0:
                good = 0
            else:
                self.ptr = int(p.value)
        if self.world > 1:
            dev = torch.device("cuda", device_index)
            t = torch.tensor(list(handle) + [good], dtype=torch.uint8, device=dev)
            dist.broadcast(t, src=root, group=group)
            good = int(t[64].item())
            if good and not self.is_root:
                h = (C.c_ubyte * 64)(*[int(x) for x in t[:64].tolist()])
                p = C.c_void_p()
                if self._lib.ibdgem_peer_open(C.c_int32(device_index), h, C.byref(p)) != 0:
                    good = 0
                else:
                    self.ptr = int(p.value)
            flag = torch.tensor([good], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            good = int(flag.item())
        self.ok = bool(good)
        if not self.ok:
            self.close()

    def block_ptr(self, row0: int) -> int:
        return self.ptr + int(row0) * self.shape[1] * 24

    def tensor(self):
        """torch view of the table (root only)."""
        import torch
        if not (self.is_root and self.ptr):
            raise RuntimeError("the gathered table lives on the root rank")
        return torch.as_tensor(_DevArray(self.ptr, self.shape), device=torch.device("cuda", self.device_index))

    def close(self):
        import ctypes as C
        if self.ptr:
            self._lib.ibdgem_peer_close(C.c_int32(self.device_index), C.c_void_p(self.ptr), C.c_int32(1 if self.is_root else 0))
            self.ptr = 0


def upload_window_shard_rows(engine, h_bits, d_panel, n_indiv: int, stream=None):
    """Window partition, end to end: the rank's engine has a window shard set (Engine.set_window_shard) and
    the site arrays uploaded; this hands it `d_panel` (torch int32 [S, Wh] device buffer) and copies from
    the pinned host panel `h_bits` ONLY the rows the shard reads, so every panel byte crosses PCIe once
    per node.  Returns the number of bytes copied."""
    import torch
    S, Wh = int(h_bits.shape[0]), int(h_bits.shape[1])
    engine.set_panel_device(d_panel.data_ptr(), S, n_indiv, Wh)
    prm = engine.params
    needs_panel_first = prm.min_af > 0.0 or prm.max_af < 1.0 or getattr(engine, "_sites", (None,) * 5)[4] is not None
    ctx = torch.cuda.stream(stream) if stream is not None else _null_ctx()
    if needs_panel_first:  # an AF filter decides which sites are kept: the window map itself needs every row
        with ctx:
            d_panel[:S].copy_(h_bits, non_blocking=True)
        engine.panel_rows_ready(S, stream.cuda_stream if stream is not None else 0)
        return S * Wh * 4
    _, _, sb, se = engine.window_shard()  # builds the window map from the site arrays (no panel rows needed)
    with ctx:
        if se > sb:
            d_panel[sb:se].copy_(h_bits[sb:se], non_blocking=True)
    engine.panel_rows_ready(S, stream.cuda_stream if stream is not None else 0)
    return (se - sb) * Wh * 4


class _null_ctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

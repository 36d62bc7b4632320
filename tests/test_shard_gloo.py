"""world_size-2 gloo test of the N>1 path (host logic only): targets sharded contiguously, every
rank scores its shard (here with the CPU oracle — the GPU engine is exercised by the -m gpu tests),
one all_gather of per-window scores; the gathered table must equal the unsharded run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ibdgem_b200.shard import (gather_window_columns, gather_window_scores, panel_pieces, replicate_panel, shard_bounds,
                                shard_targets, window_shard_bounds)


def test_shard_bounds_partition():
    for n in (0, 1, 5, 8, 1000, 1001):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    import oracle
    rng = np.random.default_rng(3)
    S, N, W = 600, 12, 50
    af = np.clip(rng.beta(0.5, 2.0, S), 0.01, 0.99)
    hap = (rng.random((S, 2 * N)) < af[:, None]).astype(np.uint8)
    pos = (1000 + 60 * np.arange(S)).astype(np.uint64)
    d = rng.poisson(2.0, S)
    n_alt = rng.binomial(d, 0.3).astype(np.uint8)
    n_ref = (d - n_alt).astype(np.uint8)
    keep = np.ones(S, np.uint8)
    prm = oracle.Params(window=W, ld_mode=1)
    targets = np.array([0, 3, 4, 7, 9], np.int32)  # odd count: shards of 3 and 2
    return prm, pos, keep, n_ref, n_alt, hap, targets, np.arange(N, dtype=np.int32)


def _score(targets):
    import oracle
    prm, pos, keep, n_ref, n_alt, hap, _, bg = _case()
    maxW = len(pos) // prm.window + 2
    out = np.full((len(targets), maxW, 3), np.nan)
    for k, t in enumerate(targets):
        o = oracle.compare_target(prm, pos, keep, n_ref, n_alt, hap, int(t), bg)
        out[k, : o["n_windows"]] = o["w_log"]
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        targets = _case()[6]
        mine = shard_targets(targets, world, rank)
        local = torch.from_numpy(_score(mine))
        full = gather_window_scores(local, len(targets))
        q.put((rank, full.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gather_matches_unsharded():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    want = _score(_case()[6])
    for r in range(world):
        np.testing.assert_array_equal(np.isnan(got[r]), np.isnan(want))
        np.testing.assert_array_equal(np.nan_to_num(got[r]), np.nan_to_num(want))


class _RecordingEngine:
    """Stands in for ibdgem_b200.Engine: replicate_panel only calls these two methods."""

    def __init__(self):
        self.calls = []

    def set_panel_device(self, ptr, n_sites, n_indiv, wh):
        self.calls.append(("set", n_sites, n_indiv, wh))

    def panel_rows_ready(self, row_end, stream=0):
        self.calls.append(("ready", row_end))


def _panel_worker(rank, world, port, q, S, pieces):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Wh = 5
        h_bits = torch.arange(S * Wh, dtype=torch.int32).reshape(S, Wh)  # identical on every rank
        per, padded = panel_pieces(S, world, pieces)
        d_panel = torch.full((padded, Wh), -1, dtype=torch.int32)
        eng = _RecordingEngine()
        replicate_panel(eng, h_bits, d_panel, n_indiv=40, pieces=pieces)
        q.put((rank, d_panel[:S].numpy().copy(), eng.calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("S,pieces", [(1000, 1), (1003, 4), (7, 3)])
def test_two_rank_panel_replication(S, pieces):
    """Every rank copies only its share of each piece; after the all_gathers both hold the whole
    panel, and the pieces were declared to the engine in increasing row order up to S."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_panel_worker, args=(r, world, port, q, S, pieces)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    want = np.arange(S * 5, dtype=np.int32).reshape(S, 5)
    for rank, panel, calls in got:
        np.testing.assert_array_equal(panel, want)
        assert calls[0] == ("set", S, 40, 5)
        ready = [c[1] for c in calls[1:]]
        assert ready == sorted(ready) and ready[-1] == S and len(ready) <= pieces


@pytest.mark.parametrize("S,pieces", [(10, 1), (1001, 4), (5, 8)])
def test_panel_replication_single_process(S, pieces):
    """Without a process group the helper degenerates to piece-wise copies of the whole panel."""
    Wh = 3
    h_bits = torch.arange(S * Wh, dtype=torch.int32).reshape(S, Wh)
    per, padded = panel_pieces(S, 1, pieces)
    assert per * pieces == padded >= S
    d_panel = torch.full((padded, Wh), -1, dtype=torch.int32)
    eng = _RecordingEngine()
    replicate_panel(eng, h_bits, d_panel, n_indiv=40, pieces=pieces)
    assert torch.equal(d_panel[:S], h_bits)
    ready = [c[1] for c in eng.calls[1:]]
    assert eng.calls[0] == ("set", S, 40, Wh) and ready == sorted(ready) and ready[-1] == S
    with pytest.raises(ValueError):
        replicate_panel(eng, h_bits, d_panel[: max(S - 1, 0)], n_indiv=40, pieces=pieces)


def test_window_shard_bounds_tile_the_windows():
    for n in (0, 1, 7, 1000, 1001):
        for world in (1, 2, 3, 8):
            cuts = [window_shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert max(hi - lo for lo, hi in cuts) - min(hi - lo for lo, hi in cuts) <= 1


def _window_worker(rank, world, port, q):
    """Partition by windows: windows are independent, so a rank can score its own windows from the sites of
    those windows alone (here with the CPU oracle), for every target."""
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prm, pos, keep, n_ref, n_alt, hap, targets, bg = _case()
        maxW = len(pos) // prm.window + 2
        inf = np.flatnonzero((keep == 1) & ((n_ref.astype(int) + n_alt) >= 1))
        nW = -(-len(inf) // prm.window)
        lo, hi = window_shard_bounds(nW, world, rank)
        s0 = 0 if lo == 0 else inf[lo * prm.window - 1] + 1  # rows between two windows go with the later shard
        s1 = len(pos) if hi == nW else inf[hi * prm.window - 1] + 1
        sl = slice(s0, s1)
        local = np.full((len(targets), maxW, 3), np.nan)
        for k, t in enumerate(targets):
            o = oracle.compare_target(prm, pos[sl], keep[sl], n_ref[sl], n_alt[sl], hap[sl], int(t), bg)
            assert o["n_windows"] == hi - lo
            local[k, lo:hi] = o["w_log"]
        full = gather_window_columns(torch.from_numpy(local))
        q.put((rank, full.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_window_partition_matches_unsharded():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_window_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    want = _score(_case()[6])
    for r in range(world):
        np.testing.assert_array_equal(np.isnan(got[r]), np.isnan(want))
        np.testing.assert_allclose(np.nan_to_num(got[r]), np.nan_to_num(want), rtol=0, atol=1e-9)

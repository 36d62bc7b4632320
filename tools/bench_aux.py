"""Secondary legs of bench.py (extra keys of its one JSON line, under "aux"): the other hot-path rows of
SURVEY.md §8 measured the same way as the headline — device-resident time, end to end through the C ABI
with pinned host buffers, roofline fraction on §8(d)'s algorithmic bytes, and the reference's CPU binary
timed beside it on a bounded sample.

  c2        non-LD per-site IBD0/1/2 + window sums, 1,000,000 sites x 100 targets, window 100 (configs[1])
  c4        hiddengem over 10,000 window tables x 10,000 bins (configs[3])
  ld_v      C3 with -v (the reference's recommended flag): per-target windows
  flat_rows C3 with the pileup's source OUTSIDE the background (no dominant column in any row)
  cli_e2e   bin/ibdgem against oracle/_ref/ibdgem on the SAME files, wall clock
  int8_peak cuBLASLt int8 GEMM 8192^3 on this box (the measured denominator MEASURED_PEAKS.json lacks)

`python tools/bench_aux.py c2 c4` runs legs stand-alone on a GPU box."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback 6.65 TB/s"


def source_genotype(bits, src):
    """Genotype (0/1/2) of individual `src` at every site, from the packed panel."""
    h0, h1 = 2 * src, 2 * src + 1
    a = (bits[:, h0 >> 5] >> np.uint32(h0 & 31)) & np.uint32(1)
    b = (bits[:, h1 >> 5] >> np.uint32(h1 & 31)) & np.uint32(1)
    return (a + b).astype(np.int64)


def make_counts(bits, src, seed, depth_floor, eps=0.02, depth_mean=2.0):
    """Pileup counts of synth_panel_* for another source individual / depth floor, same panel."""
    rng = np.random.default_rng(seed)
    S = bits.shape[0]
    g = source_genotype(bits, src)
    depth = np.minimum(rng.poisson(depth_mean, S) + depth_floor, 20)
    n_alt = rng.binomial(depth, np.where(g == 0, eps, np.where(g == 1, 0.5, 1 - eps)))
    return (depth - n_alt).astype(np.uint8), n_alt.astype(np.uint8)


def _pinned(torch, shape, dtype):
    return torch.empty(shape, dtype=dtype, pin_memory=True)


def _pin_np(torch, a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


def _timed(torch, stream, fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


# ------------------------------------------------------------------------------------------------
def leg_c2(torch, ib, bits, pos, N, steps=5, ref_dir=None, cpu_sites=10_000):
    """BASELINE.json configs[1].  Reference code: src/ibdgem.c:632-667, 751-756."""
    from ibdgem_b200.engine import _CScores
    S, T, W = bits.shape[0], 100, 100
    n_ref, n_alt = make_counts(bits, 0, 2, 0)  # Poisson(2): ~13.5 % zero-data sites (src/ibdgem.c:657-663)
    _, n_ref = _pin_np(torch, n_ref)
    _, n_alt = _pin_np(torch, n_alt)
    keep_t, keep = _pin_np(torch, np.ones(S, np.uint8))
    targets = np.arange(T, dtype=np.int32)
    maxW = S // W + 2
    stream = torch.cuda.current_stream()
    o_nw = _pinned(torch, (T,), torch.int32)
    o_ws, o_we = _pinned(torch, (T, maxW), torch.int64), _pinned(torch, (T, maxW), torch.int64)
    o_wn = _pinned(torch, (T, maxW), torch.int32)
    o_ll = _pinned(torch, (T, maxW, 3), torch.float64)
    cnt = [np.zeros(T, np.uint64) for _ in range(3)] + [np.zeros((T, 21), np.uint64)]
    cs = _CScores(maxW, o_nw.data_ptr(), o_ws.data_ptr(), o_we.data_ptr(), o_wn.data_ptr(), o_ll.data_ptr(),
                  cnt[0].ctypes.data, cnt[1].ctypes.data, cnt[2].ctypes.data, cnt[3].ctypes.data, None, None, None, None)
    with ib.Engine(ib.Params(window_size=W, device=torch.cuda.current_device())) as e:
        e.set_stream(stream.cuda_stream)

        def upload():
            e.upload_sites(pos, n_ref, n_alt, keep)
            e.upload_panel(bits, N)

        def score():
            e.invalidate()
            e.score_nonld_raw(targets, cs)

        upload()
        e.sync_uploads()
        e.enable_timing(True)
        for _ in range(3):
            score()
        e.reset_stats()
        ms_step = _timed(torch, stream, score, steps)
        st = {k: v[0] / steps for k, v in e.kernel_stats().items() if v[1]}
        launches = int(sum(v[1] for v in e.kernel_stats().values()) / steps)
        e.enable_timing(False)

        def e2e():
            upload()
            score()

        for _ in range(2):
            e2e()
        ms_e2e = _timed(torch, stream, e2e, max(2, min(steps, 3)))
    nW = int(o_nw[0])
    inf = int(((n_ref.astype(np.int64) + n_alt) >= 1).sum())
    assert nW == -(-inf // W) and int(cnt[0][0]) + int(cnt[1][0]) == S  # bit-exact site counts
    hbm, src = hbm_peak()
    ms_k = sum(st.values())
    b_8d = S * 6 + S * T * 2 / 8 + S * 56 + nW * T * 24  # SURVEY.md 8(d), kernel 1: 108 MB
    b_panel = bits.nbytes  # the allele frequency is a popcount over the packed row: the panel is read once
    h2d = bits.nbytes + pos.nbytes + n_ref.nbytes + n_alt.nbytes + keep.nbytes + targets.nbytes
    d2h = o_nw.numel() * 4 + o_ws.numel() * 16 + o_wn.numel() * 4 + o_ll.numel() * 8
    out = {"workload": "C2 non-LD: %d sites x %d targets, window %d, depth Poisson(2), %d-sample panel "
                       "(BASELINE.json configs[1])" % (S, T, W, N),
           "metric": "site x target likelihoods/s", "value": S * T / (ms_step * 1e-3), "ms_per_step": ms_step,
           "ms_kernels": ms_k, "kernels_ms": st, "gpu_launches_per_step": launches, "windows_per_target": nW,
           "e2e": {"value": S * T / (ms_e2e * 1e-3), "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": int(d2h), "pcie_floor_ms_at_55GBs": (h2d + d2h) / 55e9 * 1e3},
           "roofline": {"bound": "hbm", "unit": "GB/s", "peak": hbm, "peak_source": src,
                        "bytes": "packed panel read once for the allele frequencies (%.0f MB) + SURVEY.md 8(d) "
                                 "kernel-1 bytes (%.0f MB: counts, genotype bits, 7-value site table, window sums)" % (
                                     b_panel / 1e6, b_8d / 1e6),
                        "achieved": (b_panel + b_8d) / 1e9 / (ms_k * 1e-3), "frac": (b_panel + b_8d) / 1e9 / (ms_k * 1e-3) / hbm,
                        "frac_8d_bytes_only": b_8d / 1e9 / (ms_k * 1e-3) / hbm, "over": "sum of the pass's kernels"}}
    if ref_dir and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ibdgem")) and os.path.exists(os.path.join(ref_dir, "p.hap")):
        # the reference's non-LD run on the CPU sample written for the headline's cpu_baseline
        tl = [[0, 1, 2, 3]]
        t0 = time.perf_counter()
        for k, tlist in enumerate(tl):
            with open(os.path.join(ref_dir, "c2_targets.txt"), "w") as fh:
                fh.write("".join("i%d\n" % t for t in tlist))
            os.makedirs(os.path.join(ref_dir, "c2_out"), exist_ok=True)
            subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ibdgem"), "-H", "p.hap", "-L", "p.legend", "-I", "p.indv", "-P",
                            "u.pileup", "-w", str(W), "-S", "c2_targets.txt", "-O", "c2_out"], cwd=ref_dir, check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        wall = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": cpu_sites * 4 / wall, "unit": "site x target likelihoods/s", "cores": 1,
                               "kind": "reference", "sample": "oracle/_ref/ibdgem (-O0 as shipped), non-LD -w %d, first %d "
                               "sites x 4 targets, %.2f s wall (text in, tab.txt + summary.txt out)" % (W, cpu_sites, wall)}
    return out


# ------------------------------------------------------------------------------------------------
def leg_c4(torch, ib, steps=3, n_tables=10_000, n_bins=10_000, ref_dir=None):
    """BASELINE.json configs[3].  Reference code: src/hiddengem.c:51-147, 246-283."""
    dev = torch.device("cuda", torch.cuda.current_device())
    nb = n_tables * n_bins
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    seg = torch.randint(0, 3, (nb // 200 + 1,), generator=g, device=dev).repeat_interleave(200)[:nb]
    ll = torch.randn((nb, 3), generator=g, device=dev, dtype=torch.float64) * 10.0 - 150.0
    ll[torch.arange(nb, device=dev), seg] += 8.0  # planted IBD0/1/2 segments
    h_ll = _pinned(torch, (nb, 3), torch.float64)
    h_ll.copy_(ll)
    seg_h = seg.to(torch.uint8).cpu().numpy()
    del ll, seg
    torch.cuda.empty_cache()
    off = np.arange(n_tables + 1, dtype=np.int64) * n_bins
    h_state = _pinned(torch, (nb,), torch.uint8)
    h_score = _pinned(torch, (nb, 3), torch.float64)
    counts = np.zeros((n_tables, 3), np.int64)
    stream = torch.cuda.current_stream()
    with ib.Engine(ib.Params(device=torch.cuda.current_device())) as e:
        e.set_stream(stream.cuda_stream)
        e.enable_timing(True)

        def run():
            e.viterbi_batch_raw(h_ll.data_ptr(), off, True, h_state.data_ptr(), h_score.data_ptr(), counts.ctypes.data)

        run()
        flagged = e.viterbi_last_flagged()
        e.reset_stats()
        ms_e2e = _timed(torch, stream, run, steps)
        st = {k: v[0] / steps for k, v in e.kernel_stats().items() if v[1]}
        launches = int(sum(v[1] for v in e.kernel_stats().values()) / steps)
        # device-resident: the same tables already in HBM, results left in HBM
        d_ll = torch.empty((nb, 3), dtype=torch.float64, device=dev)
        d_ll.copy_(h_ll, non_blocking=True)
        d_state = torch.empty((nb,), dtype=torch.uint8, device=dev)
        d_score = torch.empty((nb, 3), dtype=torch.float64, device=dev)
        d_counts = torch.empty((n_tables, 3), dtype=torch.int64, device=dev)
        ms_dev = None
        if hasattr(e, "viterbi_batch_device"):
            def run_dev():
                e.viterbi_batch_device(d_ll.data_ptr(), off, True, d_state.data_ptr(), d_score.data_ptr(), d_counts.data_ptr())
            run_dev()
            e.reset_stats()
            ms_dev = _timed(torch, stream, run_dev, steps)
            st = {k: v[0] / steps for k, v in e.kernel_stats().items() if v[1]}
            launches = int(sum(v[1] for v in e.kernel_stats().values()) / steps)
            assert torch.equal(d_state.cpu(), h_state)
    hbm, src = hbm_peak()
    ms_k = sum(st.values())
    byt = nb * 49.0  # 24 B in, 24 + 1 B out per bin (kernel 3)
    rec = float((h_state.numpy() == seg_h).mean())
    out = {"workload": "C4 hiddengem: %d tables x %d bins, window log-likelihoods straight from the engine "
                       "(BASELINE.json configs[3])" % (n_tables, n_bins),
           "metric": "bins/s", "value": nb / ((ms_dev or ms_k) * 1e-3), "ms_per_step": ms_dev or ms_k, "ms_kernels": ms_k,
           "kernels_ms": st, "gpu_launches_per_step": launches, "planted_state_recovery": rec,
           "near_tie_tables_reevaluated_in_long_double": int(flagged),
           "e2e": {"value": nb / (ms_e2e * 1e-3), "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(nb * 24 + off.nbytes),
                   "d2h_bytes_per_step": int(nb * 25 + counts.nbytes),
                   "pcie_floor_ms_at_55GBs_full_duplex": nb * 25 / 55e9 * 1e3, "pcie_floor_ms_at_55GBs_combined": nb * 49 / 55e9 * 1e3,
                   "note": "host pointers through hiddengem_viterbi_batch, page-locked; batches are pipelined over three streams "
                           "(upload, kernels, download).  On these boxes upload and download do not add up: ~55 GB/s is what "
                           "both directions get TOGETHER (92-97 ms measured for 4.9 GB whatever the batching), so the combined "
                           "figure is the floor that applies"},
           "roofline": {"bound": "hbm", "unit": "GB/s", "peak": hbm, "peak_source": src, "bytes": "49 B per bin (SURVEY.md 8d)",
                        "achieved": byt / 1e9 / (ms_k * 1e-3), "frac": byt / 1e9 / (ms_k * 1e-3) / hbm, "over": "sum of the pass's kernels"}}
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "hiddengem")
    if ref_dir and os.path.exists(ref_bin):
        l = np.exp(h_ll[:n_bins].numpy())
        path = os.path.join(ref_dir, "c4.summary.txt")
        with open(path, "w") as fh:
            fh.write("# SEGMENT\tSTART\tEND\tLIBD0\tLIBD1\tLIBD2\tNUM_SITES\n")
            fh.write("".join("%d\t%d\t%d\t%e\t%e\t%e\t100\n" % (i + 1, 1000 + 6000 * i, 6940 + 6000 * i, l[i, 0], l[i, 1], l[i, 2])
                             for i in range(n_bins)))
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            subprocess.run([ref_bin, "-s", path], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        wall = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": reps * n_bins / wall, "unit": "bins/s", "cores": 1, "kind": "reference",
                               "sample": "oracle/_ref/hiddengem (-O0 as shipped), %d runs of one %d-bin table (text in, text "
                                         "out), %.2f s wall" % (reps, n_bins, wall)}
    return out


# ------------------------------------------------------------------------------------------------
def leg_ld(torch, ib, bits, pos, n_ref, n_alt, N, T, W, steps, variable_sites_only=0, bg=None, what=""):
    """One more C3-shaped --LD workload on its own engine: device-resident ms per full pass."""
    from ibdgem_b200.engine import _CScores
    S = bits.shape[0]
    keep_t, keep = _pin_np(torch, np.ones(S, np.uint8))
    targets = np.arange(T, dtype=np.int32)
    bg = np.arange(N, dtype=np.int32) if bg is None else np.asarray(bg, np.int32)
    maxW = S // W + 2
    stream = torch.cuda.current_stream()
    o_nw = _pinned(torch, (T,), torch.int32)
    o_ws, o_we = _pinned(torch, (T, maxW), torch.int64), _pinned(torch, (T, maxW), torch.int64)
    o_wn = _pinned(torch, (T, maxW), torch.int32)
    o_ll = _pinned(torch, (T, maxW, 3), torch.float64)
    cs = _CScores(maxW, o_nw.data_ptr(), o_ws.data_ptr(), o_we.data_ptr(), o_wn.data_ptr(), o_ll.data_ptr(),
                  None, None, None, None, None, None, None, None)
    with ib.Engine(ib.Params(window_size=W, variable_sites_only=variable_sites_only, device=torch.cuda.current_device())) as e:
        e.set_stream(stream.cuda_stream)
        e.upload_sites(pos, n_ref, n_alt, keep)
        e.upload_panel(bits, N)
        e.sync_uploads()
        e.enable_timing(True)

        def score():
            e.invalidate()
            e.score_ld_raw(targets, bg, -1, cs)

        score()
        path = e.last_ld_path()
        slow = path == 0  # the general CUDA-core path takes seconds per pass at this size
        for _ in range(0 if slow else 2):
            score()
        e.reset_stats()
        n = 1 if slow else steps
        ms = _timed(torch, stream, score, n)
        st = {k: v[0] / n for k, v in e.kernel_stats().items() if v[1]}
    nw = o_nw.numpy()
    wn = o_wn.numpy()
    sites_scored = int(sum(int(wn[t, :nw[t]].sum()) for t in range(T)))
    return {"workload": what, "ms_per_step": ms, "ld_path": int(path), "kernels_ms": st,
            "windows_per_target_mean": float(nw.mean()), "site_x_target_scored": sites_scored,
            "comparisons_per_s": sites_scored * len(bg) * 4 / (ms * 1e-3)}


def leg_window_sharded(torch, ib, world, rank, h_bits, pos, n_ref, n_alt, N, targets, bg, W, steps, what):
    """--LD with the WINDOWS of the run partitioned across the ranks (shared window maps): every rank scores
    all targets on its 1/world of the windows, reads only its own panel rows, and stores its columns of the
    score table straight into the root's HBM (PeerTable).  Strong scaling: the total work is fixed.
    All ranks call this; rank 0 returns the record."""
    import torch.distributed as dist
    from ibdgem_b200.engine import _CScores
    from ibdgem_b200.shard import PeerTable, upload_window_shard_rows
    dev_i = torch.cuda.current_device()
    dev = torch.device("cuda", dev_i)
    bits = h_bits.numpy().view(np.uint32)
    S, T = bits.shape[0], len(targets)
    keep_t, keep = _pin_np(torch, np.ones(S, np.uint8))
    maxW = S // W + 2
    stream = torch.cuda.current_stream()
    o_nw = _pinned(torch, (T,), torch.int32)
    # the host score table of a window shard is compact and window-major: [this rank's windows][T][3]
    # (ibdgem_engine_set_shard_compact_output)
    # (one rank: no shard, the usual [T][maxW][3] table)
    compact = world > 1
    o_ll = _pinned(torch, ((S // W + world) // world + 1, T, 3) if compact else (T, maxW, 3), torch.float64)
    book = rank == 0  # the bookkeeping arrays are the same on every rank: only the root fetches them
    o_ws = _pinned(torch, (T, maxW), torch.int64) if book else None
    o_we = _pinned(torch, (T, maxW), torch.int64) if book else None
    o_wn = _pinned(torch, (T, maxW), torch.int32) if book else None
    table = PeerTable(T, maxW, dev_i)
    if table.is_root and table.ok:
        table.tensor().fill_(float("nan"))
    cs = _CScores(maxW, o_nw.data_ptr(), o_ws.data_ptr() if book else None, o_we.data_ptr() if book else None,
                  o_wn.data_ptr() if book else None, o_ll.data_ptr(), None, None, None, None, None, None, None,
                  table.ptr if table.ok else None)
    d_panel = torch.empty((S, bits.shape[1]), dtype=torch.int32, device=dev)
    up_stream = torch.cuda.Stream(device=dev)

    per_rank = []  # ms per step of every rank, one list per timed loop (the reported figure is the maximum)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
        if world > 1:
            every = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(every, ms)
            per_rank.append([round(float(x.item()), 3) for x in every])
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # IBDGEM_TIMELINE_SHARDED=prefix: kernel timeline of the resident steps, one file per rank (prefix.T<targets>.<rank>)
    tl = os.environ.get("IBDGEM_TIMELINE_SHARDED")
    if tl:
        os.environ["IBDGEM_TIMELINE"] = "%s.T%d.%d" % (tl, T, rank)
    with ib.Engine(ib.Params(window_size=W, device=dev_i)) as e:
        e.set_stream(stream.cuda_stream)
        e.set_window_shard(rank, world)
        e.set_shard_compact_output(compact)
        if tl:
            e.enable_timing(True)  # before the sites go up: that is where the timeline's zero is taken
        e.upload_sites(pos, n_ref, n_alt, keep)
        e.upload_panel(bits, N)
        e.sync_uploads()
        e.enable_timing(True)

        def score():
            e.invalidate()
            e.score_ld_raw(targets, bg, -1, cs)

        for _ in range(3):
            score()
        assert e.last_ld_path() == 1
        e.reset_stats()
        ms = timed(score, steps)
        st = {k: v[0] / steps for k, v in e.kernel_stats().items() if v[1]}
        wb, we, sb, se = e.window_shard()
        e.enable_timing(False)
        h2d = [0]

        def e2e():
            e.upload_sites(pos, n_ref, n_alt, keep)
            up_stream.wait_stream(stream)
            h2d[0] = upload_window_shard_rows(e, h_bits, d_panel, N, stream=up_stream)
            score()

        for _ in range(2):
            e2e()
        n2 = max(2, min(steps, 3))
        ms_e2e = timed(e2e, n2)
    if tl:
        os.environ.pop("IBDGEM_TIMELINE", None)
    nW = int(o_nw[0])
    # compact, window-major: [we - wb][T][3] -> [T][we - wb][3]
    mine = o_ll.numpy()[: we - wb].transpose(1, 0, 2) if compact else o_ll.numpy()[:, wb:we]
    ok_cols = bool(np.isfinite(mine).all())
    gathered_ok = None
    if table.ok:
        barrier()
        if table.is_root:
            g = table.tensor()[:, :nW].cpu().numpy()
            gathered_ok = bool(np.isfinite(g).all()) and bool(np.array_equal(g[:, wb:we], mine))
        barrier()
    table.close()
    inf = int(((n_ref.astype(np.int64) + n_alt) >= 1).sum())
    nbg = len(bg) - int(np.isin(np.asarray(bg), np.asarray(targets)).any())  # -1 when the target is in the background
    comps = float(inf) * T * nbg * 4
    if rank != 0:
        return None
    return {"workload": what, "partition": "windows: %d window shard(s) of %d windows, all %d targets on every rank; score "
            "columns stored into the root's HBM over NVLink (CUDA IPC peer stores%s)" % (world, nW, T, "" if table.ok else
            " UNAVAILABLE here: no gather was made"),
            "scaling": "strong", "n_gpus": world, "ms_per_step": ms, "comparisons_per_s": comps / (ms * 1e-3),
            "ms_per_step_by_rank": per_rank[0] if per_rank else [ms],
            "kernels_ms_rank0": st, "rank0_windows": [wb, we], "rank0_rows": [sb, se],
            "shard_columns_finite": ok_cols, "gathered_table_complete_and_equal": gathered_ok,
            "gather_bytes_per_rank": int(T) * (we - wb) * 24,
            "e2e": {"ms_per_step": ms_e2e, "comparisons_per_s": comps / (ms_e2e * 1e-3),
                    "h2d_bytes_per_step_rank0": int(h2d[0] + pos.nbytes + n_ref.nbytes + n_alt.nbytes + keep.nbytes + targets.nbytes + bg.nbytes),
                    "d2h_bytes_per_step_rank0": int(T * (we - wb) * 24 + T * maxW * 20 + T * 4)},
            "host_table": ("compact, window-major [this rank's windows][T][3] per rank; contiguous copies, three shrinking sub-ranges"
                           if compact else "[T][max_windows][3]")}


def c5_inputs(torch, S, N, seed=5, src=None):
    """15,000-individual panel for BASELINE.json configs[4]; the reads are drawn from background
    individual 10,000."""
    from ibdgem_b200.synth import synth_panel_torch
    d = synth_panel_torch(S, N, seed=seed, src=10_000 if src is None else src, device=torch.device("cuda", torch.cuda.current_device()),
                          chunk=20_000)
    return d


def leg_int8_peak(torch):
    """Dense int8 GEMM 8192^3 through cuBLASLt (torch._int_mm): burst (best of 10) and sustained (2 s)."""
    try:
        dev = torch.device("cuda", torch.cuda.current_device())
        n = 8192
        a = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
        b = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
        for _ in range(3):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(10, int(2000.0 / best))
        e0.record()
        for _ in range(reps):
            torch._int_mm(a, b)
        e1.record()
        torch.cuda.synchronize()
        sus = e0.elapsed_time(e1) / reps
        ops = 2.0 * n ** 3
        return {"burst_tops": ops / (best * 1e-3) / 1e12, "sustained_tops": ops / (sus * 1e-3) / 1e12,
                "how": "torch._int_mm (cuBLASLt) int8 x int8 -> int32, 8192^3, random operands in [-8, 8): best of 10 and %d "
                       "back to back" % reps}
    except Exception as ex:  # noqa: BLE001 - an absent cuBLASLt int8 path is reported, not fatal
        return {"unavailable": repr(ex)[:200]}


def leg_cli(ref_dir, window, ref_wall, n_targets):
    """bin/ibdgem on the files the reference's cpu_baseline run just read: same arguments, wall clock."""
    binary = os.path.join(ROOT, "ibdgem_b200", "bin", "ibdgem")
    if not (ref_dir and os.path.exists(binary) and os.path.exists(os.path.join(ref_dir, "targets_0.txt"))):
        return {"unavailable": "bin/ibdgem or the reference sample is missing"}
    os.makedirs(os.path.join(ref_dir, "out_ours"), exist_ok=True)
    cmd = [binary, "-H", "p.hap", "-L", "p.legend", "-I", "p.indv", "-P", "u.pileup", "--LD", "-w", str(window), "-S",
           "targets_0.txt", "-O", "out_ours"]
    walls = []
    for _ in range(2):  # the second run has the CUDA context / page cache warm, like the reference's repeated runs
        t0 = time.perf_counter()
        r = subprocess.run(cmd, cwd=ref_dir, capture_output=True, text=True)
        walls.append(time.perf_counter() - t0)
        if r.returncode != 0:
            return {"unavailable": "bin/ibdgem failed: " + r.stderr[-300:]}
    same = None
    try:  # the integer columns of every summary row must be the reference's
        same = True
        for f in sorted(os.listdir(os.path.join(ref_dir, "out_0"))):
            if not f.endswith(".summary.txt"):
                continue
            a = open(os.path.join(ref_dir, "out_0", f)).read().splitlines()
            b = open(os.path.join(ref_dir, "out_ours", f)).read().splitlines()
            same = same and len(a) == len(b) and all(
                x.split("\t")[:3] == y.split("\t")[:3] and x.split("\t")[6] == y.split("\t")[6] for x, y in zip(a[1:], b[1:]))
    except OSError:
        same = False
    return {"ours_wall_s": min(walls), "ours_wall_s_first_run": walls[0], "reference_wall_s": ref_wall,
            "speedup_wall": ref_wall / min(walls), "summary_integer_columns_identical": same,
            "what": "same IMPUTE triple + pileup + -S list (%d targets), --LD -w %d, text in / tab.txt + summary.txt out; "
                    "ours includes process start, CUDA context creation and the panel parse" % (n_targets, window)}


if __name__ == "__main__":
    import torch
    import ibdgem_b200 as ib
    from ibdgem_b200.synth import synth_panel_torch
    which = sys.argv[1:] or ["c2", "c4"]
    if "c2" in which:
        d = synth_panel_torch(1_000_000, 2504, seed=1, device="cuda")
        print(json.dumps(leg_c2(torch, ib, d["bits"].numpy().view(np.uint32), d["pos"].numpy().view(np.uint64), 2504)))
    if "c4" in which:
        print(json.dumps(leg_c4(torch, ib)))
    if "int8" in which:
        print(json.dumps(leg_int8_peak(torch)))

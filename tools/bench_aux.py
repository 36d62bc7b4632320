"""Secondary measurements for the other hot-path rows of SURVEY.md §8 (not the bench.py headline):
C2 — non-LD per-site IBD0/1/2 + window sums, 1,000,000 sites x 100 targets, window 100;
C4 — hiddengem over 10,000 window tables x 10,000 bins.
Prints one JSON object per workload with device times per kernel (CUDA events on the engine stream)
and the HBM roofline fraction against MEASURED_PEAKS.json.  Run on a GPU box."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ibdgem_b200 as ib  # noqa: E402


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError, ValueError):
        return 6650.0


def c2(steps=5):
    import torch
    from ibdgem_b200.synth import synth_panel_torch
    S, N, T, W = 1_000_000, 2504, 100, 100
    d = synth_panel_torch(S, N, seed=1, depth_floor=0, device="cuda")  # Poisson(2): ~13.5 % zero-data sites
    bits = d["bits"].numpy().view(np.uint32)
    pos = d["pos"].numpy().view(np.uint64)
    n_ref, n_alt, keep = d["n_ref"].numpy(), d["n_alt"].numpy(), d["keep"].numpy()
    targets = np.arange(T, dtype=np.int32)
    out = {}
    with ib.Engine(ib.Params(window_size=W)) as e:
        e.upload_sites(pos, n_ref, n_alt, keep)
        e.upload_panel(bits, N)
        e.enable_timing(True)
        for _ in range(3):
            e.invalidate()
            sc = e.score_nonld(targets)
        e.reset_stats()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            e.invalidate()
            sc = e.score_nonld(targets)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / steps
        st = {k: v[0] / steps for k, v in e.kernel_stats().items() if v[1]}
    nW = int(sc.n_windows[0])
    hbm = peaks()
    # algorithmic bytes: the packed panel must be read once for the allele frequencies
    tbl_in = bits.nbytes + S * 3
    tbl_out = S * (8 + 2 + 56 + 56 + 4)
    win_bytes = S * 56 + S * T * 2 / 8 + nW * T * (24 + 4 + 16)
    out = {"workload": "C2 non-LD: %d sites x %d targets, window %d, depth Poisson(2)" % (S, T, W),
           "site_targets_per_s": S * T / (sum(st.values()) * 1e-3), "ms_device": sum(st.values()), "ms_wall_incl_d2h": wall * 1e3,
           "kernels_ms": st, "windows_per_target": nW,
           "site_table": {"algorithmic_GB": (tbl_in + tbl_out) / 1e9, "achieved_GBs": (tbl_in + tbl_out) / 1e9 / (st["site_table"] * 1e-3),
                          "frac_of_measured_hbm": (tbl_in + tbl_out) / 1e9 / (st["site_table"] * 1e-3) / hbm},
           "window_nonld": {"algorithmic_GB": win_bytes / 1e9, "achieved_GBs": win_bytes / 1e9 / (st["window_nonld"] * 1e-3),
                            "frac_of_measured_hbm": win_bytes / 1e9 / (st["window_nonld"] * 1e-3) / hbm},
           "hbm_peak_GBs": hbm}
    print(json.dumps(out))


def c4(steps=3, n_tables=10_000, n_bins=10_000):
    import torch
    rng = np.random.default_rng(0)
    nb = n_tables * n_bins
    seg = np.repeat(rng.integers(0, 3, nb // 200 + 1), 200)[:nb]
    ll = rng.normal(-150.0, 10.0, (nb, 3))
    ll[np.arange(nb), seg] += 8.0  # planted IBD0/1/2 segments
    off = (np.arange(n_tables + 1, dtype=np.int64) * n_bins)
    with ib.Engine(ib.Params()) as e:
        e.enable_timing(True)
        e.viterbi_batch(ll, off, True)
        e.reset_stats()
        t0 = time.perf_counter()
        for _ in range(steps):
            state, score, counts = e.viterbi_batch(ll, off, True)
        wall = (time.perf_counter() - t0) / steps
        st = {k: v[0] / steps for k, v in e.kernel_stats().items() if v[1]}
    hbm = peaks()
    ms = sum(st.values())
    byt = nb * 49.0  # 24 B in, 24 + 1 B out per bin
    frac_planted = float((state == seg).mean())
    print(json.dumps({"workload": "C4 hiddengem: %d tables x %d bins (log-likelihood front-end)" % (n_tables, n_bins),
                      "bins_per_s": nb / (ms * 1e-3), "ms_device": ms, "ms_wall_incl_copies": wall * 1e3, "kernels_ms": st,
                      "algorithmic_GB": byt / 1e9, "achieved_GBs": byt / 1e9 / (ms * 1e-3),
                      "frac_of_measured_hbm": byt / 1e9 / (ms * 1e-3) / hbm, "hbm_peak_GBs": hbm,
                      "planted_state_recovery": frac_planted}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c4"]
    if "c2" in which:
        c2()
    if "c4" in which:
        c4()

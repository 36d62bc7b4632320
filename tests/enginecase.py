"""Drives the CUDA engine (through the C ABI) for a refcases.Case and returns results in the
same shape the oracle produces, so the same formatters / comparisons apply to both."""
from __future__ import annotations

import ctypes
import math

import numpy as np

import ibdgem_b200 as ib


def draw_downsampled_counts(case, f_site, keep_site):
    """-D thinning exactly as one reference process draws it (src/ibdgem.c:126-137, 627-628):
    glibc rand() from its default seed, consumed target by target, site by site, n_ref bases
    then n_alt bases, only at sites that pass every filter for that target."""
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    RAND_MAX = 2147483647
    pk, prm = case.pk, case.params
    S = len(pk.pos)
    out = np.zeros((len(case.targets), S, 2), np.uint8)
    for k, t in enumerate(case.targets):
        for s in range(S):
            if not keep_site[s]:
                continue
            if prm.opt_v and pk.hap[s, 2 * t] == 0 and pk.hap[s, 2 * t + 1] == 0:
                continue
            for j, c in enumerate((int(pk.n_ref[s]), int(pk.n_alt[s]))):
                kept = 0
                for _ in range(c):
                    if libc.rand() / RAND_MAX < prm.cull_p:
                        kept += 1
                out[k, s, j] = kept
    return out


def run_engine(case, device=0, force_general=False, expanded=True, align_words=4):
    pk, prm = case.pk, case.params
    ep = ib.Params(epsilon=prm.epsilon, max_cov=prm.max_cov, window_size=prm.window, min_af=prm.min_af,
                   max_af=prm.max_af, variable_sites_only=prm.opt_v, device=device)
    with ib.Engine(ep) as e:
        e.upload_sites(pk.pos, pk.n_ref, pk.n_alt, pk.host_keep, case.af_user)
        e.upload_panel(ib.pack_bits(pk.hap, align_words), len(pk.names))
        e.prepare()
        f, st_shared, lik7 = e.get_site_table()
        tc = None
        if prm.cull_p != 1.0:
            tc = draw_downsampled_counts(case, f, st_shared != 0)
        if force_general:
            e.force_general_ld(True)
        linear = not (prm.ld_mode and (prm.opt_v or tc is not None))  # not offered for --LD with per-target windows
        if prm.ld_mode:
            sc = e.score_ld(case.targets, case.bg, prm.pu_idx, tgt_counts=tc, expanded=expanded, linear=linear)
        else:
            sc = e.score_nonld(case.targets, tgt_counts=tc, expanded=expanded, linear=linear)
        stats = e.kernel_stats()
    results = []
    for k, t in enumerate(case.targets):
        nw = int(sc.n_windows[k])
        ll = sc.w_loglik[k, :nw]
        with np.errstate(over="ignore", under="ignore", invalid="ignore"):
            lin = np.exp(ll)
        if tc is not None:
            nr, na = tc[k, :, 0], tc[k, :, 1]
        else:
            nr, na = pk.n_ref, pk.n_alt
        res = dict(status=sc.site_status[k] if expanded else None, f=f, n_ref=nr, n_alt=na,
                   ibd0=sc.site_lik[k, :, 0] if expanded else None,
                   ibd1=sc.site_lik[k, :, 1] if expanded else None,
                   ibd2=sc.site_lik[k, :, 2] if expanded else None,
                   n_windows=nw, w_start=sc.w_start[k, :nw], w_end=sc.w_end[k, :nw],
                   w_nsites=sc.w_nsites[k, :nw], w_log=ll, w_lin=lin, processed=int(sc.processed[k]),
                   skipped=int(sc.skipped[k]), final_total_cov=int(sc.final_total_cov[k]),
                   final_dist=sc.final_dist[k], ld_path=sc.extra.get("ld_path"), lik7=lik7,
                   w_lin_exact=(sc.w_lik_linear[k, :nw] if linear else None), ld_mode=bool(prm.ld_mode),
                   st_shared=st_shared, kernel_stats=stats)
        results.append(res)
    return results


def assert_matches_oracle(res, ora, ll_atol=1e-6, site_rtol=1e-9):
    """Parity bar of BASELINE.json: integers bit-exact; per-site likelihoods |d| <= 1e-9 |LL|;
    per-window log-likelihood sums |d| <= 1e-6 absolute."""
    assert res["n_windows"] == ora["n_windows"]
    np.testing.assert_array_equal(res["w_start"], ora["w_start"])
    np.testing.assert_array_equal(res["w_end"], ora["w_end"])
    np.testing.assert_array_equal(res["w_nsites"], ora["w_nsites"])
    assert res["processed"] == ora["processed"]
    assert res["skipped"] == ora["skipped"]
    assert res["final_total_cov"] == ora["final_total_cov"]
    np.testing.assert_array_equal(np.asarray(res["final_dist"], np.uint64), ora["final_dist"])
    if res["status"] is not None:
        np.testing.assert_array_equal(res["status"], ora["status"])
        m = ora["status"] != 0
        np.testing.assert_array_equal(np.asarray(res["n_ref"])[m], ora["n_ref"][m])
        np.testing.assert_array_equal(np.asarray(res["n_alt"])[m], ora["n_alt"][m])
        np.testing.assert_array_equal(res["f"][m], ora["f"][m])  # popcount / 2N: bit-exact
        for k in ("ibd0", "ibd1", "ibd2"):
            a, b = res[k][m], ora[k][m]
            # |d ln L| <= 1e-9 * |ln L|  (and the values themselves to 1e-9 relative)
            np.testing.assert_allclose(a, b, rtol=site_rtol, atol=0)
            la, lb = np.log(a), np.log(b)
            assert np.all(np.abs(la - lb) <= site_rtol * np.abs(lb) + 1e-300)
    if res.get("w_lin_exact") is not None:
        # the reference's own linear window products, bit for bit (denormals and zeros included): all three columns
        # of a non-LD row, the LIBD2 column of an --LD row
        cols = [2] if res["ld_mode"] else [0, 1, 2]
        np.testing.assert_array_equal(res["w_lin_exact"][:, cols], ora["w_lin"][:, cols])
    a, b = res["w_log"], ora["w_log"]
    assert a.shape == b.shape
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    np.testing.assert_array_equal(nan_a, nan_b)
    np.testing.assert_allclose(a[~nan_a], b[~nan_b], rtol=0, atol=ll_atol)

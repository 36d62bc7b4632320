"""Pins the oracle (oracle/ibdgem_oracle.c) against the reference's own golden vectors.

SURVEY.md §8(c): the 18 files under supplementary/ibdgem-test/output are the only known-answer
vectors the reference ships; they pin M1-M5, A1, F1, F2, W1, W2 (incl. an underflow-to-zero row).
"""
import math
import os

import numpy as np
import pytest

import oracle
import refio


def _pack(fixture_dir, k):
    i = os.path.join(fixture_dir, "input")
    return refio.pack_impute(os.path.join(i, "test.hap"), os.path.join(i, "test.legend"),
                             os.path.join(i, "test.indv"), os.path.join(i, f"test{k}.pileup"))


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("t", [0, 1, 2])
def test_fixture_replay_byte_exact(fixture_dir, k, t):
    pk = _pack(fixture_dir, k)
    names = pk.names
    pu_idx = names.index(f"sample{k}")
    prm = oracle.Params(pu_idx=pu_idx)
    res = oracle.compare_target(prm, pk.pos, pk.host_keep, pk.n_ref, pk.n_alt, pk.hap, t,
                                np.arange(len(names)))
    out = os.path.join(fixture_dir, "output")
    with open(os.path.join(out, f"sample{k}.{names[t]}.tab.txt")) as fh:
        gold_tab = fh.read().split("\n", 2)[2]  # drop "# Entered command" line + blank line
    with open(os.path.join(out, f"sample{k}.{names[t]}.summary.txt")) as fh:
        gold_sum = fh.read()
    assert refio.format_tab(pk, res, t, prm.max_cov) == gold_tab
    assert refio.format_summary(res) == gold_sum


def test_known_answer_snp1(fixture_dir):
    """sample1.sample1 row SNP1 — n_ref 5, n_alt 2, GT 0|1, f 0.5."""
    pk = _pack(fixture_dir, 1)
    res = oracle.compare_target(oracle.Params(pu_idx=0), pk.pos, pk.host_keep, pk.n_ref, pk.n_alt,
                                pk.hap, 0, np.arange(3))
    assert (pk.n_ref[0], pk.n_alt[0]) == (5, 2)
    assert res["ibd2"][0] == 21.0 / 128.0
    assert "%e" % res["ibd0"][0] == "8.392950e-02"


def test_ncK_table_matches_math_comb():
    L = oracle.lib()
    n = 62  # exact while n <= 62
    tab = L.orc_init_nCk(n)
    for i in range(n + 1):
        for j in range(n + 1):
            want = math.comb(i, j) if j <= i else 0
            assert L.orc_retrieve_nCk(tab, i, j) == want
    L.orc_destroy_nCk(tab, n)


def test_log_space_matches_linear_where_normal(fixture_dir):
    pk = _pack(fixture_dir, 2)
    res = oracle.compare_target(oracle.Params(pu_idx=1, window=10), pk.pos, pk.host_keep, pk.n_ref,
                                pk.n_alt, pk.hap, 2, np.arange(3))
    lin, lg = res["w_lin"], res["w_log"]
    ok = lin > 1e-300
    assert ok.any()
    np.testing.assert_allclose(np.log(lin[ok]), lg[ok], rtol=0, atol=1e-9)

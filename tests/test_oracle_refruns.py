"""Pins the oracle against outputs of the reference itself (oracle/_ref/ibdgem, hiddengem) for
the parts of the path no shipped fixture covers: --LD, -v, -D, -B, -S/-s, -A, -p, -F/-f, -M, -e,
-N naming a panel member, and hiddengem (fixtures: tests/golden/ref_runs, made by
tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import oracle
import refcases
import refio


@pytest.mark.parametrize("run", refcases.ALL_RUNS)
def test_oracle_vs_reference_run_byte_exact(golden_dir, run):
    case = refcases.load_case(os.path.join(golden_dir, "ref_runs", "caseA"), run)
    results = refcases.oracle_run(case)
    assert len(results) > 0
    for t, res in zip(case.targets, results):
        tab, summ = refcases.golden_texts(case, t)
        assert refio.format_tab(case.pk, res, t, case.params.max_cov, case.params.cull_p) == tab
        assert refio.format_summary(res) == summ
        # the log-space values agree with the printed linear ones wherever those are normal
        lin, lg = res["w_lin"], res["w_log"]
        ok = lin > 1e-290
        np.testing.assert_allclose(np.log(lin[ok]), lg[ok], rtol=0, atol=1e-9)


@pytest.mark.parametrize("k", [1, 2])
def test_oracle_vs_reference_fixture_ld(golden_dir, fixture_dir, k):
    inp = os.path.join(fixture_dir, "input")
    case = refcases.load_case(inp, "", hap="test.hap", legend="test.legend", indv="test.indv",
                              pileup=f"test{k}.pileup", args=["--LD", "-w", "10"], pileup_name=f"sample{k}")
    case.out_dir = os.path.join(golden_dir, "ref_runs", "fixtureLD", f"test{k}_ld_w10")
    for t, res in zip(case.targets, refcases.oracle_run(case)):
        tab, summ = refcases.golden_texts(case, t)
        assert refio.format_tab(case.pk, res, t, 20) == tab
        assert refio.format_summary(res) == summ


def _fmt_hidden(state, score_ld):
    o = ["Segment\tIBD0_Score\tIBD1_Score\tIBD2_Score\tInferred_State"]
    for i in range(len(state)):
        o.append("%d\t%s\t%s\t%s\t%d" % (i + 1, *[_le(score_ld[i, s]) for s in range(3)], state[i]))
    n = len(state)
    for s in range(3):
        c = int((state == s).sum())
        o.append("#%% IBD%d (n = %d): %.2f" % (s, c, c / n * 100))
    return "\n".join(o) + "\n"


def _le(x):
    """%.5Le of an x87 long double."""
    x = np.longdouble(x)
    if x != x:
        return "-nan"
    if x == 0:
        return "0.00000e+00"
    e = int(np.floor(np.log10(x)))
    m = x / np.longdouble(10) ** e
    s = "%.5f" % float(m)
    if s.startswith("10."):
        e += 1
        s = "%.5f" % float(m / 10)
    return "%se%s%02d" % (s, "-" if e < 0 else "+", abs(e))


@pytest.mark.parametrize("name,pen", [("ind2.default", ()), ("ind3.default", ()), ("ind5.default", ()),
                                      ("ind2.loose", (0.2, 0.05, 0.3)), ("ind3.loose", (0.2, 0.05, 0.3)),
                                      ("ind5.loose", (0.2, 0.05, 0.3)), ("ind3.ld_underflow", ())])
def test_oracle_hiddengem_vs_reference(golden_dir, name, pen):
    ca = os.path.join(golden_dir, "ref_runs", "caseA")
    tgt = name.split(".")[0]
    run = "ld_w100_underflow" if "underflow" in name else "nonld_w10"
    rows = refio.read_summary(os.path.join(ca, run, f"UNKWN.{tgt}.summary.txt"))
    l = np.array([[r[2], r[3], r[4]] for r in rows])
    state, score, score_ld = oracle.hiddengem(l, *pen)
    with open(os.path.join(ca, "hiddengem", name + ".txt")) as fh:
        gold = fh.read()
    assert _fmt_hidden(state, score_ld) == gold
    fin = np.isfinite(score)
    np.testing.assert_allclose(score[fin], np.log(score_ld[fin].astype(np.float64)), atol=1e-12, rtol=1e-12)

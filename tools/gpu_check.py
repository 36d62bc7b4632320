"""Quick GPU-side diagnosis of the --LD paths against the oracle (prints max |dLL| per column
instead of asserting).  Test infrastructure; run under `timeout` on a GPU box."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import enginecase as ec  # noqa: E402
import refcases  # noqa: E402
from test_gpu_parity import _synth_case  # noqa: E402


def check(name, case):
    ora = refcases.oracle_run(case)
    for fg in (True, False):
        t0 = time.time()
        res = ec.run_engine(case, force_general=fg, expanded=False)
        dt = time.time() - t0
        worst = np.zeros(3)
        bad_int = 0
        for r, o in zip(res, ora):
            if r["n_windows"] != o["n_windows"] or not np.array_equal(r["w_nsites"], o["w_nsites"]) or \
                    not np.array_equal(r["w_start"], o["w_start"]) or not np.array_equal(r["w_end"], o["w_end"]):
                bad_int += 1
                continue
            a, b = r["w_log"], o["w_log"]
            if not np.array_equal(np.isnan(a), np.isnan(b)):
                bad_int += 1
                continue
            d = np.abs(np.where(np.isnan(a), 0, a) - np.where(np.isnan(b), 0, b))
            d = np.where(np.isfinite(d), d, 1e300)
            if d.size:
                worst = np.maximum(worst, d.max(axis=0))
        print(f"{name:28s} force_general={fg!s:5s} path={res[0]['ld_path']} max|dLL| IBD0/1/2 = "
              f"{worst[0]:.3g} {worst[1]:.3g} {worst[2]:.3g}  int_mismatch={bad_int}  ({dt:.2f}s)", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["a", "b", "c", "d", "e"]
    if "a" in which:
        check("w200 S4000 N64 T8", _synth_case(1, 4000, 64, 200, True, range(8)))
    if "b" in which:
        check("w100 S3000 N40 T5 pu2", _synth_case(107, 3000, 40, 100, True, range(5), pu_idx=2))
    if "c" in which:
        check("w1000 S6100 N70 T9 pu2", _synth_case(1007, 6100, 70, 1000, True, range(9), pu_idx=2))
    if "d" in which:
        check("w37 S1500 N33 T33", _synth_case(44, 1500, 33, 37, True, range(33), pu_idx=2))
    if "e" in which:
        check("w500 S20000 N300 T100 bgsub", _synth_case(5, 20000, 300, 500, True, range(100), bg=list(range(50, 300)) + [3, 3, 7]))

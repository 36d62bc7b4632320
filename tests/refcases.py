"""Turns an `ibdgem` command line (as recorded in tests/golden/ref_runs/*/ARGS.json) into the
packed arrays + parameters the oracle and the engine take — the host-side steps of
src/ibdgem.c:868-1171 that precede the hot path.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass

import numpy as np

import oracle
import refio


@dataclass
class Case:
    pk: refio.Packed
    params: oracle.Params
    targets: list      # individual ordinals, in output order
    bg: np.ndarray     # background ordinals
    af_user: np.ndarray | None
    pileup_name: str
    out_dir: str


def _names_file(path):
    with open(path) as fh:
        return [ln.rstrip("\n") for ln in fh]


def load_case(case_dir: str, run: str, hap="panel.hap", legend="panel.legend", indv="panel.indv",
              pileup="unk.pileup", args=None, pileup_name=None) -> Case:
    out_dir = os.path.join(case_dir, run)
    if args is None:
        with open(os.path.join(out_dir, "ARGS.json")) as fh:
            meta = json.load(fh)
        args, pileup_name = meta["args"], meta["pileup_name"]
    opt = {"w": 100, "M": 20, "e": 0.02, "f": 0.0, "F": 1.0}
    flags = set()
    i = 0
    while i < len(args):
        a = args[i]
        if a in ("--LD", "-v"):
            flags.add(a)
            i += 1
        else:
            opt[a.lstrip("-")] = args[i + 1]
            i += 2
    chrom = opt.get("c")
    positions = None
    if "p" in opt:  # src/ibd-parse.c:361-421
        positions = []
        for ln in _names_file(os.path.join(case_dir, opt["p"])):
            f = ln.split()
            if len(f) >= 2 and (chrom is None or f[0] == chrom):
                positions.append(int(f[2]) if len(f) >= 3 else int(f[1]))
    pk = refio.pack_impute(os.path.join(case_dir, hap), os.path.join(case_dir, legend),
                           os.path.join(case_dir, indv), os.path.join(case_dir, pileup), chrom, positions)
    names = pk.names
    max_cov = int(opt["M"])
    prm = oracle.Params(epsilon=float(opt["e"]), max_cov=max_cov, window=int(opt["w"]),
                        min_af=float(opt["f"]), max_af=float(opt["F"]), ld_mode=int("--LD" in flags),
                        opt_v=int("-v" in flags),
                        pu_idx=names.index(pileup_name) if pileup_name in names else -1)
    if "D" in opt:  # src/ibdgem.c:83-106
        _, mean = refio.input_cov_dist(pk.pileup, max_cov)
        tgt = float(opt["D"])
        prm.cull_p = 1.0 if tgt > mean else tgt / mean
    if "S" in opt:
        targets = [names.index(n) for n in _names_file(os.path.join(case_dir, opt["S"])) if n in names]
    elif "s" in opt:
        targets = [names.index(n) for n in opt["s"].split(",") if n in names]
    else:
        targets = list(range(len(names)))
    if "B" in opt:
        bg = [names.index(n) for n in _names_file(os.path.join(case_dir, opt["B"])) if n in names]
    else:
        bg = list(range(len(names)))
    af_user = None
    if "A" in opt:  # src/ibd-parse.c:311-358, src/ibdgem.c:609-614
        table = {}
        for ln in _names_file(os.path.join(case_dir, opt["A"])):
            f = ln.split()
            if len(f) >= 3 and (chrom is None or f[0] == chrom):
                table.setdefault(int(f[1]), float(f[2]))
        af_user = np.array([table.get(int(p), math.nan) for p in pk.pos])
    return Case(pk, prm, targets, np.asarray(bg, np.int32), af_user, pileup_name, out_dir)


def oracle_run(case: Case):
    """Oracle results for every target, rand() stream continuing across targets like one
    reference process (src/ibdgem.c:522, 627-628)."""
    res = []
    for k, t in enumerate(case.targets):
        res.append(oracle.compare_target(case.params, case.pk.pos, case.pk.host_keep, case.pk.n_ref,
                                         case.pk.n_alt, case.pk.hap, t, case.bg, af_user=case.af_user,
                                         reseed=(k == 0)))
    return res


def golden_texts(case: Case, t: int):
    name = case.pk.names[t]
    with open(os.path.join(case.out_dir, f"{case.pileup_name}.{name}.tab.txt")) as fh:
        tab = fh.read().split("\n", 2)[2]
    with open(os.path.join(case.out_dir, f"{case.pileup_name}.{name}.summary.txt")) as fh:
        summ = fh.read()
    return tab, summ


ALL_RUNS = ["nonld_w10", "ld_w10", "ld_w10_self", "ld_w25_bg", "ld_v_w10", "nonld_v_w7", "ld_D1_w10",
            "nonld_D05_v", "filters", "af_pos", "ld_w100_underflow"]

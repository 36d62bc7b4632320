#!/usr/bin/env python
"""Generates tests/golden/ref_runs/* by running the UNMODIFIED reference (oracle/_ref/ibdgem and
oracle/_ref/hiddengem, built from /root/reference/src by oracle/Makefile) on small seeded
synthetic inputs.  These outputs pin the parts of the hot path that the reference's shipped
fixtures do not cover: --LD, -v, -D, -B, -S/-s, -A, -p, -F/-f, -M, -e, -N in
panel, and hiddengem.

Run from the repo root in the build container:  python tests/golden/make_golden.py
The inputs are written next to the outputs so the tests can re-pack them on the GPU box.
"""
import json
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refio  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(HERE, "ref_runs")


def synth(seed, S, N, depth_mean=2.0, src=0):
    rng = np.random.default_rng(seed)
    af = np.clip(rng.beta(0.5, 2.0, S), 0.02, 0.98)
    hap = (rng.random((S, 2 * N)) < af[:, None]).astype(np.uint8)
    pos = 1000 + 60 * np.arange(S) + rng.integers(0, 50, S)
    bases = "ACGT"
    ref = [bases[i] for i in rng.integers(0, 4, S)]
    alt = [bases[(bases.index(r) + 1 + int(k)) % 4] for r, k in zip(ref, rng.integers(0, 3, S))]
    depth = rng.poisson(depth_mean, S)
    g = hap[:, 2 * src] + hap[:, 2 * src + 1]
    p_alt = np.where(g == 0, 0.02, np.where(g == 1, 0.5, 0.98))
    n_alt = rng.binomial(depth, p_alt)
    n_ref = depth - n_alt
    extra = (rng.random(S) < 0.1).astype(int)
    return dict(hap=hap, pos=pos, ref=ref, alt=alt, n_ref=n_ref, n_alt=n_alt, extra=extra, rng=rng)


def write_case(case_dir, d, names, drop_sites=(), indel_sites=(), deep_sites=()):
    os.makedirs(case_dir, exist_ok=True)
    ref = list(d["ref"]); alt = list(d["alt"])
    for s in indel_sites:
        alt[s] = alt[s] + "T"  # not a SNP -> legend filter (src/ibdgem.c:593)
    refio.write_impute(case_dir, "panel", d["hap"], d["pos"], names, ref, alt)
    keep = np.ones(len(d["pos"]), bool)
    keep[list(drop_sites)] = False
    n_ref = d["n_ref"].copy(); n_alt = d["n_alt"].copy()
    for s in deep_sites:
        n_ref[s] += 25  # exceeds -M
    idx = np.nonzero(keep)[0]
    refio.write_pileup(os.path.join(case_dir, "unk.pileup"), "7", d["pos"][idx], n_ref[idx], n_alt[idx],
                       [d["ref"][i] for i in idx], [d["alt"][i] for i in idx], d["extra"][idx])


def run(case_dir, name, args, pileup_name):
    out = os.path.join(case_dir, name)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    cmd = [os.path.join(REF, "ibdgem"), "-H", "panel.hap", "-L", "panel.legend", "-I", "panel.indv",
           "-P", "unk.pileup", "-O", name] + args
    r = subprocess.run(cmd, cwd=case_dir, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with open(os.path.join(out, "ARGS.json"), "w") as fh:
        json.dump({"args": args, "pileup_name": pileup_name}, fh)
    return out


def hidden(summary, args, dst):
    r = subprocess.run([os.path.join(REF, "hiddengem"), "-s", summary] + args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with open(dst, "w") as fh:
        fh.write(r.stdout)


def main():
    assert os.path.exists(os.path.join(REF, "ibdgem")), "build oracle/_ref first (make -C oracle ref)"
    shutil.rmtree(OUT, ignore_errors=True)

    # ---- case A: 8 individuals x 240 sites ---------------------------------------------------
    N, S = 8, 240
    names = ["ind%d" % i for i in range(N)]
    d = synth(11, S, N, src=2)
    ca = os.path.join(OUT, "caseA")
    write_case(ca, d, names, drop_sites=range(5, S, 17), indel_sites=range(3, S, 29), deep_sites=range(7, S, 41))
    with open(os.path.join(ca, "targets.txt"), "w") as fh:
        fh.write("ind1\nind2\nnosuch\nind5\n")
    with open(os.path.join(ca, "bg.txt"), "w") as fh:
        fh.write("ind0\nind2\nind3\nind4\nind6\nind7\n")
    with open(os.path.join(ca, "af.txt"), "w") as fh:
        for s in range(0, S, 3):
            fh.write("7 %d %f\n" % (d["pos"][s], 0.05 + 0.9 * ((s * 7) % 100) / 100.0))
    with open(os.path.join(ca, "pos.txt"), "w") as fh:
        for s in range(0, S, 2):
            fh.write("7 %d\n" % d["pos"][s])

    run(ca, "nonld_w10", ["-w", "10"], "UNKWN")
    run(ca, "ld_w10", ["--LD", "-w", "10"], "UNKWN")
    run(ca, "ld_w10_self", ["--LD", "-w", "10", "-N", "ind2", "-s", "ind1,ind2,ind7"], "ind2")
    run(ca, "ld_w25_bg", ["--LD", "-w", "25", "-B", "bg.txt", "-S", "targets.txt"], "UNKWN")
    run(ca, "ld_v_w10", ["--LD", "-v", "-w", "10", "-s", "ind0,ind2,ind3"], "UNKWN")
    run(ca, "nonld_v_w7", ["-v", "-w", "7"], "UNKWN")
    run(ca, "ld_D1_w10", ["--LD", "-D", "1.0", "-w", "10", "-s", "ind2,ind4"], "UNKWN")
    run(ca, "nonld_D05_v", ["-D", "0.5", "-v", "-w", "5", "-s", "ind2,ind6,ind1"], "UNKWN")
    run(ca, "filters", ["-f", "0.1", "-F", "0.8", "-M", "4", "-e", "0.01", "-w", "6", "-s", "ind2,ind3"], "UNKWN")
    run(ca, "af_pos", ["--LD", "-A", "af.txt", "-p", "pos.txt", "-c", "7", "-w", "8", "-s", "ind2,ind0"], "UNKWN")
    run(ca, "ld_w100_underflow", ["--LD", "-w", "100", "-s", "ind2,ind3"], "UNKWN")

    # ---- VCF input of the same panel (src/ibdgem.c:185-476).  One target per run: the reference
    # frees its genotype regex inside the target loop and crashes on the second target.
    with open(os.path.join(ca, "panel.vcf"), "w") as fh:
        fh.write("##fileformat=VCFv4.2\n##source=make_golden\n")
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n")
        ref = list(d["ref"]); alt = list(d["alt"])
        for s_ in range(3, S, 29):
            alt[s_] = alt[s_] + "T"
        for s_ in range(S):
            a = alt[s_]
            if s_ % 37 == 11:
                a = a + ",C"  # multi-allelic -> skipped (src/ibdgem.c:275)
            gts = ["%d|%d:%d" % (d["hap"][s_, 2 * i], d["hap"][s_, 2 * i + 1], 20 + i) for i in range(N)]
            if s_ % 43 == 5:
                gts[3] = "./.:0"  # unparsable genotype -> whole site skipped (src/ibdgem.c:280-286)
            qual = "." if s_ % 5 == 0 else "%d" % (10 + (s_ * 13) % 60)
            fh.write("7\t%d\trs%d\t%s\t%s\t%s\tPASS\tNS=%d\tGT:GQ\t%s\n" % (d["pos"][s_], s_, ref[s_], a, qual, N, "\t".join(gts)))

    def run_vcf(name, args):
        out = os.path.join(ca, name)
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
        cmd = [os.path.join(REF, "ibdgem"), "-V", "panel.vcf", "-P", "unk.pileup", "-O", name] + args
        r = subprocess.run(cmd, cwd=ca, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        with open(os.path.join(out, "ARGS.json"), "w") as fh:
            json.dump({"args": args, "pileup_name": "UNKWN", "vcf": "panel.vcf"}, fh)

    run_vcf("vcf_nonld_w10", ["-w", "10", "-s", "ind2"])
    run_vcf("vcf_ld_w10", ["--LD", "-w", "10", "-s", "ind3"])
    run_vcf("vcf_q30_v", ["-q", "30", "-v", "-w", "6", "-s", "ind5"])

    # hiddengem on the non-LD window tables (short windows keep the values normal doubles)
    hg = os.path.join(ca, "hiddengem")
    os.makedirs(hg)
    for t in ("ind2", "ind3", "ind5"):
        src = os.path.join(ca, "nonld_w10", "UNKWN.%s.summary.txt" % t)
        hidden(src, [], os.path.join(hg, "%s.default.txt" % t))
        hidden(src, ["--p01", "0.2", "--p02", "0.05", "--p12", "0.3"], os.path.join(hg, "%s.loose.txt" % t))
    hidden(os.path.join(ca, "ld_w100_underflow", "UNKWN.ind3.summary.txt"), [],
           os.path.join(hg, "ind3.ld_underflow.txt"))

    # ---- case B: the shipped fixture under --LD ----------------------------------------------
    fx = os.path.join(HERE, "ibdgem-test", "input")
    cb = os.path.join(OUT, "fixtureLD")
    os.makedirs(cb)
    for k in (1, 2):
        out = os.path.join(cb, "test%d_ld_w10" % k)
        os.makedirs(out)
        r = subprocess.run([os.path.join(REF, "ibdgem"), "-H", os.path.join(fx, "test.hap"), "-L",
                            os.path.join(fx, "test.legend"), "-I", os.path.join(fx, "test.indv"), "-P",
                            os.path.join(fx, "test%d.pileup" % k), "-N", "sample%d" % k, "--LD", "-w", "10",
                            "-O", out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    # strip the machine-specific "# Entered command" line so the fixtures are reproducible
    for root, _, files in os.walk(OUT):
        for fn in files:
            if fn.endswith(".tab.txt"):
                p = os.path.join(root, fn)
                with open(p) as fh:
                    lines = fh.readlines()
                lines[0] = "# Entered command: (stripped)\n"
                with open(p, "w") as fh:
                    fh.writelines(lines)
    print("golden reference runs written to", OUT)


if __name__ == "__main__":
    main()

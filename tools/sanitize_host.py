"""Runs the host packers (IMPUTE sequential / threaded / cached, VCF, pileup) under AddressSanitizer and
UBSan on the shipped fixtures and on small hostile inputs: haplotype counts around the vector widths,
files without a final newline, truncated lines.  CPU only; scratch under gpurun_out/asan.
python tools/sanitize_host.py   -> one summary line, exit code 1 on any sanitizer report or hash mismatch"""
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "ibdgem_b200", "csrc", "host")
WORK = os.path.join(ROOT, "gpurun_out", "asan")
DRIVER = r"""
#include <cstdio>
#include <cstdlib>
#include "panel.h"
#include "pileup_store.h"
using namespace ibdhost;
int main(int argc, char **argv) {  // mode (0 impute, 1 vcf, 2 impute through the cache, 3 vcf through the cache) files...
    const int mode = atoi(argv[1]);
    PileupStore pu;
    PackOptions po;
    PackedPanel panel;
    int rc = 0;
    if (load_pileup(argv[(mode == 1 || mode == 3) ? 3 : 5], nullptr, &pu)) return 3;
    if (mode == 0) {
        std::vector<std::string> names;
        rc = read_indv(argv[4], &names) || pack_impute(argv[2], argv[3], names, pu, po, &panel);
    } else if (mode == 1) {
        rc = pack_vcf(argv[2], pu, po, &panel);
    } else if (mode == 3) {
        bool hit;
        rc = pack_vcf_cached(argv[2], argv[4], pu, po, &panel, &hit);
    } else {
        bool hit;
        rc = pack_impute_cached(argv[2], argv[3], argv[4], argv[6], pu, po, &panel, &hit);
    }
    unsigned long h = 0;
    for (uint32_t w : panel.bits) h = h * 1315423911u + w;
    for (uint8_t k : panel.host_keep) h = h * 31 + k;
    printf("%d %ld %lx\n", rc, (long)panel.S, h);
    return rc;
}
"""


def main():
    shutil.rmtree(WORK, ignore_errors=True)
    os.makedirs(WORK)
    os.chdir(WORK)
    open("drv.cpp", "w").write(DRIVER)
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I" + HOST,
                    "drv.cpp"] + [os.path.join(HOST, f) for f in ("panel.cpp", "pileup_store.cpp", "textio.cpp")] +
                   ["-lz", "-lpthread", "-o", "drv"], check=True)
    reports, runs = [], 0

    def run(args, env=None):
        nonlocal runs
        runs += 1
        r = subprocess.run(["./drv"] + args, capture_output=True, text=True, env=dict(os.environ, **(env or {})))
        if "AddressSanitizer" in r.stderr or "runtime error" in r.stderr:
            reports.append((args, r.stderr[:1500]))
        return r.stdout.strip()

    mt = {"IBDGEM_PACK_MT_MIN_BYTES": "1"}
    fx = os.path.join(ROOT, "tests", "golden", "ibdgem-test", "input")
    a = [os.path.join(fx, f) for f in ("test.hap", "test.legend", "test.indv", "test1.pileup")]
    assert run(["0"] + a) == run(["0"] + a, mt)
    rng = np.random.default_rng(3)
    for N in (1, 3, 4, 7, 8, 9, 16, 17, 37):
        S = 53
        hap = (rng.random((S, 2 * N)) < 0.4).astype(int)
        open("o.hap", "w").write("\n".join(" ".join(map(str, r)) for r in hap))  # no final newline
        open("o.legend", "w").write("id position a0 a1\n" + "".join(f"rs{s} {100 + 7 * s} A C\n" for s in range(S)))
        open("o.indv", "w").write("".join(f"i{i}\n" for i in range(N)))
        open("o.pileup", "w").write("".join(f"1\t{100 + 7 * s}\tA\t2\t.c\tII\t]]\n" for s in range(S))[:-1])
        lines = ["##x", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"i{i}" for i in range(N))]
        for s in range(S):
            lines.append(f"1\t{100 + 7 * s}\trs{s}\tA\tC\t50\tPASS\t.\tGT\t" + "\t".join(
                f"{hap[s, 2 * i]}|{hap[s, 2 * i + 1]}" + (":9" if (s + i) % 11 == 0 else "") for i in range(N)))
        open("o.vcf", "w").write("\n".join(lines))
        for f in ("o.cache", "v.cache"):
            if os.path.exists(f):
                os.remove(f)
        imp = ["o.hap", "o.legend", "o.indv", "o.pileup"]
        outs = {run(["0"] + imp), run(["0"] + imp, mt), run(["2"] + imp + ["o.cache"]), run(["2"] + imp + ["o.cache"]),
                run(["1", "o.vcf", "o.pileup"]), run(["1", "o.vcf", "o.pileup"], mt), run(["3", "o.vcf", "o.pileup", "v.cache"]),
                run(["3", "o.vcf", "o.pileup", "v.cache"])}
        assert len(outs) == 1, (N, outs)  # same bits and keep flags by every route
    open("t.hap", "w").write("0 1 0\n0 1")
    open("t.legend", "w").write("id position a0 a1\nrs1 100 A C\nrs2 107 A C\n")
    open("t.indv", "w").write("a\nb\n")
    run(["0", "t.hap", "t.legend", "t.indv", "o.pileup"], mt)
    open("t.pileup", "w").write("1\t100\tA\t2\t..\tII\n1\t\n\t\t\t\n1\t107\tA\t3\t.^\tIII\t]]]\n1\t114\tA\t1\t+\tI\t]\n1\t121\tA\t1\t.+99\tI\t]")
    run(["0", "o.hap", "o.legend", "o.indv", "t.pileup"])
    open("t.vcf", "w").write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ta\tb\tc\td\te\tf\tg\th\n"
                             "1\t100\tr\tA\tC\t5\tP\t.\tGT\t0|1\t0|1\t0|1\t0|1\t0|1\t0|1\t0|1\t0|\n"
                             "1\t107\tr\tA\tC\t5\tP\t.\tGT\t\n1\t114\tr\tA\tC\t5\tP\t.\tGT")
    run(["1", "t.vcf", "o.pileup"])
    run(["1", "t.vcf", "o.pileup"], mt)
    os.chdir(ROOT)
    shutil.rmtree(WORK, ignore_errors=True)
    for args, text in reports:
        print("SANITIZER REPORT for", args, "\n", text)
    print(f"{runs} runs, {len(reports)} sanitizer reports")
    return 1 if reports else 0


if __name__ == "__main__":
    sys.exit(main())

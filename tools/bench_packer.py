"""Host packer throughput: parse of IMPUTE text against the binary panel cache — CPU
only — and, with --cli on a GPU box, the wall time of bin/ibdgem with and without the tab.txt tables
(8f-2).  Writes its synthetic panel under $IBDGEM_BENCH_TMP (default gpurun_out/packer_tmp) and removes it.
python tools/bench_packer.py [sites] [samples] [--cli [targets]]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hostlib  # noqa: E402

import shutil  # noqa: E402
import subprocess  # noqa: E402

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
S = int(argv[0]) if len(argv) > 0 else 50_000
N = int(argv[1]) if len(argv) > 1 else 2504
CLI = "--cli" in sys.argv
T_CLI = int(argv[2]) if len(argv) > 2 else 64
work = os.environ.get("IBDGEM_BENCH_TMP") or os.path.join(ROOT, "gpurun_out", "packer_tmp")
os.makedirs(work, exist_ok=True)
hap, leg, indv, pu, cache = (os.path.join(work, f) for f in ("p.hap", "p.legend", "p.indv", "p.pileup", "p.cache"))
rng = np.random.default_rng(1)
t0 = time.time()
with open(hap, "wb") as fh:
    for s0 in range(0, S, 2000):
        n = min(2000, S - s0)
        a = (rng.random((n, 2 * N)) < 0.2).astype(np.uint8) + ord("0")
        line = np.empty((n, 4 * N), np.uint8)
        line[:, 0::2] = a
        line[:, 1::2] = ord(" ")
        line[:, -1] = ord("\n")
        fh.write(line.tobytes())
with open(leg, "w") as fh:
    fh.write("id position a0 a1\n")
    for s in range(S):
        fh.write(f"rs{s} {1000 + 60 * s} A G\n")
with open(indv, "w") as fh:
    for i in range(N):
        fh.write(f"i{i}\n")
with open(pu, "w") as fh:
    for s in range(0, S, 2):
        fh.write(f"1\t{1000 + 60 * s}\tA\t2\t.g\tII\t]]\n")
print(f"synthetic panel: {S} sites x {N} samples, .hap {os.path.getsize(hap) / 1e9:.2f} GB (written in {time.time() - t0:.1f} s)")
hostlib.build()
if os.path.exists(cache):
    os.remove(cache)
for label in ("parse text + write cache", "load cache", "load cache"):
    t0 = time.time()
    got, hit = hostlib.pack_cached(hap, leg, indv, cache, pu)
    dt = time.time() - t0
    print(f"{label:26s} hit={hit} {dt:7.3f} s  ({os.path.getsize(hap) / 1e9 / dt:.2f} GB of .hap text per second)")
t0 = time.time()
got = hostlib.pack(0, hap, leg, indv, pu)
dt = time.time() - t0
print(f"{'parse text (no cache)':26s}       {dt:7.3f} s  ({os.path.getsize(hap) / 1e9 / dt:.2f} GB/s), cache file {os.path.getsize(cache) / 1e6:.0f} MB")
if CLI:
    with open(os.path.join(work, "targets.txt"), "w") as fh:
        for i in range(min(T_CLI, N)):
            fh.write(f"i{i}\n")
    runs = [("ibdgem", "summaries only (--no-tab)", ["--no-tab", "--panel-cache", cache]),
            ("ibdgem", "tab.txt + summaries", ["--panel-cache", cache])]
    if os.path.exists(os.path.join(hostlib.BIN, "ibdgem_prev")):  # an earlier build of the front-end, for A/B
        runs.append(("ibdgem_prev", "tab.txt + summaries", []))
    for binary, label, extra in runs:
        out = os.path.join(work, "out")
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
        t0 = time.time()
        r = subprocess.run([os.path.join(hostlib.BIN, binary), "-H", hap, "-L", leg, "-I", indv, "-P", pu, "-S",
                            os.path.join(work, "targets.txt"), "-O", out] + extra, capture_output=True, text=True)
        dt = time.time() - t0
        nbytes = sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out))
        print(f"bin/{binary:12s} {label:28s} rc={r.returncode} {dt:7.2f} s wall, {min(T_CLI, N)} targets, {nbytes / 1e9:.2f} GB written")
shutil.rmtree(work, ignore_errors=True)

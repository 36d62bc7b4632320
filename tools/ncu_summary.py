"""Summarises ncu captures for profiles/: `--page raw` metrics of every kernel in a .ncu-rep, and the
per-kernel totals of a `--metrics gpu__time_duration.sum` launch list.  Runs where ncu is installed
(no GPU needed):  python tools/ncu_summary.py raw <file.ncu-rep> | list <launches.csv>"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("| kernel | " + " | ".join(k.split(".")[0] + "." + k.split(".")[1] if "." in k else k for k in KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])
        name = name.replace("void ", "").replace("ibdgem::", "")[:48]
        cells = []
        for k in KEYS:
            hit = [j for j, h in enumerate(hdr) if h == k or h.endswith("." + k)]  # some carry a section prefix
            if hit:
                i = hit[0]
                cells.append("%s %s" % (r[i], units[i]))
            else:
                cells.append("n/a")
        print("| %s | %s |" % (name, " | ".join(cells)))


def launches(path):
    lines = open(path).read().splitlines()
    i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines[i:]):
        n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")[:60]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", ""))
    ours = {k: v for k, v in agg.items() if "native::" not in k and "at::" not in k and "elementwise" not in k}
    tot = sum(v[1] for v in ours.values())
    print("| kernel | launches | total ms | ms / launch | share of engine kernels |")
    print("|---|---|---|---|---|")
    for k, (c, t) in ours.items():
        print("| %s | %d | %.3f | %.4f | %.1f %% |" % (k.replace("ibdgem::", ""), c, t / 1e6, t / 1e6 / c, 100 * t / tot))


if __name__ == "__main__":
    {"raw": raw, "list": launches}[sys.argv[1]](sys.argv[2])

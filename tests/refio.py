"""Test-side readers/writers for the reference's text formats — TEST INFRASTRUCTURE ONLY.

Small pure-Python restatements of the host-side parsing the reference does before the hot
path (F1/F2 rows of SURVEY.md §8a) and of its two output tables, so that the oracle can be
driven from the shipped fixture files and its results diffed byte-for-byte against the
golden outputs.  File:line citations are relative to /root/reference.
"""
from __future__ import annotations

import math
import os
import re
from dataclasses import dataclass, field

import numpy as np

MAX_COV_PILEUP = 128  # src/pileup.h:12


@dataclass
class Pileup:
    """In-memory pileup store, src/pileup.c:487-559."""
    chrom: list = field(default_factory=list)
    pos: list = field(default_factory=list)
    cov: list = field(default_factory=list)
    bases: list = field(default_factory=list)

    def index(self):
        self._by_pos = {}
        for i, p in enumerate(self.pos):
            self._by_pos.setdefault(p, i)
        return self

    def fetch(self, pos):  # src/pileup.c:472-485
        return self._by_pos.get(pos)


def parse_pileup_line(line: str):
    """src/pileup.c:206-415.  Returns (chr, pos, cov, bases) or None if the line is dropped."""
    fields = line.split()
    if len(fields) < 4:
        return None
    try:
        pos = int(fields[1])
        cov = int(fields[3])
    except ValueError:
        return None
    ref = fields[2][0]
    if len(fields[2]) != 1:
        return None
    if cov >= MAX_COV_PILEUP:  # :223
        return None
    if len(fields) < 7:  # :232-240 needs base, base-qual and map-qual fields
        return None
    if cov == 0:
        return fields[0], pos, 0, ""
    raw = fields[4]
    out = []
    i = 0
    while i < len(raw):
        c = raw[i]
        if c in ".,":  # :253-265
            out.append(ref)
            i += 1
        elif c in "ACGTNacgtn":
            out.append(c.upper())
            i += 1
        elif c in "+-":  # :337-361
            i += 1
            n = 0
            while i < len(raw) and raw[i].isdigit():
                n = n * 10 + int(raw[i])
                i += 1
            i += n
        elif c == "$":
            i += 1
        elif c == "^":  # :367-369
            i += 2
        elif c == "*":  # :371-376
            out.append("*")
            i += 1
        else:
            return None
    if len(out) != cov:  # :386-391
        return None
    if len(fields[5]) != cov and len(fields[6]) != cov:  # :395-401
        return None
    return fields[0], pos, cov, "".join(out)


def read_pileup(path: str, chrom: str | None = None) -> Pileup:
    pu = Pileup()
    with open(path) as fh:
        for line in fh:
            rec = parse_pileup_line(line)
            if rec is None:
                continue
            if chrom is not None and rec[0] != chrom:  # :523-527
                continue
            pu.chrom.append(rec[0])
            pu.pos.append(rec[1])
            pu.cov.append(rec[2])
            pu.bases.append(rec[3])
    return pu.index()


def read_indv(path: str):
    with open(path) as fh:
        return [ln.rstrip("\n") for ln in fh if ln.rstrip("\n") != ""]


def read_hap(path: str) -> np.ndarray:
    """[S, 2N] uint8.  Alleles are the characters at even offsets, src/ibdgem.c:638-639."""
    rows = []
    with open(path) as fh:
        for ln in fh:
            ln = ln.rstrip("\n")
            rows.append(np.frombuffer(ln[0::2].encode(), dtype=np.uint8) - ord("0"))
    return np.stack(rows).astype(np.uint8)


_SNP = set("ACGT")


def read_legend(path: str):
    """Returns per-line (ok, id, pos, ref, alt); header skipped, src/ibdgem.c:554,589-596."""
    out = []
    with open(path) as fh:
        fh.readline()
        for ln in fh:
            f = ln.split()
            ok = len(f) >= 4 and re.fullmatch(r"\d+", f[1]) is not None
            if ok:
                ok = f[2] in _SNP and f[3] in _SNP
                out.append((ok, f[0], int(f[1]), f[2], f[3]))
            else:
                out.append((False, "", 0, "", ""))
    return out


@dataclass
class Packed:
    """What the host packer hands to the engine / oracle for one chromosome."""
    names: list
    hap: np.ndarray        # [S, 2N] 0/1
    pos: np.ndarray        # [S] uint64
    host_keep: np.ndarray  # [S] uint8
    n_ref: np.ndarray      # [S] uint8
    n_alt: np.ndarray      # [S] uint8
    cov: np.ndarray        # [S] uint32 raw pileup coverage (DP column)
    chrom: list
    ids: list
    ref: list
    alt: list
    pileup: Pileup


def pack_impute(hap_path, legend_path, indv_path, pileup_path, chrom=None, positions=None) -> Packed:
    names = read_indv(indv_path)
    hap = read_hap(hap_path)
    leg = read_legend(legend_path)
    pu = read_pileup(pileup_path, chrom)
    S = min(len(leg), hap.shape[0])
    hap = hap[:S]
    pos = np.zeros(S, np.uint64)
    keep = np.zeros(S, np.uint8)
    nr = np.zeros(S, np.uint8)
    na = np.zeros(S, np.uint8)
    cov = np.zeros(S, np.uint32)
    chroms, ids, refs, alts = [], [], [], []
    posset = set(positions) if positions is not None else None
    for s in range(S):
        ok, sid, p, r, a = leg[s]
        pos[s] = p
        ids.append(sid); refs.append(r); alts.append(a)
        j = pu.fetch(p) if ok else None
        if ok and j is not None and (posset is None or p in posset):
            keep[s] = 1
        if j is not None and ok:
            b = pu.bases[j]
            nr[s] = min(b.count(r), 255)  # src/pileup.c:442-450
            na[s] = min(b.count(a), 255)
            cov[s] = pu.cov[j]
            chroms.append(pu.chrom[j])
        else:
            chroms.append("")
    return Packed(names, hap, pos, keep, nr, na, cov, chroms, ids, refs, alts, pu)


def input_cov_dist(pu: Pileup, max_cov: int):
    """src/ibdgem.c:83-96."""
    dist = [0] * (max_cov + 1)
    total = 0
    for c in pu.cov:
        if c <= max_cov:
            dist[c] += 1
            total += c
    mean = total / len(pu.cov) if pu.cov else float("nan")
    return dist, mean


def c_e(x: float) -> str:
    """C's %e for a double (Python's %e is the same correctly-rounded conversion)."""
    if math.isnan(x):
        return "-nan" if math.copysign(1.0, x) < 0 else "nan"
    return "%e" % x


def format_tab(pk: Packed, res: dict, target: int, max_cov: int, cull_p: float = 1.0) -> str:
    """Body of <pileup>.<target>.tab.txt from line 3 on, src/ibdgem.c:535-547, 731-733, 761-768."""
    dist, mean = input_cov_dist(pk.pileup, max_cov)
    o = ["# INPUT COVERAGE DISTRIBUTION:", "# COVERAGE N_SITES"]
    o += ["# %d %d" % (c, n) for c, n in enumerate(dist)]
    o += ["# MEAN DEPTH = %f" % mean, "# CULL DEPTH RATIO = %f" % cull_p]
    o.append("# CHR\trsID\tPOS\tREF\tALT\tAF\tDP\tSQ_NREF\tSQ_NALT\tGT_A0\tGT_A1\tLIBD0\tLIBD1\tLIBD2")
    st = res["status"]
    for s in np.nonzero(st)[0]:
        o.append("%s\t%s\t%d\t%s\t%s\t%f\t%d\t%d\t%d\t%d\t%d\t%s\t%s\t%s" % (
            pk.chrom[s], pk.ids[s], int(pk.pos[s]), pk.ref[s], pk.alt[s], res["f"][s], int(pk.cov[s]),
            int(res["n_ref"][s]), int(res["n_alt"][s]), int(pk.hap[s, 2 * target]),
            int(pk.hap[s, 2 * target + 1]), c_e(res["ibd0"][s]), c_e(res["ibd1"][s]), c_e(res["ibd2"][s])))
    o += ["# FINAL COVERAGE DISTRIBUTION:", "# COVERAGE N_SITES"]
    o += ["# %d %d" % (c, int(n)) for c, n in enumerate(res["final_dist"])]
    fm = res["final_total_cov"] / res["processed"] if res["processed"] else float("nan")
    o.append("# FINAL MEAN DEPTH = %s" % ("%f" % fm if not math.isnan(fm) else "-nan"))
    o.append("## Number of sites processed: %d" % res["processed"])
    o.append("## Number of sites skipped: %d" % res["skipped"])
    return "\n".join(o) + "\n"


def format_summary(res: dict, values=None) -> str:
    """<pileup>.<target>.summary.txt, src/ibdgem.c:548, 751-756."""
    o = ["# SEGMENT\tSTART\tEND\tLIBD0\tLIBD1\tLIBD2\tNUM_SITES"]
    v = res["w_lin"] if values is None else values
    for w in range(res["n_windows"]):
        o.append("%d\t%d\t%d\t%s\t%s\t%s\t%d" % (w + 1, int(res["w_start"][w]), int(res["w_end"][w]),
                                                 c_e(v[w][0]), c_e(v[w][1]), c_e(v[w][2]),
                                                 int(res["w_nsites"][w])))
    return "\n".join(o) + "\n"


def read_summary(path: str):
    """Rows of a summary file the way hiddengem reads them, src/hiddengem.c:59-80."""
    rows = []
    with open(path) as fh:
        for ln in fh:
            if ln.startswith("#"):
                continue
            f = ln.split()
            if len(f) >= 7:
                rows.append((int(f[1]), int(f[2]), float(f[3]), float(f[4]), float(f[5]), int(f[6])))
    return rows


def write_impute(dirpath, prefix, hap, pos, names, ref=None, alt=None, ids=None):
    """Write a synthetic IMPUTE triple (for driving oracle/_ref in golden generation)."""
    os.makedirs(dirpath, exist_ok=True)
    S, H = hap.shape
    with open(os.path.join(dirpath, prefix + ".hap"), "w") as fh:
        for s in range(S):
            fh.write(" ".join("%d" % v for v in hap[s]) + "\n")
    with open(os.path.join(dirpath, prefix + ".legend"), "w") as fh:
        fh.write("ID pos allele0 allele1\n")
        for s in range(S):
            fh.write("%s %d %s %s\n" % (ids[s] if ids else "rs%d" % s, int(pos[s]),
                                        ref[s] if ref else "A", alt[s] if alt else "G"))
    with open(os.path.join(dirpath, prefix + ".indv"), "w") as fh:
        for n in names:
            fh.write(n + "\n")


def write_pileup(path, chrom, pos, n_ref, n_alt, ref="A", alt="G", extra=None):
    """Write a 7-column pileup whose REF/ALT-matching base counts are n_ref/n_alt; `extra[s]`
    other bases ('T') are appended to exercise DP != n_ref + n_alt."""
    with open(path, "w") as fh:
        for s in range(len(pos)):
            r = ref[s] if not isinstance(ref, str) else ref
            a = alt[s] if not isinstance(alt, str) else alt
            other = "T" if "T" not in (r, a) else "C" if "C" not in (r, a) else "A"
            b = r * int(n_ref[s]) + a * int(n_alt[s]) + other * (int(extra[s]) if extra is not None else 0)
            cov = len(b)
            if cov == 0:
                fh.write("%s\t%d\tN\t0\t*\t*\t*\n" % (chrom, int(pos[s])))
            else:
                fh.write("%s\t%d\tN\t%d\t%s\t%s\t%s\n" % (chrom, int(pos[s]), cov, b, "I" * cov, "]" * cov))

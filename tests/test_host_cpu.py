"""CPU tests of the native host side (ibdgem_b200/csrc/host): the packer against the Python
restatement of the reference's parsers (tests/refio.py) on the shipped fixtures and on the
reference-run inputs, and the command-line surface that needs no GPU (help, validation messages,
exit codes — src/ibdgem.c:779-827, 966-1041)."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

import hostlib
import ibdgem_b200 as ib
import refio


@pytest.fixture(scope="module", autouse=True)
def _built():
    hostlib.build()


def _check_against_refio(got, pk, af_user=None):
    assert got["S"] == len(pk.pos) and got["N"] == len(pk.names) and got["names"] == pk.names
    np.testing.assert_array_equal(got["keep"], pk.host_keep)
    m = pk.host_keep == 1
    np.testing.assert_array_equal(got["pos"][m], pk.pos[m])
    np.testing.assert_array_equal(got["n_ref"][m], pk.n_ref[m])
    np.testing.assert_array_equal(got["n_alt"][m], pk.n_alt[m])
    np.testing.assert_array_equal(got["dp"][m], pk.cov[m])
    np.testing.assert_array_equal(got["bits"], ib.pack_bits(pk.hap))
    np.testing.assert_array_equal(got["pileup_cov"], np.asarray(pk.pileup.cov, np.uint32))
    if af_user is not None:
        a, b = got["af_user"][m], af_user[m]
        np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
        np.testing.assert_array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


@pytest.mark.parametrize("k", [1, 2, 3])
def test_pack_shipped_fixture(fixture_dir, k):
    inp = os.path.join(fixture_dir, "input")
    paths = [os.path.join(inp, f) for f in ("test.hap", "test.legend", "test.indv", f"test{k}.pileup")]
    got = hostlib.pack(0, *paths)
    _check_against_refio(got, refio.pack_impute(*paths))


def test_pack_reference_run_inputs_with_positions_af_chromosome(golden_dir):
    import refcases
    ca = os.path.join(golden_dir, "ref_runs", "caseA")
    case = refcases.load_case(ca, "af_pos")
    got = hostlib.pack(0, os.path.join(ca, "panel.hap"), os.path.join(ca, "panel.legend"), os.path.join(ca, "panel.indv"),
                       os.path.join(ca, "unk.pileup"), chrom="7", positions=os.path.join(ca, "pos.txt"),
                       af=os.path.join(ca, "af.txt"))
    _check_against_refio(got, case.pk, case.af_user)
    # a chromosome that is not in the pileup leaves nothing to parse: the reference's fatal path
    assert hostlib.pack(0, os.path.join(ca, "panel.hap"), os.path.join(ca, "panel.legend"), os.path.join(ca, "panel.indv"),
                        os.path.join(ca, "unk.pileup"), chrom="8") is None


def test_pack_gzip_inputs(golden_dir, tmp_path):
    ca = os.path.join(golden_dir, "ref_runs", "caseA")
    for f in ("panel.hap", "panel.legend", "unk.pileup"):
        with open(os.path.join(ca, f), "rb") as src, gzip.open(tmp_path / (f + ".gz"), "wb") as dst:
            shutil.copyfileobj(src, dst)
    plain = hostlib.pack(0, os.path.join(ca, "panel.hap"), os.path.join(ca, "panel.legend"), os.path.join(ca, "panel.indv"),
                         os.path.join(ca, "unk.pileup"))
    gz = hostlib.pack(0, str(tmp_path / "panel.hap.gz"), str(tmp_path / "panel.legend.gz"), os.path.join(ca, "panel.indv"),
                      str(tmp_path / "unk.pileup.gz"))
    for key in ("pos", "keep", "n_ref", "n_alt", "dp", "bits"):
        np.testing.assert_array_equal(plain[key], gz[key])


def test_pack_vcf_matches_impute_up_to_vcf_only_filters(golden_dir):
    """tests/golden/make_golden.py writes panel.vcf from the same haplotypes: every 37th (+11) record is
    multi-allelic, every 43rd (+5) has an unparsable genotype, QUAL is '.' or 10 + 13 s mod 60."""
    ca = os.path.join(golden_dir, "ref_runs", "caseA")
    imp = hostlib.pack(0, os.path.join(ca, "panel.hap"), os.path.join(ca, "panel.legend"), os.path.join(ca, "panel.indv"),
                       os.path.join(ca, "unk.pileup"))
    S = imp["S"]
    s = np.arange(S)
    multi, badgt = s % 37 == 11, s % 43 == 5
    qual = np.where(s % 5 == 0, 0.0, 10 + (s * 13) % 60)
    for q in (0.0, 30.0):
        vcf = hostlib.pack(1, os.path.join(ca, "panel.vcf"), None, None, os.path.join(ca, "unk.pileup"), min_qual=q)
        assert vcf["S"] == S and vcf["names"] == imp["names"]
        want = imp["keep"].astype(bool) & ~multi & ~badgt & (qual >= q)
        np.testing.assert_array_equal(vcf["keep"].astype(bool), want)
        np.testing.assert_array_equal(vcf["bits"][~badgt & ~multi], imp["bits"][~badgt & ~multi])
        for key in ("pos", "n_ref", "n_alt", "dp"):
            np.testing.assert_array_equal(vcf[key][want], imp[key][want])


def _run(args, cwd=None):
    return subprocess.run([os.path.join(hostlib.BIN, "ibdgem")] + args, capture_output=True, text=True, cwd=cwd)


def test_cli_help_and_validation_exit_codes(fixture_dir):
    r = _run([])
    assert r.returncode == 0 and r.stderr.startswith("IBDGem-2.0: Compares low-coverage sequencing data")
    assert "-w, --window-size  INT          Number of sites per genomic segment" in r.stderr
    r = _run(["-w", "1"])
    assert r.returncode == 0 and r.stderr == "[::] ERROR: Invalid window size (-w) of 1 (must be >= 2).\n"
    r = _run(["-M", "0"])
    assert r.returncode == 0 and "Invalid maximum estimated coverage (-M) of 0 (must be >= 1)" in r.stderr
    r = _run(["-D", "-1"])
    assert r.returncode == 0 and "Invalid down-sample coverage (-D) of -1.00 (must be > 0)" in r.stderr
    r = _run(["-F", "1.5"])
    assert r.returncode == 0 and "(-F) of 1.50 (must be <= 1)" in r.stderr
    r = _run(["-P"])
    assert r.returncode == 0 and r.stderr == "Option -P missing required argument.\n"
    r = _run(["-P", "/nonexistent.pileup"])
    assert r.returncode == 1 and "[::] ERROR parsing Pileup data; make sure input is valid." in r.stderr
    inp = os.path.join(fixture_dir, "input")
    r = _run(["-P", os.path.join(inp, "test1.pileup")])
    assert r.returncode == 1 and "[::] ERROR: Missing genotype files." in r.stderr
    r = _run(["-P", os.path.join(inp, "test1.pileup"), "-V", "x.vcf", "-H", "y.hap"])
    assert r.returncode == 1 and "2 types of genotype inputs detected" in r.stderr


def test_cli_fails_loudly_without_a_gpu(fixture_dir, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    inp = os.path.join(fixture_dir, "input")
    r = _run(["-H", os.path.join(inp, "test.hap"), "-L", os.path.join(inp, "test.legend"), "-I", os.path.join(inp, "test.indv"),
              "-P", os.path.join(inp, "test1.pileup"), "-N", "sample1", "-O", str(tmp_path)])
    assert r.returncode == 1 and "no usable CUDA device" in r.stderr and "no CPU fallback" in r.stderr
    assert not list(tmp_path.iterdir())  # nothing is written by a host-side stand-in


def test_panel_cache_round_trip_and_invalidation(fixture_dir, tmp_path):
    """--panel-cache: the first run parses the text and writes the cache, later runs (here against
    another pileup) load it and pack the same arrays as the uncached packer; a changed input file or
    a damaged cache is noticed and the cache rebuilt."""
    import shutil
    inp = os.path.join(fixture_dir, "input")
    work = tmp_path / "panel"
    work.mkdir()
    for f in ("test.hap", "test.legend", "test.indv"):
        shutil.copy(os.path.join(inp, f), work / f)
    hap, leg, indv = (str(work / f) for f in ("test.hap", "test.legend", "test.indv"))
    cache = str(tmp_path / "panel.cache")

    def same(a, b):
        assert a["names"] == b["names"] and (a["S"], a["N"], a["Wh"]) == (b["S"], b["N"], b["Wh"])
        for key in ("pos", "n_ref", "n_alt", "keep", "dp", "bits"):
            np.testing.assert_array_equal(a[key], b[key])

    pu1, pu2 = os.path.join(inp, "test1.pileup"), os.path.join(inp, "test2.pileup")
    got, hit = hostlib.pack_cached(hap, leg, indv, cache, pu1)
    assert hit == 0 and os.path.exists(cache)
    same(got, hostlib.pack(0, hap, leg, indv, pu1))
    got, hit = hostlib.pack_cached(hap, leg, indv, cache, pu2)
    assert hit == 1
    same(got, hostlib.pack(0, hap, leg, indv, pu2))
    _check_against_refio(got, refio.pack_impute(hap, leg, indv, pu2))
    # the .hap changes (one allele flipped, same size, new mtime): the cache no longer matches
    text = open(hap).read()
    flipped = ("1" if text[0] == "0" else "0") + text[1:]
    open(hap, "w").write(flipped)
    st = os.stat(hap)
    os.utime(hap, ns=(st.st_atime_ns, st.st_mtime_ns + 1_000_000_000))
    got, hit = hostlib.pack_cached(hap, leg, indv, cache, pu1)
    assert hit == 0
    same(got, hostlib.pack(0, hap, leg, indv, pu1))
    # a truncated cache is not trusted
    data = open(cache, "rb").read()
    open(cache, "wb").write(data[: len(data) // 2])
    got, hit = hostlib.pack_cached(hap, leg, indv, cache, pu1)
    assert hit == 0
    same(got, hostlib.pack(0, hap, leg, indv, pu1))
    got, hit = hostlib.pack_cached(hap, leg, indv, cache, pu1)
    assert hit == 1


_MT_SCRIPT = r"""
import gzip, os, shutil, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import hostlib
inp, work = sys.argv[2], sys.argv[3]
rng = np.random.default_rng(5)
N, S = 37, 900                      # 74 haplotypes: AVX2 groups, an 8-byte group and scalar leftovers
hap = (rng.random((S, 2 * N)) < 0.3).astype(np.uint8)
def write(tag, rows, legend_rows, trailing_newline=True, poison=None):
    lines = [" ".join(map(str, r)) for r in hap[:rows]]
    if poison is not None:
        lines[poison] = lines[poison][:10] + "2" + lines[poison][11:]
    text = "\n".join(lines) + ("\n" if trailing_newline else "")
    p = os.path.join(work, tag)
    open(p + ".hap", "w").write(text)
    with gzip.open(p + ".seq.hap.gz", "wt") as fh:
        fh.write(text)
    with open(p + ".legend", "w") as fh:
        fh.write("id position a0 a1\n")
        for s in range(legend_rows):
            fh.write(f"rs{s} {100 + 7 * s} A C\n" if s % 11 else f"rs{s} {100 + 7 * s} AT C\n")
    with open(p + ".indv", "w") as fh:
        fh.write("".join(f"i{i}\n" for i in range(N)))
    with open(p + ".pileup", "w") as fh:
        for s in range(0, legend_rows, 3):
            fh.write(f"1\t{100 + 7 * s}\tA\t2\t.c\tII\t]]\n")
    return p
for tag, kw in (("plain", dict(rows=S, legend_rows=S)), ("nonl", dict(rows=S, legend_rows=S, trailing_newline=False)),
                ("shortleg", dict(rows=S, legend_rows=S - 13)), ("shorthap", dict(rows=S - 29, legend_rows=S)),
                ("bad", dict(rows=S, legend_rows=S, poison=611))):
    p = write(tag, **kw)
    mt = hostlib.pack(0, p + ".hap", p + ".legend", p + ".indv", p + ".pileup")           # mapped, threaded
    seq = hostlib.pack(0, p + ".seq.hap.gz", p + ".legend", p + ".indv", p + ".pileup")   # gz: sequential reader
    if tag == "bad":
        assert mt is None and seq is None
        continue
    assert mt["S"] == seq["S"] == min(kw["rows"], kw["legend_rows"]), (tag, mt["S"], seq["S"])
    for key in ("pos", "n_ref", "n_alt", "keep", "dp", "bits"):
        assert np.array_equal(mt[key], seq[key]), (tag, key)
    assert np.array_equal(mt["bits"][:, :3].view(np.uint8)[:, : (2 * N + 7) // 8],
                          np.packbits(hap[: mt["S"]], axis=1, bitorder="little"))
# VCF: records of every kind the parser distinguishes, no final newline
NV, SV = 19, 300
hv = (rng.random((SV, 2 * NV)) < 0.3).astype(int)
lines = ["##a", "##b", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{i}" for i in range(NV))]
for s in range(SV):
    p, ref, alt, q, idv = 100 + 7 * s, "A", "C", "50", f"rs{s}"
    m = s % 13
    if m == 1: alt = "C,G"
    if m == 2: ref = "AT"
    if m == 3: q = "5"
    if m == 4: idv = ""
    if m == 5: q = "."
    cols = [f"{hv[s, 2 * i]}|{hv[s, 2 * i + 1]}" + (":1" if (s * i) % 17 == 3 else "") for i in range(NV)]
    if m == 6: cols[7] = ".|."
    if m == 7: cols = cols[:5]
    line = f"1\t{p}\t{idv}\t{ref}\t{alt}\t{q}\tPASS\t.\tGT\t" + "\t".join(cols)
    if m == 8: line = f"1\tx{p}\trs\tA\tC\t50\tPASS\t.\tGT\t" + "\t".join(cols)
    if m == 9: line = f"1\t{p}\trs\tA\tC"
    lines.append(line)
text = "\n".join(lines)
open(os.path.join(work, "v.vcf"), "w").write(text)
with gzip.open(os.path.join(work, "v.seq.vcf.gz"), "wt") as fh:
    fh.write(text)
open(os.path.join(work, "v.pileup"), "w").write("".join(f"1\t{100 + 7 * s}\tA\t2\t.c\tII\t]]\n" for s in range(SV)))
for q in (0.0, 20.0):
    mt = hostlib.pack(1, os.path.join(work, "v.vcf"), None, None, os.path.join(work, "v.pileup"), min_qual=q)
    seq = hostlib.pack(1, os.path.join(work, "v.seq.vcf.gz"), None, None, os.path.join(work, "v.pileup"), min_qual=q)
    assert mt["S"] == seq["S"] == SV and mt["names"] == seq["names"] and mt["labels"] == seq["labels"]
    for key in ("pos", "n_ref", "n_alt", "keep", "dp", "bits"):
        assert np.array_equal(mt[key], seq[key]), ("vcf", key)
    assert 50 < mt["keep"].sum() < SV
print("ok")
"""


def test_threaded_hap_parse_matches_sequential(fixture_dir, tmp_path):
    """Large plain .hap and VCF files are mapped and packed by several threads; IBDGEM_PACK_MT_MIN_BYTES=1
    sends small files down that path.  Same arrays as the sequential reader (the .gz route) for a missing
    final newline, a legend shorter / longer than the .hap, the same refusal of a bad allele, and every
    kind of VCF record the parser distinguishes."""
    import subprocess
    import sys
    env = dict(os.environ, IBDGEM_PACK_MT_MIN_BYTES="1")
    r = subprocess.run([sys.executable, "-c", _MT_SCRIPT, hostlib.ROOT, os.path.join(fixture_dir, "input"), str(tmp_path)],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
    assert "allele '2' of haplotype 5" in r.stderr and ".hap line 612" in r.stderr


def test_pileup_fast_splitter_matches_sscanf_route(fixture_dir, tmp_path):
    """Well-formed mpileup lines skip the two sscanf calls of line2pul; every other shape still goes
    through them.  Same store from both routes on a file that mixes ordinary lines with blanks instead
    of tabs, doubled tabs, signs, long numbers, coverage >= 128, indels, '^' markers, bad characters,
    wrong counts, short lines and CRLF endings."""
    rng = np.random.default_rng(11)
    inp = os.path.join(fixture_dir, "input")
    legend = open(os.path.join(inp, "test.legend")).read().splitlines()[1:]
    positions = [int(l.split()[1]) for l in legend]
    odd = ["1 {p} A 2 .. II ]]", "1\t\t{p}\tA\t2\t..\tII\t]]", "1\t+{p}\tA\t2\t.,\tII\t]]", "1\t{p}\tA\t130\t..\tII\t]]",
           "1\t{p}\tA\t3\t.+2AC,^]g\tIII\t]]]", "1\t{p}\tA\t2\t.$-1c*\tII\t]]", "1\t{p}\tA\t2\t.x\tII\t]]", "1\t{p}\tA\t3\t..\tII\t]]",
           "1\t{p}\tA\t2\t..\tIII\t]]]", "1\t{p}\tA\t2\t..", "1\t{p}\tAT\t2\t..\tII\t]]", "1\t0000000000{p}\tA\t1\t.\tI\t]",
           "1\t{p}\tA\t2\t..\tII\t]]\r", "1\t{p}\tA\t2\t..\tII\t]] \t", "  1\t{p}\tA\t2\t,.\tII\t]]", "1\t{p}\tA\t0\t*\t*\t*",
           "garbage", ""]
    lines = []
    for k, p in enumerate(positions):
        if k % 3 == 0:
            lines.append(odd[(k // 3) % len(odd)].format(p=p))
        else:
            n = int(rng.integers(1, 6))
            reads = "".join(rng.choice(list(".,ACGTacgt*"), n))
            lines.append(f"1\t{p}\tA\t{n}\t{reads}\t{'I' * n}\t{']' * n}")
    pu = tmp_path / "odd.pileup"
    pu.write_text("\n".join(lines) + "\n")
    paths = [os.path.join(inp, f) for f in ("test.hap", "test.legend", "test.indv")]
    fast = hostlib.pack(0, *paths, str(pu))
    os.environ["IBDGEM_PILEUP_NO_FAST"] = "1"
    try:
        slow = hostlib.pack(0, *paths, str(pu))
    finally:
        del os.environ["IBDGEM_PILEUP_NO_FAST"]
    assert fast is not None and slow is not None
    assert len(fast["pileup_cov"]) == len(slow["pileup_cov"]) > len(positions) // 2
    for key in ("pileup_cov", "pos", "n_ref", "n_alt", "keep", "dp"):
        np.testing.assert_array_equal(fast[key], slow[key])
    assert fast["keep"].sum() > 50


def test_vcf_genotype_groups_and_column_by_column_agree(tmp_path):
    """pack_vcf takes eight plain "a|b" columns per AVX2 step and falls back to one column at a time for
    anything else; the bits must be the genotypes whatever mix a line has: '/' separators, columns with
    extra FORMAT fields, a line with a missing genotype (skipped, src/ibd-parse.c:150-173)."""
    rng = np.random.default_rng(17)
    N, S = 37, 400
    hap = (rng.random((S, 2 * N)) < 0.35).astype(np.uint8)
    lines = ["##fileformat=VCFv4.2",
             "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"i{i}" for i in range(N))]
    bad = {123}
    for s in range(S):
        cols = []
        for i in range(N):
            sep = "|" if (s + i) % 5 else "/"
            g = f"{hap[s, 2 * i]}{sep}{hap[s, 2 * i + 1]}"
            if s % 7 == 3 and rng.random() < 0.2:
                g += ":35:0.9"
            if s in bad and i == 20:
                g = "./."
            cols.append(g)
        lines.append(f"1\t{100 + 7 * s}\trs{s}\tA\tC\t50\tPASS\t.\tGT\t" + "\t".join(cols))
    vcf = tmp_path / "p.vcf"
    vcf.write_text("\n".join(lines) + "\n")
    pu = tmp_path / "p.pileup"
    pu.write_text("".join(f"1\t{100 + 7 * s}\tA\t2\t.c\tII\t]]\n" for s in range(S)))
    got = hostlib.pack(1, str(vcf), None, None, str(pu))
    assert got is not None and got["S"] == S and got["N"] == N
    want = np.packbits(hap, axis=1, bitorder="little")
    want[list(bad)] = 0
    np.testing.assert_array_equal(got["bits"].view(np.uint8)[:, : want.shape[1]], want)
    keep = np.ones(S, np.uint8)
    keep[list(bad)] = 0
    np.testing.assert_array_equal(got["keep"], keep)


_LEGEND_SCRIPT = r"""
import hashlib, os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import hostlib
w = sys.argv[2]
g = hostlib.pack(0, w + "/o.hap", w + "/o.legend", w + "/o.indv", w + "/o.pileup")
h = hashlib.sha256()
for key in ("pos", "keep", "n_ref", "n_alt", "dp", "bits"):
    h.update(g[key].tobytes())
print(g["S"], int(g["keep"].sum()), h.hexdigest())
"""


def test_legend_fast_splitter_matches_sscanf_route(tmp_path):
    """Legend lines of the plain shape skip sscanf; the others (signs, letters after the digits, tokens
    over 128 characters, too few columns, tabs, leading blanks, CRLF, extra columns) go through it.
    Same packed arrays from both routes."""
    import subprocess
    import sys
    odd = ["rs{s} +{p} A C", "rs{s} {p}x A C", "rs{s} {p} A", "rs{s}\t{p}\tA\tC", "   rs{s}   {p}  A  C  extra  columns",
           "rs{s} {p} A C\r", "L" * 130 + " {p} A C", "rs{s} {p} AC G", "rs{s} 00000000000000000000{p} A C", "", "rs{s} {p} A " + "G" * 129]
    S = 330
    leg = ["id position a0 a1"]
    for s in range(S):
        p = 100 + 7 * s
        leg.append(odd[(s // 3) % len(odd)].format(s=s, p=p) if s % 3 == 0 else f"rs{s} {p} A C")
    (tmp_path / "o.legend").write_text("\n".join(leg) + "\n")
    (tmp_path / "o.hap").write_text("0 1 1 0 0 0\n" * S)
    (tmp_path / "o.indv").write_text("a\nb\nc\n")
    (tmp_path / "o.pileup").write_text("".join(f"1\t{100 + 7 * s}\tA\t2\t.c\tII\t]]\n" for s in range(S)))
    outs = []
    for env in ({}, {"IBDGEM_LEGEND_NO_FAST": "1"}):
        r = subprocess.run([sys.executable, "-c", _LEGEND_SCRIPT, hostlib.ROOT, str(tmp_path)], capture_output=True, text=True,
                           env=dict(os.environ, **env), timeout=120)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout.strip())
    assert outs[0] == outs[1]
    n_sites, n_kept = int(outs[0].split()[0]), int(outs[0].split()[1])
    assert n_sites == S and 200 < n_kept < S


def test_vcf_panel_cache_applies_quality_filter_at_join_time(golden_dir, tmp_path):
    """--panel-cache with VCF input: the cache holds the parsed records including QUAL, so a later run
    with another -q (or another pileup) loads it and still filters like the uncached packer."""
    import shutil
    ca = os.path.join(golden_dir, "ref_runs", "caseA")
    vcf = str(tmp_path / "panel.vcf")
    shutil.copy(os.path.join(ca, "panel.vcf"), vcf)
    pu = os.path.join(ca, "unk.pileup")
    cache = str(tmp_path / "vcf.cache")

    def same(a, b):
        assert a["names"] == b["names"] and a["labels"] == b["labels"]
        for key in ("pos", "n_ref", "n_alt", "keep", "dp", "bits"):
            np.testing.assert_array_equal(a[key], b[key])

    got, hit = hostlib.pack_vcf_cached(vcf, cache, pu)
    assert hit == 0
    same(got, hostlib.pack(1, vcf, None, None, pu))
    got30, hit = hostlib.pack_vcf_cached(vcf, cache, pu, min_qual=30.0)
    assert hit == 1
    same(got30, hostlib.pack(1, vcf, None, None, pu, min_qual=30.0))
    assert got30["keep"].sum() < got["keep"].sum()
    # an IMPUTE cache is not mistaken for a VCF cache of other files
    hap, leg, indv = (os.path.join(ca, f) for f in ("panel.hap", "panel.legend", "panel.indv"))
    _, hit = hostlib.pack_cached(hap, leg, indv, cache, pu)
    assert hit == 0
    _, hit = hostlib.pack_vcf_cached(vcf, cache, pu)
    assert hit == 0

"""ctypes wrapper over oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference's scoring path (oracle/ibdgem_oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")


def build_oracle() -> str:
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "ibdgem_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)
    return so


class _Params(C.Structure):
    _fields_ = [
        ("epsilon", C.c_double),
        ("max_cov", C.c_uint32),
        ("window", C.c_int32),
        ("min_af", C.c_double),
        ("max_af", C.c_double),
        ("ld_mode", C.c_int32),
        ("opt_v", C.c_int32),
        ("cull_p", C.c_double),
        ("pu_idx", C.c_int32),
    ]


class _Result(C.Structure):
    _fields_ = [
        ("status", C.c_void_p),
        ("f", C.c_void_p),
        ("n_ref", C.c_void_p),
        ("n_alt", C.c_void_p),
        ("ibd0", C.c_void_p),
        ("ibd1", C.c_void_p),
        ("ibd2", C.c_void_p),
        ("n_windows", C.c_int32),
        ("w_start", C.c_void_p),
        ("w_end", C.c_void_p),
        ("w_nsites", C.c_void_p),
        ("w_lin", C.c_void_p),
        ("w_log", C.c_void_p),
        ("processed", C.c_uint64),
        ("skipped", C.c_uint64),
        ("final_total_cov", C.c_uint64),
        ("final_dist", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        _lib.orc_find_pDgG.restype = C.c_double
        _lib.orc_find_pDgG.argtypes = [C.c_void_p, C.c_double, C.c_ushort, C.c_ushort, C.c_uint, C.c_uint]
        _lib.orc_find_pDgf.restype = C.c_double
        _lib.orc_find_pDgf.argtypes = [C.c_double] * 4
        _lib.orc_find_pDgIBD1.restype = C.c_double
        _lib.orc_find_pDgIBD1.argtypes = [C.c_ushort, C.c_ushort] + [C.c_double] * 4
        _lib.orc_init_nCk.restype = C.c_void_p
        _lib.orc_init_nCk.argtypes = [C.c_uint]
        _lib.orc_retrieve_nCk.restype = C.c_ulong
        _lib.orc_retrieve_nCk.argtypes = [C.c_void_p, C.c_uint, C.c_uint]
        _lib.orc_destroy_nCk.argtypes = [C.c_void_p, C.c_uint]
        _lib.orc_compare_target.restype = C.c_int
        _lib.orc_hiddengem.restype = C.c_int
        _lib.orc_ld_loop_bench.restype = C.c_uint64
    return _lib


@dataclass
class Params:
    epsilon: float = 0.02
    max_cov: int = 20
    window: int = 100
    min_af: float = 0.0
    max_af: float = 1.0
    ld_mode: int = 0
    opt_v: int = 0
    cull_p: float = 1.0
    pu_idx: int = -1

    def c(self) -> _Params:
        return _Params(self.epsilon, self.max_cov, self.window, self.min_af, self.max_af,
                       self.ld_mode, self.opt_v, self.cull_p, self.pu_idx)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def pDgG_table(epsilon: float, max_cov: int) -> np.ndarray:
    """P(D|G) for every (n_ref, n_alt) with both <= max_cov: [max_cov+1, max_cov+1, 3]."""
    L = lib()
    n = max_cov
    # rows with n_ref + n_alt > max_cov are never used by the reference; the nCk table only
    # goes to max_cov, so leave them NaN.
    nck = L.orc_init_nCk(n)
    tab = np.full((n + 1, n + 1, 3), np.nan)
    for r in range(n + 1):
        for a in range(n + 1 - r):
            tab[r, a, 0] = L.orc_find_pDgG(nck, epsilon, 0, 0, r, a)
            tab[r, a, 1] = L.orc_find_pDgG(nck, epsilon, 0, 1, r, a)
            tab[r, a, 2] = L.orc_find_pDgG(nck, epsilon, 1, 1, r, a)
    L.orc_destroy_nCk(nck, n)
    return tab


def compare_target(params: Params, pos, host_keep, n_ref, n_alt, hap, target, bg,
                   af_user=None, reseed=True, max_windows=None):
    """Run the oracle for one target.  hap: [S, 2N] uint8 of 0/1."""
    L = lib()
    S, H = hap.shape
    N = H // 2
    pos = np.ascontiguousarray(pos, dtype=np.uint64)
    host_keep = np.ascontiguousarray(host_keep, dtype=np.uint8)
    n_ref = np.ascontiguousarray(n_ref, dtype=np.uint8)
    n_alt = np.ascontiguousarray(n_alt, dtype=np.uint8)
    hap = np.ascontiguousarray(hap, dtype=np.uint8)
    bg = np.ascontiguousarray(bg, dtype=np.int32)
    if af_user is not None:
        af_user = np.ascontiguousarray(af_user, dtype=np.float64)
    if max_windows is None:
        max_windows = S // max(params.window, 1) + 2
    out = {
        "status": np.zeros(S, np.uint8), "f": np.zeros(S), "n_ref": np.zeros(S, np.uint8),
        "n_alt": np.zeros(S, np.uint8), "ibd0": np.zeros(S), "ibd1": np.zeros(S), "ibd2": np.zeros(S),
        "w_start": np.zeros(max_windows, np.uint64), "w_end": np.zeros(max_windows, np.uint64),
        "w_nsites": np.zeros(max_windows, np.int32), "w_lin": np.zeros((max_windows, 3)),
        "w_log": np.zeros((max_windows, 3)), "final_dist": np.zeros(params.max_cov + 1, np.uint64),
    }
    r = _Result()
    for k in ("status", "f", "n_ref", "n_alt", "ibd0", "ibd1", "ibd2", "w_start", "w_end",
              "w_nsites", "w_lin", "w_log", "final_dist"):
        setattr(r, k, out[k].ctypes.data)
    cp = params.c()
    rc = L.orc_compare_target(C.byref(cp), C.c_int64(S), C.c_int32(N), _p(pos), _p(host_keep),
                              _p(af_user) if af_user is not None else None, _p(n_ref), _p(n_alt),
                              _p(hap), C.c_int32(int(target)), _p(bg), C.c_int32(len(bg)),
                              C.c_int32(max_windows), C.c_int(1 if reseed else 0), C.byref(r))
    if rc != 0:
        raise RuntimeError(f"orc_compare_target failed rc={rc}")
    nw = r.n_windows
    for k in ("w_start", "w_end", "w_nsites", "w_lin", "w_log"):
        out[k] = out[k][:nw]
    out.update(n_windows=nw, processed=int(r.processed), skipped=int(r.skipped),
               final_total_cov=int(r.final_total_cov))
    return out


def hiddengem(l, p01=1e-3, p02=1e-6, p12=1e-3):
    """l: [n_bins, 3] likelihoods as parsed from a summary file."""
    L = lib()
    l = np.ascontiguousarray(l, dtype=np.float64)
    n = l.shape[0]
    state = np.zeros(n, np.int32)
    score = np.zeros((n, 3))
    score_ld = np.zeros((n, 3), dtype=np.longdouble)
    rc = L.orc_hiddengem(_p(l), C.c_int32(n), C.c_double(p01), C.c_double(p02), C.c_double(p12),
                         _p(state), _p(score), _p(score_ld))
    if rc != 0:
        raise RuntimeError("orc_hiddengem failed")
    return state, score, score_ld


def ld_loop_bench(params: Params, n_ref, n_alt, hap, targets, bg):
    L = lib()
    S, H = hap.shape
    n_ref = np.ascontiguousarray(n_ref, dtype=np.uint8)
    n_alt = np.ascontiguousarray(n_alt, dtype=np.uint8)
    hap = np.ascontiguousarray(hap, dtype=np.uint8)
    targets = np.ascontiguousarray(targets, dtype=np.int32)
    bg = np.ascontiguousarray(bg, dtype=np.int32)
    sink = C.c_double(0)
    cp = params.c()
    return int(L.orc_ld_loop_bench(C.byref(cp), C.c_int64(S), C.c_int32(H // 2), _p(n_ref), _p(n_alt),
                                   _p(hap), _p(targets), C.c_int32(len(targets)), _p(bg),
                                   C.c_int32(len(bg)), C.byref(sink)))

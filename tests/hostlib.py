"""ctypes view of ibdgem_b200/libibdgem_host.so (the native packer, no CUDA) for the CPU tests."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_DIR = os.path.join(ROOT, "ibdgem_b200", "csrc", "host")
LIB = os.path.join(ROOT, "ibdgem_b200", "libibdgem_host.so")
BIN = os.path.join(ROOT, "ibdgem_b200", "bin")


def build():
    r = subprocess.run(["make", "-C", HOST_DIR], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stdout[-2000:] + r.stderr[-2000:])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.ibdhost_pack.restype = C.c_void_p
        _lib.ibdhost_pack.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p,
                                      C.c_char_p, C.c_double]
        for name, rt in (("n_sites", C.c_int64), ("n_indiv", C.c_int32), ("words", C.c_int64), ("n_pileup", C.c_int64),
                         ("pos", C.c_void_p), ("n_ref", C.c_void_p), ("n_alt", C.c_void_p), ("keep", C.c_void_p),
                         ("dp", C.c_void_p), ("bits", C.c_void_p), ("af_user", C.c_void_p), ("pileup_cov", C.c_void_p)):
            f = getattr(_lib, "ibdhost_" + name)
            f.restype = rt
            f.argtypes = [C.c_void_p]
        _lib.ibdhost_pack_cached.restype = C.c_void_p
        _lib.ibdhost_pack_cached.argtypes = [C.c_char_p] * 6 + [C.POINTER(C.c_int)]
        _lib.ibdhost_pack_vcf_cached.restype = C.c_void_p
        _lib.ibdhost_pack_vcf_cached.argtypes = [C.c_char_p] * 4 + [C.c_double, C.POINTER(C.c_int)]
        _lib.ibdhost_site_label.restype = C.c_int
        _lib.ibdhost_site_label.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int]
        _lib.ibdhost_name.restype = C.c_char_p
        _lib.ibdhost_name.argtypes = [C.c_void_p, C.c_int32]
        _lib.ibdhost_free.argtypes = [C.c_void_p]
    return _lib


def _b(s):
    return None if s is None else s.encode()


def pack_cached(hap, legend, indv, cache, pileup, chrom=None):
    """IMPUTE panel through the binary cache: (arrays as pack() returns them, cache_hit)."""
    L = lib()
    hit = C.c_int(-1)
    h = L.ibdhost_pack_cached(_b(hap), _b(legend), _b(indv), _b(cache), _b(pileup), _b(chrom), C.byref(hit))
    return (_arrays(L, h) if h else None), hit.value


def pack_vcf_cached(vcf, cache, pileup, chrom=None, min_qual=0.0):
    L = lib()
    hit = C.c_int(-1)
    h = L.ibdhost_pack_vcf_cached(_b(vcf), _b(cache), _b(pileup), _b(chrom), min_qual, C.byref(hit))
    return (_arrays(L, h) if h else None), hit.value


def pack(mode, a, b, c, pileup, chrom=None, positions=None, af=None, min_qual=0.0):
    """Returns a dict of numpy copies of the packed arrays, or None if the packer reported an error."""
    L = lib()
    h = L.ibdhost_pack(mode, _b(a), _b(b), _b(c), _b(pileup), _b(chrom), _b(positions), _b(af), min_qual)
    if not h:
        return None
    return _arrays(L, h)


def _arrays(L, h):
    try:
        S, N, Wh = L.ibdhost_n_sites(h), L.ibdhost_n_indiv(h), L.ibdhost_words(h)

        def arr(ptr, dtype, n):
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(dtype)), shape=(n,)).copy()

        out = dict(S=S, N=N, Wh=Wh, pos=arr(L.ibdhost_pos(h), C.c_uint64, S), n_ref=arr(L.ibdhost_n_ref(h), C.c_uint8, S),
                   n_alt=arr(L.ibdhost_n_alt(h), C.c_uint8, S), keep=arr(L.ibdhost_keep(h), C.c_uint8, S),
                   dp=arr(L.ibdhost_dp(h), C.c_uint32, S), bits=arr(L.ibdhost_bits(h), C.c_uint32, S * Wh).reshape(S, Wh),
                   names=[L.ibdhost_name(h, i).decode() for i in range(N)],
                   pileup_cov=arr(L.ibdhost_pileup_cov(h), C.c_uint32, L.ibdhost_n_pileup(h)))
        buf = C.create_string_buffer(1024)
        out["labels"] = [buf.value.decode() if L.ibdhost_site_label(h, s, buf, 1024) > 0 else None for s in range(S)] \
            if S <= 100_000 else None  # chr, rsID, REF, ALT of the kept sites (small panels only)
        af_ptr = L.ibdhost_af_user(h)
        out["af_user"] = arr(af_ptr, C.c_double, S) if af_ptr else None
        return out
    finally:
        L.ibdhost_free(h)

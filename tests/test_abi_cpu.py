"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol
include/ibdgem_b200.h declares; the host mirror fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest

import ibdgem_b200 as ib
from ibdgem_b200._lib import ABI_SYMBOLS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "ibdgem_b200.h")) as fh:
        hdr = fh.read()
    declared = set(re.findall(r"\b((?:ibdgem|hiddengem)_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(ABI_SYMBOLS)
    lib = ib.load_library()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.ibdgem_abi_version() == 2


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ib.EngineError, match="no usable CUDA device"):
        ib.Engine(ib.Params())


def test_param_validation_messages_follow_reference():
    import torch
    if torch.cuda.is_available():
        with pytest.raises(ib.EngineError, match=r"window size \(-w\)"):
            ib.Engine(ib.Params(window_size=1))
    with pytest.raises(ib.EngineError, match=r"maximum estimated coverage \(-M\)"):
        ib.Engine(ib.Params(max_cov=0))


def test_pack_bits_layout():
    rng = np.random.default_rng(0)
    hap = rng.integers(0, 2, (7, 70)).astype(np.uint8)
    bits = ib.pack_bits(hap)
    assert bits.shape == (7, 4) and bits.dtype == np.uint32
    for s in range(7):
        for h in range(70):
            assert (int(bits[s, h >> 5]) >> (h & 31)) & 1 == hap[s, h]
    assert (bits[:, 2] >> 6).max() == 0 and bits[:, 3].max() == 0

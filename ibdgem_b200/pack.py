"""Host packer helpers: alleles -> the site-major bit panel the engine uploads
(include/ibdgem_b200.h, ibdgem_engine_upload_panel)."""
from __future__ import annotations

import numpy as np


def pack_bits(hap: np.ndarray, align_words: int = 4) -> np.ndarray:
    """hap: [S, 2N] array of 0/1 alleles -> [S, Wh] uint32, haplotype h in bit (h & 31) of word
    (h >> 5).  Rows are padded to a multiple of `align_words` words (16 bytes) so the device can
    use 128-bit loads."""
    hap = np.ascontiguousarray(hap, dtype=np.uint8)
    S, H = hap.shape
    wh = (H + 31) // 32
    wh = (wh + align_words - 1) // align_words * align_words
    padded = np.zeros((S, wh * 32), dtype=np.uint8)
    padded[:, :H] = hap & 1
    by = np.packbits(padded, axis=1, bitorder="little")  # [S, wh*4] bytes, LSB-first
    return np.ascontiguousarray(by).view("<u4").reshape(S, wh)

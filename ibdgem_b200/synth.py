"""Seeded synthetic inputs of the shapes BASELINE.json names.

C3 (north star): S = 1,000,000 sites, N = 2,504 panel individuals, allele frequency
Beta(0.5, 2) clipped to [0.01, 0.99], haplotypes Bernoulli(AF), pileup depth Poisson(2) + 1
(the depth >= 1 variant: every site informative, exactly S / W windows) with reads drawn from
individual 0's genotype at error 0.02.  Generation runs on the GPU when one is present (torch
is plumbing here), otherwise in numpy chunks; both are seeded and deterministic.
"""
from __future__ import annotations

import numpy as np


def synth_panel_numpy(S: int, N: int, seed: int = 1, depth_mean: float = 2.0, depth_floor: int = 1,
                      eps: float = 0.02, src: int = 0, chunk: int = 50_000, want_hap: bool = False):
    """Returns dict(pos, n_ref, n_alt, keep, bits[S, Wh] uint32, af, hap (optional))."""
    from .pack import pack_bits
    rng = np.random.default_rng(seed)
    H = 2 * N
    af = np.clip(rng.beta(0.5, 2.0, S), 0.01, 0.99)
    pos = (1000 + 60 * np.arange(S, dtype=np.uint64)).astype(np.uint64)
    bits_chunks, hap_chunks = [], []
    g_src = np.zeros(S, np.int64)
    for s0 in range(0, S, chunk):
        s1 = min(S, s0 + chunk)
        hap = (rng.random((s1 - s0, H)) < af[s0:s1, None]).astype(np.uint8)
        g_src[s0:s1] = hap[:, 2 * src].astype(np.int64) + hap[:, 2 * src + 1]
        bits_chunks.append(pack_bits(hap))
        if want_hap:
            hap_chunks.append(hap)
    depth = rng.poisson(depth_mean, S) + depth_floor
    depth = np.minimum(depth, 20)
    p_alt = np.where(g_src == 0, eps, np.where(g_src == 1, 0.5, 1 - eps))
    n_alt = rng.binomial(depth, p_alt)
    n_ref = depth - n_alt
    out = dict(pos=pos, n_ref=n_ref.astype(np.uint8), n_alt=n_alt.astype(np.uint8),
               keep=np.ones(S, np.uint8), bits=np.concatenate(bits_chunks), af=af, N=N, S=S)
    if want_hap:
        out["hap"] = np.concatenate(hap_chunks)
    return out


def synth_panel_torch(S: int, N: int, seed: int = 1, depth_mean: float = 2.0, depth_floor: int = 1,
                      eps: float = 0.02, src: int = 0, device="cuda", chunk: int = 100_000):
    """Same distribution, generated on the GPU (fast for the 1M x 5008 panel) and returned as
    PINNED host tensors, which is where the end-to-end measurement starts from."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    H = 2 * N
    wh = (H + 31) // 32
    wh = (wh + 3) // 4 * 4
    beta = torch.distributions.Beta(torch.tensor(0.5), torch.tensor(2.0))
    torch.manual_seed(seed)
    af = beta.sample((S,)).clamp_(0.01, 0.99).to(device)
    bits_host = torch.empty((S, wh), dtype=torch.int32, pin_memory=True)
    g_src = torch.empty(S, dtype=torch.int64, device=device)
    weights = (2 ** torch.arange(32, device=device, dtype=torch.int64))
    for s0 in range(0, S, chunk):
        s1 = min(S, s0 + chunk)
        u = torch.rand((s1 - s0, H), generator=g, device=device)
        hap = (u < af[s0:s1, None])
        g_src[s0:s1] = hap[:, 2 * src].long() + hap[:, 2 * src + 1].long()
        padded = torch.zeros((s1 - s0, wh * 32), dtype=torch.int64, device=device)
        padded[:, :H] = hap
        words = (padded.view(s1 - s0, wh, 32) * weights).sum(dim=2)  # < 2^32
        words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
        bits_host[s0:s1].copy_(words)
        del u, hap, padded, words
    depth = (torch.poisson(torch.full((S,), depth_mean, device=device), generator=g) + depth_floor).clamp_(max=20)
    p_alt = torch.where(g_src == 0, torch.tensor(eps, device=device),
                        torch.where(g_src == 1, torch.tensor(0.5, device=device), torch.tensor(1 - eps, device=device)))
    n_alt = torch.binomial(depth, p_alt, generator=g)
    n_ref = depth - n_alt
    torch.cuda.synchronize()

    def pin(t):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        return h

    pos = pin((1000 + 60 * torch.arange(S, dtype=torch.int64)))
    return dict(pos=pos, n_ref=pin(n_ref.to(torch.uint8).cpu()), n_alt=pin(n_alt.to(torch.uint8).cpu()),
                keep=pin(torch.ones(S, dtype=torch.uint8)), bits=bits_host, N=N, S=S)


def unpack_rows(bits: np.ndarray, H: int, rows) -> np.ndarray:
    """Rows of a packed panel back to [len(rows), H] 0/1 alleles (for the CPU checker)."""
    sub = np.ascontiguousarray(bits[rows]).view(np.uint8)
    return np.unpackbits(sub, axis=1, bitorder="little")[:, :H]

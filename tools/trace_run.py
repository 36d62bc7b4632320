"""One C3 --LD pass with the panel resident (single window range), for IBDGEM_MMA_TRACE captures."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ibdgem_b200 as ib  # noqa: E402
from ibdgem_b200.synth import synth_panel_torch  # noqa: E402

S, N, T, W = 1_000_000, 2504, 1000, 1000
SRC = int(os.environ.get("SRC", "0"))        # individual the reads are drawn from
NBG = int(os.environ.get("NBG", str(N)))     # background = individuals 0 .. NBG-1
d = synth_panel_torch(S, N, seed=1, src=SRC, device="cuda")
with ib.Engine(ib.Params(window_size=W)) as e:
    e.upload_sites(d["pos"].numpy().view(np.uint64), d["n_ref"].numpy(), d["n_alt"].numpy(), d["keep"].numpy())
    e.upload_panel(d["bits"].numpy().view(np.uint32), N)
    e.sync_uploads()
    e.enable_timing(True)
    for _ in range(3):
        e.invalidate()
        sc = e.score_ld(np.arange(T, dtype=np.int32), np.arange(NBG, dtype=np.int32), -1)
    print({k: round(v[0] / v[1], 3) for k, v in e.kernel_stats().items() if v[1]})

"""C3 -v --LD passes for timing experiments on the per-target-window path (IBDGEM_VMMA_* knobs)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ibdgem_b200 as ib  # noqa: E402
from ibdgem_b200.synth import synth_panel_torch  # noqa: E402

S, N, T, W = int(os.environ.get("S", 1_000_000)), 2504, int(os.environ.get("T", 1000)), 1000
d = synth_panel_torch(S, N, seed=1, device="cuda")
with ib.Engine(ib.Params(window_size=W, variable_sites_only=1)) as e:
    e.upload_sites(d["pos"].numpy().view(np.uint64), d["n_ref"].numpy(), d["n_alt"].numpy(), d["keep"].numpy())
    e.upload_panel(d["bits"].numpy().view(np.uint32), N)
    e.sync_uploads()
    e.enable_timing(True)
    for _ in range(3):
        e.invalidate()
        sc = e.score_ld(np.arange(T, dtype=np.int32), np.arange(N, dtype=np.int32), -1)
    print(os.environ.get("IBDGEM_VMMA_BPROD", "0"), os.environ.get("IBDGEM_VMMA_DEBUG", "0"),
          {k: round(v[0] / v[1], 3) for k, v in e.kernel_stats().items() if v[1] and k.startswith(("ld_vmma", "v_"))})

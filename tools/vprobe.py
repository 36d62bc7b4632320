"""C3 -v --LD passes for timing experiments on the per-target-window path (IBDGEM_VMMA_* knobs)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ibdgem_b200 as ib  # noqa: E402
from ibdgem_b200.synth import synth_panel_torch  # noqa: E402

S, N, T, W = int(os.environ.get("S", 1_000_000)), 2504, int(os.environ.get("T", 1000)), 1000
import torch  # noqa: E402
from ibdgem_b200.engine import _CScores  # noqa: E402
d = synth_panel_torch(S, N, seed=1, device="cuda")
maxW = S // W + 2


def pinned(shape, dt):
    return torch.empty(shape, dtype=dt).pin_memory()


o_nw = pinned((T,), torch.int32)
o_ws, o_we, o_wn = pinned((T, maxW), torch.int64), pinned((T, maxW), torch.int64), pinned((T, maxW), torch.int32)
o_ll = pinned((T, maxW, 3), torch.float64)
cs = _CScores(maxW, o_nw.data_ptr(), o_ws.data_ptr(), o_we.data_ptr(), o_wn.data_ptr(), o_ll.data_ptr(),
              None, None, None, None, None, None, None, None)
targets, bg = np.arange(T, dtype=np.int32), np.arange(N, dtype=np.int32)
stream = torch.cuda.current_stream()
with ib.Engine(ib.Params(window_size=W, variable_sites_only=1)) as e:
    e.set_stream(stream.cuda_stream)
    e.upload_sites(d["pos"].numpy().view(np.uint64), d["n_ref"].numpy(), d["n_alt"].numpy(), d["keep"].numpy())
    e.upload_panel(d["bits"].numpy().view(np.uint32), N)
    e.sync_uploads()
    e.enable_timing(True)
    for _ in range(3):
        e.invalidate()
        e.score_ld_raw(targets, bg, -1, cs)
    e.reset_stats()
    n = 5
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n):
        e.invalidate()
        e.score_ld_raw(targets, bg, -1, cs)
    e1.record(stream)
    torch.cuda.synchronize()
    print("side=%s direct=%s debug=%s step %.3f ms" % (os.environ.get("IBDGEM_V_SIDE", "1"), os.environ.get("IBDGEM_LD_DIRECT_STORE", "1"),
                                                     os.environ.get("IBDGEM_VMMA_DEBUG", "0"), e0.elapsed_time(e1) / n),
          {k: round(v[0] / v[1], 3) for k, v in e.kernel_stats().items() if v[1] and k.startswith(("ld_", "v_", "site", "fill"))})

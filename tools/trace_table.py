"""Prints the per-tile timeline of an IBDGEM_MMA_TRACE dump (CTA 0: MMA issuer + two epilogue warps)."""
import collections
import sys

import numpy as np

t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(4, 1024)
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 70)
ev = []
for role in range(3):
    for x in t[role]:
        if x == 0:
            continue
        x = int(x)
        ev.append((x >> 16, role, (x >> 12) & 0xF, x & 0xFFF))
ev.sort()
t0 = ev[0][0]
d = collections.defaultdict(dict)
for c, role, e, tile in ev:
    d[tile].setdefault((role, e), c - t0)
print("tile  mma_wait   mma_got (wait)    r1:got     E1    r2:got     E2")
for tile in sorted(d):
    if not lo <= tile < hi:
        continue
    x = d[tile]
    g = lambda k: x.get(k, -1)
    print(f"{tile:4d} {g((0,1)):9d} {g((0,2)):9d} ({g((0,2))-g((0,1)):6d})  {g((1,2)):8d} {g((1,3))-g((1,2)):6d}  "
          f"{g((2,2)):8d} {g((2,3))-g((2,2)):6d}")

"""Prints the last steps of an IBDGEM_TIMELINE file (engine.cu: resolve_timers) as a table:
kernel / event, start and end in ms since the step's upload_sites.  Diagnostic tool."""
import sys

path = sys.argv[1]
last = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps, cur = [], []
for line in open(path):
    line = line.strip()
    if line == "---":
        steps.append(cur)
        cur = []
    elif line:
        n, a, b = line.split()
        cur.append((float(a), float(b), n))
for st in steps[-last:]:
    print("step with %d entries" % len(st))
    for a, b, n in sorted(st):
        print("  %-22s %8.3f %8.3f  (%.3f)" % (n, a, b, b - a))

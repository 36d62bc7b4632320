import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
        k = d["roofline"]["kernels_ms_per_step"]
        print(f, "ms/step=%.2f value=%.3e mma=%.2f e2e_ms=%.2f" % (d["ms_per_step"], d["value"], k.get("ld_mma", 0), d["e2e"]["ms_per_step"]),
              {a: round(b, 3) for a, b in k.items() if a != "ld_mma"})
    except Exception as ex:
        print(f, "ERR", ex)

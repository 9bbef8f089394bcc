"""Print the headline and the per-kernel components of a bench.py JSON line: python scripts/show_bench.py FILE..."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        print(path, round(d["value"], 1), "expr/s", round(d["ms_per_step"], 3), "ms/step | e2e", round(d["e2e"]["value"], 1),
              "| launches/step", d.get("gpu_launches", 0) // max(1, d.get("steps", 1)))
        for c in d.get("components") or []:
            print("   %-72s %8.3f ms  frac %.3f" % (c["kernel"], c["ms"], c["frac"]))
    except Exception as e:      # noqa: BLE001
        print(path, ": no line", e)

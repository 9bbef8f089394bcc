#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` export: stall reasons and the hottest instructions.

    python scripts/sass_stalls.py gpurun_out/prof_X_sass.csv [top]
"""
import csv
import io
import sys


def main(path, top=30):
    txt = open(path).read()
    for sec in txt.split('"Kernel Name",')[1:]:
        lines = sec.split("\n")
        name = lines[0][:100]
        rdr = csv.reader(io.StringIO("\n".join(lines[1:])))
        hdr = next(rdr)
        rows = [r for r in rdr if len(r) == len(hdr)]
        idx = {h: i for i, h in enumerate(hdr)}
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        f = lambda r, h: float(r[idx[h]] or 0)
        tot = {h: sum(f(r, h) for r in rows) for h in stalls}
        alls = sum(f(r, "# Samples") for r in rows)
        inst = sum(f(r, "Instructions Executed") for r in rows)
        print("\n== %s\n   samples %.0f, warp instructions %.0f, SASS lines %d" % (name, alls, inst, len(rows)))
        for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
            print("   %-28s %8.0f %5.1f%%" % (h, v, 100 * v / max(alls, 1)))
        print("   -- hottest by samples")
        for r in sorted(rows, key=lambda r: -f(r, "# Samples"))[:top]:
            st = {h: f(r, h) for h in stalls}
            main_ = max(st.items(), key=lambda kv: kv[1])
            print("   %6s %10s  %-64s %s" % (r[idx["# Samples"]], r[idx["Instructions Executed"]], r[idx["Source"]][:64], main_[0]))
        # instruction histogram by execution count
        print("   -- executed-instruction mass by opcode")
        ops = {}
        for r in rows:
            src = r[idx["Source"]].split()
            op = src[1] if src and src[0].startswith("@") and len(src) > 1 else (src[0] if src else "?")
            op = op.split(".")[0]
            ops[op] = ops.get(op, 0) + f(r, "Instructions Executed")
        for op, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]:
            print("   %-10s %12.0f %5.1f%%" % (op, v, 100 * v / max(inst, 1)))
        break     # the export repeats each kernel


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)

#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv` output: stall-reason totals and the hottest SASS instructions.
usage: ncu -i rep --page source --csv > src.csv ; python tools/ncu_src_summary.py src.csv [top_n]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hi = his[0]
    end = his[1] - 1 if len(his) > 1 else len(rows)
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
    col = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {s: 0 for s in stalls}
    for r in data:
        for s in stalls:
            tot[s] += int(r[col[s]] or 0)
    T = sum(tot.values())
    print("kernel:", rows[0][1] if rows[0] else "?")
    print("total warp-stall samples", T, " instructions", len(data))
    for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]:
        print(f"  {s:26s} {v:8d} {100 * v / max(T, 1):5.1f}%")
    ops = {}
    for r in data:
        op = r[col["Source"]].strip().split()
        op = [t for t in op if not t.startswith("@")]
        name = op[0].split(".")[0] if op else "?"
        ops.setdefault(name, [0, 0])
        ops[name][0] += int(r[col["Instructions Executed"]] or 0)
        ops[name][1] += int(r[col["# Samples"]] or 0)
    print("executed warp-instructions by opcode:")
    for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"  {k:10s} {v[0]:12d}  samples {v[1]:8d}")
    print("hottest instructions:")
    for r in sorted(data, key=lambda r: -int(r[col["# Samples"]] or 0))[:topn]:
        st = {s: int(r[col[s]] or 0) for s in stalls}
        big = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(r[col["# Samples"]].rjust(7), r[col["Instructions Executed"]].rjust(10), r[col["Source"]].strip()[:64].ljust(64), big)


if __name__ == "__main__":
    main()

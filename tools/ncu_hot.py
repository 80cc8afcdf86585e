#!/usr/bin/env python
"""Summarises `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` output: stall mix and
the instructions that collect the most warp-stall samples."""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.012
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    si, src, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot, data = {}, []
    for k, r in enumerate(rows[hi + 1:]):
        if len(r) < len(h) or not r[si].isdigit():
            continue
        data.append((k, int(r[si]), r))
        for i in stall_cols:
            tot[h[i]] = tot.get(h[i], 0) + int(r[i] or 0)
    total = sum(d[1] for d in data)
    print("total samples", total, "instructions", sum(int(d[2][ie] or 0) for d in data))
    print(sorted(tot.items(), key=lambda x: -x[1])[:8])
    cum = 0
    for k, s, r in data:
        cum += s
        if s > total * frac:
            st = {h[i]: int(r[i] or 0) for i in stall_cols if int(r[i] or 0) > 0}
            top = sorted(st.items(), key=lambda x: -x[1])[:2]
            print(k, s, f"{cum / total:.2f}", r[ie], r[src][:64], top)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-kernel mean duration and share of an ncu launch list
(`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`). Prints a markdown table."""
import collections
import csv
import re
import sys


def main():
    fn = sys.argv[1]
    rows = list(csv.reader(line for line in open(fn) if line.startswith('"')))
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        name = re.sub(r"^(smb|vis)::(<unnamed>::)?", "", name)
        name = re.sub(r"<.*", "", name)
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[r[ui]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    launches = sum(a[0] for a in agg.values())
    print(f"| kernel | launches | mean us | share |\n|---|---:|---:|---:|")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| {n} | {a[0]} | {a[1] / a[0]:.2f} | {100 * a[1] / total:.1f} % |")
    print(f"| **total** | {launches} | {total / launches:.2f} | {total:.0f} us |")


if __name__ == "__main__":
    main()

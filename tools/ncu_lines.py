#!/usr/bin/env python
"""Per-source-line warp-stall samples and executed instructions of one kernel.

  ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv
  cuobjdump -xelf all lib.so; nvdisasm -gi -c file.cubin > file.dis
  tools/ncu_lines.py src.csv file.dis K source.cu

The SASS rows of the ncu page and of nvdisasm are in the same order, so the n-th instruction of
the kernel in the disassembly gives the source line of the n-th row."""
import csv
import re
import sys


def lines_of_kernel(dis_path, kernel, source_name):
    out, inside, cur, locked = [], False, None, False
    for ln in open(dis_path):
        if ln.startswith("\t.section\t.text.") or ln.startswith("//-----"):
            inside = inside
            if ln.startswith("\t.section"):
                inside = (kernel + "E") in ln
            continue
        if not inside:
            continue
        if "//## File" in ln:
            # innermost location first, then the chain of call sites it was inlined at
            # (the call sites follow as further "File" lines of their own: ignored until the next instruction)
            if locked:
                continue
            locs = re.findall(r'"([^"]+)", line (\d+)', ln)
            for path, line in locs:
                if path.endswith(source_name):
                    cur = int(line)
                    break
            locked = "inlined at" in ln
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            out.append(cur)
            locked = False
    return out


def main():
    src_csv, dis, kernel, source = sys.argv[1:5]
    rows = list(csv.reader(open(src_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    si, ie = h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    body = [r for r in rows[hi + 1:] if len(r) >= len(h) and r[si].isdigit()]
    lines = lines_of_kernel(dis, kernel, source.split("/")[-1])
    n = len(lines)
    body = body[:n]  # the page may repeat the kernel
    print(f"{len(body)} SASS rows, {n} disassembled instructions")
    agg = {}
    for r, line in zip(body, lines):
        a = agg.setdefault(line, [0, 0, {}])
        a[0] += int(r[si]); a[1] += int(r[ie] or 0)
        for i in stall_cols:
            v = int(r[i] or 0)
            if v: a[2][h[i]] = a[2].get(h[i], 0) + v
    total = sum(a[0] for a in agg.values()); total_i = sum(a[1] for a in agg.values())
    text = open(source).read().split("\n")
    print(f"samples {total}  warp instructions {total_i}")
    for line in sorted(k for k in agg if k is not None):
        s, ins, st = agg[line]
        if s < total * 0.01 and ins < total_i * 0.01:
            continue
        top = ",".join(f"{k[6:]}:{v}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
        print(f"{line:5d} {100*s/total:5.1f}% smp {100*ins/total_i:5.1f}% ins  {text[line-1].strip()[:70]:70s} {top}")


if __name__ == "__main__":
    main()

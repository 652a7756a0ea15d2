#!/usr/bin/env python
"""Join an `ncu --page source --csv` dump (SASS rows with executed counts and stall samples) with the line table
of the same cubin (`nvdisasm -g -c`) and print the hottest source lines.

  cuobjdump -xelf all optable_b200/liboptb.so && nvdisasm -g -c optb.sm_100a.cubin > sass.txt
  ncu -i prof.ncu-rep --page source --csv > src.csv
  python tools/sass_lines.py src.csv sass.txt 'trace_kernelILb1' [top]
"""
import collections
import csv
import re
import sys


def line_table(path, func):
    table, cur, active = {}, None, False
    for ln in open(path):
        if ln.startswith("//-") and ".text." in ln:
            active = func in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    src, sass, func = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = line_table(sass, func)
    rows = list(csv.reader(open(src)))[2:]
    base = int(rows[0][0], 16)
    inst, samp, thr = collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows:
        try:
            off, n, t, s = int(r[0], 16) - base, int(r[5]), int(r[6]), int(r[4])
        except ValueError:
            continue
        loc = table.get(off, (None, ""))[0]
        inst[loc] += n
        thr[loc] += t
        samp[loc] += s
    ti, ts = sum(inst.values()), sum(samp.values())
    print(f"total warp instructions {ti}, stall samples {ts}")
    print(f"{'file:line':32s} {'warp inst':>12s} {'%':>6s} {'samples%':>8s} {'avg thr':>8s}")
    for loc, n in inst.most_common(top):
        name = f"{loc[0]}:{loc[1]}" if loc else "?"
        print(f"{name:32s} {n:12d} {100 * n / ti:6.2f} {100 * samp[loc] / ts:8.2f} {thr[loc] / max(n, 1):8.1f}")


if __name__ == "__main__":
    main()

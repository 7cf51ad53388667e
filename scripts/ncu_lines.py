#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export:
instructions executed (warp level) and stall samples by CUDA source line, top N."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
rows = []
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] and r[0].isdigit():
        i_inst = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        try:
            rows.append((cur_file, int(r[0]), r[1].strip()[:110], int(r[i_inst]), int(r[i_samp])))
        except ValueError:
            pass
tot_i = sum(x[3] for x in rows) or 1
tot_s = sum(x[4] for x in rows) or 1
print(f"total warp instructions {tot_i}, samples {tot_s}")
for x in sorted(rows, key=lambda x: -x[3])[:top]:
    print(f"{100*x[3]/tot_i:5.1f}% inst {100*x[4]/tot_s:5.1f}% smp  {x[0]}:{x[1]:<5d} {x[2]}")

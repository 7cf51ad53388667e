#!/usr/bin/env python
"""Per-kernel, per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export of SEVERAL
kernels: warp instructions and stall samples by CUDA source line (top N per kernel), plus the stall-reason totals.
Usage: ncu_kernel_lines.py source.csv [top] [kernel-substring]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
only = sys.argv[3] if len(sys.argv) > 3 else None
cur_file = kern = hdr = None
lines = defaultdict(lambda: defaultdict(lambda: [0, 0, ""]))  # kern -> (file,line) -> [inst, samples, text]
stalls = defaultdict(lambda: defaultdict(int))
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        kern = r[1].split("(")[0].replace("void cssm::", "")
    elif r[0] == "Line No":
        hdr = r
        i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        st_idx = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    elif r[0].isdigit():
        try:
            e = lines[kern][(cur_file, int(r[0]))]
            e[0] += int(r[i_inst]); e[1] += int(r[i_samp]); e[2] = r[1].strip()[:120]
            for i, h in st_idx:
                stalls[kern][h] += int(r[i] or 0)
        except ValueError:
            pass
for k, d in lines.items():
    if only and only not in k:
        continue
    ti = sum(v[0] for v in d.values()) or 1
    ts = sum(v[1] for v in d.values()) or 1
    print(f"\n=== {k}: warp instructions {ti}, samples {ts}")
    print("   stalls: " + ", ".join(f"{h[6:]} {100*v/ts:.0f}%" for h, v in sorted(stalls[k].items(), key=lambda x: -x[1])[:8]))
    for (f, ln), v in sorted(d.items(), key=lambda x: -x[1][0])[:top]:
        print(f"{100*v[0]/ti:5.1f}% inst {100*v[1]/ts:5.1f}% smp  {f}:{ln:<5d} {v[2]}")

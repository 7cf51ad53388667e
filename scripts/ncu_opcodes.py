#!/usr/bin/env python
"""Dynamic opcode mix from an `ncu --page source --csv --print-source sass` export."""
import csv, sys, collections
hdr=None; cnt=collections.Counter(); tot=0
for r in csv.reader(open(sys.argv[1])):
    if not r: continue
    if r[0] in ("Address","#") or (hdr is None and "Instructions Executed" in r):
        hdr=r; continue
    if hdr is None: continue
    try:
        src=r[hdr.index("Source")]; n=int(r[hdr.index("Instructions Executed")])
    except (ValueError, IndexError):
        continue
    toks=src.split()
    if not toks: continue
    op=toks[1] if toks[0].startswith("@") and len(toks)>1 else toks[0]
    op=op.split(".")[0] if len(sys.argv)<3 else op
    cnt[op]+=n; tot+=n
print("total",tot)
for op,n in cnt.most_common(30): print(f"{100*n/tot:5.1f}%  {n:>11d}  {op}")

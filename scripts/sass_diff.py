#!/usr/bin/env python3
"""Which kernels of two objects / libraries differ in their machine code?  (cuobjdump -sass, instruction encodings only.)
Usage: sass_diff.py old.o new.o [substring]   -- the check behind "the single-rank kernels are bit-identical" """
import hashlib
import re
import subprocess
import sys


def sass(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    d, cur = {}, None
    for l in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", l)
        if m:
            cur = m.group(1)
            d[cur] = []
            continue
        if cur:
            d[cur] += re.findall(r"/\* (0x[0-9a-f]{16}) \*/", l)
    return {k: (hashlib.md5("".join(v).encode()).hexdigest(), len(v)) for k, v in d.items()}


a, b = sass(sys.argv[1]), sass(sys.argv[2])
only = sys.argv[3] if len(sys.argv) > 3 else ""
same = diff = 0
for k in sorted(set(a) | set(b)):
    if only not in k:
        continue
    if a.get(k) == b.get(k):
        same += 1
    else:
        diff += 1
        print("DIFF", k[:110], a.get(k, (0, 0))[1], b.get(k, (0, 0))[1])
print(f"{same} kernels identical, {diff} differ")

#!/bin/bash
# flat (atomics-free) K2/K3 tables against the two-level ones at mid-size clouds; CSSM_FLAT_MAX_NT is the switch
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for n in 1048576 2097152 4194304; do
for v in 4096 0; do
CSSM_FLAT_MAX_NT=$v timeout 300 python bench.py --no-cpu --workload c2 --particles $n --obs 300 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=$n flat_max=$v', '%.4g'%j['value'], {k:round(x,4) for k,x in j['roofline']['kernel_ms_per_launch'].items() if x})"
done; done

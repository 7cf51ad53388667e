#!/bin/bash
# round 2: the certified fp64 scan (K3 fast path) -- parity suites, then A/B on the SAME library (CSSM_K3_FAST=1/0)
TAG=${1:-r02_n}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
for fast in 1 0 1 0; do
  for args in "--obs 300" "--workload c5 --particles 16777216 --obs 100"; do
    CSSM_K3_FAST=$fast timeout 300 python bench.py --no-cpu --no-extra $args 2>gpurun_out/${TAG}_err.txt | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fast=$fast', '$args', '%.4g'%j['value'], {k:round(x,4) for k,x in j['roofline']['kernel_ms_per_launch'].items() if x}, j.get('scan_tiles_last_run'), 'll', j['log_likelihood_mean'])
except Exception as e:
    print('ERR', e, open('gpurun_out/${TAG}_err.txt').read()[-1500:])"
  done
done

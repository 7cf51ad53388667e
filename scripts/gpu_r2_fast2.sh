#!/bin/bash
# round 2: the certified scan as its own kernel -- parity suites, then builds with other launch bounds, fast on / off
TAG=${1:-r02_q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
for lib in "" $(ls composablestatespacemodels_b200/csrc/alt/*.so 2>/dev/null); do
 for fast in 1 0; do
  for args in "--obs 300" "--workload c5 --particles 16777216 --obs 100" "--workload c2 --obs 300"; do
    L=""; [ -n "$lib" ] && L=$PWD/$lib
    CSSM_LIB=$L CSSM_K3_FAST=$fast timeout 300 python bench.py --no-cpu --no-extra $args 2>gpurun_out/${TAG}_err.txt | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=$(basename "$lib") fast=$fast', '$args', '%.4g'%j['value'], {k:round(x,4) for k,x in j['roofline']['kernel_ms_per_launch'].items() if x}, j.get('scan_tiles_last_run'), 'launches', j['gpu_launches'])
except Exception as e:
    print('ERR', e, open('gpurun_out/${TAG}_err.txt').read()[-1500:])"
  done
 done
done

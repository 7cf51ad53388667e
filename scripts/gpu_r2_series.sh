#!/bin/bash
# round 2: the flagged-slot series kernel -- parity suite, c4 with the slot kernel and with the grid-barrier kernel, cycle stamps
TAG=${1:-r02_d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
for ll in 1 0; do
  CSSM_SERIES_LL=$ll CSSM_SERIES_DEBUG=1 timeout 300 python bench.py --workload c4 --steps 30 --no-cpu --chains 2,3 2>gpurun_out/${TAG}_c4err_$ll.txt | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4 LL=$ll', j['value'], j['roofline']['us_per_observation'], [c['value'] for c in j['concurrent_chains']], j['log_likelihood_mean'])"
  tail -2 gpurun_out/${TAG}_c4err_$ll.txt
  CSSM_SERIES_LL=$ll timeout 300 python bench.py --workload c4 --steps 30 --no-cpu --chains 2 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4 (no stamps) LL=$ll', j['value'], j['roofline']['us_per_observation'], [c['value'] for c in j['concurrent_chains']])"
done
for args in "--obs 300" "--workload c2 --obs 300" "--workload c1"; do
  timeout 300 python bench.py --no-cpu --no-extra $args 2>gpurun_out/${TAG}_err.txt | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$args', '%.4g'%j['value'], {k:round(x,4) for k,x in (j['roofline'].get('kernel_ms_per_launch') or {}).items() if x}, j['roofline'].get('us_per_observation'), 'll', j['log_likelihood_mean'])
except Exception as e:
    print('ERR', e, open('gpurun_out/${TAG}_err.txt').read()[-1500:])"
done

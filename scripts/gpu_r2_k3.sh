#!/bin/bash
# round 2: the warp-synchronous K3 -- parity suite, then the target / c2 lines for several launch bounds (CSSM_LIB builds)
TAG=${1:-r02_b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
for lib in "" alt/libcssm_mb4.so alt/libcssm_mb2.so; do
  for args in "--obs 300" "--workload c2 --obs 300"; do
    L=""; [ -n "$lib" ] && L=$PWD/composablestatespacemodels_b200/csrc/$lib
    CSSM_LIB=$L timeout 300 python bench.py --no-cpu --no-extra $args 2>gpurun_out/${TAG}_err.txt | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=$lib', '$args', '%.4g'%j['value'], {k:round(x,4) for k,x in j['roofline']['kernel_ms_per_launch'].items() if x}, 'll', j['log_likelihood_mean'])
except Exception as e:
    print('ERR', e, open('gpurun_out/${TAG}_err.txt').read()[-1500:])"
  done
done
timeout 300 python bench.py --workload c4 --steps 30 --no-cpu --chains 2 2>gpurun_out/${TAG}_c4err.txt | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4', j['value'], j['roofline']['us_per_observation'], [c['value'] for c in j['concurrent_chains']])"
tail -2 gpurun_out/${TAG}_c4err.txt

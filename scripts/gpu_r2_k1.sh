#!/bin/bash
# round 2: K1 without end-of-kernel barriers (+ Philox4x32-7 as an alternative build): parity suite, target / c2 / c5 lines
TAG=${1:-r02_i}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -12 gpurun_out/${TAG}_pytest.log
for lib in "" alt/libcssm_philox7.so; do
  for args in "--obs 300" "--workload c2 --obs 300" "--workload c5 --particles 16777216 --obs 100"; do
    L=""; [ -n "$lib" ] && L=$PWD/composablestatespacemodels_b200/csrc/$lib
    CSSM_LIB=$L timeout 300 python bench.py --no-cpu --no-extra $args 2>gpurun_out/${TAG}_err.txt | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=$lib', '$args', '%.4g'%j['value'], {k:round(x,4) for k,x in j['roofline']['kernel_ms_per_launch'].items() if x}, 'frac', round(j['roofline']['frac'],3), 'll', j['log_likelihood_mean'])
except Exception as e:
    print('ERR', e, open('gpurun_out/${TAG}_err.txt').read()[-1500:])"
  done
done

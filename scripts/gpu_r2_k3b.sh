#!/bin/bash
# round 2: K3 variants -- parity subset, target (degenerate Poisson weights) and c5-model (mild Normal weights) at 2^24, ncu of K3
TAG=${1:-r02_c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
for lib in "" alt/libcssm_v1.so; do
  for args in "--obs 300" "--workload c5 --particles 16777216 --obs 100" "--workload c2 --obs 300"; do
    L=""; [ -n "$lib" ] && L=$PWD/composablestatespacemodels_b200/csrc/$lib
    CSSM_LIB=$L timeout 300 python bench.py --no-cpu --no-extra $args 2>gpurun_out/${TAG}_err.txt | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=$lib', '$args', '%.4g'%j['value'], {k:round(x,4) for k,x in j['roofline']['kernel_ms_per_launch'].items() if x}, 'll', j['log_likelihood_mean'])
except Exception as e:
    print('ERR', e, open('gpurun_out/${TAG}_err.txt').read()[-1500:])"
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_scan" -s 6 -c 1 -f -o /tmp/${TAG}_prof python bench.py --steps 1 --warmup 1 --obs 12 --no-cpu --no-extra > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_source.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_source.csv

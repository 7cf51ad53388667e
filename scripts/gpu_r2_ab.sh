#!/bin/bash
# round 2: interleaved A/B of library builds (CSSM_LIB) on the target line -- each build twice, alternating, so that box
# drift cannot masquerade as a code effect; then the series kernel (c4) with cycle stamps
TAG=${1:-r02_j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pmmh.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
for rep in 1 2; do
for lib in "" $(ls composablestatespacemodels_b200/csrc/alt/*.so); do
  L=""; [ -n "$lib" ] && L=$PWD/$lib
  CSSM_LIB=$L timeout 300 python bench.py --no-cpu --no-extra --obs 300 2>gpurun_out/${TAG}_err.txt | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rep $rep lib=$(basename "$lib")', '%.4g'%j['value'], {k:round(x,4) for k,x in j['roofline']['kernel_ms_per_launch'].items() if x}, 'frac', round(j['roofline']['frac'],3), 'sm', j['clocks']['sm_mhz'])
except Exception as e:
    print('ERR', e, open('gpurun_out/${TAG}_err.txt').read()[-1500:])"
done
done
CSSM_SERIES_DEBUG=1 timeout 300 python bench.py --workload c4 --steps 30 --no-cpu --chains 2,3 2>gpurun_out/${TAG}_c4err.txt | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4', j['value'], j['roofline']['us_per_observation'], [c['value'] for c in j['concurrent_chains']], j['log_likelihood_mean'])"
tail -2 gpurun_out/${TAG}_c4err.txt
timeout 300 python bench.py --workload c4 --steps 50 --no-cpu --chains 2 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4 (no stamps)', j['value'], j['roofline']['us_per_observation'], [c['value'] for c in j['concurrent_chains']])"

#!/bin/bash
# the one-block series kernel: its parity tests, c1 with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "one_block" 2>&1 | tail -8
for env in "X=1" "CSSM_SERIES_ONE=0"; do
  env $env timeout 300 python bench.py --workload c1 --no-cpu --steps 20 2>gpurun_out/one_err.txt | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$env] c1 %.4g'%j['value'], 'e2e %.4g'%j['e2e']['value'], j['roofline'].get('us_per_observation'), j.get('log_likelihood_mean'))
except Exception as e:
    print('ERR', e, open('gpurun_out/one_err.txt').read()[-1500:])"
done

#!/bin/bash
# the GPU suite, the three bench lines that matter, the series stamps
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/${TAG}_pytest.log | head -20
for wl in target c2; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-extra --obs 300 2>gpurun_out/${TAG}_err.txt | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=j['roofline']['kernel_ms_per_launch']; print('$wl %.4g'%j['value'], {a:round(x,4) for a,x in k.items() if x}, 'll', j.get('log_likelihood_mean'))
except Exception as e:
    print('ERR', e, open('gpurun_out/${TAG}_err.txt').read()[-800:])"
done
bash scripts/gpu_series2.sh $TAG $2

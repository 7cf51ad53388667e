#!/bin/bash
# the repeated-key memo in the exact scan path: GPU suite, target + c2 kernel times, the resampling micro-benchmark
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/${TAG}_pytest.log | head -10
for wl in target c2; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-extra --obs 300 2>gpurun_out/${TAG}_err.txt | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=j['roofline']['kernel_ms_per_launch']; print('$wl %.4g'%j['value'], {a:round(x,4) for a,x in k.items() if x}, 'll', j.get('log_likelihood_mean'))"
done
timeout 600 python bench.py --workload resample > gpurun_out/${TAG}_bench_resample.json 2> gpurun_out/${TAG}_bench_resample.err
python - <<PY
import json
j=json.loads(open('gpurun_out/${TAG}_bench_resample.json').read().strip().splitlines()[-1])
print('resample headline %.4g'%j['value'])
for c in j['cases']:
    if c['n'] >= (1<<20): print(c['n'], c['weights'], c['kind'], 'gpu %.4g /s %.2f ms'%(c['gpu_particles_per_s'], c['gpu_ms']), 'cpu', c.get('cpu_ms'), c.get('same_ancestors'))
PY

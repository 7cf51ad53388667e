#!/bin/bash
# One gpurun call that produces everything profiles/ carries for a version: the GPU test suite, smoke(), the default
# bench line (with cpu_baseline), the reference arm, every other workload, and the ncu captures (CSV exported on the box).
TAG=${1:-cur}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_target.json 2> gpurun_out/${TAG}_bench_target.err; tail -c 1500 gpurun_out/${TAG}_bench_target.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 900 gpurun_out/${TAG}_bench_reference.json
for wl in c2 c4 c3 c5 c1; do
  timeout 600 python bench.py --workload $wl --warmup 3 > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${TAG}_bench_${wl}.json').read().strip().splitlines()[-1])
    r=j['roofline'] or {}
    print('${wl}', '%.4g'%j['value'], j['unit'], 'e2e %.4g'%j['e2e']['value'], 'cpu', (j['cpu_baseline'] or {}).get('value'), r.get('kernel_ms_per_launch'), r.get('us_per_observation'))
except Exception as e:
    print('${wl} ERR', e, open('gpurun_out/${TAG}_bench_${wl}.err').read()[-800:])
PY
done
bash scripts/gpu_prof.sh ${TAG} series > gpurun_out/${TAG}_prof.log 2>&1
ls -la gpurun_out | grep ${TAG} | head -40

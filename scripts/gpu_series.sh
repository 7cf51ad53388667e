#!/bin/bash
# single-launch series kernels: parity tests, then c2-type clouds of 2^18..2^22 particles with and without them
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/${TAG}_pytest.log | head -30
for lg in 18 19 20 21 22; do
  for mx in 0 100000000; do
    CSSM_SERIES_MAX_N=$mx timeout 300 python bench.py --workload c2 --particles $((1<<lg)) --obs 300 --warmup 3 --steps 3 --no-cpu > gpurun_out/${TAG}_c2_${lg}_${mx}.json 2> gpurun_out/${TAG}_c2_${lg}_${mx}.err
    python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${TAG}_c2_${lg}_${mx}.json').read().strip().splitlines()[-1])
    print('2^${lg} series_max=${mx}', '%.4g'%j['value'], 'us/obs %.2f'%(j['ms_per_step']*1000/300), 'launches', j['gpu_launches'])
except Exception as e:
    print('ERR', e, open('gpurun_out/${TAG}_c2_${lg}_${mx}.err').read()[-800:])
PY
  done
done
timeout 600 python bench.py --workload c4 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
python -c "
import json;j=json.loads(open('gpurun_out/${TAG}_bench_c4.json').read().strip().splitlines()[-1]);print('c4',j['value'],j['e2e']['value'],j['roofline']['us_per_observation'])"

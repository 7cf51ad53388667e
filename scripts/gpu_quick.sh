#!/bin/bash
# parity tests + target and PMMH bench lines (no profiler)
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/${TAG}_pytest.log | head -30
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_target.json 2> gpurun_out/${TAG}_bench_target.err
timeout 600 python bench.py --workload c4 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
timeout 600 python bench.py --workload c2 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
python - <<PY
import json
for w in ("target","c4","c2"):
    try:
        j=json.loads(open('gpurun_out/${TAG}_bench_%s.json'%w).read().strip().splitlines()[-1])
        r=j['roofline']
        print(w, '%.4g'%j['value'], 'e2e %.4g'%j['e2e']['value'], r.get('kernel_ms_per_launch'), r.get('us_per_observation'))
    except Exception as e:
        print(w,'ERR',e, open('gpurun_out/${TAG}_bench_%s.err'%w).read()[-500:])
PY

#!/bin/bash
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/${TAG}_pytest.log | head -40
CSSM_SERIES_DEBUG=1 timeout 600 python bench.py --workload c4 --warmup 2 --steps 5 --no-cpu > gpurun_out/${TAG}_bench_c4d.json 2> gpurun_out/${TAG}_bench_c4d.err
tail -2 gpurun_out/${TAG}_bench_c4d.err
timeout 600 python bench.py --workload c4 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
python -c "
import json;j=json.loads(open('gpurun_out/${TAG}_bench_c4.json').read().strip().splitlines()[-1]);print(j['value'],j['e2e']['value'],j['roofline']['us_per_observation'])"

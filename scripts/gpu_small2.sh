#!/bin/bash
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "series or parity_systematic" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|FAILED|assert " gpurun_out/${TAG}_pytest.log | head -40
CSSM_SERIES_DEBUG=1 timeout 600 python bench.py --workload c4 --warmup 2 --steps 5 --no-cpu > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
tail -c 700 gpurun_out/${TAG}_bench_c4.json; tail -3 gpurun_out/${TAG}_bench_c4.err
CSSM_SERIES_DEBUG=1 timeout 600 python bench.py --workload c1 --warmup 2 --steps 2 --no-cpu > gpurun_out/${TAG}_bench_c1.json 2> gpurun_out/${TAG}_bench_c1.err
tail -2 gpurun_out/${TAG}_bench_c1.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_target.json 2> gpurun_out/${TAG}_bench_target.err
python -c "
import json;j=json.loads(open('gpurun_out/${TAG}_bench_target.json').read().strip().splitlines()[-1]);print(j['value'],j['roofline']['kernel_ms_per_launch'])"

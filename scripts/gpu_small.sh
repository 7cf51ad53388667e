#!/bin/bash
# small-cloud path: GPU parity tests, then the PMMH (c4) and c1 bench lines
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
for wl in c4 c1; do
  timeout 600 python bench.py --workload $wl --warmup 3 > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err
  tail -c 2500 gpurun_out/${TAG}_bench_${wl}.json; tail -3 gpurun_out/${TAG}_bench_${wl}.err
done

#!/bin/bash
# GPU parity tests + the non-default workloads (c3 LGCP, c4 PMMH, c5 single-GPU leg)
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
for wl in c4 c3 c5 c1; do
  timeout 600 python bench.py --workload $wl --warmup 3 > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err
  tail -c 2500 gpurun_out/${TAG}_bench_${wl}.json; tail -3 gpurun_out/${TAG}_bench_${wl}.err
done

#!/bin/bash
# quick check + the series kernel's cycle stamps (block 0, per observation)
TAG=${1:-cur}
bash scripts/gpu_quick.sh $TAG
CSSM_SERIES_DEBUG=1 timeout 300 python bench.py --workload c4 --steps 20 --no-cpu --chains "" 2>gpurun_out/${TAG}_stamps.txt | tail -c 300
tail -5 gpurun_out/${TAG}_stamps.txt

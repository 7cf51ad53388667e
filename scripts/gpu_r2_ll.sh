#!/bin/bash
# round 2: polling policies of the flagged-slot series kernel (CSSM_LIB builds), cycle stamps
TAG=${1:-r02_e}
mkdir -p gpurun_out
for lib in "" $(ls composablestatespacemodels_b200/csrc/alt/*.so); do
  L=""; [ -n "$lib" ] && L=$PWD/$lib
  CSSM_LIB=$L CSSM_SERIES_DEBUG=1 timeout 300 python bench.py --workload c4 --steps 20 --no-cpu --chains 2 2>gpurun_out/${TAG}_err.txt | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4 lib=$lib', round(j['value'],1), round(j['roofline']['us_per_observation'],2), [round(c['value'],1) for c in j['concurrent_chains']], j['log_likelihood_mean'])"
  tail -1 gpurun_out/${TAG}_err.txt
done

#!/bin/bash
# series kernel: c4 line (one and two chains), per-block cycle stamps; optional ncu source page
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 300 python bench.py --workload c4 --steps 40 --no-cpu --chains 2 2>gpurun_out/${TAG}_c4.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4', round(j['value'],1), round(j['roofline']['us_per_observation'],2), [round(c['value'],1) for c in j['concurrent_chains']], j['log_likelihood_mean'])"
CSSM_SERIES_DEBUG=1 timeout 300 python bench.py --workload c4 --steps 5 --no-cpu --chains "" 2>gpurun_out/${TAG}_stamps.txt >/dev/null
tail -1 gpurun_out/${TAG}_stamps.txt
if [ "$2" == "ncu" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_series" -s 1 -c 1 -f -o /tmp/${TAG}_prof_series python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu --chains "" > gpurun_out/${TAG}_ncu_series.log 2>&1
  ncu -i /tmp/${TAG}_prof_series.ncu-rep --page raw --csv > gpurun_out/${TAG}_series_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_prof_series.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_series_source.csv 2>/dev/null
  gzip -f gpurun_out/${TAG}_series_source.csv
fi

#!/bin/bash
# round 2: full GPU suite, then run-time A/B inside ONE library: K2-tail prefix on/off (target, c2), series kernel with one
# or two particles per thread (c4 + block-0 stamps); optional ncu source page of the series kernel
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/${TAG}_pytest.log | head -20
for env in "X=1" "CSSM_K2_PREFIX=0"; do
  for wl in target c2; do
  env $env timeout 300 python bench.py --workload $wl --no-cpu --no-extra --obs 300 2>gpurun_out/${TAG}_err.txt | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=j['roofline']['kernel_ms_per_launch']; print('[$env] $wl %.4g'%j['value'], {a:round(x,4) for a,x in k.items() if x}, 'll', j.get('log_likelihood_mean'))
except Exception as e:
    print('ERR', e, open('gpurun_out/${TAG}_err.txt').read()[-800:])"
  done
done
for env in "X=1" "CSSM_SERIES_ITEMS=2"; do
  env $env timeout 300 python bench.py --workload c4 --steps 40 --no-cpu --chains 2 2>gpurun_out/${TAG}_c4.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$env] c4', round(j['value'],1), round(j['roofline']['us_per_observation'],2), [round(c['value'],1) for c in j['concurrent_chains']], j['log_likelihood_mean'])"
  env $env CSSM_SERIES_DEBUG=1 timeout 300 python bench.py --workload c4 --steps 5 --no-cpu --chains "" 2>gpurun_out/${TAG}_stamps.txt >/dev/null
  tail -1 gpurun_out/${TAG}_stamps.txt
done
if [ "$2" == "ncu" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_series" -s 1 -c 1 -f -o /tmp/${TAG}_prof_series python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu --chains "" > gpurun_out/${TAG}_ncu_series.log 2>&1
  ncu -i /tmp/${TAG}_prof_series.ncu-rep --page raw --csv > gpurun_out/${TAG}_series_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_prof_series.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_series_source.csv 2>/dev/null
  gzip -f gpurun_out/${TAG}_series_source.csv
fi

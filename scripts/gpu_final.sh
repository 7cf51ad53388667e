#!/bin/bash
# One gpurun call that produces everything profiles/ carries for a version: the GPU test suite, smoke(), the default
# bench line (with cpu_baseline, pmmh sub-record), the reference arm, every other workload, and the ncu captures
# (CSV exported on the box).
TAG=${1:-cur}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_target.json 2> gpurun_out/${TAG}_bench_target.err; tail -c 600 gpurun_out/${TAG}_bench_target.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 600 gpurun_out/${TAG}_bench_reference.json
for wl in c2 c4 c3 c5 c1; do
  timeout 600 python bench.py --workload $wl --warmup 3 > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${TAG}_bench_${wl}.json').read().strip().splitlines()[-1])
    r=j['roofline'] or {}
    print('${wl}', '%.4g'%j['value'], j['unit'], 'e2e %.4g'%j['e2e']['value'], 'cpu', (j['cpu_baseline'] or {}).get('value'), r.get('kernel_ms_per_launch'), r.get('us_per_observation'), r.get('frac'))
except Exception as e:
    print('${wl} ERR', e, open('gpurun_out/${TAG}_bench_${wl}.err').read()[-800:])
PY
done
# ncu: launch list of the default command, full capture of the three step kernels, full capture of the series kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --obs 20 --no-cpu --no-extra > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_propagate|k_weight_sums|k_scan" -s 12 -c 3 -f -o /tmp/${TAG}_prof python bench.py --steps 1 --warmup 1 --obs 12 --no-cpu --no-extra > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_source.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_series" -s 1 -c 1 -f -o /tmp/${TAG}_prof_series python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu --chains "" > gpurun_out/${TAG}_ncu_series.log 2>&1
ncu -i /tmp/${TAG}_prof_series.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_series_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_prof_series.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_series_source.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_series_source.csv
CSSM_SERIES_DEBUG=1 timeout 300 python bench.py --workload c4 --steps 5 --no-cpu --chains "" 2>gpurun_out/${TAG}_series_stamps.txt >/dev/null
gzip -f gpurun_out/${TAG}_source.csv
ls -la gpurun_out | grep ${TAG} | head -40

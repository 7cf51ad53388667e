#!/bin/bash
# ncu: launch list + full capture of the step kernels (target); the .ncu-rep files are exported to CSV on the
# box (raw page + per-line source page) and removed, gpurun_out/ only carries 64 MiB back.
# Usage (under gpurun): bash scripts/gpu_prof.sh <tag> [series]
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --obs 20 --no-cpu > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_propagate|k_weight_sums|k_scan" -s 12 -c 3 -f -o /tmp/${TAG}_prof python bench.py --steps 1 --warmup 1 --obs 12 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_source.csv 2>/dev/null
if [ "$2" == "series" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_series" -s 1 -c 1 -f -o /tmp/${TAG}_prof_series python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_series.log 2>&1
  ncu -i /tmp/${TAG}_prof_series.ncu-rep --page raw --csv > gpurun_out/${TAG}_series_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_prof_series.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_series_source.csv 2>/dev/null
fi
gzip -f gpurun_out/${TAG}_*source.csv
ls -la gpurun_out | grep ${TAG}

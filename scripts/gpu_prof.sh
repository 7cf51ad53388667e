#!/bin/bash
# ncu: launch list + full capture of the step kernels (target) and of the series kernel (c4)
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --obs 20 --no-cpu > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_propagate|k_weight_sums|k_scan" -s 12 -c 3 -f -o gpurun_out/${TAG}_prof python bench.py --steps 1 --warmup 1 --obs 12 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_series" -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_series python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_series.log 2>&1
ls -la gpurun_out | grep ${TAG}

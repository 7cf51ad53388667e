#!/bin/bash
# ncu --set full of one k_scan_search launch of the target line (source-level CSV exported on the box)
TAG=${1:-r02_ncu}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_scan" -s 6 -c 1 -f -o /tmp/${TAG}_prof python bench.py --steps 1 --warmup 1 --obs 12 --no-cpu --no-extra > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_source.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_source.csv
tail -3 gpurun_out/${TAG}_ncu_full.log

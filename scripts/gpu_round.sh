#!/bin/bash
# One gpurun call: GPU parity tests, bench lines, ncu launch list and a full capture of the hot kernels.
# Usage (under gpurun): bash scripts/gpu_round.sh <tag>
TAG=${1:-cur}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_target.json 2> gpurun_out/${TAG}_bench_target.err; tail -c 3000 gpurun_out/${TAG}_bench_target.json
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; tail -c 600 gpurun_out/${TAG}_bench_c2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --obs 20 --no-cpu > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_" -s 12 -c 6 -f -o gpurun_out/${TAG}_prof python bench.py --steps 1 --warmup 1 --obs 12 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -20

#!/bin/bash
# final check of a tree: GPU suite, smoke, default bench line, c3 bench line + ncu of the LGCP kernel
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_target.json 2> gpurun_out/${TAG}_bench_target.err; cut -c1-300 gpurun_out/${TAG}_bench_target.json
timeout 600 python bench.py --workload c3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; cut -c1-300 gpurun_out/${TAG}_bench_c3.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_lgcp" -s 20 -c 1 -f -o /tmp/${TAG}_lgcp python bench.py --workload c3 --steps 1 --warmup 1 --obs 40 --no-cpu > gpurun_out/${TAG}_ncu_lgcp.log 2>&1
ncu -i /tmp/${TAG}_lgcp.ncu-rep --page raw --csv > gpurun_out/${TAG}_lgcp_raw.csv 2>/dev/null
ls -la gpurun_out | grep ${TAG}

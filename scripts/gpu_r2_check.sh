#!/bin/bash
# round 2: GPU suite (new PMMH / driver tests first), default bench line with the pmmh sub-record, c4 line
TAG=${1:-r02_a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_pmmh.py -m gpu -q -x > gpurun_out/${TAG}_pytest_pmmh.log 2>&1; echo "pytest pmmh rc=$?" >> gpurun_out/${TAG}_pytest_pmmh.log
tail -15 gpurun_out/${TAG}_pytest_pmmh.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_pmmh.py > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --no-cpu > gpurun_out/${TAG}_bench_target.json 2> gpurun_out/${TAG}_bench_target.err; tail -c 3000 gpurun_out/${TAG}_bench_target.json; tail -5 gpurun_out/${TAG}_bench_target.err
CSSM_SERIES_DEBUG=1 timeout 600 python bench.py --workload c4 --steps 30 --no-cpu > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err; tail -c 1500 gpurun_out/${TAG}_bench_c4.json; tail -3 gpurun_out/${TAG}_bench_c4.err

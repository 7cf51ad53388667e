#!/bin/bash
# N-GPU run (gpurun --gpus N): the two-process sharding test, then the c5 (sharded) and target (replica) bench lines
N=${1:-2}; TAG=${2:-cur}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -k multiprocess > gpurun_out/${TAG}_pytest_mp.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_mp.log
tail -15 gpurun_out/${TAG}_pytest_mp.log
for wl in c5 target c4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $N --workload $wl --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_${wl}_g${N}.json 2> gpurun_out/${TAG}_bench_${wl}_g${N}.err
  echo "$wl rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_${wl}_g${N}.json; tail -5 gpurun_out/${TAG}_bench_${wl}_g${N}.err
done

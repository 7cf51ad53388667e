#!/bin/bash
# N-GPU run (gpurun --gpus N): the two-process sharding test, then the c5 (sharded), target and c4 (replica) bench lines
N=${1:-2}; TAG=${2:-cur}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/${TAG}_pytest_mp.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_mp.log
tail -3 gpurun_out/${TAG}_pytest_mp.log
for wl in c5 target c4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $N --workload $wl --warmup 3 --steps 3 --no-cpu > gpurun_out/${TAG}_bench_${wl}_g${N}.json 2> gpurun_out/${TAG}_bench_${wl}_g${N}.err
  echo "$wl rc=$?"; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${TAG}_bench_${wl}_g${N}.json').read().strip().splitlines()[-1])
    print('${wl}', j['n_gpus'], '%.4g'%j['value'], j['unit'], 'e2e %.4g'%j['e2e']['value'], 'ms/step %.2f'%j['ms_per_step'])
except Exception as e:
    print('${wl} ERR', e, open('gpurun_out/${TAG}_bench_${wl}_g${N}.err').read()[-1500:])
PY
done

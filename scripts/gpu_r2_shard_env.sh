#!/bin/bash
# N GPUs (gpurun --gpus N): c5 (one 2^27-particle filter sharded over the ranks) under a list of run-time switches
N=${1:-8}; TAG=${2:-r02_senv}; shift 2
mkdir -p gpurun_out
i=0
for env in "X=1" "CSSM_PDL=0" "CSSM_SHARD_LOCAL=0"; do
  i=$((i+1))
  env $env timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$i \
     bench.py --gpus $N --workload c5 --steps 4 --warmup 2 --no-cpu > gpurun_out/${TAG}_c5_${i}_g${N}.json 2> gpurun_out/${TAG}_c5_${i}_g${N}.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${TAG}_c5_${i}_g${N}.json').read().strip().splitlines()[-1])
    r=j['roofline']
    print('[$env] n', j['n_gpus'], 'value %.4g'%j['value'], 'ms/step', round(j['ms_per_step'],2), {k:round(v,4) for k,v in r.get('kernel_ms_per_launch').items() if v}, 'll', j.get('log_likelihood_mean'))
except Exception as e:
    print('ERR', e, open('gpurun_out/${TAG}_c5_${i}_g${N}.err').read()[-2000:])
PY
done

#!/bin/bash
# N GPUs (gpurun --gpus N): c5 (one 2^27-particle filter sharded over the ranks) with interior K1 blocks starting without
# the peers (default) and with the unconditional rendezvous (CSSM_SHARD_LOCAL=0); then the two-process IPC test
N=${1:-2}; TAG=${2:-r02_sab}
mkdir -p gpurun_out
for loc in 1 0; do
  CSSM_SHARD_LOCAL=$loc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$loc \
     bench.py --gpus $N --workload c5 --steps 4 --warmup 2 --no-cpu > gpurun_out/${TAG}_c5_local${loc}_g${N}.json 2> gpurun_out/${TAG}_c5_local${loc}_g${N}.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${TAG}_c5_local${loc}_g${N}.json').read().strip().splitlines()[-1])
    r=j['roofline']
    print('LOCAL=$loc n', j['n_gpus'], 'value %.4g'%j['value'], 'ms/step', round(j['ms_per_step'],2), r.get('kernel_ms_per_launch'), 'll', j.get('log_likelihood_mean'), j['clocks'])
except Exception as e:
    print('ERR', e, open('gpurun_out/${TAG}_c5_local${loc}_g${N}.err').read()[-2000:])
PY
done
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/${TAG}_pytest_mp.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_mp.log
tail -3 gpurun_out/${TAG}_pytest_mp.log

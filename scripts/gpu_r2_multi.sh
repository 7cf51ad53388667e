#!/bin/bash
# round 2, N GPUs (gpurun --gpus N): the two-process IPC test, then the DEFAULT bench line under torchrun (carries the
# pmmh and sharded sub-records), then c5 alone
N=${1:-2}; TAG=${2:-r02_g2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/${TAG}_pytest_mp.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_mp.log
tail -3 gpurun_out/${TAG}_pytest_mp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --pmmh-iters 30 > gpurun_out/${TAG}_bench_default_g${N}.json 2> gpurun_out/${TAG}_bench_default_g${N}.err
echo "default rc=$?"; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${TAG}_bench_default_g${N}.json').read().strip().splitlines()[-1])
    print('target', j['n_gpus'], '%.4g'%j['value'], 'clocks', j['clocks'])
    p=j['pmmh']; print('pmmh', p['value'], p['e2e']['value'], [(c['chains_per_gpu'], round(c['value'],1)) for c in p['concurrent_chains']], p['clocks'])
    s=j['sharded']; print('sharded', '%.4g'%s['value'], 'speedup', round(s['speedup_vs_one_gpu'],3), 'eff', round(s['strong_scaling_efficiency'],3), 'exposed us/obs', round(s['exposed_exchange_us_per_observation'],1), s['kernel_ms_per_launch'], 'one gpu', s['one_gpu']['kernel_ms_per_launch'], 'same bits', s['same_bits_as_one_gpu'], s['clocks'])
except Exception as e:
    print('ERR', e, open('gpurun_out/${TAG}_bench_default_g${N}.err').read()[-3000:])
PY

#!/bin/bash
# round 2: does the leading dimension of the cloud (stride between coordinate arrays) explain the bimodal kernel times?
mkdir -p gpurun_out
for pad in 0 64 448 4160 69696 0; do
  CSSM_NS_PAD=$pad timeout 300 python bench.py --no-cpu --no-extra --obs 300 2>gpurun_out/pad_err.txt | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=j['roofline']['kernel_ms_per_launch']; print('pad=$pad', '%.4g'%j['value'], {a:round(x,4) for a,x in k.items() if x}, 'sum', round(sum(k.values()),4))
except Exception as e:
    print('ERR', e, open('gpurun_out/pad_err.txt').read()[-800:])"
done

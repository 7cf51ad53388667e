#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,serial,uuid,vbios_version,power.limit,clocks.max.sm,clocks.max.mem --format=csv
for lib in "" $(ls composablestatespacemodels_b200/csrc/alt/*.so 2>/dev/null) ""; do
  L=""; [ -n "$lib" ] && L=$PWD/$lib
  CSSM_LIB=$L timeout 300 python bench.py --no-cpu --no-extra --obs 200 2>gpurun_out/modes_err.txt | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=j['roofline']['kernel_ms_per_launch']; print('lib=$(basename "$lib")', '%.4g'%j['value'], {a:round(x,4) for a,x in k.items() if x}, 'sum', round(sum(k.values()),4))
except Exception as e:
    print('ERR', e, open('gpurun_out/modes_err.txt').read()[-800:])"
done

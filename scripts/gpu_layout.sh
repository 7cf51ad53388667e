#!/bin/bash
# code-placement candidates (CSSM_LAYOUT_PAD builds under csrc/alt/): target line kernel times and the PMMH line for each
TAG=${1:-layout}
mkdir -p gpurun_out
for lib in "" $(ls composablestatespacemodels_b200/csrc/alt/*.so 2>/dev/null); do
  L=""; [ -n "$lib" ] && L=$PWD/$lib
  CSSM_LIB=$L timeout 300 python bench.py --no-cpu --no-extra --obs 300 2>gpurun_out/${TAG}_err.txt | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=j['roofline']['kernel_ms_per_launch']; print('lib=$(basename "$lib")', 'target %.4g'%j['value'], {a:round(x,4) for a,x in k.items() if x}, 'sum', round(sum(k.values()),4))
except Exception as e:
    print('ERR', e, open('gpurun_out/${TAG}_err.txt').read()[-800:])"
  CSSM_LIB=$L timeout 300 python bench.py --workload c4 --steps 40 --no-cpu --chains 2 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   c4', round(j['value'],1), round(j['roofline']['us_per_observation'],2), [round(c['value'],1) for c in j['concurrent_chains']])"
  CSSM_LIB=$L timeout 300 python bench.py --workload c2 --obs 300 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=j['roofline']['kernel_ms_per_launch']; print('   c2 %.4g'%j['value'], {a:round(x,4) for a,x in k.items() if x})"
done

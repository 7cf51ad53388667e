#!/bin/bash
# tuning sweep of the launch knobs on the mid-size cloud (c2, 2^20) and neighbours; prints one line per variant
mkdir -p gpurun_out
run() { # label, env..., -- bench args
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  out=$(env "${envs[@]}" timeout 300 python bench.py --no-cpu --warmup 3 --steps 5 "$@" 2>/dev/null | tail -1)
  python - "$label" "$out" <<'PY'
import json,sys
try:
    j=json.loads(sys.argv[2]); r=j['roofline'] or {}
    print(sys.argv[1], '%.4g'%j['value'], 'ms/step %.3f'%j['ms_per_step'], {k:round(v,4) for k,v in (r.get('kernel_ms_per_launch') or {}).items() if v}, r.get('us_per_observation'))
except Exception as e:
    print(sys.argv[1],'ERR',e,sys.argv[2][:200])
PY
}
for n in 1048576 2097152 4194304; do
  run "N=$n base" -- --workload c2 --particles $n --obs 300
  run "N=$n ppt2" CSSM_K1_PPT=2 -- --workload c2 --particles $n --obs 300
  run "N=$n items2" CSSM_TILE_ITEMS=2 -- --workload c2 --particles $n --obs 300
  run "N=$n series" CSSM_SERIES_MAX_N=8388608 -- --workload c2 --particles $n --obs 300
  run "N=$n nopdl" CSSM_PDL=0 -- --workload c2 --particles $n --obs 300
done
run "target ppt2" CSSM_K1_PPT=2 -- --obs 300
run "target base" -- --obs 300

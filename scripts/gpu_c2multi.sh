for n in 0 2097152; do
CSSM_SERIES_MAX_N=$n timeout 300 python bench.py --workload c2 --no-cpu --no-extra --obs 300 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MAXN=$n c2 %.4g'%j['value'], j['roofline'].get('us_per_observation'), j['gpu_launches'], j.get('log_likelihood_mean'))"
done

#!/bin/bash
# GPU suite + the default, c2 and 2^22 bench lines (kernel times per launch)
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for args in "--obs 300" "--workload c2 --obs 300" "--workload c2 --particles 4194304 --obs 300"; do
timeout 300 python bench.py --no-cpu $args 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$args', '%.4g'%j['value'], {k:round(x,4) for k,x in j['roofline']['kernel_ms_per_launch'].items() if x})"
done

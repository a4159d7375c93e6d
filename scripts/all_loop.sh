#!/bin/bash
# kernel times of the all-sets step (100 000 nuclei, P = 64)
timeout 200 python bench.py --workload all --quick --no-cpu-baseline --no-e2e --steps 4 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('all ms/step', d['ms_per_step'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"

#!/bin/bash
# kernel times of BASELINE config 5 (P = 256, 500-vertex rings, all sets, 20 000 nuclei)
timeout 300 python bench.py --workload stress --quick --no-cpu-baseline --no-e2e --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('stress ms/step', round(d['ms_per_step'],2), {k:round(v['avg_ms'],2) for k,v in d['kernels'].items()}, d['clocks'])"

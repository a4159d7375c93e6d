#!/bin/bash
# quick loop for the colour kernels: parity tests of the colour set, then the headline workload
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "color or batch or chunk or hue or all_418 or odd or patch_sizes" 2>&1 | tail -4
timeout 120 python bench.py --quick --no-cpu-baseline --no-e2e --steps 30 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('color ms/step', round(d['ms_per_step'],4), {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})"

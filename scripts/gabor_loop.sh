#!/bin/bash
# quick loop for k_gabor work: parity tests of the gabor set, then the gabor workload alone
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gabor or all_418 or odd or patch_sizes" 2>&1 | tail -5
timeout 120 python bench.py --workload gabor --quick --no-cpu-baseline --no-e2e --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('gabor ms/step', d['ms_per_step'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"

#!/bin/bash
# quick loop for k_glrlm work: parity tests of the GLRLM set, then the glrlm workload alone
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "glrlm or all_418 or odd or patch_sizes or stress" 2>&1 | tail -5
timeout 120 python bench.py --workload glrlm --quick --no-cpu-baseline --no-e2e --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('glrlm ms/step', d['ms_per_step'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"

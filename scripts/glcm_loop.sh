#!/bin/bash
# quick loop for k_glcm64 work: parity tests of the GLCM set, then the glcm workload alone
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "glcm or grey or all_418 or odd or patch_sizes" 2>&1 | tail -5
timeout 120 python bench.py --workload glcm --nuclei 100000 --quick --no-cpu-baseline --no-e2e --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('glcm ms/step', d['ms_per_step'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"

#!/bin/bash
# quick loop for k_geom<raster,shape>: parity tests of the geometry set, then the geometry workload at 100 000 nuclei
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "shape or ellipse or masks or centroid or stress_p256_masks or all_418" 2>&1 | tail -3
timeout 120 python bench.py --workload shape --nuclei 100000 --tile 16384 --quick --no-cpu-baseline --no-e2e --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('shape ms/step', round(d['ms_per_step'],4), {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})"

#!/bin/bash
# ncu capture of the texture kernels (k_gabor, k_glcm64, k_glrlm) with source counters, 20 000 nuclei; plain run first.
set -u
O=gpurun_out
mkdir -p $O
Q="--workload all --nuclei 20000 --quick --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py $Q > $O/r2_tex_plain.json 2>$O/r2_tex_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_gabor|k_glcm|k_glrlm" -s 9 -c 3 \
    -o $O/r2_tex -f python bench.py $Q > $O/ncu_t.log 2>&1
tail -n 2 $O/ncu_t.log

#!/bin/bash
# ncu capture of k_geom<raster,shape> with source counters (geometry set, 100 000 nuclei); plain run first.
set -u
O=gpurun_out
mkdir -p $O
Q="--workload shape --nuclei 100000 --tile 16384 --quick --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py $Q > $O/r2_shape_plain.json 2>$O/r2_shape_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_geom" -s 4 -c 1 \
    -o $O/r2_shape3 -f python bench.py $Q > $O/ncu_g.log 2>&1
tail -n 2 $O/ncu_g.log

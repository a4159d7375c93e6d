#!/bin/bash
# Round-2 evidence for profiles/ (run through gpurun from the repo root; plain runs first, ncu afterwards).
set -u
O=gpurun_out
mkdir -p $O
Q="--quick --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
NFX_BENCH_EXACT_WARMUP=1 python bench.py $Q > $O/r2_plain_color.json 2>$O/r2_plain_color.err || exit 1
python bench.py --workload staged --steps 5 > $O/r2_bench_staged.json 2>/dev/null || exit 1
NFX_BENCH_EXACT_WARMUP=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_color_launches.csv \
    python bench.py $Q > $O/ncu_l.log 2>&1
NFX_BENCH_EXACT_WARMUP=1 ncu --set full --clock-control none --import-source on -k regex:"k_color|k_hue_batch|k_geom" -s 12 -c 4 \
    -o $O/r2_color_final -f python bench.py $Q > $O/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_gather|k_geom" -s 4 -c 2 \
    -o $O/r2_staged -f python bench.py --workload staged --steps 3 > $O/ncu_s.log 2>&1
tail -n 2 $O/ncu_c.log; tail -n 2 $O/ncu_s.log

#!/bin/bash
# Partial refresh after a change to the colour kernels only (run through gpurun from the repo root): the per-feature-set
# table, the default bench line, the stress line, then (ncu last) the launch list and the full capture of the colour kernels.
set -u
O=gpurun_out
mkdir -p $O
bash scripts/per_set_table.sh > $O/r1_set_table.md
python bench.py > $O/r1j_bench_color.json 2>/dev/null
python bench.py --workload stress --steps 3 --no-cpu-baseline > $O/r1j_bench_stress_p256.json 2>/dev/null
python bench.py --workload pipeline --steps 3 --warmup 1 > $O/r1j_bench_pipeline_color.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1j_color_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu1.log 2>&1
NFX_BENCH_EXACT_WARMUP=1 ncu --set full --clock-control none --import-source on -k regex:"k_color|k_hue_batch|k_geom" -s 9 -c 3 \
    -o $O/r1j_color -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu2.log 2>&1
cat $O/r1_set_table.md; tail -2 $O/ncu2.log

#!/bin/bash
# Re-measures everything profiles/ holds for this round on one B200 (run through gpurun from the repo root):
#   bash scripts/refresh_profiles.sh            -> files under gpurun_out/r1h_*
# Plain runs first; ncu only afterwards (a number printed under ncu is never a bench value).
set -u
O=gpurun_out
mkdir -p $O
python bench.py > $O/r1h_bench_color.json 2> $O/r1h_bench_color.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r1h_bench_color_reference_arm.json 2>> $O/r1h_bench_color.err
python bench.py --workload shape --no-cpu-baseline > $O/r1h_bench_shape.json 2>/dev/null
python bench.py --workload glcm --nuclei 200000 --tile 16384 --steps 5 --no-cpu-baseline > $O/r1h_bench_glcm_200k.json 2>/dev/null
python bench.py --workload glcm --steps 3 --no-cpu-baseline --no-e2e > $O/r1h_bench_glcm_1M.json 2>/dev/null
python bench.py --workload all --steps 5 --no-cpu-baseline > $O/r1h_bench_all.json 2>/dev/null
python bench.py --workload staged --steps 10 > $O/r1h_bench_staged.json 2>/dev/null
python bench.py --workload stress --steps 3 --no-cpu-baseline > $O/r1h_bench_stress_p256.json 2>/dev/null
python bench.py --workload pipeline --steps 3 --warmup 1 > $O/r1h_bench_pipeline_color.json 2>/dev/null
python bench.py --workload pipeline --sets all --nuclei 100000 --steps 2 --warmup 1 > $O/r1h_bench_pipeline_all.json 2>/dev/null
python scripts/bench_csv.py 200000 > $O/r1h_csv.jsonl 2>/dev/null
python scripts/bench_geojson.py 500000 > $O/r1h_geojson.jsonl 2>/dev/null
# launch list of the default bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1h_color_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu1.log 2>&1
# full capture of the colour kernels (same command) and of the texture / csv kernels
NFX_BENCH_EXACT_WARMUP=1 ncu --set full --clock-control none --import-source on -k regex:"k_color|k_hue_batch|k_geom" -s 9 -c 3 \
    -o $O/r1h_color -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu2.log 2>&1
NFX_BENCH_EXACT_WARMUP=1 ncu --set full --clock-control none --import-source on -k regex:"k_glcm|k_gabor|k_glrlm|k_geom" -s 12 -c 4 \
    -o $O/r1h_all -f python bench.py --workload all --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --nuclei 20000 > $O/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_csv" -c 2 \
    -o $O/r1h_csv -f python scripts/bench_csv.py 100000 > $O/ncu4.log 2>&1
tail -2 $O/ncu2.log $O/ncu3.log $O/ncu4.log
ls -la $O | tail -30

#!/bin/bash
# Final round-2 evidence (run through gpurun from the repo root): the whole GPU test suite, the default bench line and the
# reference arm, then launch list + ncu captures of the colour step and of the texture kernels (plain runs first).
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > $O/r2f_gputests.txt; cat $O/r2f_gputests.txt
python bench.py > $O/r2f_bench_default.json 2>$O/r2f_bench_default.err || exit 1
python bench.py --impl reference --steps 5 --warmup 3 > $O/r2f_bench_reference_arm.json 2>/dev/null
python __graft_entry__.py --smoke > $O/r2f_smoke.txt 2>&1; tail -2 $O/r2f_smoke.txt
python scripts/decode_rate.py > $O/r2_decode_rate.txt 2>/dev/null
bash scripts/refresh_r2.sh
bash scripts/refresh_tex.sh

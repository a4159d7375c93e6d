#!/bin/bash
# A/B of the colour kernels (run through gpurun from the repo root): old/new k_hue_batch x old/new k_color_warp.
O=gpurun_out
mkdir -p $O
for h in 1 2; do for c in 1 2; do
  NFX_HUE_V=$h NFX_CW_V=$c python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > $O/ab_h${h}_c${c}.json 2>$O/ab_h${h}_c${c}.err
  python - <<PY
import json
d=json.load(open("$O/ab_h${h}_c${c}.json"))
print("hue v$h cw v$c: ms/step %.4f" % d["ms_per_step"], {k: round(v["avg_ms"],4) for k,v in d["kernels"].items()})
PY
done; done

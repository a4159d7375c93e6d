#!/bin/bash
# BASELINE metric "nuclei/s per feature set ... vs ref CPU": one bench line per feature set on one B200, each with its own
# CPU baseline (the oracle on the box's host cores) and end-to-end number. Run through gpurun from the repo root.
set -u
O=gpurun_out
mkdir -p $O
python bench.py --workload shape --steps 20 > $O/r1_set_geometry.json 2>/dev/null
python bench.py --workload color --steps 20 > $O/r1_set_color.json 2>/dev/null
python bench.py --workload glcm --nuclei 200000 --tile 16384 --steps 5 > $O/r1_set_glcm.json 2>/dev/null
python bench.py --workload glrlm --steps 10 > $O/r1_set_glrlm.json 2>/dev/null
python bench.py --workload gabor --steps 5 > $O/r1_set_gabor.json 2>/dev/null
python bench.py --workload all --steps 5 > $O/r1_set_all.json 2>/dev/null
python - <<'PY'
import json
print("| feature set | columns | nuclei/s resident | nuclei/s end to end | CPU oracle nuclei/s (cores) | dominant kernel | frac of HBM peak |")
print("|---|---|---|---|---|---|---|")
for n in ("geometry", "color", "glcm", "glrlm", "gabor", "all"):
    try:
        d = json.loads(open(f"gpurun_out/r1_set_{n}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f"| {n} | failed: {e} |"); continue
    cb, e2e, rf = d.get("cpu_baseline") or {}, d.get("e2e") or {}, d.get("roofline") or {}
    print(f"| {n} | {d['config']['workload'][:40]} | {d['value']:.3g} | {e2e.get('value', 0):.3g} | {cb.get('value', 0):.4g} ({cb.get('cores')}) | {rf.get('kernel')} | {rf.get('frac', 0):.3f} |")
PY

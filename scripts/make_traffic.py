"""profiles/r2_traffic.json: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the captured kernels, tied to
the sources of those kernels (bench.TRAFFIC_SOURCES, bench.source_hash()): bench.py reports roofline.traffic only while the
workload's hash still matches.
usage: python scripts/make_traffic.py <workload> <nuclei> <patch> <ncu --page raw --csv file> [...more "workload nuclei patch file"]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

NAMES = {"k_geom<1, 0": "k_geom<raster>", "k_geom<1, 1": "k_geom<raster,shape>", "k_hue_batch": "k_hue_batch", "k_color_warp": "k_color_warp",
         "k_color(": "k_color", "k_glcm": "k_glcm", "k_glrlm": "k_glrlm", "k_gabor": "k_gabor", "k_gather": "k_gather"}
out = {"unit": "bytes per launch", "source": "ncu --set full --clock-control none (scripts/refresh_r2.sh)"}
a = sys.argv[1:]
for k in range(0, len(a), 4):
    wl, nuclei, P, path = a[k], int(a[k + 1]), int(a[k + 2]), a[k + 3]
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    jn, jr, jw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    acc = {}
    for r in data:
        nm = next((v for key, v in NAMES.items() if key in r[jn]), None)
        if nm is None:
            continue
        b = float(r[jr]) * scale[units[jr]] + float(r[jw]) * scale[units[jw]]
        acc.setdefault(nm, []).append(b)
    out.setdefault(wl, {"src_sha16": bench.source_hash(bench.TRAFFIC_SOURCES[wl]), "sources": list(bench.TRAFFIC_SOURCES[wl])})
    for nm, v in acc.items():
        out[wl][nm] = {"nuclei": nuclei, "patch": P, "dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v)}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

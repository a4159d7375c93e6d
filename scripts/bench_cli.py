"""Wall clock of the whole C++ driver on synthetic files: nfx-cli <geojson> <slide.tif> out.csv <sets>
(GeoJSON read + parse, TIFF read + nvJPEG decode, kernels, CSV formatted on the GPU and written to /dev/shm)."""
import json, os, subprocess, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "nuclei-feature-extraction_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from nfx import synth
from test_slide_tiff import write_tiled_tiff
import bench_geojson

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
sets = sys.argv[2:] or ["all"]
side = 8192
d = "/dev/shm/nfx_cli_bench"
os.makedirs(d, exist_ok=True)
img = synth.synth_tile(side, side, 3)
data, tiles = write_tiled_tiff(img, 256, big=True, quality=90)
open(f"{d}/slide.tif", "wb").write(data)
xy, off = synth.synth_polygons(n, side, side, 3)
open(f"{d}/cells.geojson", "wb").write(bench_geojson.make_text_from(xy, off))
cli = os.path.join(ROOT, "nuclei-feature-extraction_b200", "nfx-cli")
for rep in range(2):
    t0 = time.perf_counter()
    r = subprocess.run([cli, "-o", f"{d}/cells.geojson", f"{d}/slide.tif", f"{d}/out.csv", *sets], capture_output=True, text=True, env=dict(os.environ, NFX_CLI_TIMING="1"))
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr
print(r.stderr[-900:])
print(json.dumps({"cli": "nfx-cli", "nuclei": n, "sets": sets, "geojson_MB": round(os.path.getsize(f"{d}/cells.geojson") / 1e6, 1),
                  "tiff_MB": round(len(data) / 1e6, 1), "tiles": len(tiles), "csv_MB": round(os.path.getsize(f"{d}/out.csv") / 1e6, 1),
                  "wall_s": round(dt, 3), "nuclei_per_s": round(n / dt)}))

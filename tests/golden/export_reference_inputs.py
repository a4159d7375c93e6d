"""Writes the inputs of tests/golden/make_golden.py in the REFERENCE's own input formats (a GeoJSON FeatureCollection,
src/geojson.rs:8-24, and a PNG, src/main.rs:28-33) for tools/emit_reference_golden.rs:
    python tests/golden/export_reference_inputs.py  ->  tests/golden/reference/case.geojson, case.png
Coordinates are printed with repr(float(f32)) so that serde_json's f32 parse returns the same bits."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import golden_case  # noqa: E402

if __name__ == "__main__":
    from PIL import Image
    tile, xy, off = golden_case()
    out = os.path.join(HERE, "reference")
    os.makedirs(out, exist_ok=True)
    feats = []
    for i in range(len(off) - 1):
        ring = xy[off[i]:off[i + 1]]
        coords = [[float(np.float32(x)), float(np.float32(y))] for x, y in ring]
        xs, ys = ring[:, 0], ring[:, 1]
        feats.append({"type": "Feature", "bbox": [float(xs.min()), float(ys.min()), float(xs.max()), float(ys.max())],
                      "geometry": {"type": "Polygon", "coordinates": [coords]}, "properties": {}})
    with open(os.path.join(out, "case.geojson"), "w") as f:
        json.dump({"type": "FeatureCollection", "features": feats}, f)
    Image.fromarray(tile, "RGB").save(os.path.join(out, "case.png"))
    print("wrote", out)

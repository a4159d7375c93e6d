import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(R, "nuclei-feature-extraction_b200")); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, nfx
from cases import small_case
tile, rings = small_case()
xy, off = nfx.pack_polygons(rings)
ex = nfx.Extractor(0, 64, 100)
ex.upload_tile(tile); ex.upload_polygons(xy, off)
which = sys.argv[1]
if which == "gather":
    print(ex.gather_patches().sum())
elif which == "color":
    print(ex.extract(xy, off, ["color"])[2][:2])
elif which == "glcm":
    print(ex.extract(xy, off, ["glcm"])[2][:2, :8])

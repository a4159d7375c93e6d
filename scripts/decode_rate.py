"""Tile rate of the two slide decoders on a synthetic 8192 x 8192 tiled TIFF (1024 JPEG tiles of 256 x 256, 4:4:4 and 4:2:0)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nuclei-feature-extraction_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nfx
from nfx import synth
from test_slide_tiff import write_tiled_tiff
img = synth.synth_tile(8192, 8192, 3)
for sub in (0, 2):
    data, tiles = write_tiled_tiff(img, 256, True, False, sub)
    with nfx.Extractor(0) as ex:
        for fast in (False, True):
            ex.load_tiff(data, 0, fast=fast)
            t0 = time.perf_counter(); ex.load_tiff(data, 0, fast=fast); dt = time.perf_counter() - t0
            print(f"subsampling {sub} {'nvJPEG' if fast else 'exact '}: {len(tiles)/dt:8.0f} tiles/s  {img.nbytes/dt/1e9:5.2f} GB/s of pixels  ({len(data)/1e6:.0f} MB file, {os.cpu_count()} host threads)")

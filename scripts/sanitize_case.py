"""Smallest inputs that launch EVERY kernel of libnfx.so, for `compute-sanitizer --tool {racecheck,memcheck,initcheck,synccheck}`
(one tool per gpurun call; logs kept under profiles/r2_sanitizer_*.txt):
  P = 64  : k_geom<raster>, k_geom<raster,shape>, k_color_warp, k_hue_batch, k_hue_finalize, k_glcm64, k_glrlm, k_gabor,
            k_gather, k_expand, k_csv_measure / k_csv_write, k_pack_batch + k_geom<shape> (trait-level entry)
  P = 128 : k_color (two slabs), k_glcm_generic, tiled k_gabor + k_gabor_finalize
  P = 256 : k_glcm_large, k_glrlm<1024>
  P = 48  : partial mask words / partial slabs
plus a border case (windows partly outside the tile) and the slide-row copy between two contexts."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nuclei-feature-extraction_b200"))
import nfx  # noqa: E402
from nfx import synth  # noqa: E402


def run(P, n, side, batch, **kw):
    tile = synth.synth_tile(side, side, P)
    xy, off = synth.synth_polygons(n, side, side, P, patch=P, border_frac=0.25, **kw)
    with nfx.Extractor(0, P, batch) as ex:
        ex.upload_tile(tile)
        keys, cents, feats, names = ex.extract(xy, off, ["all"])
        ex.rasterize()
        ex.gather_patches()
        ex.debug_ellipses()
        ex.debug_glcm_counts(254, (1, -1))
        ex.debug_grey_levels(32)
        ex.csv_rows()
        # trait-level entry (reference Batch layout)
        patch_u8 = ex.gather_patches()
        masks_u8 = ex.rasterize()
        patchs = np.transpose(patch_u8.reshape(n, P, P, 3), (0, 3, 1, 2)).astype(np.float32) / np.float32(255.0)
        masks = masks_u8.reshape(n, 1, P, P).astype(np.float32)
        rings = [xy[off[i]:off[i + 1]] - xy[off[i]:off[i + 1]].mean(0) for i in range(n)]
        ex.compute_features_batched(nfx.FS_ALL, cents, rings, patchs, masks)
    print(f"P={P}: {n} nuclei, {feats.shape[1]} columns, finite share {np.isfinite(feats).mean():.3f}", flush=True)
    return tile


if __name__ == "__main__":
    t = run(64, 45, 256, 20)
    run(48, 9, 160, 4, r0_range=(4.0, 18.0))
    run(128, 5, 300, 3, r0_range=(12.0, 50.0), v_range=(40, 90))
    run(256, 3, 420, 2, r0_range=(40.0, 110.0), v_range=(500, 500), harmonics=(3, 7, 19))
    with nfx.Extractor(0, 64, 10) as a, nfx.Extractor(0, 64, 10) as b:
        a.upload_tile(t)
        a.sync()
        b.slide_alloc(256, 256)
        b.write_tile(np.ascontiguousarray(t[:100]), 0, 0)
        b.slide_copy_rows(a, 100, 156)
        b.sync()
    print("sanitize_case done", flush=True)

"""Regenerates tests/golden/oracle_small.npz: oracle outputs on a small seeded case.

PARITY UNPINNED: the reference (Rust + un-vendored tch-utils / geometric-features) cannot be built
or imported here, so these vectors come from oracle/nfx_oracle.py itself (see oracle/SPEC.md). They
pin the oracle against drift and give the GPU tests a fixture that does not need the oracle's
slow paths.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "nuclei-feature-extraction_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import nfx_oracle as o  # noqa: E402
from nfx import pack_polygons, synth  # noqa: E402


def golden_case():
    tile = synth.synth_tile(256, 256, 3)
    xy, off = synth.synth_polygons(48, 256, 256, 3, border_frac=0.15)
    return tile, xy, off


if __name__ == "__main__":
    tile, xy, off = golden_case()
    rings = synth.rings_of(xy, off)
    keys, cents, feats, names = o.extract(rings, tile, ["all"], 64, 20)
    masks = np.stack([o.polygon_mask(64, 64, o.preprocess_polygon(r)[1].astype(np.float64)) for r in rings])
    patches = np.stack([o.gather_patch_u8(tile, c, 64) for c in cents])
    np.savez_compressed(os.path.join(HERE, "oracle_small.npz"),
                        tile_crc=np.array([int(tile.astype(np.uint64).sum())]), xy=xy, off=off, centroids=cents,
                        keys=np.array(keys), features=feats, names=np.array(names),
                        masks=np.packbits(masks, axis=-1), patches_sum=patches.reshape(len(rings), -1).sum(1))
    print("wrote", os.path.join(HERE, "oracle_small.npz"), feats.shape)

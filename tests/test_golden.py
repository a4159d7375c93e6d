"""The oracle against the committed golden vectors (tests/golden/make_golden.py regenerates them).
PARITY UNPINNED: the vectors come from the oracle itself (the reference cannot run here), so this
test guards the oracle -- the GPU tests' yardstick -- against drift."""
import os

import numpy as np

import nfx_oracle as o
from nfx import synth

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_reproduces_golden_vectors():
    from golden.make_golden import golden_case
    g = np.load(os.path.join(HERE, "golden", "oracle_small.npz"))
    tile, xy, off = golden_case()
    assert int(tile.astype(np.uint64).sum()) == int(g["tile_crc"][0])
    assert np.array_equal(xy, g["xy"]) and np.array_equal(off, g["off"])
    rings = synth.rings_of(xy, off)
    keys, cents, feats, names = o.extract(rings, tile, ["all"], 64, 20)
    assert list(g["names"]) == names and list(g["keys"]) == keys
    assert np.array_equal(cents, g["centroids"])
    assert np.allclose(feats, g["features"], rtol=1e-5, atol=1e-7, equal_nan=True)
    masks = np.stack([o.polygon_mask(64, 64, o.preprocess_polygon(r)[1].astype(np.float64)) for r in rings])
    assert np.array_equal(np.packbits(masks, axis=-1), g["masks"])

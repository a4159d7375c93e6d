"""GPU tests at BASELINE.json's full size for the headline configuration (configs[1]: 100 000 nuclei, 64x64
windows, batch 100, one 16 384^2 tile), through properties that do not need the oracle on every nucleus:
whole reference chunks sampled and compared with the oracle, closed forms on a constant tile, and bit-exact
invariances (chunk permutation, a second context)."""
import numpy as np
import pytest
import torch

import nfx
import nfx_oracle as o
from nfx import synth
from parity_checks import check_all_columns, check_color, check_shape

pytestmark = pytest.mark.gpu

N, SIDE, P, B = 100_000, 16384, 64, 100


@pytest.fixture(scope="module")
def full(libnfx):
    tile = synth.synth_tile(SIDE, SIDE, 2)
    xy, off = synth.synth_polygons(N, SIDE, SIDE, 2)
    ex = nfx.Extractor(0, P, B)
    ex.upload_tile(tile)
    keys, cents, feats, names = ex.extract(xy, off, ["geometry", "color"])
    yield dict(tile=tile, xy=xy, off=off, ex=ex, cents=cents, feats=feats, names=names, keys=keys)
    ex.close()


def test_sampled_reference_chunks_match_oracle(full):
    rng = np.random.default_rng(0)
    chunks = sorted(rng.choice(N // B, size=4, replace=False).tolist() + [0, N // B - 1])
    names = full["names"]
    assert names == o.SHAPE_COLUMNS + o.COLOR_COLUMNS
    for k in chunks:
        lo, hi = k * B, (k + 1) * B
        rings = [full["xy"][full["off"][i]:full["off"][i + 1]] for i in range(lo, hi)]
        cents, polys, patches, masks = o.load_image_dataset(rings, full["tile"], P)
        assert [o.centroid_key(c) for c in cents] == full["keys"][lo:hi]
        want_s, dbg = o.shape_features(polys, masks, return_debug=True)
        check_shape(full["feats"][lo:hi, :12], want_s, dbg, names[:12])
        want_c = o.color_features(patches.clone(), masks)
        check_color(full["feats"][lo:hi, 12:], want_c, names[12:], patches, masks, B)


def test_constant_tile_closed_forms(full):
    """Every window of a constant tile holds one colour: means are that colour's conversion, stds are 0, and the
    co-occurrence matrix has one cell (contrast 0, ASM 1, entropy 0) for 100 000 nuclei."""
    rgb = (183, 96, 152)
    tile = np.empty((4096, 4096, 3), np.uint8)
    tile[:] = rgb
    xy, off = synth.synth_polygons(N, 4096, 4096, 5, border_frac=0.0)
    with nfx.Extractor(0, P, B) as ex:
        ex.upload_tile(tile)
        keys, cents, feats, names = ex.extract(xy, off, ["color", "glcm"])
    px = torch.tensor(rgb, dtype=torch.float32).view(1, 3, 1, 1) / 255.0
    hsv, hed = o.hsv_from_rgb(px).flatten().numpy(), o.hed_from_rgb(px).flatten().numpy()
    want = {"mean_r": rgb[0] / 255, "mean_g": rgb[1] / 255, "mean_b": rgb[2] / 255, "mean_h": hsv[0], "mean_s": hsv[1], "mean_v": hsv[2],
            "mean_haematoxylin": hed[0], "mean_eosin": hed[1], "mean_dab": hed[2]}
    ok = ~np.isnan(feats[:, names.index("mean_r")])           # (no empty masks expected, but do not assume)
    assert ok.sum() > 0.999 * N
    for nm, w in want.items():
        col = feats[ok, names.index(nm)]
        assert np.allclose(col, w, rtol=1e-5, atol=(2e-3 if nm == "mean_h" else 1e-6)), (nm, w, col[:4])
    for nm in names:
        if nm.startswith("std_"):
            assert np.abs(feats[ok, names.index(nm)]).max() <= (2e-3 if nm == "std_h" else 1e-6), nm   # one-pass f32 rounding of equal values
    for L in (32, 64, 128, 254):
        for d in ("0_1", "1_1", "1_0", "1_-1"):
            g = lambda f: feats[ok, names.index(f"{f}_{d}_{L}")]
            assert np.all(g("contrast") == 0) and np.all(g("dissimilarity") == 0) and np.all(g("difference_variance") == 0)
            assert np.all(g("angular_second_moment") == 1.0) and np.all(g("entropy") == 0.0) and np.all(g("sum_entropy") == 0.0)
            assert np.allclose(g("inverse_difference_moment"), 1.0, atol=1e-6) and np.all(g("sum_of_squares") == 0.0)
            assert np.all(np.isnan(g("correlation"))) and np.all(np.isnan(g("information_measure_correlation1")))   # 0/0 like the reference
            assert np.all(g("information_measure_correlation2") == 0.0)


def test_chunk_permutation_is_bit_exact(full):
    """Reference chunks are independent units (mean_h couples only nuclei of one chunk): moving whole chunks
    permutes the rows bit for bit, and a second context (another CUDA stream, fresh buffers) reproduces them."""
    rng = np.random.default_rng(1)
    perm = rng.permutation(N // B)
    off, xy = full["off"], full["xy"]
    idx = (perm[:, None] * B + np.arange(B)[None, :]).reshape(-1)
    lens = np.diff(off)
    noff = np.concatenate([[0], np.cumsum(lens[idx])]).astype(np.int64)
    nxy = np.concatenate([xy[off[i]:off[i + 1]] for i in idx], 0)
    ex = full["ex"]
    keys, cents, feats, names = ex.extract(nxy, noff, ["geometry", "color"])
    assert feats.tobytes() == full["feats"][idx].tobytes() and cents.tobytes() == full["cents"][idx].tobytes()
    with nfx.Extractor(0, P, B) as e2:
        e2.upload_tile(full["tile"])
        k2, c2, f2, _ = e2.extract(xy, off, ["geometry", "color"])
    assert f2.tobytes() == full["feats"].tobytes() and k2 == full["keys"]


def test_sampled_chunks_of_the_texture_sets_match_oracle(full):
    """GLCM, GLRLM and Gabor at the full size (100 000 nuclei in one launch): whole reference chunks sampled against the
    oracle, every column."""
    ex = full["ex"]
    keys, cents, feats, names = ex.extract(full["xy"], full["off"], ["texture"])
    assert keys == full["keys"]
    rng = np.random.default_rng(3)
    for k in sorted(rng.choice(N // B, size=2, replace=False).tolist() + [N // B - 1]):
        lo, hi = k * B, (k + 1) * B
        rings = [full["xy"][full["off"][i]:full["off"][i + 1]] for i in range(lo, hi)]
        check_all_columns(feats[lo:hi], names, rings, full["tile"], P, B, sets=["glcm", "glrlm", "gabor"])


class _PeriodicSlide:
    """What the oracle's gather needs of an image (shape, 2-D slicing) for a slide that repeats one block: the 5 GB slide
    of the test below never has to exist on the host."""

    def __init__(self, block, side):
        self.block, self.shape = block, (side, side, 3)

    def __getitem__(self, idx):
        rs, cs = idx
        n = self.block.shape[0]
        return self.block[np.ix_(np.arange(rs.start, rs.stop) % n, np.arange(cs.start, cs.stop) % n)]


def test_slide_path_sampled_chunks_match_oracle(libnfx):
    """BASELINE config 4 mechanics at a size the test box holds quickly: a 40 960^2 slide (5 GB in HBM) written as 8192^2
    tiles from two pinned staging buffers, 400 000 nuclei over all of it, all five sets; whole chunks sampled against the oracle
    (which reads its windows from a periodic view of the staging block)."""
    side, T, n = 40960, 8192, 400_000
    block = synth.synth_tile(4096, 4096, 4)
    stage = [nfx.pinned_empty((T, T, 3), np.uint8) for _ in range(2)]
    for b in stage:
        for r in range(0, T, 4096):
            for c in range(0, T, 4096):
                b[r:r + 4096, c:c + 4096] = block
    xy, off = synth.synth_polygons_pool(n, side, side, 4)
    with nfx.Extractor(0, P, B) as ex:
        ex.slide_alloc(side, side)
        k = 0
        for y in range(0, side, T):
            for x in range(0, side, T):
                ex.write_tile(stage[k & 1], x, y)
                k += 1
        keys, cents, feats, names = ex.extract(xy, off, ["all"])
    assert feats.shape == (n, 418)
    slide = _PeriodicSlide(block, side)
    rng = np.random.default_rng(5)
    for k in sorted(rng.choice(n // B, size=2, replace=False).tolist()):
        lo, hi = k * B, (k + 1) * B
        rings = [xy[off[i]:off[i + 1]] for i in range(lo, hi)]
        assert keys[lo:hi] == [o.centroid_key(o.preprocess_polygon(r)[0]) for r in rings]
        check_all_columns(feats[lo:hi], names, rings, slide, P, B)       # whole chunks (mean_h couples them), all 418 columns

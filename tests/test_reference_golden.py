"""Pins the oracle to the REFERENCE's own results -- when someone has produced them.

oracle/SPEC.md section B rules restate un-vendored crates (tch-utils@d1c10c0, geometric-features@163ae81) from
their call sites only ("parity unpinned"). tools/emit_reference_golden.rs, built inside a checkout of the
reference by anyone who has cargo + libtorch, dumps the reference's masks, colour conversions, GLCM / GLRLM
matrices, Gabor responses, polygon geometry and the full 418-column feature matrix for the inputs of
tests/golden/make_golden.py (tests/golden/export_reference_inputs.py writes them as GeoJSON + PNG). Drop the
reference_*.npy files into tests/golden/reference/ and these tests compare the oracle against them; while the
files are absent every test here is SKIPPED (the image has no rustc, so they could not be generated in-tree).

A mismatch in one of the three highest-risk rules is a one-flag fix: see RULE_FLAGS below / oracle.RULES and
nfx_config.rule_flags (include/nfx.h)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "golden", "reference")
sys.path.insert(0, os.path.join(HERE, "golden"))

import nfx_oracle as o  # noqa: E402
from tolerances import mismatches  # noqa: E402


def ref(name):
    path = os.path.join(REF, f"reference_{name}.npy")
    if not os.path.exists(path):
        pytest.skip(f"{os.path.relpath(path, os.path.dirname(HERE))} not present: run tools/emit_reference_golden.rs in a "
                    "checkout of the reference (needs cargo + libtorch, absent from this image)")
    return np.load(path)


@pytest.fixture(scope="module")
def case():
    from make_golden import golden_case
    from nfx import synth
    tile, xy, off = golden_case()
    rings = synth.rings_of(xy, off)
    cents, polys, patches, masks = o.load_image_dataset(rings, tile, 64)
    return tile, rings, cents, polys, patches, masks


def test_polygon_raster_is_the_reference_rule(case):
    want = ref("masks")[:, 0] != 0
    got = case[5][:, 0].numpy() != 0
    if not np.array_equal(got, want):
        # try the other position of the sample-offset switch before failing, to say which flag fixes it
        with o.rules(raster_offset=0.5 if o.RULES["raster_offset"] == 0.0 else 0.0):
            alt = np.stack([o.polygon_mask(64, 64, p.astype(np.float64)) for p in case[3]])
        hint = " -- matches with the other raster_offset: flip NFX_RULE_RASTER_PIXEL_CENTRE" if np.array_equal(alt, want) else ""
        raise AssertionError(f"{int((got != want).sum())} mask pixels differ from the reference{hint}")


def test_patches_are_the_reference_gather(case):
    assert np.array_equal(case[4].numpy(), ref("patches"))


def test_hsv_hed_conversions(case):
    p4 = case[4][:4]
    for name, fn in (("hsv", o.hsv_from_rgb), ("hed", o.hed_from_rgb)):
        want = ref(name)
        got = fn(p4).numpy()
        scale = np.array([360.0, 1.0, 1.0] if name == "hsv" else [1.0, 1.0, 1.0]).reshape(1, 3, 1, 1)
        assert np.all(np.abs(got - want) <= 1e-4 * scale), f"{name}: max |diff| {np.abs(got - want).max()}"


@pytest.mark.parametrize("L,off,tag", [(32, (0, 1), "32_0_1"), (64, (1, 1), "64_1_1"), (128, (1, 0), "128_1_0"),
                                       (254, (1, -1), "254_1_-1"), (254, (0, 1), "254_0_1")])
def test_glcm_matrix(case, L, off, tag):
    want = ref(f"glcm_{tag}")
    grey = o.grey_scale(case[4][:4])
    got = o.glcm(grey, off, L, case[5][:4]).numpy()
    assert got.shape == want.shape, f"shape {got.shape} vs reference {want.shape}"
    # normalised matrices: entries are counts / total, so 1e-6 absolute separates a one-count difference (>= 1/8192)
    assert np.allclose(got, want, rtol=0, atol=1e-6, equal_nan=True), f"max |diff| {np.nanmax(np.abs(got - want))}"


@pytest.mark.parametrize("d,tag", [((1, 0), "1_0"), ((1, 1), "1_1"), ((0, 1), "0_1"), ((-1, 1), "-1_1")])
def test_glrlm_matrix(case, d, tag):
    want = ref(f"glrlm_{tag}")
    grey = o.grey_scale(case[4][:4])
    got = o.glrlm_counts(grey, o.GLRLM_LEVELS, o.GLRLM_MAX_LENGTH, d, case[5][:4]).numpy()
    assert np.array_equal(got.astype(np.float32), want.astype(np.float32))


def test_gabor_responses(case):
    want = ref("gabor")
    got = o.apply_gabor_filter(o.grey_scale(case[4][:2])).numpy()
    assert got.shape == want.shape
    assert np.allclose(got, want, rtol=1e-4, atol=2e-4), f"max |diff| {np.abs(got - want).max()}"


def test_ellipse_rule():
    want = ref("ellipse_fixed")[0] != 0
    got = o.ellipse_mask(64, 64, (1.25, -2.5), (17.0, 9.5), 0.6)
    assert np.array_equal(got, want), f"{int((got != want).sum())} ellipse pixels differ"


def test_polygon_geometry(case):
    want = ref("polygon_geometry")          # area, perimeter, equivalent_perimeter, compacity, hull area, hull perimeter, deviation
    got = np.array([[g[k] for k in ("area", "perimeter", "equivalent_perimeter", "compacity", "convex_hull_area",
                                     "convex_perimeter", "convex_deffect")]
                    for g in (o.polygon_geometry(p.astype(np.float64)) for p in case[3])])
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9), f"max rel diff {np.abs(got / want - 1).max()}"


def test_all_418_columns(case):
    want = ref("features")
    names = open(os.path.join(REF, "reference_names.txt")).read().split("\n")
    tile, rings = case[0], case[1]
    keys, cents, got, got_names = o.extract(rings, tile, ["all"], 64, 20)
    assert got_names == names
    wkeys = open(os.path.join(REF, "reference_keys.txt")).read().split("\n")
    assert keys == wkeys
    col = 0
    bad = []
    for s in o.FLAT_ORDER:
        k = len(o.SET_COLUMNS[s])
        bad += mismatches(got[:, col:col + k], want[:, col:col + k], names[col:col + k], s)
        col += k
    assert not bad, f"{len(bad)} cells outside 1e-4: {bad[:8]}"

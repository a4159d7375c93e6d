"""Known-answer tests the oracle itself must satisfy. The reference has no tests or
golden vectors, so the oracle is pinned on analytic cases, on independent implementations
(cv2, colorsys) and on torch.linalg.eig for the LAPACK convention."""
import colorsys
import math

import numpy as np
import pytest
import torch

import nfx_oracle as o


def test_centroid_is_sequential_f32_mean_including_closing_duplicate():
    ring = np.array([[0.1, 0.2], [1e7, 3.3], [0.7, -2.5], [0.1, 0.2]], np.float32)
    c, cp = o.preprocess_polygon(ring)
    ax = np.float32(0)
    ay = np.float32(0)
    for p in ring:
        ax = np.float32(ax + p[0])
        ay = np.float32(ay + p[1])
    assert c[0] == np.float32(ax / np.float32(4)) and c[1] == np.float32(ay / np.float32(4))
    assert np.array_equal(cp, (ring - c).astype(np.float32))


def test_rectangle_mask_half_open_rule():
    # centred rectangle [-8,12) x [-6,4): sample points are integers -> half-open on both axes
    pts = np.array([[-8, -6], [12, -6], [12, 4], [-8, 4], [-8, -6]], np.float64)
    m = o.polygon_mask(64, 64, pts)
    rr, cc = np.nonzero(m)
    assert rr.min() == 26 and rr.max() == 35 and cc.min() == 24 and cc.max() == 43
    assert m.sum() == 200


def test_mask_matches_winding_number_on_simple_polygons():
    """Independent algorithm: for a simple (non self-intersecting) polygon the even-odd rule equals
    winding number != 0, computed here by summing signed angles."""
    rng = np.random.default_rng(0)
    for _ in range(12):
        V = rng.integers(5, 30)
        th = np.sort(rng.uniform(0, 2 * np.pi, V))
        r = rng.uniform(8, 28, V)
        pts = np.stack([r * np.cos(th), r * np.sin(th)], 1) + rng.uniform(-3, 3, 2)
        m = o.polygon_mask(64, 64, pts)
        X, Y = np.meshgrid(np.arange(64) - 32.0, np.arange(64) - 32.0)
        a = pts[None, None] - np.stack([X, Y], -1)[:, :, None, :]          # [64,64,V,2]
        b = np.roll(a, -1, axis=2)
        ang = np.arctan2(a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0], (a * b).sum(-1))
        wn = np.rint(ang.sum(-1) / (2 * np.pi)).astype(int)
        # ignore sample points within 1e-6 of an edge (boundary convention differs by design)
        d = np.abs(a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]) / np.maximum(np.linalg.norm(b - a, axis=-1), 1e-30)
        safe = d.min(-1) > 1e-6
        assert np.array_equal(m[safe], (wn != 0)[safe])


def test_even_odd_rule_on_bow_tie_and_unclosed_ring():
    bow = np.array([[-15, -15], [15, 15], [15, -15], [-15, 15]], np.float64)      # unclosed, self-intersecting
    m = o.polygon_mask(64, 64, bow)
    assert m[32 - 10, 32 + 12] and m[32 + 10, 32 + 12]          # right lobe: 0 < x < 15, |y| < x
    assert m[32, 32 - 10]                                        # left lobe (closing edge x = -15)
    assert not m[32 - 12, 32 + 2] and not m[32 + 12, 32 - 2]    # above / below the crossing point
    closed = np.vstack([bow, bow[:1]])
    assert np.array_equal(m, o.polygon_mask(64, 64, closed))    # closing duplicate adds a null edge


def test_patch_window_truncates_toward_zero_and_pads():
    img = np.arange(40 * 50 * 3, dtype=np.uint32).reshape(40, 50, 3).astype(np.uint8)
    # cy - 8 = -2.5 -> top = -2 (not -3), bottom = trunc(13.5) = 13 -> 13 rows pasted at offset 2, row 15 zero
    p = o.gather_patch_u8(img, (25.0, 5.5), 16)
    assert o.patch_window((25.0, 5.5), 16) == (-2, 17, 13, 33)
    assert (p[:2] == 0).all() and (p[15] == 0).all()
    assert np.array_equal(p[2:15], img[0:13, 17:33])
    # fully interior
    q = o.gather_patch_u8(img, (20.2, 20.9), 16)
    assert np.array_equal(q, img[12:28, 12:28])
    # f32 patch is exactly u8/255
    t = o.gather_patch(img, (20.2, 20.9), 16)
    assert torch.equal(t, torch.from_numpy(q).permute(2, 0, 1).float() / 255.0)


def test_eig_closed_form_matches_lapack_convention():
    rng = np.random.default_rng(1)
    for t in range(3000):
        pts = rng.integers(0, 64, size=(rng.integers(2, 300), 2)).astype(np.float32)
        if t % 5 == 0:
            pts[:, 1] = pts[0, 1]
        c = pts - pts.mean(0)
        cov = torch.from_numpy((c.T @ c / len(c)).astype(np.float32))
        w, v = torch.linalg.eig(cov)
        l0, l1, V = o.eig2x2_lapack(cov[0, 0].item(), cov[0, 1].item(), cov[1, 1].item())
        scale = max(abs(w.real).max().item(), 1e-30)
        assert abs(w[0].real.item() - l0) <= 2e-6 * scale and abs(w[1].real.item() - l1) <= 2e-6 * scale
        assert np.allclose(v.real.numpy(), V, atol=2e-6)


def test_shape_features_of_axis_aligned_rectangle():
    pts = np.array([[-10, -5], [10, -5], [10, 5], [-10, 5], [-10, -5]], np.float32)
    m = torch.from_numpy(o.polygon_mask(64, 64, pts.astype(np.float64)).astype(np.float32))[None, None]
    f = o.shape_features([pts], m)[0]
    d = dict(zip(o.SHAPE_COLUMNS, f))
    assert d["area"] == 200 and d["perimeter"] == 60 and d["convex_hull_area"] == 200
    assert d["convex_perimeter"] == 60 and d["convex_deffect"] == 0
    assert math.isclose(d["equivalent_perimeter"], 2 * math.sqrt(math.pi * 200), rel_tol=1e-12)
    assert math.isclose(d["compacity"], 4 * math.pi * 200 / 3600, rel_tol=1e-12)
    # mask = 20 cols x 10 rows: var_c = (20^2-1)/12, var_r = (10^2-1)/12; axes = 2*2*sqrt(var)... = 2 sqrt(l)*2/2
    assert math.isclose(d["major_axis"], 2 * math.sqrt((400 - 1) / 12), rel_tol=1e-5)
    assert math.isclose(d["minor_axis"], 2 * math.sqrt((100 - 1) / 12), rel_tol=1e-5)
    # cov = [[var_r, 0],[0, var_c]], b == 0 -> V = I, l0 = var_r < l1 -> row 1 = (0,1) -> atan2(0,1) = 0
    assert d["orientation"] == 0.0
    assert math.isclose(d["eccentricity"], math.sqrt(1 - 99 / 399), rel_tol=1e-5)


def test_polygon_geometry_against_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    for _ in range(30):
        V = rng.integers(6, 40)
        th = np.sort(rng.uniform(0, 2 * np.pi, V))
        r = rng.uniform(5, 30, V)
        pts = np.stack([r * np.cos(th), r * np.sin(th)], 1).astype(np.float32)
        g = o.polygon_geometry(pts.astype(np.float64))
        assert math.isclose(g["area"], cv2.contourArea(pts), rel_tol=1e-4)
        assert math.isclose(g["perimeter"], cv2.arcLength(pts, True), rel_tol=1e-4)
        hull = cv2.convexHull(pts)
        assert math.isclose(g["convex_hull_area"], cv2.contourArea(hull), rel_tol=1e-4)
        assert math.isclose(g["convex_perimeter"], cv2.arcLength(hull, True), rel_tol=1e-4)


def test_regular_ngon_closed_forms():
    n, R = 12, 20.0
    th = 2 * np.pi * np.arange(n) / n
    pts = np.stack([R * np.cos(th), R * np.sin(th)], 1)
    g = o.polygon_geometry(pts)
    assert math.isclose(g["area"], 0.5 * n * R * R * math.sin(2 * math.pi / n), rel_tol=1e-12)
    assert math.isclose(g["perimeter"], 2 * n * R * math.sin(math.pi / n), rel_tol=1e-12)
    assert abs(g["convex_deffect"]) < 1e-12


def test_hsv_matches_colorsys():
    rng = np.random.default_rng(3)
    u8 = rng.integers(0, 256, size=(200, 3)).astype(np.uint8)
    u8[:5] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [10, 10, 200], [7, 200, 200]]
    rgb = torch.from_numpy(u8.astype(np.float32) / np.float32(255)).T.reshape(1, 3, 1, 200)
    hsv = o.hsv_from_rgb(rgb)[0, :, 0].T.numpy()
    for k in range(200):
        h, s, v = colorsys.rgb_to_hsv(*(u8[k] / 255.0))
        assert abs(hsv[k, 0] - 360 * h) < 2e-3 or abs(abs(hsv[k, 0] - 360 * h) - 360) < 2e-3
        assert abs(hsv[k, 1] - s) < 1e-6 and abs(hsv[k, 2] - v) < 1e-6


def test_hed_is_the_ruifrok_deconvolution():
    rgb_from_hed = np.array([[0.65, 0.70, 0.29], [0.07, 0.99, 0.11], [0.27, 0.57, 0.78]])
    conc = np.array([0.4, 0.2, 0.05])                    # stain concentrations
    od = conc @ rgb_from_hed                             # optical density per RGB channel (log_1e-6 units)
    rgb = np.exp(od * math.log(1e-6))
    t = torch.tensor(rgb, dtype=torch.float32).reshape(1, 3, 1, 1)
    hed = o.hed_from_rgb(t).reshape(3).numpy()
    assert np.allclose(hed, conc, atol=2e-6)
    white = o.hed_from_rgb(torch.ones(1, 3, 1, 1)).reshape(3).numpy()
    assert np.allclose(white, 0)


def test_constant_colour_patch_statistics_and_batch_coupled_hue():
    P, N = 16, 3
    cols = np.array([[200, 40, 40], [40, 200, 40], [40, 40, 200]], np.float32) / np.float32(255)
    patches = torch.from_numpy(cols).reshape(N, 3, 1, 1).expand(N, 3, P, P).contiguous()
    masks = torch.zeros(N, 1, P, P)
    masks[:, :, 4:12, 4:12] = 1
    f = o.color_features(patches.clone(), masks)
    d = {c: f[:, j] for j, c in enumerate(o.COLOR_COLUMNS)}
    assert np.allclose(d["std_r"], 0, atol=1e-7) and np.allclose(d["std_s"], 0, atol=1e-7)
    assert np.allclose(d["mean_r"], cols[:, 0], atol=1e-7)
    # hues are 0, 120, 240 degrees: every nucleus sees the SUM over the batch (color.rs:50-51 broadcast)
    # -> resultant ~ 0 for all three, so mean_h is the same (ill-conditioned) value for every row
    assert np.allclose(d["mean_h"], d["mean_h"][0], atol=1e-3)
    # with a single-patch batch the hue mean is that patch's own hue
    f1 = o.color_features(patches[1:2].clone(), masks[1:2])
    assert abs(f1[0, o.COLOR_COLUMNS.index("mean_h")] - 120.0) < 1e-3


def test_glcm_counts_of_a_two_level_checkerboard():
    P = 8
    g = (np.indices((P, P)).sum(0) % 2).astype(np.float32) * np.float32(0.9)          # levels 0 and floor(.9*L)
    grey = torch.from_numpy(g)[None, None]
    masks = torch.ones(1, 1, P, P)
    L = 32
    hi = int(math.floor(np.float32(0.9) * L))
    G = o.glcm_counts(grey, (0, 1), L, masks)[0].numpy()
    assert G[0, hi] == G[hi, 0] == P * (P - 1) and G.sum() == 2 * P * (P - 1)
    G = o.glcm_counts(grey, (1, 1), L, masks)[0].numpy()
    assert G[0, 0] + G[hi, hi] == 2 * (P - 1) ** 2 and G[0, hi] == 0
    # masked: only pairs with both pixels inside the mask count
    masks[:, :, :, 4:] = 0
    G = o.glcm_counts(grey, (0, 1), L, masks)[0].numpy()
    assert G.sum() == 2 * P * 3


def test_glcm_features_known_values():
    # p = [[.5, 0],[0, .5]] (perfectly correlated two-level texture)
    p = torch.tensor([[[0.5, 0.0], [0.0, 0.5]]])
    f = dict(zip(o.GLCM_FEATURES, o.glcm_features(p)[0]))
    assert math.isclose(f["correlation"], 1.0, rel_tol=1e-6) and f["contrast"] == 0 and f["dissimilarity"] == 0
    assert math.isclose(f["entropy"], math.log(2), rel_tol=1e-6)
    assert math.isclose(f["angular_second_moment"], 0.5, rel_tol=1e-6)
    assert math.isclose(f["sum_average"], 1.0, rel_tol=1e-6) and math.isclose(f["sum_variance"], 1.0, rel_tol=1e-6)
    assert math.isclose(f["sum_of_squares"], 0.25, rel_tol=1e-6)
    assert math.isclose(f["inverse_difference_moment"], 1.0, rel_tol=1e-6)
    assert math.isclose(f["information_measure_correlation1"], -1.0, rel_tol=1e-6)
    assert math.isclose(f["information_measure_correlation2"], math.sqrt(1 - math.exp(-2 * math.log(2))), rel_tol=1e-6)
    # HXY1 == HXY2 == HX + HY for any symmetric p (used by the kernel's factorisation)
    rng = np.random.default_rng(4)
    a = rng.random((1, 6, 6))
    a = a + a.transpose(0, 2, 1)
    a[0, 2] = 0
    a[0, :, 2] = 0
    a /= a.sum()
    f = dict(zip(o.GLCM_FEATURES, o.glcm_features(torch.tensor(a, dtype=torch.float32))[0]))
    px = a[0].sum(1)
    hx = -(px[px > 0] * np.log(px[px > 0])).sum()
    hxy = -(a[a > 0] * np.log(a[a > 0])).sum()
    assert math.isclose(f["information_measure_correlation1"], (hxy - 2 * hx) / hx, rel_tol=1e-4)
    # empty mask -> 0/0 -> NaN everywhere
    assert np.isnan(o.glcm_features(torch.full((1, 4, 4), float("nan")))).all()


def test_empty_mask_gives_nan_like_the_reference():
    m = torch.zeros(1, 1, 64, 64)
    pts = np.array([[0.2, 0.2], [0.6, 0.2], [0.6, 0.6], [0.2, 0.2]], np.float32)
    f = dict(zip(o.SHAPE_COLUMNS, o.shape_features([pts], m)[0]))
    assert np.isnan(f["major_axis"]) and np.isnan(f["minor_axis"]) and np.isnan(f["orientation"])
    assert np.isnan(f["eliptic_deviation"]) and f["area"] > 0


def test_key_string_is_rust_display():
    assert o.centroid_key(np.array([1024.0, 33.5], np.float32)) == "1024,33.5"
    assert o.centroid_key(np.array([0.1, 1e-7], np.float32)) == "0.1,0.0000001"
    assert o.centroid_key(np.array([16777216.0, -0.0], np.float32)) == "16777216,-0"
    assert o.centroid_key(np.array([np.nan, np.inf], np.float32)) == "NaN,inf"


def test_glrlm_counts_known_answer():
    # 4x6 patch, two grey levels in horizontal stripes of known run lengths, full mask
    g = np.zeros((4, 6), np.float32)
    g[0] = [0, 0, 0, .9, .9, 0]          # runs: 3x lvl0, 2x lvlH, 1x lvl0
    g[1] = .9                            # one run of 6
    g[2] = [0, .9, 0, .9, 0, .9]         # six runs of 1
    g[3] = 0                             # one run of 6
    grey = torch.from_numpy(g)[None, None]
    masks = torch.ones(1, 1, 4, 6)
    hi = int(np.floor(np.float32(.9) * 24))
    R = o.glrlm_counts(grey, 24, 16, (1, 0), masks)[0].numpy()
    assert R[0, 2] == 1 and R[hi, 1] == 1 and R[0, 0] == 1 + 3 and R[hi, 0] == 3 and R[hi, 5] == 1 and R[0, 5] == 1
    assert R.sum() == 3 + 1 + 6 + 1
    # vertical direction (dx,dy) = (0,1): column 0 has levels 0,H,0,0 -> runs 1,1,2
    Rv = o.glrlm_counts(grey, 24, 16, (0, 1), masks)[0].numpy()
    assert (Rv * np.arange(1, 17)[None]).sum() == 24          # every pixel belongs to exactly one run
    # a masked-out pixel breaks a run
    masks[0, 0, 1, 2] = 0
    Rm = o.glrlm_counts(grey, 24, 16, (1, 0), masks)[0].numpy()
    assert Rm[hi, 5] == 0 and Rm[hi, 1] == 2 and Rm[hi, 2] == 1
    f = o.glrlm_features(torch.from_numpy(R)[None], torch.tensor([24.0]))[0]
    d = dict(zip(o.GLRLM_FEATURES, f))
    assert math.isclose(d["run_percentage"], 11 / 24, rel_tol=1e-6)
    assert math.isclose(d["run_length_mean"], 24 / 11, rel_tol=1e-6)


def test_gabor_bank_and_same_padding():
    bank = o.gabor_bank()
    assert bank.shape == (48, 30, 30)
    # theta = 0, f = 0.5: g(u,v) = exp(-(u^2+v^2)/(2 0.45^2)) cos(pi u): symmetric in v, even in u
    assert np.allclose(bank[0], bank[0][::-1], atol=1e-7) and np.allclose(bank[0], bank[0][:, ::-1], atol=1e-7)
    # angle index 2 = 90 degrees: the carrier runs along v (rows)
    assert np.allclose(bank[2 * 6 + 3], bank[0 * 6 + 3].T, atol=1e-6)
    # separable factorisation used by the kernel
    t = np.linspace(-1, 1, 30)
    G = np.exp(-t * t / (2 * 0.45 ** 2))
    th, f = 3 * 2 * np.pi / 8, 4.0
    a, b = 2 * np.pi * f * np.cos(th), 2 * np.pi * f * np.sin(th)
    sep = np.outer(G * np.cos(b * t), G * np.cos(a * t)) - np.outer(G * np.sin(b * t), G * np.sin(a * t))
    assert np.allclose(bank[3 * 6 + 3], sep, atol=1e-6)
    # impulse response: 'same' padding puts tap (14,14) on the output pixel
    x = torch.zeros(1, 1, 64, 64)
    x[0, 0, 20, 30] = 1.0
    y = o.apply_gabor_filter(x)
    assert y.shape == (1, 48, 64, 64)
    assert np.allclose(y[0, 5, 20 - 3, 30 + 2].item(), bank[5, 14 + 3, 14 - 2], atol=1e-7)


def test_glrlm_vectorised_equals_literal_loop():
    from cases import small_case
    tile, rings = small_case(n=12)
    c, polys, patches, masks = o.load_image_dataset(rings, tile, 64)
    gs = o.grey_scale(patches)
    for d in o.GLRLM_DIRECTIONS:
        assert torch.equal(o.glrlm_counts(gs, 24, 16, d, masks), o.glrlm_counts_loop(gs, 24, 16, d, masks))


def test_rule_switches_of_the_unpinned_rules():
    """oracle.RULES (mirrored by nfx_config.rule_flags): each switch changes exactly the rule it names and is restored."""
    sq = np.array([[-2.0, -2.0], [2.0, -2.0], [2.0, 2.0], [-2.0, 2.0], [-2.0, -2.0]])
    a = o.polygon_mask(8, 8, sq)
    assert a.sum() == 16 and a[2:6, 2:6].all()                 # samples -2..1 are inside [-2, 2): half-open at the default grid
    with o.rules(raster_offset=0.5):
        b = o.polygon_mask(8, 8, sq)
        assert b.sum() == 16 and b[2:6, 2:6].all()             # centres -1.5 .. 1.5: the same 16 pixels for this square ...
        tri = np.array([[-2.0, -2.0], [2.0, -2.0], [-2.0, 2.0], [-2.0, -2.0]])
        assert o.polygon_mask(8, 8, tri).sum() != 0
        e1 = o.ellipse_mask(8, 8, (0.0, 0.0), (2.0, 2.0), 0.0)
    assert not np.array_equal(o.ellipse_mask(8, 8, (0.0, 0.0), (2.0, 2.0), 0.0), e1)   # ... but not for a disc of radius 2
    assert o.RULES["raster_offset"] == 0.0
    g = torch.tensor([[[[1.0, 0.999, 0.5, 0.0]]]])
    assert o.quantise(g, 254)[0, 0, 0].tolist() == [253, 253, 127, 0]
    assert o.quantise(g, 254, "u8")[0, 0, 0].tolist() == [253, 253, 127, 0] and o.quantise(torch.tensor([[[[0.9961]]]]), 254, "u8").item() == 253
    assert o.quantise(torch.tensor([[[[0.5]]]]), 254, "u8").item() == 127 and o.quantise(torch.tensor([[[[0.502]]]]), 254, "u8").item() == 128
    assert o.quantise(torch.tensor([[[[0.502]]]]), 254).item() == 127
    assert o.quantise(g, 128, "u8")[0, 0, 0].tolist() == o.quantise(g, 128)[0, 0, 0].tolist()
    full = o.gabor_bank()
    with o.rules(gabor_span=np.pi):
        half = o.gabor_bank()
    assert np.allclose(full[0:6], half[0:6]) and np.allclose(full[12:18], half[24:30]) and not np.allclose(full[6:12], half[6:12])
    assert np.allclose(full[24:30], full[0:6])                 # theta and theta + pi: why the full turn has 24 distinct filters
    assert o.patch_window((10.3, 11.7), 64) == (-20, -21, 43, 42)
    with o.rules(window="slide"):
        assert o.patch_window((10.3, 11.7), 64) == (0, 0, 64, 64)
        assert o.patch_window((100.5, 200.25), 64) == (168, 68, 232, 132)
    with pytest.raises(KeyError):
        o.rules(nonsense=1)


def test_extension_outputs_against_independent_implementations():
    """SPEC.md section C (extension outputs, no reference counterpart): the oracle's definitions against scipy (skew /
    kurtosis), cv2 (moments, Hu) and hand-countable boundaries."""
    import cv2
    from scipy import stats
    rng = np.random.default_rng(3)
    tile = (rng.integers(0, 256, size=(96, 96, 3))).astype(np.uint8)
    t = np.linspace(0, 2 * np.pi, 25)
    rings = [np.stack([48 + r * np.cos(t), 47.5 + 0.8 * r * np.sin(t) + 2 * np.cos(3 * t)], 1).astype(np.float32) for r in (9.0, 17.0, 24.5)]
    cents, polys, patches, masks = o.load_image_dataset(rings, tile, 64)
    cm = o.ext_color_moments(patches, masks)
    for i in range(3):
        sel = masks[i, 0].numpy() != 0
        r = patches[i, 0].numpy()[sel].astype(np.float64)
        assert np.isclose(cm[i, 0], stats.skew(r), rtol=1e-5) and np.isclose(cm[i, 1], stats.kurtosis(r), rtol=1e-5, atol=1e-6)
        v = o.hsv_from_rgb(patches[i:i + 1])[0, 2].numpy()[sel].astype(np.float64)
        assert np.isclose(cm[i, 10], stats.skew(v), rtol=1e-5) and np.isclose(cm[i, 11], stats.kurtosis(v), rtol=1e-5, atol=1e-6)
    mm = o.ext_mask_moments(masks)
    for i in range(3):
        m8 = (masks[i, 0].numpy() != 0).astype(np.uint8)
        M = cv2.moments(m8, binaryImage=True)
        want = [M[k] for k in ("m00", "m10", "m01", "m20", "m11", "m02", "m30", "m21", "m12", "m03", "mu20", "mu11", "mu02", "mu30", "mu21", "mu12", "mu03")]
        want += cv2.HuMoments(M).flatten().tolist()
        assert np.allclose(mm[i], np.array(want, dtype=np.float32), rtol=2e-5, atol=1e-12), (mm[i], want)
    sq = torch.zeros(1, 1, 8, 8)
    sq[0, 0, 2:5, 1:6] = 1                      # 3 x 5 rectangle: 16 cracks; 4 corners (Q1), 2*(2+4) edge quads (Q2)
    c = o.ext_contour(sq)[0]
    assert c[0] == 16.0 and np.isclose(c[1], 12 + 4 / np.sqrt(2))
    one = torch.zeros(1, 1, 8, 8)
    one[0, 0, 0, 0] = 1                          # a corner pixel: the window border counts as background
    assert o.ext_contour(one)[0].tolist() == [4.0, np.float32(4 / np.sqrt(2))]
    g = o.ext_glcm_d2(patches, masks)
    assert g.shape == (3, 112) and np.array_equal(np.nan_to_num(g[:, :14]), np.nan_to_num(o.glcm_features(o.glcm(o.grey_scale(patches), (0, 1), 32, masks))))
    names, allx = o.extract_ext(rings, tile, list(o.EXT_ORDER))
    assert allx.shape == (3, 18 + 24 + 2 + 112) and names[:2] == ["skew_r", "kurtosis_r"] and names[-1] == "information_measure_correlation2_2_-2_32"
    empty = torch.zeros(1, 1, 8, 8)
    assert o.ext_contour(empty)[0].tolist() == [0.0, 0.0] and o.ext_mask_moments(empty)[0, 0] == 0 and np.isnan(o.ext_mask_moments(empty)[0, 10])

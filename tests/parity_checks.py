"""Comparison rules shared by the GPU parity tests: which rows of the REFERENCE's own arithmetic are
ill-conditioned (and are therefore compared on a coarser footing) for the shape and colour sets."""
import numpy as np
import torch

import nfx_oracle as o
from tolerances import mismatches


def _report(bad, limit=12):
    return "\n".join(f"row {r} {c}: got {g!r} want {w!r}" for r, c, g, w in bad[:limit]) + f"\n({len(bad)} mismatches)"


def check_shape(got, want, dbg, names):
    """got [n,12] from the kernels, (want, dbg) = o.shape_features(polys, masks, return_debug=True)."""
    _ = o
    # ill-conditioned rows: rank-1 masks (lambda_min ~ 0) -> compare minor axis relative to major
    got = got.astype(np.float64).copy()
    w2 = want.copy()
    j_min, j_maj, j_ecc, j_ori, j_dev = 2, 1, 3, 4, 8
    with np.errstate(invalid="ignore"):
        degenerate = ~(w2[:, j_min] > 1e-2 * w2[:, j_maj])
    for j in (j_min, j_ecc, j_ori, j_dev):
        got[degenerate, j] = 0
        w2[degenerate, j] = 0
    # near-isotropic masks: the eigenvector direction is ill-conditioned in the reference's own f32
    with np.errstate(invalid="ignore"):
        iso = (w2[:, j_maj] - w2[:, j_min]) < 1e-3 * w2[:, j_maj]
    for j in (j_ori, j_dev):
        got[iso, j] = 0
        w2[iso, j] = 0
    # orientation is an angle: compare modulo pi-wrap at +-pi
    d = np.abs(got[:, j_ori] - w2[:, j_ori])
    wrap = np.isclose(d, 2 * np.pi, atol=1e-3)
    got[wrap, j_ori] = w2[wrap, j_ori]
    bad = mismatches(got, w2, names, "geometry")
    # orientation = atan2 of an eigenvector of the 2 x 2 pixel covariance: its condition number is lambda_max / (lambda_max -
    # lambda_min) = M^2 / (M^2 - m^2). The REFERENCE computes that covariance in f32 (mean subtraction + mm over K pixels,
    # ~1e-5 relative noise; the kernel's integer moments are exact), so beyond 1e-4 the angle is only defined to
    # 1e-5 * M^2 / (M^2 - m^2): a 1.7 % axis difference already means 3e-4 rad.
    M2, m2 = w2[:, j_maj] ** 2, w2[:, j_min] ** 2
    with np.errstate(invalid="ignore", divide="ignore"):
        ori_tol = 1e-5 * M2 / (M2 - m2)
    bad = [b for b in bad if not (b[1] == "orientation" and abs(b[2] - b[3]) <= ori_tol[b[0]])]
    # the deviation is a ratio of two integer counts: f32 noise upstream of the ellipse parameters may
    # flip a boundary pixel in the REFERENCE's own arithmetic; allow +-2 px on <1% of the nuclei ...
    dev_bad = [b for b in bad if b[1] == "eliptic_deviation"]
    other = [b for b in bad if b[1] != "eliptic_deviation"]
    assert not other, _report(other)
    for r, _, g, w in dev_bad:
        K = dbg[r]["area_px"]
        assert abs(g * K - w * K) <= 2.5, f"row {r}: deviation count {g*K} vs {w*K}"
    assert len(dev_bad) <= max(2, len(want) // 100), _report(dev_bad)


def check_color(got, want, names, patches, masks, batch):
    got = got.astype(np.float64).copy()
    want = want.astype(np.float64).copy()
    j = names.index("mean_h")
    # hue mean: compare as an angle (wrap) and skip rows whose resultant vector is ~0 (atan2 of noise)
    n = len(want)
    for k in range(0, n, batch):
        hsv = o.hsv_from_rgb(patches[k:k + batch])
        s, c = o.circular_mean_vectors(hsv[:, 0], masks[k:k + batch])
        area = masks[k:k + batch].sum(dim=[1, 2, 3]) * len(hsv)
        R = (torch.sqrt(s * s + c * c) / area).numpy()
        illc = ~(R > 1e-2)
        got[k:k + batch][illc, j] = 0
        want[k:k + batch][illc, j] = 0
    d = np.abs(got[:, j] - want[:, j])
    wrap = np.abs(d - 360.0) < 0.05
    got[wrap, j] = want[wrap, j]
    bad = mismatches(got, want, names, "color")
    assert not bad, _report(bad)


def check_all_columns(got, names, rings, tile, P, batch, sets=None):
    """Every column of every requested set (flat() order, default all five = 418 columns) against the oracle on the same
    rings / tile: geometry through check_shape, colour through check_color (chunk-coupled mean_h), the texture sets through
    tolerances.mismatches. Returns the oracle's masks so that callers can compare rasters too."""
    sets = list(o.FLAT_ORDER) if sets is None else sets
    cents, polys, patches, masks = o.load_image_dataset(rings, tile, P)
    assert list(names) == [c for s in sets for c in o.SET_COLUMNS[s]], "schema mismatch"
    col = 0
    for s in sets:
        cols = o.SET_COLUMNS[s]
        sub = np.asarray(got)[:, col:col + len(cols)]
        col += len(cols)
        if s == "geometry":
            want, dbg = o.shape_features(polys, masks, return_debug=True)
            check_shape(sub, want, dbg, cols)
        elif s == "color":
            want = np.concatenate([o.color_features(patches[k:k + batch].clone(), masks[k:k + batch]) for k in range(0, len(rings), batch)], 0)
            check_color(sub, want, cols, patches, masks, batch)
        else:
            fn = {"glcm": o.glcm_feature_set, "glrlm": o.glrlm_feature_set, "gabor": o.gabor_feature_set}[s]
            bad = mismatches(sub, fn(patches, masks), cols, s)
            assert not bad, f"{s} at P={P}:\n" + _report(bad)
    return masks

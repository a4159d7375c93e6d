"""Tolerance of the floating-point parity tests: |a-b| <= RTOL * max(|b|, floor) with NaN == NaN.
RTOL is the 1e-4 relative tolerance BASELINE.json's north_star states. `floor` keeps columns whose
value can legitimately be ~0 (std of a flat region, DAB of an H&E pixel, orientation near 0) from
demanding an absolute accuracy below f32 accumulation noise of the REFERENCE itself."""
import numpy as np

RTOL = 1e-4

# per-column absolute scale floors (same unit as the column)
FLOORS = {
    # shape: radians / pixels / ratios
    "orientation": 1.0, "eccentricity": 1.0, "eliptic_deviation": 0.05, "convex_deffect": 0.05,
    "compacity": 0.1,
    # colour: channels are in [0,1], hue in degrees
    "mean_h": 360.0, "std_h": 1.0,
}
# eccentricity = sqrt(1 - (m/M)^2) has an unbounded condition number as m -> M (near-isotropic masks:
# a 1e-6 relative f32 error of the reference's own covariance moves it by 1e-6/ecc^2); the
# well-conditioned quantity is its square, compared on scale 1.
TRANSFORMS = {"eccentricity": lambda v: v * v}

# gabor: a 900-tap f32 convolution of values ~0.5 with taps of alternating sign carries ~1e-5 absolute
# noise in the reference's own conv2d; outputs range from 0.04 to 70, so scale 1 is the floor.
DEFAULT_FLOOR = {"color": 0.02, "glcm": 0.05, "geometry": 1.0, "glrlm": 0.01, "gabor": 1.0}


def floor_for(name, set_name):
    if name in FLOORS:
        return FLOORS[name]
    # normalised GLCM quantities live in [-1,1]: (E[ij]-mu^2)/sigma^2 cancels ~2 digits, and the f32
    # normalised matrix of the reference carries 6e-8 per entry, so 1e-4 is meaningful on scale 1.
    if name.startswith(("correlation_", "information_measure_")):
        return 1.0
    return DEFAULT_FLOOR[set_name]


def mismatches(a, b, names, set_name, rtol=RTOL):
    """Returns list of (row, col_name, got, want) outside tolerance."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    bad = []
    for j, nm in enumerate(names):
        fl = floor_for(nm, set_name)
        x, y = a[:, j], b[:, j]
        if nm in TRANSFORMS:
            x, y = TRANSFORMS[nm](x), TRANSFORMS[nm](y)
        both_nan = np.isnan(x) & np.isnan(y)
        same_inf = np.isinf(x) & np.isinf(y) & (np.sign(x) == np.sign(y))
        with np.errstate(invalid="ignore"):
            ok = np.abs(x - y) <= rtol * np.maximum(np.abs(y), fl)
        ok = ok | both_nan | same_inf
        for i in np.where(~ok)[0]:
            bad.append((int(i), nm, float(x[i]), float(y[i])))
    return bad

"""CPU oracle for the per-nucleus feature pipeline of oxabz/nuclei-feature-extraction.

TEST INFRASTRUCTURE ONLY -- never imported by the product path (`nfx`, `libnfx.so`).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.

PARITY UNPINNED for the functions marked [B*]: their arithmetic lives in the un-vendored git
dependencies tch-utils@d1c10c0 and geometric-features@163ae81 (Cargo.lock:932-935, 2591-2598), the
reference ships no tests/golden vectors and cannot be built here (no rustc).  They follow the rules
written in oracle/SPEC.md section B.  Functions marked [A*] restate in-tree reference code
op-for-op with torch CPU float32 (the same ATen kernels `tch` binds); the cited file:line is
relative to /root/reference.
"""
from __future__ import annotations

import math

import numpy as np
import torch

F32 = np.float32

GLCM_LEVELS = (32, 64, 128, 254)                      # src/features/texture.rs:19
GLCM_OFFSETS = ((0, 1), (1, 1), (1, 0), (1, -1))      # src/features/texture.rs:20  (dy, dx)
GLRLM_LEVELS = 24                                     # src/features/texture.rs:174
GLRLM_MAX_LENGTH = 16                                 # src/features/texture.rs:175
GLRLM_DIRECTIONS = ((1, 0), (1, 1), (0, 1), (-1, 1))  # src/features/texture.rs:176 (dx, dy)
GABOR_ANGLES = 8                                      # src/features/texture.rs:319
GABOR_FREQUENCIES = (0.5, 1.0, 2.0, 4.0, 6.0, 8.0)    # src/features/texture.rs:320
GABOR_KERNEL = 30                                     # src/features/texture.rs:334
GABOR_SIGMA = 0.45                                    # src/features/texture.rs:334


# --------------------------------------------------------------------------------------------
# Switches of the highest-risk UNPINNED rules (SPEC.md section B): each has the adopted position (default) and
# the most plausible alternative, so that a mismatch against reference vectors (tests/test_reference_golden.py,
# tools/emit_reference_golden.rs) is a one-flag fix here and in the kernels (nfx_config.rule_flags, include/nfx.h).
#   raster_offset : 0.0 = pixel (r, c) samples (c - w/2, r - h/2); 0.5 = the pixel centre (c + 0.5 - w/2, ...)   [B1, B2]
#   gabor_span    : angles theta_i = i * span / 8 with span = 2 pi (default) or pi                                [B9]
#   glcm_quant    : "floor" = min(floor(g * L), L - 1) for every level count; "u8" = the 254-level matrix quantises like an
#                   8-bit image, q = min(floor(g * 255), 253) (the other level counts are unchanged: the kernels rely on
#                   floor(g*32) = floor(g*128) >> 2, which a rounding rule would break)                            [B5]
#   window        : "image" = src/utils.rs:159-192 (trunc toward zero, zero padding on every side);
#                   "slide" = src/utils.rs:96-126 (OpenSlide read at `(c - P/2) as u32`: a negative origin saturates to 0,
#                   the window is always P x P and pixels beyond the slide read as 0)                               [A2 / G3]
# --------------------------------------------------------------------------------------------
RULES = {"raster_offset": 0.0, "gabor_span": 2.0 * math.pi, "glcm_quant": "floor", "window": "image"}


class rules:
    """with rules(raster_offset=0.5): ...   -- temporarily switch rules (test infrastructure)."""

    def __init__(self, **kw):
        unknown = set(kw) - set(RULES)
        if unknown:
            raise KeyError(f"unknown rule(s) {sorted(unknown)}")
        self.kw, self.old = kw, None

    def __enter__(self):
        self.old = dict(RULES)
        RULES.update(self.kw)
        return self

    def __exit__(self, *a):
        RULES.clear()
        RULES.update(self.old)


# --------------------------------------------------------------------------------------------
# Schema (A14)
# --------------------------------------------------------------------------------------------
SHAPE_COLUMNS = [  # src/features/shape.rs:114-128
    "area", "major_axis", "minor_axis", "eccentricity", "orientation", "perimeter",
    "equivalent_perimeter", "compacity", "eliptic_deviation", "convex_hull_area",
    "convex_deffect", "convex_perimeter",
]
COLOR_COLUMNS = [  # src/features/color.rs:80-100
    "mean_r", "mean_g", "mean_b", "std_r", "std_g", "std_b", "mean_h", "mean_s", "mean_v",
    "std_h", "std_s", "std_v", "mean_haematoxylin", "mean_eosin", "mean_dab",
    "std_haematoxylin", "std_eosin", "std_dab",
]
GLCM_FEATURES = [  # src/features/texture.rs:81-157 (push order; note the two IMC column names)
    "correlation", "contrast", "dissimilarity", "entropy", "angular_second_moment",
    "sum_average", "sum_variance", "sum_entropy", "sum_of_squares",
    "inverse_difference_moment", "difference_average", "difference_variance",
    "information_measure_correlation1", "information_measure_correlation2",
]
GLCM_COLUMNS = [f"{f}_{o[0]}_{o[1]}_{L}" for L in GLCM_LEVELS for o in GLCM_OFFSETS
                for f in GLCM_FEATURES]
GLRLM_FEATURES = [  # src/features/texture.rs:243-301 (vec! order)
    "short_run_emphasis", "long_run_emphasis", "gray_level_nonuniformity",
    "run_length_nonuniformity", "low_gray_level_run_emphasis", "high_gray_level_run_emphasis",
    "short_run_low_gray_level_emphasis", "short_run_high_gray_level_emphasis",
    "long_run_low_gray_level_emphasis", "long_run_high_gray_level_emphasis",
    "short_run_mid_gray_level_emphasis", "long_run_mid_gray_level_emphasis",
    "short_run_extreme_gray_level_emphasis", "long_run_extreme_gray_level_emphasis",
    "run_percentage", "run_length_mean", "run_length_variance",
]
GLRLM_COLUMNS = [f"{f}_{d[0]}_{d[1]}" for d in GLRLM_DIRECTIONS for f in GLRLM_FEATURES]


def rust_f32_display(x) -> str:
    """Rust `Display` for f32: shortest round-trip decimal, never an exponent (utils.rs:226-228)."""
    x = np.float32(x)
    if np.isnan(x):
        return "NaN"
    if np.isinf(x):
        return "inf" if x > 0 else "-inf"
    return np.format_float_positional(x, unique=True, trim="-")


GABOR_COLUMNS = [  # src/features/texture.rs:346-361
    f"gabor_angle_{rust_f32_display(np.float32(j // len(GABOR_FREQUENCIES)) * np.float32(45.0))}"
    f"_frequency_{rust_f32_display(np.float32(GABOR_FREQUENCIES[j % len(GABOR_FREQUENCIES)]))}_{s}"
    for j in range(GABOR_ANGLES * len(GABOR_FREQUENCIES)) for s in ("mean", "variance")
]
FLAT_ORDER = ("geometry", "color", "glcm", "glrlm", "gabor")        # src/args.rs:38-44
SET_COLUMNS = {"geometry": SHAPE_COLUMNS, "color": COLOR_COLUMNS, "glcm": GLCM_COLUMNS,
               "glrlm": GLRLM_COLUMNS, "gabor": GABOR_COLUMNS}


def flat(names):
    """args::FeatureSet::flat (src/args.rs:35-49), case-insensitive FromStr (src/args.rs:18-32)."""
    out = []
    for s in names:
        s = s.lower()
        if s == "all":
            out += list(FLAT_ORDER)
        elif s == "texture":
            out += ["glcm", "glrlm", "gabor"]
        elif s in SET_COLUMNS:
            out.append(s)
        else:
            raise ValueError(f"{s} is not a valid feature set")
    return out


def centroid_key(c) -> str:
    """centroid_to_key_string (src/utils.rs:226-228)."""
    return f"{rust_f32_display(c[0])},{rust_f32_display(c[1])}"


# --------------------------------------------------------------------------------------------
# Batch builder (A1, A2, B1)
# --------------------------------------------------------------------------------------------
def preprocess_polygon(ring):
    """[A1] src/utils.rs:54-74. ring: (V,2) f32 as stored in the GeoJSON (closing duplicate kept).
    Returns (centroid f32[2], centred f32[V,2]); accumulation is sequential f32."""
    ring = np.asarray(ring, dtype=F32).reshape(-1, 2)
    acc = np.cumsum(ring, axis=0, dtype=F32)[-1]          # add.accumulate is sequential
    centroid = (acc / F32(len(ring))).astype(F32)
    return centroid, (ring - centroid).astype(F32)


def polygon_mask(w, h, pts):
    """[B1] tch_utils::shapes::polygon(w, h, &pts_f64, (Float, Cpu)) -> [h,w] bool. SPEC.md B1."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    X = (np.arange(w, dtype=np.float64) + RULES["raster_offset"] - w / 2.0)[None, :]
    Y = (np.arange(h, dtype=np.float64) + RULES["raster_offset"] - h / 2.0)[:, None]
    inside = np.zeros((h, w), dtype=bool)
    n = len(pts)
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(n):
            xi, yi = pts[i]
            xj, yj = pts[(i + 1) % n]
            cross = (yi <= Y) != (yj <= Y)                         # [h,1]
            xint = xi + (Y - yi) * (xj - xi) / (yj - yi)           # [h,1]
            inside ^= cross & (X < xint)
    return inside


def ellipse_mask(w, h, center, radii, angle):
    """[B2] tch_utils::shapes::ellipse(w, h, center, radii, angle, ..) -> [h,w] bool. SPEC.md B2."""
    cx, cy = np.float64(center[0]), np.float64(center[1])
    a, b = np.float64(radii[0]), np.float64(radii[1])
    ang = np.float64(angle)
    X = (np.arange(w, dtype=np.float64) + RULES["raster_offset"] - w / 2.0)[None, :]
    Y = (np.arange(h, dtype=np.float64) + RULES["raster_offset"] - h / 2.0)[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        cs, sn = np.cos(ang), np.sin(ang)
        dx, dy = X - cx, Y - cy
        xr = dx * cs + dy * sn
        yr = dy * cs - dx * sn
        u = xr / a
        v = yr / b
        val = u * u + v * v
        return np.broadcast_to(val <= 1.0, (h, w)).copy()


def patch_window(centroid, P):
    """[A2] src/utils.rs:159-162: f32 arithmetic, Rust `as i64` truncates toward zero."""
    half = F32(P) / F32(2.0)
    cx, cy = F32(centroid[0]), F32(centroid[1])
    if RULES["window"] == "slide":
        # src/utils.rs:96-109: Region { address: ((cx - P/2) as u32, (cy - P/2) as u32), size: P x P, level 0 }. Rust's
        # float -> u32 cast saturates: negative and NaN give 0. OpenSlide returns a full P x P image (black outside the
        # slide), so the top-left padding branch (utils.rs:113-122) never changes it.
        def as_u32(v):
            return 0 if not (v == v) or v <= 0 else int(min(np.trunc(v), 4294967295.0))
        top, left = as_u32(cy - half), as_u32(cx - half)
        return top, left, top + P, left + P
    top, left = int(np.trunc(cy - half)), int(np.trunc(cx - half))
    bottom, right = int(np.trunc(cy + half)), int(np.trunc(cx + half))
    return top, left, bottom, right


def gather_patch_u8(image_hwc, centroid, P):
    """[A2] src/utils.rs:159-192 on an interleaved [H,W,3] u8 image; returns [P,P,3] u8 (zero padded).
    The reference's f32 patch is exactly this / 255 (see `gather_patch`)."""
    H, W = image_hwc.shape[:2]
    top, left, bottom, right = patch_window(centroid, P)
    r0, r1 = max(top, 0), min(bottom, H)
    c0, c1 = max(left, 0), min(right, W)
    out = np.zeros((P, P, 3), dtype=np.uint8)
    if r1 <= r0 or c1 <= c0:
        return out
    oy, ox = -min(top, 0), -min(left, 0)
    nr, nc = min(r1 - r0, P - oy), min(c1 - c0, P - ox)       # clip (reference would panic)
    if nr > 0 and nc > 0:
        out[oy:oy + nr, ox:ox + nc] = image_hwc[r0:r0 + nr, c0:c0 + nc]
    return out


def gather_patch(image_hwc, centroid, P):
    """[A2] [3,P,P] f32 in [0,1] = u8.to_kind(Float)/255.0 (src/utils.rs:172)."""
    u8 = gather_patch_u8(image_hwc, centroid, P)
    return (torch.from_numpy(u8).permute(2, 0, 1).to(torch.float32) / 255.0)


def load_image_dataset(rings, image_hwc, P):
    """[A1,A2,B1] src/utils.rs:141-206 -> (centroids [N,2] f32, centred polygons, patches
    [N,3,P,P] f32, masks [N,1,P,P] f32)."""
    cents, polys, patches, masks = [], [], [], []
    for ring in rings:
        c, cp = preprocess_polygon(ring)
        m = polygon_mask(P, P, cp.astype(np.float64))
        cents.append(c)
        polys.append(cp)
        masks.append(torch.from_numpy(m.astype(np.float32))[None])
        patches.append(gather_patch(image_hwc, c, P))
    return (np.stack(cents).astype(F32), polys, torch.stack(patches, 0), torch.stack(masks, 0))


# --------------------------------------------------------------------------------------------
# Shape set (A3-A8, B2, B7, B8)
# --------------------------------------------------------------------------------------------
def center_of_mass(mask):
    """[A3] src/features/shape.rs:219-226. mask: [1,P,P] tensor."""
    nz = mask.nonzero()[:, -2:]
    c = nz.mean(dim=[0], keepdim=False, dtype=torch.float32)
    return [F32(c[0].item()), F32(c[1].item())]


def eig2x2_lapack(a, b, d):
    """[A5] closed form of LAPACK sgeev on the symmetric 2x2 [[a,b],[b,d]] (slanv2), float32.
    Returns (l0, l1, V) with V's COLUMNS the eigenvectors, exactly as torch.linalg.eig orders them.
    Checked against torch.linalg.eig in tests/test_oracle.py."""
    a, b, d = F32(a), F32(b), F32(d)
    if b == 0:
        return a, d, np.array([[1, 0], [0, 1]], dtype=F32)
    p = F32(0.5) * (a - d)
    r = F32(np.hypot(p, b))
    z = p + F32(math.copysign(r, p))
    l0 = d + z
    l1 = d - (b / z) * b
    tau = F32(np.hypot(b, z))
    cs, sn = z / tau, b / tau
    return F32(l0), F32(l1), np.array([[cs, -sn], [sn, cs]], dtype=F32)


def major_minor_axes_w_angle(mask):
    """[A4,A5] src/features/shape.rs:141-203. Returns (major, minor, angle) as f32."""
    nan = F32(np.nan)
    nz = mask.squeeze().nonzero()
    centroid = nz.mean(dim=[0], keepdim=False, dtype=torch.float32)
    if centroid.size(0) == 0:
        return nan, nan, nan
    points = nz - centroid
    cov = points.transpose(0, 1).mm(points) / points.size(0)
    if bool(cov.isnan().any()):
        return nan, nan, nan
    if cov.size(0) != 2 or cov.size(1) != 2:
        return nan, nan, nan
    if bool(cov.isinf().any()):
        return nan, nan, nan
    eigenvalues, eigenvector = torch.linalg.eig(cov)
    a = F32(eigenvalues[0].real.item())
    b = F32(eigenvalues[1].real.item())
    with np.errstate(invalid="ignore"):
        if a > b:
            major, minor, mc = np.sqrt(a), np.sqrt(b), eigenvector[0]
        else:
            major, minor, mc = np.sqrt(b), np.sqrt(a), eigenvector[1]
    x = F32(mc[0].real.item())
    y = F32(mc[1].real.item())
    angle = F32(np.arctan2(x, y))
    return F32(major * F32(2.0)), F32(minor * F32(2.0)), angle


def eccentricity(major, minor):
    """[A6] src/features/shape.rs:205-207 (f32)."""
    M, m = F32(major) * F32(0.5), F32(minor) * F32(0.5)
    with np.errstate(invalid="ignore", divide="ignore"):
        return F32(np.sqrt(F32(M * M - m * m)) / M)


def eliptic_deviation(mask, ellipse):
    """[A8] src/features/shape.rs:209-217."""
    mask = mask.to(torch.float32)
    ellipse = ellipse.to(torch.float32)
    mask_area = float(F32(mask.sum(dtype=torch.float32).item()))
    delta = mask - ellipse
    return F32((delta.abs().sum(dtype=torch.float32) / mask_area).item())


def polygon_area(pts):
    """[B7] shoelace, float64."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    x, y = pts[:, 0], pts[:, 1]
    xn, yn = np.roll(x, -1), np.roll(y, -1)
    return 0.5 * abs(float(np.sum(x * yn - xn * y)))


def polygon_perimeter(pts):
    """[B7] closed polyline length, float64."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    d = np.roll(pts, -1, axis=0) - pts
    return float(np.sum(np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2)))


def convex_hull(pts):
    """[B8] Andrew's monotone chain (strict turns; collinear points dropped)."""
    P = sorted(set(map(tuple, np.asarray(pts, dtype=np.float64).reshape(-1, 2).tolist())))
    if len(P) <= 2:
        return np.array(P, dtype=np.float64).reshape(-1, 2)

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lo, up = [], []
    for p in P:
        while len(lo) >= 2 and cross(lo[-2], lo[-1], p) <= 0:
            lo.pop()
        lo.append(p)
    for p in reversed(P):
        while len(up) >= 2 and cross(up[-2], up[-1], p) <= 0:
            up.pop()
        up.append(p)
    return np.array(lo[:-1] + up[:-1], dtype=np.float64)


def polygon_geometry(pts):
    """[B7,B8] the 7 polygon scalars of src/features/shape.rs:89-97, float64."""
    area = polygon_area(pts)
    per = polygon_perimeter(pts)
    hull = convex_hull(pts)
    harea = polygon_area(hull) if len(hull) >= 3 else 0.0
    hper = polygon_perimeter(hull) if len(hull) >= 2 else 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        return dict(
            area=area, perimeter=per,
            equivalent_perimeter=2.0 * math.sqrt(math.pi * area),
            compacity=float(np.float64(4.0 * math.pi * area) / np.float64(per * per)),
            convex_hull_area=harea, convex_perimeter=hper,
            convex_deffect=float((np.float64(harea) - np.float64(area)) / np.float64(harea)),
        )


def shape_features(centred_polygons, masks, return_debug=False):
    """[A3-A8] ShapeFeatureSet::compute_features_batched (src/features/shape.rs:16-130).
    Returns [N,12] float64 in SHAPE_COLUMNS order (f32-valued where the reference is f32)."""
    N = masks.shape[0]
    P = masks.shape[3]
    out = np.zeros((N, 12), dtype=np.float64)
    dbg = []
    for i in range(N):
        poly = np.asarray(centred_polygons[i], dtype=F32).astype(np.float64)   # utils.rs:24-31
        mask = masks[i]
        com = center_of_mass(mask)
        com = [F32(com[0] - F32(P) / F32(2.0)), F32(com[1] - F32(P) / F32(2.0))]
        major, minor, angle = major_minor_axes_w_angle(mask)
        ecc = eccentricity(major, minor)
        ell = ellipse_mask(P, P, (float(com[1]), float(com[0])), (float(major), float(minor)),
                           float(angle))
        ell_t = torch.from_numpy(ell.astype(np.float32))[None]
        g = polygon_geometry(poly)
        dev = eliptic_deviation(mask, ell_t)
        out[i] = [g["area"], major, minor, ecc, angle, g["perimeter"], g["equivalent_perimeter"],
                  g["compacity"], dev, g["convex_hull_area"], g["convex_deffect"],
                  g["convex_perimeter"]]
        if return_debug:
            dbg.append(dict(com=com, ellipse=ell,
                            abs_diff=int(np.sum(ell != (mask[0].numpy() != 0))),
                            area_px=int((mask[0] != 0).sum().item())))
    return (out, dbg) if return_debug else out


# --------------------------------------------------------------------------------------------
# Colour set (A9-A11, B3, B4)
# --------------------------------------------------------------------------------------------
def hsv_from_rgb(rgb):
    """[B3] tch_utils::color::hsv_from_rgb, [N,3,P,P] f32 in [0,1] -> HSV, H in degrees."""
    r, g, b = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    maxc = torch.maximum(torch.maximum(r, g), b)
    minc = torch.minimum(torch.minimum(r, g), b)
    delta = maxc - minc
    zero = torch.zeros_like(maxc)
    safe = torch.where(delta == 0, torch.ones_like(delta), delta)
    hr = torch.remainder((g - b) / safe, 6.0)
    hg = (b - r) / safe + 2.0
    hb = (r - g) / safe + 4.0
    h = torch.where(maxc == r, hr, torch.where(maxc == g, hg, hb)) * 60.0
    h = torch.where(delta == 0, zero, h)
    s = torch.where(maxc > 0, delta / torch.where(maxc > 0, maxc, torch.ones_like(maxc)), zero)
    return torch.stack([h, s, maxc], dim=1)


_RGB_FROM_HED = np.array([[0.65, 0.70, 0.29], [0.07, 0.99, 0.11], [0.27, 0.57, 0.78]],
                         dtype=np.float64)
HED_FROM_RGB = np.linalg.inv(_RGB_FROM_HED).astype(np.float32)      # M[k][c]
LOG_1E6 = np.float32(math.log(1e-6))


def hed_from_rgb(rgb):
    """[B4] tch_utils::color::hed_from_rgb (scikit-image rgb2hed), [N,3,P,P] f32 -> H,E,D."""
    od = torch.log(torch.clamp(rgb, min=1e-6)) / float(LOG_1E6)
    M = HED_FROM_RGB
    chans = []
    for c in range(3):
        v = od[:, 0] * float(M[0, c]) + od[:, 1] * float(M[1, c]) + od[:, 2] * float(M[2, c])
        chans.append(torch.clamp(v, min=0.0))
    return torch.stack(chans, dim=1)


def mean_std(img, mask):
    """[A9] src/features/color.rs:117-134."""
    masked = img * mask
    mask_area = mask.sum(dim=[-1, -2], keepdim=True, dtype=torch.float32)
    mean = masked.sum(dim=[-1, -2], keepdim=True, dtype=torch.float32)
    mean /= mask_area
    std = img - mean
    std = std.square() * mask
    std = std.sum(dim=[-1, -2], keepdim=True, dtype=torch.float32)
    std /= mask_area
    std = std.sqrt()
    std.squeeze_()
    mean.squeeze_()
    return mean, std


def circular_mean(image, mask):
    """[A10] src/features/color.rs:144-155 -- called with image [N,P,P], mask [N,1,P,P]: the
    product broadcasts to [N,N,P,P] (batch-coupled hue mean)."""
    mask_area = mask.sum(dim=[-1, -2, -3], keepdim=False, dtype=torch.float32)
    img = image.deg2rad()
    cos = img.cos() * mask
    sin = img.sin() * mask
    cos = cos.sum(dim=[-1, -2, -3], keepdim=False, dtype=torch.float32) / mask_area
    sin = sin.sum(dim=[-1, -2, -3], keepdim=False, dtype=torch.float32) / mask_area
    return (sin.atan2(cos).rad2deg_() + 360.0).fmod_(360.0)


def circular_mean_vectors(image, mask):
    """The two resultant components (Σ sin, Σ cos) of A10 before atan2 -- the well-conditioned
    quantity the parity test falls back to when the resultant is ~0."""
    img = image.deg2rad()
    cos = (img.cos() * mask).sum(dim=[-1, -2, -3], dtype=torch.float32)
    sin = (img.sin() * mask).sum(dim=[-1, -2, -3], dtype=torch.float32)
    return sin, cos


def color_features(patchs, masks):
    """[A9-A11] ColorFeatureSet::compute_features_batched (src/features/color.rs:10-102) for ONE
    batch. Returns [N,18] f32 in COLOR_COLUMNS order."""
    N = patchs.shape[0]
    hsv = hsv_from_rgb(patchs)
    hed = hed_from_rgb(patchs)
    mean_rgb, std_rgb = mean_std(patchs, masks)
    mean_hed, std_hed = mean_std(hed, masks)
    h = hsv.select(-3, 0)
    mean_h = circular_mean(h, masks)
    h -= mean_h.view(-1, 1, 1)
    mean_hsv, std_hsv = mean_std(hsv, masks)

    def col(t, c):
        return t.select(-1, c).reshape(N)

    cols = [col(mean_rgb, 0), col(mean_rgb, 1), col(mean_rgb, 2),
            col(std_rgb, 0), col(std_rgb, 1), col(std_rgb, 2),
            mean_h.reshape(N), col(mean_hsv, 1), col(mean_hsv, 2),
            col(std_hsv, 0), col(std_hsv, 1), col(std_hsv, 2),
            col(mean_hed, 0), col(mean_hed, 1), col(mean_hed, 2),
            col(std_hed, 0), col(std_hed, 1), col(std_hed, 2)]
    return torch.stack(cols, dim=1).numpy().astype(F32)


# --------------------------------------------------------------------------------------------
# GLCM set (A12, A13, B5, B6)
# --------------------------------------------------------------------------------------------
def grey_scale(patchs):
    """[A12] src/features/texture.rs:36."""
    return patchs.mean(dim=[-3], keepdim=True, dtype=torch.float32)


def quantise(grey, levels, rule="floor"):
    """[B5] q = min(floor(grey*L), L-1) with an f32 multiply (rule "floor"); rule "u8" (GLCM only, see RULES) scales the
    254-level case by 255 instead."""
    scale = 255.0 if (rule == "u8" and int(levels) == 254) else float(levels)
    q = torch.floor(grey * scale).to(torch.int64)
    return torch.clamp(q, max=int(levels) - 1)


def glcm_counts(grey, offset, levels, masks):
    """[B5] symmetric masked co-occurrence COUNTS G = C + C^T, [N,L,L] int64."""
    N, _, H, W = grey.shape
    L = int(levels)
    dy, dx = offset
    q = quantise(grey, L, RULES["glcm_quant"])[:, 0]
    m = masks[:, 0] != 0
    r0, r1 = max(0, -dy), min(H, H - dy)
    c0, c1 = max(0, -dx), min(W, W - dx)
    src_q, dst_q = q[:, r0:r1, c0:c1], q[:, r0 + dy:r1 + dy, c0 + dx:c1 + dx]
    valid = m[:, r0:r1, c0:c1] & m[:, r0 + dy:r1 + dy, c0 + dx:c1 + dx]
    n_idx = torch.arange(N).view(N, 1, 1).expand_as(src_q)
    flat_idx = (n_idx * L * L + src_q * L + dst_q)[valid]
    C = torch.bincount(flat_idx, minlength=N * L * L).view(N, L, L)
    return C + C.transpose(1, 2)


def glcm(grey, offset, levels, masks):
    """[B5] normalised symmetric masked GLCM, [N,L,L] f32."""
    G = glcm_counts(grey, offset, levels, masks).to(torch.float32)
    return G / G.sum(dim=[-1, -2], keepdim=True)


def _xlogx(p):
    return torch.where(p > 0, p * torch.log(torch.where(p > 0, p, torch.ones_like(p))),
                       torch.zeros_like(p))


def glcm_features(p):
    """[B6] tch_utils::glcm::features::glcm_features; p [N,L,L] f32 -> [N,14] f32 in
    GLCM_FEATURES order. SPEC.md B6. Evaluated in float64 from the f32 matrix so that the oracle
    value is the mathematically defined one (the kernel accumulates in f32/f64, tolerance 1e-4)."""
    N, L, _ = p.shape
    nanrow = torch.isnan(p).flatten(1).any(dim=1)
    p = torch.nan_to_num(p, nan=0.0).to(torch.float64)
    i = torch.arange(L, dtype=torch.float64).view(1, L, 1)
    j = torch.arange(L, dtype=torch.float64).view(1, 1, L)
    px = p.sum(dim=2)
    py = p.sum(dim=1)
    lv = torch.arange(L, dtype=torch.float64).view(1, L)
    mux = (lv * px).sum(1)
    muy = (lv * py).sum(1)
    varx = (((lv - mux[:, None]) ** 2) * px).sum(1)
    vary = (((lv - muy[:, None]) ** 2) * py).sum(1)
    corr = ((i * j * p).sum(dim=[1, 2]) - mux * muy) / torch.sqrt(varx * vary)
    contrast = (((i - j) ** 2) * p).sum(dim=[1, 2])
    dissim = ((i - j).abs() * p).sum(dim=[1, 2])
    entropy = -_xlogx(p).sum(dim=[1, 2])
    asm = (p * p).sum(dim=[1, 2])
    # p_{x+y}, p_{x-y}
    ks = (torch.arange(L).view(L, 1) + torch.arange(L).view(1, L)).flatten()
    kd = (torch.arange(L).view(L, 1) - torch.arange(L).view(1, L)).abs().flatten()
    pf = p.flatten(1)
    psum = torch.zeros(N, 2 * L - 1, dtype=torch.float64).index_add_(1, ks, pf)
    pdif = torch.zeros(N, L, dtype=torch.float64).index_add_(1, kd, pf)
    k2 = torch.arange(2 * L - 1, dtype=torch.float64).view(1, -1)
    k1 = torch.arange(L, dtype=torch.float64).view(1, -1)
    sum_avg = (k2 * psum).sum(1)
    sum_var = (((k2 - sum_avg[:, None]) ** 2) * psum).sum(1)
    sum_ent = -_xlogx(psum).sum(1)
    sos = (((i - mux.view(N, 1, 1)) ** 2) * p).sum(dim=[1, 2])
    idm = (p / (1.0 + (i - j) ** 2)).sum(dim=[1, 2])
    dif_avg = (k1 * pdif).sum(1)
    dif_var = (((k1 - dif_avg[:, None]) ** 2) * pdif).sum(1)
    hx = -_xlogx(px).sum(1)
    hy = -_xlogx(py).sum(1)
    pxpy = px[:, :, None] * py[:, None, :]
    logpxpy = torch.log(torch.where(pxpy > 0, pxpy, torch.ones_like(pxpy)))
    hxy1 = -(p * logpxpy).sum(dim=[1, 2])
    hxy2 = -(pxpy * logpxpy).sum(dim=[1, 2])
    imc1 = (entropy - hxy1) / torch.maximum(hx, hy)
    imc2 = torch.sqrt(torch.clamp(1.0 - torch.exp(-2.0 * (hxy2 - entropy)), min=0.0))
    out = torch.stack([corr, contrast, dissim, entropy, asm, sum_avg, sum_var, sum_ent, sos, idm,
                       dif_avg, dif_var, imc1, imc2], dim=1)
    out[nanrow] = float("nan")
    return out.to(torch.float32).numpy()


def glcm_feature_set(patchs, masks, levels=GLCM_LEVELS, offsets=GLCM_OFFSETS):
    """[A12,A13] GlcmFeatureSet::compute_features_batched (src/features/texture.rs:24-168):
    [N, 14*len(levels)*len(offsets)] f32, loop order levels outer, offsets inner."""
    grey = grey_scale(patchs)
    cols = []
    for L in levels:
        for off in offsets:
            cols.append(glcm_features(glcm(grey, off, L, masks)))
    return np.concatenate(cols, axis=1).astype(F32)


# --------------------------------------------------------------------------------------------
# GLRLM set (SPEC.md B9) -- "next" row of SURVEY.md 8f; every definition below is UNPINNED
# --------------------------------------------------------------------------------------------
def glrlm_counts(grey, levels, max_length, direction, masks):
    """[B9] tch_utils::glrlm::glrlm(&gs, levels, max_length, direction, Some(masks)) -> [N,levels,max_length]
    int64. q = min(floor(grey*levels), levels-1); direction = (dx, dy) (texture.rs:243 names them so);
    a run is a maximal sequence p, p+d, p+2d, ... of masked-in pixels with equal level (masked-out
    pixels and the patch border break runs); runs longer than max_length fall in the last bin."""
    N, _, H, W = grey.shape
    q = quantise(grey, levels)[:, 0].numpy()
    m = (masks[:, 0] != 0).numpy()
    dx, dy = direction

    def shifted(a, k):
        """out[n, r, c] = a[n, r + k*dy, c + k*dx] (zero outside the patch)."""
        out = np.zeros_like(a)
        r0, r1 = max(0, -k * dy), min(H, H - k * dy)
        c0, c1 = max(0, -k * dx), min(W, W - k * dx)
        if r1 > r0 and c1 > c0:
            out[:, r0:r1, c0:c1] = a[:, r0 + k * dy:r1 + k * dy, c0 + k * dx:c1 + k * dx]
        return out

    # a pixel CONTINUES a run when its predecessor along the direction is masked-in with the same level
    cont = m & shifted(m, -1) & (q == shifted(q + 1, -1) - 1)
    start = m & ~cont
    length = start.astype(np.int64)
    alive = start.copy()
    k = 1
    while alive.any():
        alive &= shifted(cont, k)
        length += alive
        k += 1
    R = np.zeros((N, levels, max_length), dtype=np.int64)
    n_idx, rr, cc = np.nonzero(start)
    np.add.at(R, (n_idx, q[n_idx, rr, cc], np.minimum(length[n_idx, rr, cc], max_length) - 1), 1)
    return torch.from_numpy(R)


def glrlm_counts_loop(grey, levels, max_length, direction, masks):
    """Literal per-pixel restatement of glrlm_counts (kept as the cross-check of the vectorised form)."""
    N, _, H, W = grey.shape
    q = quantise(grey, levels)[:, 0].numpy()
    m = (masks[:, 0] != 0).numpy()
    dx, dy = direction
    R = np.zeros((N, levels, max_length), dtype=np.int64)
    for n in range(N):
        rr, cc = np.nonzero(m[n])
        for r, c in zip(rr.tolist(), cc.tolist()):
            pr, pc = r - dy, c - dx
            if 0 <= pr < H and 0 <= pc < W and m[n, pr, pc] and q[n, pr, pc] == q[n, r, c]:
                continue                                   # not the start of a run
            lv, l, r2, c2 = q[n, r, c], 1, r + dy, c + dx
            while 0 <= r2 < H and 0 <= c2 < W and m[n, r2, c2] and q[n, r2, c2] == lv:
                l, r2, c2 = l + 1, r2 + dy, c2 + dx
            R[n, lv, min(l, max_length) - 1] += 1
    return torch.from_numpy(R)


def glrlm_features(R, pixel_count):
    """[B9] tch_utils::glrlm::features::glrlm_features(&glrlm, Some(&pixel_count)) -> [N,17] f32 in
    GLRLM_FEATURES order. i = 1-based grey level, j = 1-based run length, Nr = sum R:
    Galloway / Chu / Dasarathy emphases; the reference's non-standard "mid"/"extreme" grey-level
    emphases use m(i) = 1 - u^2 and e(i) = u^2 with u = (i - c)/(c - 1), c = (L+1)/2."""
    R = R.to(torch.float64)
    N, L, M = R.shape
    i = torch.arange(1, L + 1, dtype=torch.float64).view(1, L, 1)
    j = torch.arange(1, M + 1, dtype=torch.float64).view(1, 1, M)
    nr = R.sum(dim=[1, 2])
    c = (L + 1) / 2.0
    u2 = ((i - c) / (c - 1.0)) ** 2

    def S(w):
        return (R * w).sum(dim=[1, 2]) / nr

    sre, lre = S(1 / j ** 2), S(j ** 2)
    gln = (R.sum(dim=2) ** 2).sum(dim=1) / nr
    rln = (R.sum(dim=1) ** 2).sum(dim=1) / nr
    lgre, hgre = S(1 / i ** 2), S(i ** 2)
    srlge, srhge = S(1 / (i ** 2 * j ** 2)), S(i ** 2 / j ** 2)
    lrlge, lrhge = S(j ** 2 / i ** 2), S(i ** 2 * j ** 2)
    srmge, lrmge = S((1 - u2) / j ** 2), S((1 - u2) * j ** 2)
    srege, lrege = S(u2 / j ** 2), S(u2 * j ** 2)
    rp = nr / pixel_count.to(torch.float64)
    rlm = S(j + 0 * i)
    rlv = (R * (j - rlm.view(N, 1, 1)) ** 2).sum(dim=[1, 2]) / nr
    out = torch.stack([sre, lre, gln, rln, lgre, hgre, srlge, srhge, lrlge, lrhge, srmge, lrmge, srege, lrege,
                       rp, rlm, rlv], dim=1)
    return out.to(torch.float32).numpy()


def glrlm_feature_set(patchs, masks):
    """GLRLMFeatureSet::compute_features_batched (src/features/texture.rs:178-310): [N,68] f32."""
    gs = grey_scale(patchs)
    pixel_count = masks.sum(dim=[-3, -2, -1])
    cols = []
    for d in GLRLM_DIRECTIONS:
        R = glrlm_counts(gs, GLRLM_LEVELS, GLRLM_MAX_LENGTH, d, masks)
        cols.append(glrlm_features(R, pixel_count))
    return np.concatenate(cols, axis=1).astype(F32)


# --------------------------------------------------------------------------------------------
# Gabor set (SPEC.md B9) -- "next" row; UNPINNED
# --------------------------------------------------------------------------------------------
def gabor_bank(angles=GABOR_ANGLES, ksize=GABOR_KERNEL, freqs=GABOR_FREQUENCIES, sigma=GABOR_SIGMA):
    """[B9] the 48 kernels of tch_utils::gabor::apply_gabor_filter, filter j = angle_idx*len(freqs)+freq_idx
    (texture.rs:349-350), float32 [48, ksize, ksize]. Taps on the normalised grid u,v in
    linspace(-1,1,ksize) (sigma and frequency are in those units), theta = angle_idx * 2*pi/angles:
        x' = u cos(theta) + v sin(theta) ;  g = exp(-(u^2+v^2)/(2 sigma^2)) * cos(2 pi f x')
    (isotropic envelope, real part, no normalisation). v = row axis, u = column axis."""
    t = np.linspace(-1.0, 1.0, ksize)
    U, V = np.meshgrid(t, t)          # U varies along columns, V along rows
    bank = []
    for a in range(angles):
        th = a * RULES["gabor_span"] / angles
        xr = U * np.cos(th) + V * np.sin(th)
        for f in freqs:
            bank.append(np.exp(-(U * U + V * V) / (2.0 * sigma * sigma)) * np.cos(2.0 * np.pi * f * xr))
    return np.stack(bank).astype(np.float32)


def apply_gabor_filter(gs):
    """[B9] -> [N,48,P,P]: cross-correlation (torch conv2d) with zero padding 'same'; for the even
    30-tap kernel that is 14 taps before and 15 after the output pixel."""
    k = torch.from_numpy(gabor_bank())[:, None]                       # [48,1,30,30]
    pad_lo = (GABOR_KERNEL - 1) // 2
    pad_hi = GABOR_KERNEL - 1 - pad_lo
    x = torch.nn.functional.pad(gs, (pad_lo, pad_hi, pad_lo, pad_hi))
    return torch.nn.functional.conv2d(x, k)


def gabor_feature_set(patchs, masks):
    """GaborFilterFeatureSet::compute_features_batched (src/features/texture.rs:322-369): [N,96] f32,
    (mean, variance) interleaved per filter."""
    gs = grey_scale(patchs)
    filtered = apply_gabor_filter(gs)
    N, Fc = patchs.shape[0], filtered.shape[1]
    areas = masks.sum(dim=[-3, -2, -1])
    mean = (filtered * masks).sum(dim=[-2, -1]) / areas.unsqueeze(-1)
    var = ((filtered - mean.view(N, Fc, 1, 1)).square() * masks).sum(dim=[-2, -1]) / areas.unsqueeze(-1)
    return torch.stack([mean, var], dim=2).reshape(N, 2 * Fc).numpy().astype(F32)


# --------------------------------------------------------------------------------------------
# Whole pipeline, batch by batch (src/main.rs:146-158, 47-91)
# --------------------------------------------------------------------------------------------
def extract(rings, image_hwc, feature_sets, patch_size=64, batch_size=100):
    """Reference pipeline restated: par_chunks(batch_size) -> load_image_dataset ->
    compute_features_batched per set -> hstack in flat() order. Rows are returned in INPUT order
    (the reference's completion order is non-deterministic; join on the key).
    Returns (keys list[str], centroids [N,2] f32, features [N,F] float64, column names)."""
    sets = flat(feature_sets)
    if len(set(sets)) != len(sets):
        raise ValueError("duplicate feature set (reference: DataFrame::new fails, main.rs:89)")
    names = [c for s in sets for c in SET_COLUMNS[s]]
    keys, cents, rows = [], [], []
    for k in range(0, len(rings), batch_size):
        chunk = rings[k:k + batch_size]
        c, polys, patches, masks = load_image_dataset(chunk, image_hwc, patch_size)
        blocks = []
        for s in sets:
            if s == "geometry":
                blocks.append(shape_features(polys, masks))
            elif s == "color":
                blocks.append(color_features(patches, masks).astype(np.float64))
            elif s == "glcm":
                blocks.append(glcm_feature_set(patches, masks).astype(np.float64))
            elif s == "glrlm":
                blocks.append(glrlm_feature_set(patches, masks).astype(np.float64))
            elif s == "gabor":
                blocks.append(gabor_feature_set(patches, masks).astype(np.float64))
            else:
                raise NotImplementedError(f"oracle for feature set {s!r} not written yet")
        rows.append(np.concatenate(blocks, axis=1))
        cents.append(c)
        keys += [centroid_key(x) for x in c]
    return keys, np.concatenate(cents, 0), np.concatenate(rows, 0), names


# --------------------------------------------------------------------------------------------
# EXTENSION outputs (BASELINE.json north_star items the REFERENCE does not compute: skew / kurtosis, raw / central /
# Hu moments, a contour perimeter, GLCM at distance 2). They have no reference counterpart -- the definitions below
# ARE the specification (SPEC.md section C) -- and are never mixed into the 418-column drop-in schema:
# separate NFX_EXT_* bits, separate matrix (include/nfx.h nfx_compute_ext).
# --------------------------------------------------------------------------------------------
EXT_COLOR_CHANNELS = ("r", "g", "b", "grey", "s", "v", "haematoxylin", "eosin", "dab")
EXT_COLOR_COLUMNS = [f"{m}_{c}" for c in EXT_COLOR_CHANNELS for m in ("skew", "kurtosis")]
EXT_MASK_COLUMNS = ["m00", "m10", "m01", "m20", "m11", "m02", "m30", "m21", "m12", "m03",
                    "mu20", "mu11", "mu02", "mu30", "mu21", "mu12", "mu03"] + [f"hu{k}" for k in range(1, 8)]
EXT_CONTOUR_COLUMNS = ["contour_crack_length", "contour_perimeter"]
EXT_GLCM_OFFSETS = ((0, 1), (1, 1), (1, 0), (1, -1), (0, 2), (2, 2), (2, 0), (2, -2))   # BASELINE config 3: 32 levels, d in {1, 2}
EXT_GLCM_COLUMNS = [f"{f}_{o[0]}_{o[1]}_32" for o in EXT_GLCM_OFFSETS for f in GLCM_FEATURES]
EXT_COLUMNS = {"color_moments": EXT_COLOR_COLUMNS, "mask_moments": EXT_MASK_COLUMNS, "contour": EXT_CONTOUR_COLUMNS,
               "glcm_d2": EXT_GLCM_COLUMNS}
EXT_ORDER = ("color_moments", "mask_moments", "contour", "glcm_d2")


def ext_color_moments(patchs, masks):
    """[C1] masked population skewness m3 / m2^1.5 and excess kurtosis m4 / m2^2 - 3 (scipy.stats.skew / kurtosis with
    their defaults) of r, g, b, grey = ((r+g)+b)/3, s, v, haematoxylin, eosin, dab; float64; m2 == 0 -> NaN. [N,18]."""
    hsv, hed = hsv_from_rgb(patchs), hed_from_rgb(patchs)
    chans = torch.cat([patchs, grey_scale(patchs), hsv[:, 1:3], hed], dim=1).to(torch.float64)   # [N,9,P,P]
    m = masks.to(torch.float64)
    K = m.sum(dim=[1, 2, 3]).view(-1, 1, 1, 1)
    mean = (chans * m).sum(dim=[2, 3], keepdim=True) / K
    d = (chans - mean) * m
    m2, m3, m4 = [(d ** k).sum(dim=[2, 3]) / K.view(-1, 1) for k in (2, 3, 4)]
    with np.errstate(invalid="ignore", divide="ignore"):
        skew, kurt = (m3 / m2 ** 1.5).numpy(), (m4 / (m2 * m2) - 3.0).numpy()
    bad = ~(m2.numpy() > 0)
    skew[bad], kurt[bad] = np.nan, np.nan
    return np.stack([skew, kurt], axis=2).reshape(len(patchs), 18).astype(F32)


def ext_mask_moments(masks):
    """[C2] cv2.moments(mask, binaryImage=True) + cv2.HuMoments on the raster mask: x = column, y = row, float64.
    Raw m_pq = sum x^p y^q (exact integers), central mu_pq about (m10/m00, m01/m00), nu_pq = mu_pq / m00^((p+q)/2+1),
    Hu's seven invariants. Empty mask -> m00 = 0 and NaN for everything that divides by it. [N,24]."""
    out = []
    for m in masks[:, 0].numpy() != 0:
        ys, xs = np.nonzero(m)
        x, y = xs.astype(np.float64), ys.astype(np.float64)
        raw = [float(len(x)), x.sum(), y.sum(), (x * x).sum(), (x * y).sum(), (y * y).sum(), (x ** 3).sum(), (x * x * y).sum(),
               (x * y * y).sum(), (y ** 3).sum()]
        if len(x) == 0:
            out.append(raw + [np.nan] * 14)
            continue
        cx, cy = raw[1] / raw[0], raw[2] / raw[0]
        dx, dy = x - cx, y - cy
        mu = {"20": (dx * dx).sum(), "11": (dx * dy).sum(), "02": (dy * dy).sum(), "30": (dx ** 3).sum(), "21": (dx * dx * dy).sum(),
              "12": (dx * dy * dy).sum(), "03": (dy ** 3).sum()}
        n = {k: v / raw[0] ** ((int(k[0]) + int(k[1])) / 2.0 + 1.0) for k, v in mu.items()}
        a, b = n["30"] + n["12"], n["21"] + n["03"]
        c, d = n["30"] - 3 * n["12"], 3 * n["21"] - n["03"]
        hu = [n["20"] + n["02"],
              (n["20"] - n["02"]) ** 2 + 4 * n["11"] ** 2,
              c * c + d * d,
              a * a + b * b,
              c * a * (a * a - 3 * b * b) + d * b * (3 * a * a - b * b),
              (n["20"] - n["02"]) * (a * a - b * b) + 4 * n["11"] * a * b,
              d * a * (a * a - 3 * b * b) - c * b * (3 * a * a - b * b)]
        out.append(raw + [mu[k] for k in ("20", "11", "02", "30", "21", "12", "03")] + hu)
    return np.array(out, dtype=np.float64).astype(F32)


def ext_contour(masks):
    """[C3] boundary length of the raster mask, two local definitions (pixels outside the window count as background):
    contour_crack_length = number of pixel edges between a set and an unset 4-neighbour (the exact length of the
    boundary cracks); contour_perimeter = Pratt's bit-quad estimate n(Q2) + (n(Q1) + n(Q3) + 2 n(QD)) / sqrt(2) over all
    2 x 2 windows of the zero-padded mask (Q_k = k set pixels, QD = the two diagonal ones). [N,2]."""
    out = []
    for m in masks[:, 0].numpy() != 0:
        p = np.pad(m.astype(np.int64), 1)
        crack = (np.abs(np.diff(p, axis=0)).sum() + np.abs(np.diff(p, axis=1)).sum())
        q = p[:-1, :-1] + p[:-1, 1:] + p[1:, :-1] + p[1:, 1:]
        diag = (q == 2) & (p[:-1, :-1] == p[1:, 1:])
        n1, n2, n3, nd = (q == 1).sum(), ((q == 2) & ~diag).sum(), (q == 3).sum(), diag.sum()
        out.append([float(crack), n2 + (n1 + n3 + 2 * nd) / math.sqrt(2.0)])
    return np.array(out, dtype=np.float64).astype(F32)


def ext_glcm_d2(patchs, masks):
    """[C4] BASELINE config 3's literal variant: 32 grey levels, distances 1 and 2, 4 angles = 8 symmetric masked matrices
    (rules B5 / B6 unchanged, offsets EXT_GLCM_OFFSETS) x 14 Haralick features. [N,112]."""
    grey = grey_scale(patchs)
    return np.concatenate([glcm_features(glcm(grey, off, 32, masks)) for off in EXT_GLCM_OFFSETS], axis=1).astype(F32)


def extract_ext(rings, image_hwc, ext_sets, patch_size=64):
    """Extension matrix for the given sets (EXT_ORDER order): (names, [N,F] f32)."""
    cents, polys, patches, masks = load_image_dataset(rings, image_hwc, patch_size)
    fn = {"color_moments": lambda: ext_color_moments(patches, masks), "mask_moments": lambda: ext_mask_moments(masks),
          "contour": lambda: ext_contour(masks), "glcm_d2": lambda: ext_glcm_d2(patches, masks)}
    sets = [s for s in EXT_ORDER if s in ext_sets]
    return [c for s in sets for c in EXT_COLUMNS[s]], np.concatenate([fn[s]() for s in sets], axis=1)

"""Seeded synthetic cases shared by the CPU and GPU tests."""
import numpy as np

from nfx import synth


def small_case(n=230, size=640, seed=7, patch=64):
    """Tile + star polygons incl. border-touching ones, plus hand-made edge cases appended."""
    tile = synth.synth_tile(size, size, seed)
    xy, off = synth.synth_polygons(n, size, size, seed, patch=patch, border_frac=0.08)
    rings = synth.rings_of(xy, off)
    extra = []
    # axis-aligned rectangle, integer vertices (edges exactly on sample points)
    extra.append(np.array([[100, 200], [120, 200], [120, 210], [100, 210], [100, 200]], np.float32))
    # triangle with half-integer vertices
    extra.append(np.array([[300.5, 300.5], [320.5, 305.5], [305.5, 322.5], [300.5, 300.5]], np.float32))
    # non-convex "C" shape
    extra.append(np.array([[400, 400], [430, 400], [430, 410], [410, 410], [410, 430], [430, 430],
                           [430, 440], [400, 440], [400, 400]], np.float32))
    # self-intersecting bow-tie (even-odd rule matters)
    extra.append(np.array([[200, 100], [230, 130], [230, 100], [200, 130], [200, 100]], np.float32))
    # larger than the patch (clipped by the window)
    t = np.linspace(0, 2 * np.pi, 41)
    extra.append(np.stack([320 + 50 * np.cos(t), 320 + 45 * np.sin(t)], 1).astype(np.float32))
    # tiny polygon that covers no sample point (empty mask -> NaN features)
    extra.append(np.array([[500.2, 500.2], [500.6, 500.2], [500.6, 500.6], [500.2, 500.2]], np.float32))
    # one-pixel mask
    extra.append(np.array([[150.5, 150.5], [151.5, 150.5], [151.5, 151.5], [150.5, 151.5], [150.5, 150.5]], np.float32))
    # thin horizontal sliver (rank-1 covariance)
    extra.append(np.array([[250, 260.25], [275, 260.25], [275, 260.75], [250, 260.75], [250, 260.25]], np.float32))
    # corner nuclei: negative window origin with a fractional part (trunc-toward-zero quirk)
    extra.append(np.stack([10.3 + 8 * np.cos(t), 11.7 + 8 * np.sin(t)], 1).astype(np.float32))
    extra.append(np.stack([size - 9.6 + 7 * np.cos(t), size - 12.2 + 9 * np.sin(t)], 1).astype(np.float32))
    # unclosed ring (no closing duplicate): centroid uses the ring as stored
    extra.append(np.array([[350, 100], [370, 102], [375, 120], [355, 125], [345, 110]], np.float32))
    rings = rings + extra
    return tile, rings


def stress_case(n=24, size=1024, seed=5, patch=256):
    """config-5-like: large irregular nuclei, 500-vertex polygons, P=256."""
    tile = synth.synth_tile(size, size, seed)
    xy, off = synth.synth_polygons(n, size, size, seed, patch=patch, r0_range=(40.0, 110.0),
                                   v_range=(500, 500), border_frac=0.1, rough=0.25,
                                   harmonics=(3, 7, 19))
    return tile, synth.rings_of(xy, off)

"""Host logic that needs no GPU: the C-ABI library loads and exports every symbol include/nfx.h
declares, the schema strings equal the reference's (hard-coded from the cited source lines in the
oracle), key strings follow Rust's f32 Display, the partition rule keeps chunks whole."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import nfx
import nfx_oracle as o
from nfx._lib import LIB_PATH, SYMBOLS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(libnfx):
    hdr = open(os.path.join(ROOT, "include", "nfx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nfx_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(SYMBOLS), declared ^ set(SYMBOLS)
    raw = C.CDLL(LIB_PATH)
    for name in declared:
        assert getattr(raw, name) is not None
    assert libnfx.nfx_version().decode() == "nfx 0.1.0 (sm_100a)"


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_gpu_means_error_not_fallback(libnfx):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(nfx.NfxError) as e:
        nfx.Extractor(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_column_names_match_the_reference(libnfx):
    assert nfx.feature_names(nfx.FS_GEOMETRY) == o.SHAPE_COLUMNS
    assert nfx.feature_names(nfx.FS_COLOR) == o.COLOR_COLUMNS
    assert nfx.feature_names(nfx.FS_GLCM) == o.GLCM_COLUMNS
    assert nfx.feature_names(nfx.FS_GLRLM) == o.GLRLM_COLUMNS
    assert nfx.feature_names(nfx.FS_GABOR) == o.GABOR_COLUMNS
    allc = nfx.feature_names(nfx.FS_ALL)
    assert len(allc) == 418 and len(set(allc)) == 418
    assert allc == [c for s in o.FLAT_ORDER for c in o.SET_COLUMNS[s]]
    # spot checks against the literal strings of the source
    assert "information_measure_correlation1_1_-1_254" in allc      # texture.rs:146
    assert "gray_level_nonuniformity_-1_1" in allc                  # texture.rs:252, 176
    assert "gabor_angle_315_frequency_0.5_variance" in allc         # texture.rs:346-361
    assert libnfx.nfx_feature_name(nfx.FS_ALL, 418) is None and libnfx.nfx_feature_name(nfx.FS_ALL, -1) is None


def test_feature_set_names_and_parsing(libnfx):
    assert nfx.parse_feature_sets(["ALL"]) == nfx.FS_ALL                       # args.rs:21 to_lowercase
    assert nfx.parse_feature_sets(["texture"]) == nfx.FS_TEXTURE
    assert nfx.parse_feature_sets(["Geometry", "color"]) == nfx.FS_GEOMETRY | nfx.FS_COLOR
    with pytest.raises(nfx.NfxError):
        nfx.parse_feature_sets(["glcm", "texture"])                            # duplicate -> main.rs:89 fails
    with pytest.raises(nfx.NfxError):
        nfx.parse_feature_sets(["colour"])                                     # args.rs:29
    with pytest.raises(nfx.NfxError):
        nfx.parse_feature_sets([])                                             # main.rs:76 features[0]
    names = [libnfx.nfx_feature_set_name(b).decode() for b in (1, 2, 4, 8, 16)]
    assert names == ["geometry", "color", "GLCM", "GLRLM", "gabor filter"]
    assert o.flat(["all"]) == list(o.FLAT_ORDER)


def test_centroid_key_matches_rust_display(libnfx):
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        rng.uniform(0, 1e5, 3000).astype(np.float32), rng.uniform(-1, 1, 500).astype(np.float32),
        (rng.integers(0, 200000, 500) / 4).astype(np.float32),
        np.array([0, -0.0, 1e-7, 1e10, 123456.79, 16777216, 0.1, 1 / 3, 2 ** -10, 3.4e38, 1e-45], np.float32)])
    for a, b in zip(vals[::2], vals[1::2]):
        assert nfx.centroid_key(a, b) == o.centroid_key((a, b))
    assert nfx.centroid_key(1024.0, 33.5) == "1024,33.5"
    buf = C.create_string_buffer(4)
    assert libnfx.nfx_centroid_key(1024.0, 33.5, buf, 4) < 0


def test_partition_keeps_reference_chunks_whole(libnfx):
    for n, B, parts in [(1050, 100, 4), (99, 100, 8), (5_000_000, 100, 8), (0, 100, 2), (1234, 37, 3), (800, 100, 8)]:
        b = nfx.partition(n, B, parts)
        assert b[0] == 0 and b[-1] == n and len(b) == parts + 1
        assert all(x <= y for x, y in zip(b, b[1:]))
        assert all(x % B == 0 for x in b[:-1])            # every chunk [kB,(k+1)B) lives on one GPU
        sizes = [y - x for x, y in zip(b, b[1:])]
        assert max(sizes) - min(sizes) <= 2 * B
    with pytest.raises(nfx.NfxError):
        nfx.partition(10, 0, 2)


def test_pack_polygons_roundtrip():
    rings = [np.arange(8, dtype=np.float32).reshape(4, 2), np.ones((3, 2), np.float32)]
    xy, off = nfx.pack_polygons(rings)
    assert off.tolist() == [0, 4, 7] and xy.dtype == np.float32 and xy.shape == (7, 2)


def test_extension_schema_is_separate_from_the_drop_in_schema():
    """NFX_EXT_*: north_star items the reference does not compute. Their names come from nfx_ext_feature_name, match the
    oracle's EXT_COLUMNS, and never appear among the 418 drop-in columns (except the distance-1 GLCM columns at 32 levels,
    which BASELINE config 3's literal variant repeats)."""
    import nfx
    import nfx_oracle as o
    from nfx._lib import lib
    L = lib()
    assert L.nfx_ext_feature_count(nfx.EXT_ALL) == 18 + 24 + 2 + 112
    assert L.nfx_ext_feature_count(0x10) == -1 and L.nfx_ext_feature_name(nfx.EXT_ALL, 156) is None
    names = [L.nfx_ext_feature_name(nfx.EXT_ALL, i).decode() for i in range(156)]
    assert names == [c for s in o.EXT_ORDER for c in o.EXT_COLUMNS[s]]
    bits = {"color_moments": nfx.EXT_COLOR_MOMENTS, "mask_moments": nfx.EXT_MASK_MOMENTS, "contour": nfx.EXT_CONTOUR, "glcm_d2": nfx.EXT_GLCM_D2}
    for s, b in bits.items():
        assert [L.nfx_ext_feature_name(b, i).decode() for i in range(L.nfx_ext_feature_count(b))] == o.EXT_COLUMNS[s]
    drop_in = set(nfx.feature_names(nfx.FS_ALL))
    shared = [n for n in names if n in drop_in]
    assert len(shared) == 56 and all(n.endswith("_32") for n in shared)

"""Output assembly: Rust `Display` of f32 (the polars 0.32 CSV cell and the centroid
key, oracle/SPEC.md B12) on the host and on the GPU, and the CSV rows formatted on the device."""
import numpy as np
import pytest

import nfx
from nfx import synth


def _shortest(v: np.float32) -> str:
    """Independent statement of the rule: shortest digits that round-trip (numpy's Dragon4, `unique`), positional,
    no trailing '.0' -- what Rust's Grisu/Dragon `Display` prints."""
    if np.isnan(v):
        return "NaN"
    if np.isinf(v):
        return "inf" if v > 0 else "-inf"
    return np.format_float_positional(v, unique=True, trim="-")


def _patterns(rng, n):
    bits = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    special = [0, 0x80000000, 1, 2, 0x007FFFFF, 0x00800000, 0x00800001, 0x7F7FFFFF, 0x7F800000, 0xFF800000, 0x7FC00000,
               0x3F800000, 0x3DCCCCCD, 0x4B800000, 0x4B7FFFFF, 0x5F000000, 0x41200000, 0x42C80000, 0x447A0000, 0x501502F9]
    for e in range(0, 255):                      # around every power of two: the gap below is half as wide
        special += [e << 23, (e << 23) + 1, max((e << 23) - 1, 0)]
    return np.concatenate([np.array(special, np.uint32), bits])


def test_f32_display_is_shortest_roundtrip(libnfx):
    rng = np.random.default_rng(7)
    vals = _patterns(rng, 60000).view(np.float32)
    feat = (rng.uniform(-1, 1, 20000) * 10.0 ** rng.integers(-8, 9, 20000)).astype(np.float32)      # feature-like magnitudes
    cents = (rng.integers(0, 400000, 20000) / 4.0).astype(np.float32)                                # centroid-like
    for v in np.concatenate([vals, feat, cents]):
        got = nfx.format_f32(v)
        assert got == _shortest(v), (v.view(np.uint32), got)
        if np.isfinite(v):
            assert np.float32(got).tobytes() == np.float32(v).tobytes()
            assert "e" not in got and not got.endswith(".0")


def test_csv_header(libnfx):
    h = nfx.csv_header(nfx.parse_feature_sets(["color", "geometry"])).decode()
    cols = h.rstrip("\n").split(",")
    assert h.endswith("\n") and cols[0] == "centroid" and cols[1] == "area" and cols[13] == "mean_r" and len(cols) == 31
    assert len(nfx.csv_header(nfx.parse_feature_sets(["all"])).decode().split(",")) == 419


def _rows(cent, feat):
    out = []
    for c, f in zip(cent, feat):
        out.append('"%s,%s"' % (nfx.format_f32(c[0]), nfx.format_f32(c[1])) + "".join("," + nfx.format_f32(v) for v in f) + "\n")
    return "".join(out).encode()


@pytest.mark.gpu
def test_device_formatter_matches_host_on_bit_patterns(libnfx):
    rng = np.random.default_rng(3)
    bits = _patterns(rng, 200000)
    F = 37                                           # more than one 32-cell chunk per row, ragged tail
    n = len(bits) // (F + 2)
    m = bits[:n * (F + 2)].view(np.float32).reshape(n, F + 2)
    with nfx.Extractor(0) as ex:
        got = ex.csv_format(m[:, :2], m[:, 2:])
        assert got == _rows(m[:, :2], m[:, 2:])
        assert ex.csv_format(m[:5, :2], np.zeros((5, 0), np.float32)) == _rows(m[:5, :2], np.zeros((5, 0), np.float32))
        assert ex.csv_format(np.zeros((0, 2), np.float32), np.zeros((0, 3), np.float32)) == b""


@pytest.mark.gpu
def test_csv_rows_of_resident_result(libnfx):
    tile = synth.synth_tile(512, 512, 2)
    xy, off = synth.synth_polygons(700, 512, 512, 2, border_frac=0.05)
    with nfx.Extractor(0, 64, 100) as ex:
        ex.upload_tile(tile)
        keys, cent, feat, names = ex.extract(xy, off, ["all"])
        want = _rows(cent, feat)
        assert ex.csv_rows() == want
        parts = b"".join(ex.csv_rows(lo, min(lo + 129, 700)) for lo in range(0, 700, 129))
        assert parts == want
        assert ex.csv_rows(5, 5) == b""
        with pytest.raises(nfx.NfxError):
            ex.csv_rows(0, 701)
        first = want.split(b"\n")[0].decode()
        assert first.startswith('"' + keys[0] + '",') and first.count(",") == 419

"""GeoJSON -> CSR packing: nfx_geojson_parse / nfx_parse_f32 against the restatement of
the reference's reader (oracle/geojson_ref.py: serde_json 1.0.107 number rule + the model of
src/geojson.rs:8-24). Host code: no GPU needed. Bit-exact on every coordinate."""
import json

import numpy as np
import pytest

import geojson_ref as gr
import nfx
from nfx import synth


def _same(a, b):
    return a.shape == b.shape and a.tobytes() == b.tobytes()


def _check(text, threads=(1, 3, 8)):
    want = gr.load_text(text)
    for t in threads:
        got = nfx.geojson_pack(text, t)
        for g, w in zip(got, want):
            assert _same(g, w), (t, g[:4], w[:4])
    return want


TOKENS = ["0", "-0", "1", "7", "-13", "16777217", "9007199254740993", "18446744073709551615", "18446744073709551616",
          "184467440737095516159.5", "-9223372036854775808", "-9223372036854775809", "1e-400", "0e999999999999",
          "123456789012345678901234567890", "0.1", "12345.678901234567", "4.9e-324", "3.4028235e38", "1E+2", "1e-0",
          "0." + "0" * 44 + "1", "123456789012345678.5e-3", "1.17549435e-38", "0.30000000000000004", "5e-46", "7.0e-46"]
BAD_TOKENS = ["01", "1.", "-", ".5", "1e", "+1", "1e+", "0x10", "NaN", "Infinity", "1e400", "--1", "1.e3", ""]


def test_number_rule_known_tokens(libnfx):
    for t in TOKENS:
        assert nfx.parse_f32(t).tobytes() == gr.serde_f32(t).tobytes(), t
    for t in BAD_TOKENS:
        with pytest.raises(Exception):
            gr.serde_f32(t)
        with pytest.raises(nfx.NfxError):
            nfx.parse_f32(t)
    # the rule is not a correctly rounded strtod for long significands; find one such token and pin both results
    rng = np.random.default_rng(5)
    found = 0
    for _ in range(40000):
        t = "%d.%019d" % (rng.integers(1, 99999), rng.integers(0, 10**18))
        a = nfx.parse_f32(t)
        assert a.tobytes() == gr.serde_f32(t).tobytes(), t
        found += a.tobytes() != np.float32(float(t)).tobytes()
    assert found < 40            # rare (double rounding), and identical to the restatement when it happens


def test_number_rule_random_tokens(libnfx):
    rng = np.random.default_rng(11)
    for k in range(60000):
        m = k % 5
        if m == 0:
            t = repr(float(rng.uniform(0, 1e5)))                       # 15-17 digit doubles (QuPath exports)
        elif m == 1:
            t = "%.2f" % rng.uniform(-1e5, 1e5)
        elif m == 2:
            t = repr(float(np.float32(rng.uniform(0, 1e5))))
        elif m == 3:
            t = "%d.%d" % (rng.integers(0, 10**9), rng.integers(0, 10**12)) + ("e%d" % rng.integers(-40, 40) if rng.random() < .5 else "")
        else:
            t = "%de%d" % (rng.integers(0, 2**63), rng.integers(-60, 30))
        assert nfx.parse_f32(t).tobytes() == gr.serde_f32(t).tobytes(), t


def _doc(rings, extra_top=None, props=True, indent=None, holes=False, z=False):
    feats = []
    for i, r in enumerate(rings):
        pts = [[float(x), float(y)] + ([0.5] if z else []) for x, y in r]
        coords = [pts] + ([[[1.0, 2.0], [3.0, 4.0], [1.0, 2.0]]] if holes and i % 3 == 0 else [])
        ft = {"type": "Feature", "id": "%08x-cell" % i,
              "geometry": {"type": "Polygon", "coordinates": coords},
              "bbox": [float(r[:, 0].min()), float(r[:, 1].min()), float(r[:, 0].max()), float(r[:, 1].max())]}
        if props:
            ft["properties"] = {"objectType": "detection", "name": 'say "[{" \\ ]}', "coordinates": [[1, 2]],
                                "measurements": [{"name": "a[", "value": 1e-3}, None, True],
                                "classification": {"name": "Tumor", "colorRGB": -3670016}}
        feats.append(ft)
    d = {"type": "FeatureCollection"}
    d.update(extra_top or {})
    d["features"] = feats
    return json.dumps(d, indent=indent)


def test_pack_matches_reference_reader(libnfx):
    xy, off = synth.synth_polygons(1500, 4096, 4096, 3)
    rings = [np.asarray(r, np.float64) * 1.0000001 for r in synth.rings_of(xy, off)]     # long decimal expansions
    text = _doc(rings)
    assert len(text) > 400_000                                                          # several chunks per thread
    w = _check(text)
    assert len(w[1]) == 1501 and w[3].max() == 1
    _check(_doc(rings[:200], indent=2, holes=True, z=True))
    _check(_doc(rings[:50], props=False, extra_top={"crs": {"properties": {"name": "x"}}, "bbox": [0, 0, 1, 1],
                                                     "other": [{"a": [1, 2]}, {"b": {}}]}))
    _check(json.dumps({"features": [], "type": "FeatureCollection"}))
    _check('  {"features"\n:\t[ ]\r\n }  ')
    # keys in any order, features not last
    d = json.loads(_doc(rings[:30]))
    _check(json.dumps({"features": d["features"], "z": [[{"q": 1}]], "type": "FeatureCollection"}))
    # ring 0 may be empty; rings i32 counts holes
    t = '{"features":[{"bbox":[1,2],"geometry":{"coordinates":[[],[[1,2],[3,4]]],"type":"Polygon"}}]}'
    got = _check(t)
    assert got[1].tolist() == [0, 0] and got[3].tolist() == [2] and np.isnan(got[2][0, 2:]).all()


@pytest.mark.parametrize("text,msg", [
    ('{"type":"FeatureCollection"}', "missing field `features`"),
    ('{"features":[{"geometry":{"type":"Polygon","coordinates":[[[1,2]]]}}]}', "missing field `bbox`"),
    ('{"features":[{"bbox":[],"geometry":{"coordinates":[[[1,2]]]}}]}', "missing field `type`"),
    ('{"features":[{"bbox":[],"geometry":{"type":"Polygon"}}]}', "missing field `coordinates`"),
    ('{"features":[{"bbox":[]}]}', "missing field `geometry`"),
    ('{"features":[{"bbox":[],"bbox":[],"geometry":{"type":"P","coordinates":[[[1,2]]]}}]}', "duplicate field `bbox`"),
    ('{"features":[{"bbox":[],"geometry":{"type":"Point","coordinates":[1,2]}}]}', "invalid type"),
    ('{"features":[{"bbox":[],"geometry":{"type":"MultiPolygon","coordinates":[[[[1,2],[3,4]]]]}}]}', "invalid type"),
    ('{"features":[{"bbox":null,"geometry":{"type":"P","coordinates":[[[1,2]]]}}]}', "invalid type"),
    ('{"features":[{"bbox":["1"],"geometry":{"type":"P","coordinates":[[[1,2]]]}}]}', "invalid type"),
    ('{"features":[{"bbox":[],"geometry":{"type":"P","coordinates":[[[1]]]}}]}', "fewer than two"),
    ('{"features":[{"bbox":[],"geometry":{"type":"P","coordinates":[]}}]}', "without a ring"),
    ('{"features":[{"bbox":[],"geometry":{"type":"P","coordinates":[[[1,2]]]}}', "EOF"),
    ('{"features":[{"bbox":[],"geometry":{"type":"P","coordinates":[[[1,2]]]}}]} x', "trailing characters"),
    ('{"features":[{"bbox":[],"geometry":{"type":"P","coordinates":[[[1,2]]]}} {"bbox":[]}]}', "expected `,` or `]`"),
    ('{"features":[{"bbox":[],"geometry":{"type":"P","coordinates":[[[01,2]]]}}]}', "invalid number"),
    ('{"features":[1,2]}', "expected struct Feature"),
    ('{"features":[],"features":[]}', "duplicate field `features`"),
])
def test_pack_rejects_what_serde_rejects(libnfx, text, msg):
    with pytest.raises(Exception):
        gr.load_text(text)
    for t in (1, 4):
        with pytest.raises(nfx.NfxError) as e:
            nfx.geojson_pack(text, t)
        assert msg in str(e.value), str(e.value)
    assert "line 1 column" in str(e.value)


def test_pack_error_position_and_order(libnfx):
    xy, off = synth.synth_polygons(3000, 4096, 4096, 9)
    rings = [np.asarray(r, np.float64) for r in synth.rings_of(xy, off)]
    d = json.loads(_doc(rings, props=False))
    del d["features"][2500]["bbox"]
    d["features"][700]["geometry"]["coordinates"] = [[[1]]]
    text = json.dumps(d, indent=1)
    msgs = set()
    for t in (1, 2, 7, 16):
        with pytest.raises(nfx.NfxError) as e:
            nfx.geojson_pack(text, t)
        msgs.add(str(e.value))
    assert len(msgs) == 1 and "fewer than two" in msgs.pop()      # the first error in input order, whatever the thread count


def test_pack_fuzz_accepts_and_rejects_like_the_reference_reader(libnfx):
    """Random one- and two-byte mutations of a valid document: the packer must accept exactly what the restated
    serde reader accepts, and produce the same arrays when both accept (strings with escapes, nested unknown values,
    holes, 3-D positions and an empty bbox are all in the seed document)."""
    base = ('{"type":"FeatureCollection","features":[{"type":"Feature","id":"a","bbox":[1.5,2,3e1,4],"geometry":{"type":"Polygon",'
            '"coordinates":[[[10.25,20.5],[30,20.125],[25.5,40],[10.25,20.5]],[[1,2],[3,4],[1,2]]]},"properties":{"name":"x \\"q\\" [{",'
            '"v":[1,{"k":null},true]}},{"bbox":[],"geometry":{"coordinates":[[[0.1,0.2],[3,4,5]]],"type":"Polygon"}}],"extra":{"a":[1,2,{"b":"]"}]}}')
    rng = np.random.default_rng(123)
    alphabet = list('{}[]",:.-+eE0123456789 \n\\tnulrfas')
    accepted = 0
    for _ in range(2500):
        s = list(base)
        for _ in range(rng.integers(1, 3)):
            k = rng.integers(0, len(s))
            op = rng.integers(0, 3)
            if op == 0:
                del s[k]
            elif op == 1:
                s.insert(k, alphabet[rng.integers(0, len(alphabet))])
            else:
                s[k] = alphabet[rng.integers(0, len(alphabet))]
        t = "".join(s)
        try:
            want = gr.load_text(t)
        except Exception:
            want = None
        try:
            got = nfx.geojson_pack(t, 2)
        except nfx.NfxError:
            got = None
        assert (want is None) == (got is None), t
        if want is not None:
            accepted += 1
            assert all(_same(g, w) for g, w in zip(got, want)), t
    assert accepted > 300

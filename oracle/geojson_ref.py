"""TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU restatement of how the reference turns a GeoJSON text into polygons:
`serde_json::from_reader::<FeatureCollection>` (src/main.rs:37-42) with the model of src/geojson.rs:8-24,
then ring 0 of every feature (src/utils.rs:54-60).

PARITY UNPINNED for the number rule: it lives in serde_json 1.0.107 (Cargo.lock:2298), which is not under
/root/reference. Restated from its published source (src/de.rs: parse_integer, parse_long_integer,
parse_decimal, parse_decimal_overflow, parse_exponent, f64_from_parts; `float_roundtrip` is off because
Cargo.toml:25 enables no feature) -- oracle/SPEC.md B11:
  * integer digits accumulate in a u64; the first digit that would overflow and all further integer digits
    only bump a decimal exponent;
  * fraction digits accumulate too (exponent - 1 each) until they would overflow, the rest are dropped;
  * value = (significand as f64) * 10^e or / 10^-e with ONE f64 operation (|e| <= 308), i.e. NOT correctly
    rounded for 16+ digit significands;
  * an f32 field takes `v as f32` of that f64; tokens without '.', 'e' are u64 / i64 and cast directly.
The structure (objects, arrays, strings) is read with Python's json module.
"""
import json
import math
import re

import numpy as np

U64_MAX = (1 << 64) - 1
_NUM = re.compile(r"-?(0|[1-9][0-9]*)(\.[0-9]+)?([eE][+-]?[0-9]+)?\Z")


def _int_to_f32(n: int) -> np.float32:
    """Exact integer -> nearest f32, ties to even (Rust `u64 as f32` / `i64 as f32`)."""
    if n == 0:
        return np.float32(0.0)
    s, a = (-1 if n < 0 else 1), abs(n)
    bl = a.bit_length()
    if bl > 24:
        sh = bl - 24
        q, r = a >> sh, a & ((1 << sh) - 1)
        half = 1 << (sh - 1)
        if r > half or (r == half and (q & 1)):
            q += 1
        a = q << sh
    return np.float32(s * float(a))      # a has <= 24 significant bits: exact in f64 and in f32


def _f64_from_parts(positive: bool, sig: int, exp: int) -> float:
    f = float(sig)                        # u64 as f64: nearest, ties to even (Python int -> float is the same)
    while True:
        a = abs(exp)
        if a <= 308:
            p = float("1e%d" % a)
            if exp >= 0:
                f = f * p
                if math.isinf(f):
                    raise ValueError("number out of range")
            else:
                f = f / p
            break
        if f == 0.0:
            break
        if exp >= 0:
            raise ValueError("number out of range")
        f = f / 1e308
        exp += 308
    return f if positive else -f


def serde_f32(token: str) -> np.float32:
    """One JSON number token -> the f32 serde_json 1.0.107 hands to a `f32` field."""
    if not _NUM.match(token):
        raise ValueError("invalid number")
    positive = not token.startswith("-")
    t = token.lstrip("-")
    m = re.match(r"([0-9]+)(?:\.([0-9]+))?(?:[eE]([+-]?[0-9]+))?\Z", t)
    ip, fp, ep = m.group(1), m.group(2), m.group(3)
    sig, exp, overflowed = 0, 0, False
    for ch in ip:
        d = ord(ch) - 48
        if overflowed:
            exp += 1
        elif sig * 10 + d > U64_MAX:
            overflowed = True
            exp += 1
        else:
            sig = sig * 10 + d
    is_float = overflowed or fp is not None or ep is not None
    if fp is not None:
        for ch in fp:
            d = ord(ch) - 48
            if sig * 10 + d > U64_MAX:
                break                     # parse_decimal_overflow: the remaining digits are skipped
            sig = sig * 10 + d
            exp -= 1
    if ep is not None:
        e = int(ep)
        if abs(e) > 2**31 - 1:
            if sig != 0 and e > 0:
                raise ValueError("number out of range")
            return np.float32(0.0 if positive else -0.0)
        exp = max(-2**31, min(2**31 - 1, exp + e))
    if not is_float:
        if positive:
            return _int_to_f32(sig)
        if sig == 0 or sig > (1 << 63):   # -0 and values below i64::MIN take the f64 road
            return np.float32(-float(sig))
        return _int_to_f32(-sig)
    with np.errstate(over="ignore"):
        return np.float32(_f64_from_parts(positive, sig, exp))


class _Tok(str):
    """A number token kept as text through json.loads."""


def load_text(text: str):
    """-> (xy f32 [nv,2], off i64 [n+1], bbox f32 [n,4] NaN-padded, rings i32 [n]); raises ValueError like serde would."""
    def pairs(items):
        keys = [k for k, _ in items]
        return {"__dups__": [k for k in set(keys) if keys.count(k) > 1], **dict(items)}

    doc = json.loads(text, parse_float=_Tok, parse_int=_Tok, parse_constant=lambda c: (_ for _ in ()).throw(ValueError("expected value")),
                     object_pairs_hook=pairs)
    if not isinstance(doc, dict) or "features" not in doc:
        raise ValueError("missing field `features`")
    if "features" in doc["__dups__"]:
        raise ValueError("duplicate field `features`")
    if not isinstance(doc["features"], list):
        raise ValueError("invalid type: expected a sequence")
    xy, off, bbox, rings = [], [0], [], []

    def f32(v):
        if not isinstance(v, _Tok):
            raise ValueError("invalid type: expected f32")
        return serde_f32(str(v))

    for ft in doc["features"]:
        if not isinstance(ft, dict):
            raise ValueError("invalid type: expected struct Feature")
        for k in ("bbox", "geometry"):
            if k in ft["__dups__"]:
                raise ValueError("duplicate field `%s`" % k)
            if k not in ft:
                raise ValueError("missing field `%s`" % k)
        if not isinstance(ft["bbox"], list):
            raise ValueError("invalid type: expected a sequence")
        bb = [f32(v) for v in ft["bbox"]]
        g = ft["geometry"]
        if not isinstance(g, dict):
            raise ValueError("invalid type: expected struct Geometry")
        for k in ("type", "coordinates"):
            if k in g["__dups__"]:
                raise ValueError("duplicate field `%s`" % k)
            if k not in g:
                raise ValueError("missing field `%s`" % k)
        if not isinstance(g["type"], str) or isinstance(g["type"], _Tok):
            raise ValueError("invalid type: expected a string")
        co = g["coordinates"]
        if not isinstance(co, list) or any(not isinstance(r, list) or any(not isinstance(p, list) for p in r) for r in co):
            raise ValueError("invalid type: expected a sequence")
        pts = [[f32(v) for v in p] for r in co for p in r]   # every number must deserialise, not only ring 0
        if not co:
            raise ValueError("feature without a ring")
        n0 = len(co[0])
        for p in pts[:n0]:
            if len(p) < 2:
                raise ValueError("a position of ring 0 has fewer than two numbers")
            xy.append((p[0], p[1]))
        off.append(off[-1] + n0)
        bbox.append((bb + [np.float32("nan")] * 4)[:4])
        rings.append(len(co))
    return (np.array(xy, np.float32).reshape(-1, 2), np.array(off, np.int64),
            np.array(bbox, np.float32).reshape(-1, 4), np.array(rings, np.int32))

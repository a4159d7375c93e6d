"""Throughput of the GeoJSON -> CSR packer (nfx_geojson_parse) on a synthetic QuPath-like export.
usage: python scripts/bench_geojson.py [nuclei] -- prints one JSON line per thread count."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "nuclei-feature-extraction_b200"))
import numpy as np

import nfx
from nfx import synth


def make_text(n, seed=4):
    xy, off = synth.synth_polygons(n, 100_000, 100_000, seed)
    return make_text_from(xy, off)


def make_text_from(xy, off):
    """A QuPath-like export of these rings (coordinates printed as shortest-repr doubles)."""
    n = len(off) - 1
    parts = ['{"type":"FeatureCollection","features":[']
    xs = xy.astype(np.float64)
    for i in range(n):
        r = xs[off[i]:off[i + 1]]
        pts = ",".join("[%r,%r]" % (float(x), float(y)) for x, y in r)
        parts.append('%s{"type":"Feature","id":"%08x","geometry":{"type":"Polygon","coordinates":[[%s]]},"bbox":[%r,%r,%r,%r],'
                     '"properties":{"objectType":"detection","classification":{"name":"Tumor","colorRGB":-3670016}}}'
                     % ("," if i else "", i, pts, float(r[:, 0].min()), float(r[:, 1].min()), float(r[:, 0].max()), float(r[:, 1].max())))
    parts.append("]}")
    return "".join(parts).encode()


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    text = make_text(n)
    cores = os.cpu_count()
    t0 = time.perf_counter(); json.loads(text); t_py = time.perf_counter() - t0
    print(json.dumps({"impl": "python json.loads (1 thread, no f32 rule)", "MB": len(text) / 1e6, "s": t_py, "MB_per_s": len(text) / 1e6 / t_py}))
    for t in sorted({1, 2, 4, 8, 16, 32, cores}):
        if t > cores:
            continue
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); out = nfx.geojson_pack(text, t); best = min(best, time.perf_counter() - t0)
        print(json.dumps({"impl": "nfx_geojson_parse", "threads": t, "nuclei": n, "vertices": int(out[1][-1]), "MB": round(len(text) / 1e6, 1),
                          "s": round(best, 4), "MB_per_s": round(len(text) / 1e6 / best, 1), "nuclei_per_s": round(n / best)}))

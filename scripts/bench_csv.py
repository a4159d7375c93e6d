"""Output assembly throughput: CSV rows formatted on the GPU from the resident result (nfx_csv_rows) against
pandas.to_csv on the host cores. usage: python scripts/bench_csv.py [nuclei] [feature sets...]"""
import io
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "nuclei-feature-extraction_b200"))
import numpy as np

import nfx
from nfx import synth

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    sets = sys.argv[2:] or ["all"]
    tile = synth.synth_tile(8192, 8192, 1)
    xy, off = synth.synth_polygons(n, 8192, 8192, 1)
    with nfx.Extractor(0, 64, 100) as ex:
        ex.upload_tile(tile)
        keys, cent, feat, names = ex.extract(xy, off, sets)
        F = feat.shape[1]
        block = max(1, min(n, (256 << 20) // (12 * (F + 2))))
        buf = nfx.pinned_empty((block * (F + 2) * 14,), np.uint8)
        ex.csv_rows(0, min(block, n), buf)                      # warm-up (allocations)
        ex.profile(True); ex.profile_reset()
        t0 = time.perf_counter()
        total = 0
        for lo in range(0, n, block):
            total += len(ex.csv_rows(lo, min(lo + block, n), buf, view=True))
        dt = time.perf_counter() - t0
        prof = {k: v for k, v in ex.profile_get().items() if k.startswith("k_csv")}
        kern_ms = sum(v[1] for v in prof.values())
        print(json.dumps({"impl": "nfx_csv_rows (GPU format + D2H of text)", "nuclei": n, "cols": F + 1, "MB": round(total / 1e6, 1),
                          "s": round(dt, 4), "rows_per_s": round(n / dt), "cells_per_s": round(n * (F + 2) / dt), "MB_per_s": round(total / 1e6 / dt, 1),
                          "kernels_ms": {k: round(v[1], 3) for k, v in prof.items()},
                          "kernel_text_GB_per_s": round(total / 1e9 / (kern_ms / 1e3), 1) if kern_ms else None}))
    import pandas as pd
    m = min(n, 20000)
    df = pd.DataFrame(feat[:m], columns=names)
    df.insert(0, "centroid", keys[:m])
    t0 = time.perf_counter(); s = io.StringIO(); df.to_csv(s, index=False); dt = time.perf_counter() - t0
    print(json.dumps({"impl": "pandas.to_csv (1 host thread)", "nuclei": m, "s": round(dt, 3), "rows_per_s": round(m / dt), "cells_per_s": round(m * (F + 2) / dt)}))

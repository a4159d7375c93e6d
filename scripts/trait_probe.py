"""Where the time of one trait-level call goes: the C call alone (prepared buffers) for several chunk sizes."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "nuclei-feature-extraction_b200"))
import ctypes as C
import numpy as np
import nfx
from nfx import synth
from nfx._lib import lib

P = 64
for B in (100, 400, 1600):
    tile = synth.synth_tile(4096, 4096, 2)
    xy, off = synth.synth_polygons(B, 4096, 4096, 2)
    ex = nfx.Extractor(0, P, B)
    ex.upload_tile(tile); ex.upload_polygons(xy, off)
    m8 = ex.rasterize(); p8 = ex.gather_patches()
    patchs = nfx.pinned_empty((B, 3, P, P), np.float32); masks = nfx.pinned_empty((B, 1, P, P), np.float32)
    patchs[:] = np.transpose(p8.reshape(B, P, P, 3), (0, 3, 1, 2)).astype(np.float32) / np.float32(255)
    masks[:, 0] = m8.reshape(B, P, P)
    cents = np.zeros((B, 2), np.float32)
    out = nfx.pinned_empty((B, 18), np.float32)
    pp = lambda a: a.ctypes.data_as(C.c_void_p)
    L = lib()
    def call():
        rc = L.nfx_compute_features_batched(ex._h, nfx.FS_COLOR, B, pp(cents), pp(xy), pp(off), pp(patchs), pp(masks), pp(out))
        assert rc == 0
    for _ in range(5): call()
    n = 100
    t0 = time.perf_counter()
    for _ in range(n): call()
    dt = (time.perf_counter() - t0) / n
    ex.profile(True); ex.profile_reset(); call(); prof = ex.profile_get()
    print(f"B={B}: {dt*1e6:.0f} us per call, {B/dt:.0f} nuclei/s, H2D {patchs.nbytes+masks.nbytes} B -> {(patchs.nbytes+masks.nbytes)/dt/1e9:.1f} GB/s effective; kernels",
          {k: round(v[1], 3) for k, v in prof.items()})
    ex.close()

"""csrc/jpeg_exact.cpp against libjpeg-turbo (PIL decodes with it): the default slide decoder has to hand the kernels the
pixels OpenSlide hands the reference, bit for bit. Host only (no GPU): every chroma layout the decoder claims (4:4:4, 4:2:2,
4:2:0), odd sizes (partial MCUs, the one-column / one-row edge rules of the triangle filter), quality 30 ... 100 (16-bit
products in the IDCT, saturating range limit), optimised Huffman tables, restart intervals, grey scale, RGB components
without colour transform; streams outside its scope are refused, not guessed."""
import io

import numpy as np
import pytest
from PIL import Image

import nfx


def _img(rng, h, w, kind):
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    a = np.stack([(xx * 3 + yy) % 256, (yy * 5 + xx * 2) % 256, (xx * yy // 7) % 256], -1)
    return (a // 2 + rng.integers(0, 128, (h, w, 3))).astype(np.uint8)


def _save(a, **kw):
    b = io.BytesIO()
    Image.fromarray(a).save(b, "JPEG", **kw)
    return b.getvalue()


@pytest.mark.parametrize("size", [(64, 64), (63, 65), (17, 33), (256, 256), (100, 37), (8, 8), (31, 31), (240, 241), (2, 2), (1, 9)])
def test_bit_identical_to_libjpeg(libnfx, size):
    rng = np.random.default_rng(size[0] * 1000 + size[1])
    n = 0
    for kind in ("smooth", "noise"):
        a = _img(rng, size[0], size[1], kind)
        for sub in (0, 1, 2):
            for q in (30, 75, 95, 100):
                for extra in ({}, {"optimize": True}, {"restart_marker_blocks": 3}, {"restart_marker_rows": 1}):
                    try:
                        data = _save(a, quality=q, subsampling=sub, **extra)
                    except Exception:      # an option this Pillow / libjpeg build does not take at this size
                        continue
                    want = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
                    got = nfx.jpeg_decode(data)
                    assert got.shape == want.shape and np.array_equal(got, want), (size, kind, sub, q, extra)
                    n += 1
    assert n >= 48


def test_grey_rgb_components_and_refusals(libnfx):
    rng = np.random.default_rng(5)
    a = _img(rng, 90, 70, "smooth")
    g = _save(a[:, :, 0], quality=85)                                   # one component: R = G = B = Y
    assert np.array_equal(nfx.jpeg_decode(g), np.asarray(Image.open(io.BytesIO(g)).convert("RGB")))
    data = _save(a, quality=90, subsampling=0)
    im = Image.open(io.BytesIO(data))
    im.draft("YCbCr", im.size)                                          # libjpeg hands out the three components untransformed
    ycc = np.asarray(im)
    assert im.mode == "YCbCr"
    assert np.array_equal(nfx.jpeg_decode(data, 0), ycc)                # colourspace 0: components taken as R, G, B
    assert np.array_equal(nfx.jpeg_decode(data, 1), nfx.jpeg_decode(data, -1))
    prog = _save(a, quality=90, progressive=True)
    with pytest.raises(nfx.NfxError, match="progressive"):
        nfx.jpeg_decode(prog)
    for bad in (b"", b"\xff\xd8", data[:200], b"PNG" + data[3:]):
        with pytest.raises(nfx.NfxError):
            nfx.jpeg_decode(bad)
    with pytest.raises(nfx.NfxError):
        nfx.jpeg_decode(data, 7)                                        # unknown colourspace argument
    # damaged streams: refused or decoded to SOME picture of the right size, never a crash or a hang
    small = _save(a[:40, :48], quality=80, subsampling=2, restart_marker_blocks=2)
    cases = [small[: len(small) // 2] + b"\xff\xd9", small[:-2]]
    for _ in range(300):
        m = bytearray(small)
        for _ in range(int(rng.integers(1, 4))):
            m[int(rng.integers(2, len(m)))] = int(rng.integers(0, 256))
        cases.append(bytes(m))
    for c in cases:
        try:
            out = nfx.jpeg_decode(c)
            assert out.ndim == 3 and out.shape[2] == 3 and out.size <= 1 << 26
        except nfx.NfxError:
            pass

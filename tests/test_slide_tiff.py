"""Slide decode: the TIFF / BigTIFF container parser (host) and the decode of its JPEG blocks into the
resident slide (GPU), on synthetic files: strips written by libtiff through PIL (shared JPEGTables, RGB components) and
tiles written by the helper below (stand-alone YCbCr JPEG streams, classic and BigTIFF, both byte orders). The default
decoder (csrc/jpeg_exact.cpp) must give libjpeg-turbo's pixels (PIL) bit for bit -- what OpenSlide hands the reference;
the nvJPEG path (fast=True) differs by a few grey levels (its IDCT / upsampling are not libjpeg's)."""
import io
import struct

import numpy as np
import pytest
from PIL import Image

import nfx
from nfx import synth


def _jpeg(arr, quality=92, subsampling=0):
    b = io.BytesIO()
    Image.fromarray(arr).save(b, format="JPEG", quality=quality, subsampling=subsampling)
    return b.getvalue()


def write_tiled_tiff(img, tile=256, big=False, be=False, subsampling=0, quality=92):
    """Minimal tiled TIFF / BigTIFF writer: one directory, JPEG tiles (complete streams), Photometric = YCbCr."""
    h, w, _ = img.shape
    across, down = (w + tile - 1) // tile, (h + tile - 1) // tile
    E = ">" if be else "<"
    tiles = []
    for ty in range(down):
        for tx in range(across):
            t = np.zeros((tile, tile, 3), np.uint8)
            blk = img[ty * tile:(ty + 1) * tile, tx * tile:(tx + 1) * tile]
            t[:blk.shape[0], :blk.shape[1]] = blk
            t[blk.shape[0]:, :blk.shape[1]] = blk[-1:]          # replicate edges like real writers do
            t[:, blk.shape[1]:] = t[:, blk.shape[1] - 1:blk.shape[1]]
            tiles.append(_jpeg(t, quality, subsampling))
    hdr = 16 if big else 8
    data = bytearray(b"\0" * hdr)
    offs, cnts = [], []
    for t in tiles:
        offs.append(len(data))
        cnts.append(len(t))
        data += t
        if len(data) & 1:
            data += b"\0"
    ot, oc = ("Q", 16) if big else ("I", 4)                      # LONG8 / LONG
    def put_array(fmt, vals):
        at = len(data)
        data.extend(struct.pack(E + fmt * len(vals), *vals))
        if len(data) & 1:
            data.extend(b"\0")
        return at
    n = len(tiles)
    fs = 8 if big else 4                                         # size of the value field of a directory entry
    tsz, tfmt = {3: 2, 4: 4, 16: 8}, {3: "H", 4: "I", 16: "Q"}
    # (tag, type, values): values that fit the field are stored inline, others behind an offset
    ent = [(256, 4, [w]), (257, 4, [h]), (258, 3, [8, 8, 8]), (259, 3, [7]), (262, 3, [6]), (277, 3, [3]), (284, 3, [1]),
           (322, 3, [tile]), (323, 3, [tile]), (324, oc, offs), (325, oc, cnts)]
    fields = []
    for tag, typ, vals in ent:
        if len(vals) * tsz[typ] <= fs:
            fields.append(struct.pack(E + tfmt[typ] * len(vals), *vals).ljust(fs, b"\0"))
        else:
            fields.append(struct.pack(E + ("Q" if big else "I"), put_array(tfmt[typ], vals)))
    ifd = len(data)
    data.extend(struct.pack(E + ("Q" if big else "H"), len(ent)))
    for (tag, typ, vals), field in zip(ent, fields):
        data.extend(struct.pack(E + ("HHQ" if big else "HHI"), tag, typ, len(vals)) + field)
    data.extend(struct.pack(E + ("Q" if big else "I"), 0))
    if big:
        data[:16] = (b"MM" if be else b"II") + struct.pack(E + "HHHQ", 43, 8, 0, ifd)
    else:
        data[:8] = (b"MM" if be else b"II") + struct.pack(E + "HI", 42, ifd)
    return bytes(data), tiles


@pytest.fixture(scope="module")
def image():
    return synth.synth_tile(450, 600, 13)


@pytest.mark.parametrize("big,be", [(False, False), (True, False), (False, True), (True, True)])
def test_container_parser_tiles(libnfx, image, big, be):
    data, tiles = write_tiled_tiff(image, 256, big, be)
    info = nfx.tiff_info(data)
    assert (info["width"], info["height"], info["block_width"], info["block_height"]) == (600, 450, 256, 256)
    assert info["blocks"] == 6 and info["compression"] == 7 and info["photometric"] == 6 and info["jpeg_tables_bytes"] == 0


def test_container_parser_libtiff_strips_and_errors(libnfx, image, tmp_path):
    p = tmp_path / "strips.tif"
    Image.fromarray(image).save(p, compression="jpeg", quality=90)
    data = p.read_bytes()
    info = nfx.tiff_info(data)
    ref = Image.open(p)
    assert info["width"] == 600 and info["height"] == 450 and info["block_width"] == 600
    assert info["block_height"] == ref.tag_v2[278] and info["blocks"] == len(ref.tag_v2[273])
    assert info["compression"] == 7 and info["photometric"] == ref.tag_v2[262] and info["jpeg_tables_bytes"] > 100
    for bad, msg in [(b"", "too short"), (b"PK" + data[2:], "byte-order"), (data[:2] + b"\x07\x00" + data[4:], "magic"),
                     (data[:200], "past the end|out of range"), (data[:4] + struct.pack("<I", len(data) + 50) + data[8:], "out of range")]:
        with pytest.raises(nfx.NfxError, match=msg):
            nfx.tiff_info(bad)
    q = tmp_path / "raw.tif"
    Image.fromarray(image).save(q)                                   # uncompressed: parsed, refused by the decoder only
    assert nfx.tiff_info(q.read_bytes())["compression"] == 1


def _close(got, want, max_abs, mean_abs):
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= max_abs and d.mean() <= mean_abs, (int(d.max()), float(d.mean()))


@pytest.mark.gpu
@pytest.mark.parametrize("big,be,subsampling", [(False, False, 0), (True, True, 0), (False, False, 2)])
def test_decode_tiles_matches_libjpeg(libnfx, image, big, be, subsampling):
    data, tiles = write_tiled_tiff(image, 256, big, be, subsampling)
    want = np.zeros_like(image)
    for k, t in enumerate(tiles):
        ty, tx = divmod(k, 3)
        dec = np.asarray(Image.open(io.BytesIO(t)).convert("RGB"))
        blk = want[ty * 256:(ty + 1) * 256, tx * 256:(tx + 1) * 256]
        blk[:] = dec[:blk.shape[0], :blk.shape[1]]
    with nfx.Extractor(0) as ex:
        ex.load_tiff(data, 3)
        got = ex.slide_read(0, 0, 600, 450)
        assert np.array_equal(got, want)                         # the default decoder: libjpeg's pixels, bit for bit
        _close(got, image, 60 if subsampling == 0 else 200, 6.0 if subsampling == 0 else 12.0)   # the picture, not garbage (JPEG loss)
        # the decoded slide feeds the feature kernels like an uploaded tile does
        xy, off = synth.synth_polygons(40, 450, 600, 3, border_frac=0.1)
        keys, cents, feats, names = ex.extract(xy, off, ["color"])
        ex.upload_tile(got)
        k2, c2, f2, _ = ex.extract(xy, off, ["color"])
        assert keys == k2 and feats.tobytes() == f2.tobytes()
        ex.load_tiff(data, 3, fast=True)
        fast = ex.slide_read(0, 0, 600, 450)
        # nvJPEG, 4:4:4: only the IDCT differs; 4:2:0: chroma upsampling differs too (libjpeg-turbo's "fancy" filter)
        _close(fast, want, 6 if subsampling == 0 else 48, 0.8 if subsampling == 0 else 3.0)


@pytest.mark.gpu
def test_decode_libtiff_strips_rgb_components(libnfx, image, tmp_path):
    """libtiff writes Photometric = RGB strips: abbreviated streams + JPEGTables, components are R, G, B."""
    p = tmp_path / "strips.tif"
    Image.fromarray(image).save(p, compression="jpeg", quality=92)
    want = np.asarray(Image.open(p).convert("RGB"))
    with nfx.Extractor(0) as ex:
        ex.load_tiff(p.read_bytes(), 2)
        got = ex.slide_read(0, 0, 600, 450)
        assert np.array_equal(got, want)
        ex.load_tiff(p.read_bytes(), 2, fast=True)
        _close(ex.slide_read(0, 0, 600, 450), want, 6, 0.8)
    with nfx.Extractor(0) as ex:
        q = tmp_path / "raw.tif"
        Image.fromarray(image).save(q)
        with pytest.raises(nfx.NfxError, match="compression"):
            ex.load_tiff(q.read_bytes())


@pytest.mark.gpu
def test_cli_takes_tiff_slides(libnfx, image, tmp_path):
    """nfx-cli <geojson> <slide.tif> out.csv color: the InputImage::Slide arm of src/main.rs:20-35 without OpenSlide."""
    import csv
    import json
    import os
    import subprocess
    cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nuclei-feature-extraction_b200", "nfx-cli")
    tif, png = tmp_path / "slide.tif", tmp_path / "slide.png"
    Image.fromarray(image).save(tif, compression="jpeg", quality=95)
    Image.open(tif).convert("RGB").save(png)                       # the same pixels as libjpeg decodes them
    xy, off = synth.synth_polygons(60, 450, 600, 4, border_frac=0.1)
    feats = []
    for r in synth.rings_of(xy, off):
        feats.append({"type": "Feature", "bbox": [0, 0, 1, 1], "geometry": {"type": "Polygon", "coordinates": [[[float(x), float(y)] for x, y in r]]}})
    (tmp_path / "c.geojson").write_text(json.dumps({"type": "FeatureCollection", "features": feats}))
    outs = []
    for slide in (tif, png):
        out = tmp_path / (slide.suffix[1:] + ".csv")
        r = subprocess.run([cli, "-o", str(tmp_path / "c.geojson"), str(slide), str(out), "color"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        rows = list(csv.reader(open(out)))
        outs.append((rows[0], [row[0] for row in rows[1:]], np.array([[float(v) for v in row[1:]] for row in rows[1:]])))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1] and len(outs[0][1]) == 60
    j = outs[0][0].index("mean_r") - 1
    # .tif input takes the reference's SLIDE path (src/utils.rs:96-126: a window that would start left of / above the
    # slide is shifted to 0, NFX_RULE_WINDOW_SLIDE), .png the image path (zero padding): nuclei within P/2 of the left
    # or top edge legitimately differ, all others agree up to the decoders' few grey levels
    xyk = np.array([[float(v) for v in k.split(",")] for k in outs[0][1]])
    interior = (xyk[:, 0] >= 32.0) & (xyk[:, 1] >= 32.0)
    assert interior.sum() >= 40
    d = np.abs(outs[0][2][:, j:j + 3] - outs[1][2][:, j:j + 3])
    assert np.nanmax(d[interior]) < 0.01

"""nfx -- host-side mirror of the reference's operator interface over the C ABI of libnfx.so.

Mirrors (same names, argument meaning and error behaviour):
  * `trait FeatureSet { name(); compute_features_batched(centroids, polygons, patchs, masks) }`
    (src/features/mod.rs:12-28) -> ShapeFeatureSet / ColorFeatureSet / GlcmFeatureSet;
  * `args::FeatureSet::{from_str, flat, to_fs}` (src/args.rs:7-73) -> parse_feature_sets / to_fs;
  * the stage triple patch_loader -> move_tensors_to_device -> extract_features
    (src/main.rs:149-151) -> Extractor.extract.
All compute goes through libnfx.so (sm_100a CUDA); nothing here computes features on the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import (FS_ALL, FS_COLOR, FS_GABOR, FS_GEOMETRY, FS_GLCM, FS_GLRLM, FS_TEXTURE, NFX_OK, RULE_GABOR_HALF_TURN,
                   RULE_GLCM_254_U8, RULE_RASTER_PIXEL_CENTRE, RULE_WINDOW_SLIDE, NfxConfig, NfxError, NfxKernelTime,
                   NfxTiffLevel, lib)
from ._lib import EXT_ALL, EXT_COLOR_MOMENTS, EXT_CONTOUR, EXT_GLCM_D2, EXT_MASK_MOMENTS  # noqa: F401

_FLAT_BITS = (FS_GEOMETRY, FS_COLOR, FS_GLCM, FS_GLRLM, FS_GABOR)


def parse_feature_sets(names) -> int:
    """args::FeatureSet::from_str + flat (src/args.rs:18-49). Duplicates raise like the reference's
    DataFrame::new does (src/main.rs:89); an empty list raises like `features[0]` (src/main.rs:76)."""
    if isinstance(names, str):
        names = [names]
    if len(names) == 0:
        raise NfxError(-1, "no feature set given")
    mask = 0
    for s in names:
        bits = C.c_uint32(0)
        if lib().nfx_parse_feature_set(s.encode(), C.byref(bits)) != NFX_OK:
            raise NfxError(-1, f"{s} is not a valid feature set")
        if mask & bits.value:
            raise NfxError(-1, f"duplicate feature set {s!r}")
        mask |= bits.value
    return mask


def feature_names(mask: int):
    n = lib().nfx_feature_count(mask)
    return [lib().nfx_feature_name(mask, i).decode() for i in range(n)]


def centroid_key(x, y) -> str:
    buf = C.create_string_buffer(128)
    n = lib().nfx_centroid_key(float(np.float32(x)), float(np.float32(y)), buf, 128)
    if n < 0:
        raise NfxError(n, "key buffer too small")
    return buf.value.decode()


def format_f32(v) -> str:
    """Rust `Display` of one f32 (nfx_format_f32): what polars writes into CSV cells and the key."""
    buf = C.create_string_buffer(80)
    n = lib().nfx_format_f32(C.c_float(float(np.float32(v))), buf, 80)
    if n < 0:
        raise NfxError(n, "format buffer too small")
    return buf.value.decode()


def csv_header(mask: int) -> bytes:
    """`centroid,<columns>\\n` (src/main.rs:80-88, 163-166)."""
    n = C.c_int64()
    lib().nfx_csv_header(mask, None, 0, C.byref(n))
    buf = C.create_string_buffer(n.value)
    if lib().nfx_csv_header(mask, buf, n.value, C.byref(n)) != NFX_OK:
        raise NfxError(-1, "csv header")
    return buf.raw[:n.value]


def partition(n: int, batch_size: int, parts: int):
    b = (C.c_int64 * (parts + 1))()
    rc = lib().nfx_partition(n, batch_size, parts, b)
    if rc != NFX_OK:
        raise NfxError(rc, "bad partition arguments")
    return list(b)


def pack_polygons(rings):
    """list of (V_i,2) arrays -> CSR (poly_xy f32 [sum V,2], poly_off int64 [n+1])."""
    off = np.zeros(len(rings) + 1, dtype=np.int64)
    for i, r in enumerate(rings):
        off[i + 1] = off[i] + len(r)
    xy = np.zeros((int(off[-1]), 2), dtype=np.float32)
    for i, r in enumerate(rings):
        xy[off[i]:off[i + 1]] = np.asarray(r, dtype=np.float32).reshape(-1, 2)
    return xy, off


def tiff_info(data) -> dict:
    """Directory 0 of a TIFF / BigTIFF held in memory (nfx_tiff_info): size, block size, compression."""
    buf = np.frombuffer(data, dtype=np.uint8)
    lv = NfxTiffLevel()
    if lib().nfx_tiff_info(buf.ctypes.data, buf.size, C.byref(lv)) != NFX_OK:
        raise NfxError(-1, (lib().nfx_last_error(None) or b"").decode())
    return {f: getattr(lv, f) for f, _ in NfxTiffLevel._fields_}


DECODE_FAST = 0x1


def jpeg_decode(data, colourspace: int = -1) -> np.ndarray:
    """One baseline JPEG stream -> [h, w, 3] u8 with libjpeg's pixels bit for bit (nfx_jpeg_decode, host only).
    colourspace: 0 = components are R,G,B; 1 = YCbCr; -1 = libjpeg's own rule."""
    buf = np.frombuffer(data, dtype=np.uint8)
    w, h = C.c_int32(), C.c_int32()
    if lib().nfx_jpeg_decode(buf.ctypes.data, buf.size, colourspace, None, 0, C.byref(w), C.byref(h)) != NFX_OK:
        raise NfxError(-1, (lib().nfx_last_error(None) or b"").decode())
    out = np.empty((h.value, w.value, 3), np.uint8)
    if lib().nfx_jpeg_decode(buf.ctypes.data, buf.size, colourspace, out.ctypes.data, out.size, C.byref(w), C.byref(h)) != NFX_OK:
        raise NfxError(-1, (lib().nfx_last_error(None) or b"").decode())
    return out


def parse_f32(token: str) -> np.float32:
    """One JSON number token as the reference's serde_json hands it to an f32 field (nfx_parse_f32)."""
    o = C.c_float()
    b = token.encode()
    if lib().nfx_parse_f32(b, len(b), C.byref(o)) != NFX_OK:
        raise NfxError(-1, (lib().nfx_last_error(None) or b"").decode())
    return np.float32(o.value)


class _GeojsonHandle:
    """Owns an nfx_geojson; the arrays returned by geojson_pack are views into it and keep it alive."""

    def __init__(self, h):
        self.h = h

    def __del__(self):
        if self.h:
            lib().nfx_geojson_free(self.h)
            self.h = None


def _view(ptr, shape, dtype, owner):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(C.addressof(ptr.contents))
    a = np.frombuffer(buf, dtype=dtype).reshape(shape)
    a.flags.writeable = False
    buf._owner = owner               # the ctypes array is the numpy base: tie the handle's lifetime to it
    return a


def geojson_pack(text, threads: int = 0):
    """GeoJSON text (bytes / str / uint8 array) -> CSR of ring 0 of every feature, parsed by `threads` host threads
    (src/main.rs:37-42, src/geojson.rs:8-24): (poly_xy f32 [nv,2], poly_off i64 [n+1], bbox f32 [n,4], rings i32 [n]).
    The arrays are read-only views into the library's result (no copy); it is freed with the last of them."""
    if isinstance(text, str):
        text = text.encode()
    buf = np.frombuffer(text, dtype=np.uint8)
    h = C.c_void_p()
    rc = lib().nfx_geojson_parse(C.cast(buf.ctypes.data, C.c_char_p), buf.size, threads, C.byref(h))
    if rc != NFX_OK:
        raise NfxError(rc, (lib().nfx_last_error(None) or b"").decode())
    own = _GeojsonHandle(h)
    n = lib().nfx_geojson_count(h)
    nv = lib().nfx_geojson_vertices(h)
    xy = _view(lib().nfx_geojson_xy(h), (nv, 2), np.float32, own)
    off = _view(lib().nfx_geojson_offsets(h), (n + 1,), np.int64, own)
    bbox = _view(lib().nfx_geojson_bbox(h), (n, 4), np.float32, own)
    rings = _view(lib().nfx_geojson_rings(h), (n,), np.int32, own)
    return xy, off, bbox, rings


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Extractor:
    """One context per (host thread, GPU) -- the analogue of one rayon worker (src/utils.rs:215-221)."""

    def __init__(self, device: int = 0, patch_size: int = 64, batch_size: int = 100, rule_flags: int = 0):
        self._h = C.c_void_p()
        cfg = NfxConfig(patch_size, batch_size, rule_flags)
        rc = lib().nfx_create(device, C.byref(cfg), C.byref(self._h))
        if rc != NFX_OK:
            raise NfxError(rc, (lib().nfx_last_error(None) or b"").decode())
        self.patch_size, self.batch_size, self.device = patch_size, batch_size, device
        self.n = 0
        self._mask = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().nfx_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != NFX_OK:
            raise NfxError(rc, (lib().nfx_last_error(self._h) or b"").decode())

    # -- inputs --
    def upload_tile(self, rgb, origin=(0, 0)):
        """rgb: [H,W,3] u8 (C-contiguous rows). Replaces load_input_image (src/main.rs:20-35)."""
        rgb = np.asarray(rgb)
        if rgb.dtype != np.uint8 or rgb.ndim != 3 or rgb.shape[2] != 3 or rgb.strides[2] != 1 \
                or rgb.strides[1] != 3:
            raise NfxError(-1, "tile must be [H,W,3] u8 with contiguous rows")
        self._tile_ref = rgb
        self._ck(lib().nfx_tile_upload(self._h, _ptr(rgb), rgb.shape[1], rgb.shape[0], rgb.strides[0],
                                       int(origin[0]), int(origin[1])))

    def load_tiff(self, data, threads: int = 0, fast: bool = False):
        """Level 0 of a JPEG-compressed TIFF / .svs held in memory -> the resident slide. Default: libjpeg's pixels bit for
        bit (host decoder, csrc/jpeg_exact.cpp); fast=True: nvJPEG (a few grey levels off, about twice the tile rate)."""
        buf = np.frombuffer(data, dtype=np.uint8)
        self._ck(lib().nfx_slide_load_tiff_ex(self._h, buf.ctypes.data, buf.size, threads, DECODE_FAST if fast else 0))

    def slide_read(self, x0, y0, w, h):
        out = np.empty((h, w, 3), dtype=np.uint8)
        self._ck(lib().nfx_debug_slide_read(self._h, x0, y0, w, h, _ptr(out)))
        return out

    def slide_alloc(self, w, h, origin=(0, 0)):
        """Reserve a W x H slide in HBM; fill it with write_tile (tiles / row bands, any order)."""
        self._ck(lib().nfx_slide_alloc(self._h, int(w), int(h), int(origin[0]), int(origin[1])))

    def write_tile(self, rgb, x0, y0):
        rgb = np.asarray(rgb)
        if rgb.dtype != np.uint8 or rgb.ndim != 3 or rgb.shape[2] != 3 or rgb.strides[2] != 1 or rgb.strides[1] != 3:
            raise NfxError(-1, "tile must be [h,w,3] u8 with contiguous rows")
        self._ck(lib().nfx_slide_write_tile(self._h, _ptr(rgb), int(x0), int(y0), rgb.shape[1], rgb.shape[0],
                                            rgb.strides[0]))

    # -- one slide on several GPUs (include/nfx.h: nfx_slide_export / import_rows / copy_rows) --
    def slide_export(self) -> bytes:
        """Handle of this context's slide for another PROCESS (CUDA IPC): 96 bytes, send them by any host mechanism."""
        h = np.zeros(96, dtype=np.uint8)
        self._ck(lib().nfx_slide_export(self._h, _ptr(h)))
        return h.tobytes()

    def slide_import_rows(self, handle: bytes, y0: int, rows: int):
        """Rows [y0, y0 + rows) of the peer's slide -> this context's slide, device to device (NVLink between GPUs)."""
        h = np.frombuffer(handle, dtype=np.uint8).copy()
        self._ck(lib().nfx_slide_import_rows(self._h, _ptr(h), int(y0), int(rows)))

    def slide_copy_rows(self, src: "Extractor", y0: int, rows: int):
        """Same transfer between two contexts of one process."""
        self._ck(lib().nfx_slide_copy_rows(self._h, src._h, int(y0), int(rows)))

    def upload_polygons(self, poly_xy, poly_off):
        poly_xy = np.ascontiguousarray(poly_xy, dtype=np.float32)
        poly_off = np.ascontiguousarray(poly_off, dtype=np.int64)
        self.n = len(poly_off) - 1
        self._ck(lib().nfx_polygons_upload(self._h, self.n, _ptr(poly_xy), _ptr(poly_off)))

    # -- hot path --
    def compute(self, mask: int):
        self._mask = mask
        self._ck(lib().nfx_compute(self._h, mask))

    def sync(self):
        self._ck(lib().nfx_sync(self._h))

    def compute_ext(self, ext_mask: int = EXT_ALL):
        """Extension outputs (include/nfx.h NFX_EXT_*: skew / kurtosis, mask moments, contour length, GLCM at distance 2) for
        the staged tile + polygons: (names, [n, F] f32). Never part of the drop-in schema."""
        self._ck(lib().nfx_compute_ext(self._h, ext_mask))
        F = lib().nfx_ext_feature_count(ext_mask)
        out = np.empty((self.n, F), dtype=np.float32)
        self._ck(lib().nfx_download_ext(self._h, _ptr(out)))
        return [lib().nfx_ext_feature_name(ext_mask, i).decode() for i in range(F)], out

    def download(self, centroids=None, features=None):
        F = lib().nfx_feature_count(self._mask)
        if centroids is None:
            centroids = np.empty((self.n, 2), dtype=np.float32)
        if features is None:
            features = np.empty((self.n, F), dtype=np.float32)
        self._ck(lib().nfx_download(self._h, _ptr(centroids), _ptr(features)))
        return centroids, features

    def csv_rows(self, lo: int = 0, hi: int | None = None, buf=None, view: bool = False):
        """CSV text of rows [lo, hi) of the resident result, formatted on the GPU (nfx_csv_rows). `buf` is an
        optional reusable (pinned) uint8 array; it is grown when the text does not fit."""
        hi = self.n if hi is None else hi
        if buf is None:
            buf = np.empty(max(1 << 16, (hi - lo) * 64), dtype=np.uint8)
        n = C.c_int64()
        rc = lib().nfx_csv_rows(self._h, lo, hi, _ptr(buf), buf.size, C.byref(n))
        if rc != NFX_OK and n.value > buf.size:
            buf = np.empty(n.value, dtype=np.uint8)
            rc = lib().nfx_csv_rows(self._h, lo, hi, _ptr(buf), buf.size, C.byref(n))
        self._ck(rc)
        return buf[:n.value] if view else buf[:n.value].tobytes()

    def csv_format(self, centroids, features) -> bytes:
        """CSV rows of a caller-held frame (centroids [n,2], features [n,F] f32), formatted on the GPU."""
        centroids = np.ascontiguousarray(centroids, dtype=np.float32).reshape(-1, 2)
        n = len(centroids)
        features = np.ascontiguousarray(features, dtype=np.float32)
        features = features.reshape(n, features.shape[-1] if features.ndim == 2 else (features.size // max(n, 1)))
        need = C.c_int64()
        rc = lib().nfx_csv_format(self._h, n, features.shape[1], _ptr(centroids), _ptr(features), None, 0, C.byref(need))
        if need.value == 0:
            self._ck(rc)
            return b""
        buf = np.empty(need.value, dtype=np.uint8)
        self._ck(lib().nfx_csv_format(self._h, n, features.shape[1], _ptr(centroids), _ptr(features), _ptr(buf), buf.size, C.byref(need)))
        return buf[:need.value].tobytes()

    def extract(self, poly_xy, poly_off, feature_sets):
        """chunk(s) -> features, rows in input order. Returns (keys, centroids, features, names)."""
        mask = feature_sets if isinstance(feature_sets, int) else parse_feature_sets(feature_sets)
        self.upload_polygons(poly_xy, poly_off)
        self.compute(mask)
        cents, feats = self.download()
        keys = [centroid_key(c[0], c[1]) for c in cents]
        return keys, cents, feats, feature_names(mask)

    # -- staged kernels / parity taps --
    def rasterize(self):
        out = np.empty((self.n, self.patch_size, self.patch_size), dtype=np.uint8)
        self._ck(lib().nfx_rasterize(self._h, _ptr(out)))
        return out

    def rasterize_device(self):
        """Kernel (2) only: masks stay in HBM as bitmasks (no host copy)."""
        self._ck(lib().nfx_rasterize(self._h, None))

    def gather_patches(self, want=True):
        out = np.empty((self.n, self.patch_size, self.patch_size, 3), dtype=np.uint8) if want else None
        self._ck(lib().nfx_gather_patches(self._h, _ptr(out)))
        return out

    def debug_ellipses(self):
        out = np.empty((self.n, self.patch_size, self.patch_size), dtype=np.uint8)
        self._ck(lib().nfx_debug_ellipses(self._h, _ptr(out)))
        return out

    def debug_glcm_counts(self, levels, offset):
        out = np.empty((self.n, levels, levels), dtype=np.uint32)
        self._ck(lib().nfx_debug_glcm_counts(self._h, levels, offset[0], offset[1], _ptr(out)))
        return out

    def debug_grey_levels(self, levels):
        out = np.empty((self.n, self.patch_size, self.patch_size), dtype=np.uint8)
        self._ck(lib().nfx_debug_grey_levels(self._h, levels, _ptr(out)))
        return out

    def compute_features_batched(self, bit, centroids, polygons, patchs, masks):
        n = int(patchs.shape[0])
        patchs = np.ascontiguousarray(patchs, dtype=np.float32)
        masks = np.ascontiguousarray(masks, dtype=np.float32)
        cents = np.ascontiguousarray(centroids, dtype=np.float32)
        xy, off = pack_polygons(polygons)
        out = np.empty((n, lib().nfx_feature_count(bit)), dtype=np.float32)
        self._ck(lib().nfx_compute_features_batched(self._h, bit, n, _ptr(cents), _ptr(xy), _ptr(off),
                                                    _ptr(patchs), _ptr(masks), _ptr(out)))
        return out

    # -- measurement --
    def profile(self, on=True):
        self._ck(lib().nfx_profile_enable(self._h, 1 if on else 0))

    def profile_reset(self):
        self._ck(lib().nfx_profile_reset(self._h))

    def profile_get(self):
        arr = (NfxKernelTime * 32)()
        n = lib().nfx_profile_get(self._h, arr, 32)
        if n < 0:
            self._ck(n)
        return {arr[i].name.decode(): (arr[i].launches, arr[i].total_ms) for i in range(min(n, 32))}

    def timer_start(self):
        self._ck(lib().nfx_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        self._ck(lib().nfx_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        return int(lib().nfx_launch_count(self._h))

    def flush_l2(self):
        self._ck(lib().nfx_flush_l2(self._h))


def pinned_empty(shape, dtype):
    """numpy array over cudaHostAlloc memory (for asynchronous H2D / D2H in the e2e path)."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    rc = lib().nfx_host_alloc(C.byref(p), nbytes)
    if rc != NFX_OK:
        raise NfxError(rc, (lib().nfx_last_error(None) or b"").decode())
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr


class _FeatureSetBase:
    """trait FeatureSet (src/features/mod.rs:12-28) over nfx_compute_features_batched."""
    BIT = 0

    def __init__(self, extractor: Extractor):
        self._ex = extractor

    def name(self) -> str:
        return lib().nfx_feature_set_name(self.BIT).decode()

    def columns(self):
        return feature_names(self.BIT)

    def compute_features_batched(self, centroids, polygons, patchs, masks):
        """Returns (keys, [N,F] f32): the DataFrame's `centroid` column and feature columns."""
        # the asserts of src/features/shape.rs:23-47 and color.rs:18-42
        assert patchs.ndim == 4, "The patchs tensor must be 4 dimensional"
        assert masks.ndim == 4, "The masks tensor must be 4 dimensional"
        assert patchs.shape[1] == 3, "The patchs tensor must have 3 channels"
        assert masks.shape[1] == 1, "The masks tensor must have 1 channel"
        assert patchs.shape[0] == masks.shape[0], "The number of patchs and masks must be the same"
        assert patchs.shape[0] == len(centroids), "The number of patchs and centroids must be the same"
        assert patchs.shape[0] == len(polygons), "The number of patchs and polygons must be the same"
        out = self._ex.compute_features_batched(self.BIT, centroids, polygons, patchs, masks)
        keys = [centroid_key(c[0], c[1]) for c in np.asarray(centroids, dtype=np.float32)]
        return keys, out


class ShapeFeatureSet(_FeatureSetBase):      # src/features/shape.rs:13
    BIT = FS_GEOMETRY


class ColorFeatureSet(_FeatureSetBase):      # src/features/color.rs:8
    BIT = FS_COLOR


class GlcmFeatureSet(_FeatureSetBase):       # src/features/texture.rs:22
    BIT = FS_GLCM


class GLRLMFeatureSet(_FeatureSetBase):      # src/features/texture.rs:178
    BIT = FS_GLRLM


class GaborFilterFeatureSet(_FeatureSetBase):  # src/features/texture.rs:317
    BIT = FS_GABOR


def to_fs(names, extractor):
    """args::FeatureSet::to_fs (src/args.rs:51-73)."""
    mask = parse_feature_sets(names)
    table = {FS_GEOMETRY: ShapeFeatureSet, FS_COLOR: ColorFeatureSet, FS_GLCM: GlcmFeatureSet,
             FS_GLRLM: GLRLMFeatureSet, FS_GABOR: GaborFilterFeatureSet}
    out = []
    for b in _FLAT_BITS:
        if mask & b:
            if b not in table:
                raise NfxError(-4, f"feature set {lib().nfx_feature_set_name(b).decode()} not built yet")
            out.append(table[b](extractor))
    return out

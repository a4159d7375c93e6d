"""ctypes binding of libnfx.so (include/nfx.h). Fails loudly when the CUDA library is missing:
there is no CPU fallback in the product path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NFX_LIB selects another build of the same library (e.g. the CMake build: NFX_LIB=$PWD/b/libnfx.so)
LIB_PATH = os.environ.get("NFX_LIB") or os.path.join(os.path.dirname(_HERE), "libnfx.so")

NFX_OK = 0
FS_GEOMETRY, FS_COLOR, FS_GLCM, FS_GLRLM, FS_GABOR = 0x01, 0x02, 0x04, 0x08, 0x10
FS_TEXTURE = FS_GLCM | FS_GLRLM | FS_GABOR
FS_ALL = 0x1F


class NfxConfig(C.Structure):
    _fields_ = [("patch_size", C.c_int32), ("batch_size", C.c_int32), ("rule_flags", C.c_int32), ("reserved", C.c_int32 * 5)]


# nfx_config.rule_flags (include/nfx.h): switches of the unpinned rules, mirrored by oracle.RULES
RULE_RASTER_PIXEL_CENTRE, RULE_GABOR_HALF_TURN, RULE_GLCM_254_U8, RULE_WINDOW_SLIDE = 0x1, 0x2, 0x4, 0x8
# extension outputs (never part of the drop-in schema)
EXT_COLOR_MOMENTS, EXT_MASK_MOMENTS, EXT_CONTOUR, EXT_GLCM_D2, EXT_ALL = 0x1, 0x2, 0x4, 0x8, 0xF


class NfxTiffLevel(C.Structure):
    _fields_ = [("width", C.c_int64), ("height", C.c_int64), ("block_width", C.c_int32), ("block_height", C.c_int32),
                ("blocks", C.c_int64), ("compression", C.c_int32), ("photometric", C.c_int32), ("jpeg_tables_bytes", C.c_int32)]


class NfxKernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_int64), ("total_ms", C.c_double)]


# every symbol include/nfx.h declares: (restype, argtypes)
_vp, _i, _i64, _u32, _f = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_float
SYMBOLS = {
    "nfx_create": (_i, [_i, C.POINTER(NfxConfig), C.POINTER(_vp)]),
    "nfx_destroy": (_i, [_vp]),
    "nfx_last_error": (C.c_char_p, [_vp]),
    "nfx_version": (C.c_char_p, []),
    "nfx_tile_upload": (_i, [_vp, _vp, _i64, _i64, _i64, _i64, _i64]),
    "nfx_slide_alloc": (_i, [_vp, _i64, _i64, _i64, _i64]),
    "nfx_slide_write_tile": (_i, [_vp, _vp, _i64, _i64, _i64, _i64, _i64]),
    "nfx_slide_export": (_i, [_vp, _vp]),
    "nfx_slide_import_rows": (_i, [_vp, _vp, _i64, _i64]),
    "nfx_slide_copy_rows": (_i, [_vp, _vp, _i64, _i64]),
    "nfx_polygons_upload": (_i, [_vp, _i64, _vp, _vp]),
    "nfx_ext_feature_count": (_i, [_u32]),
    "nfx_ext_feature_name": (C.c_char_p, [_u32, _i]),
    "nfx_compute_ext": (_i, [_vp, _u32]),
    "nfx_download_ext": (_i, [_vp, _vp]),
    "nfx_compute": (_i, [_vp, _u32]),
    "nfx_download": (_i, [_vp, _vp, _vp]),
    "nfx_extract": (_i, [_vp, _i64, _vp, _vp, _u32, _vp, _vp]),
    "nfx_sync": (_i, [_vp]),
    "nfx_compute_features_batched": (_i, [_vp, _u32, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nfx_gather_patches": (_i, [_vp, _vp]),
    "nfx_rasterize": (_i, [_vp, _vp]),
    "nfx_debug_ellipses": (_i, [_vp, _vp]),
    "nfx_debug_glcm_counts": (_i, [_vp, _i, _i, _i, _vp]),
    "nfx_debug_grey_levels": (_i, [_vp, _i, _vp]),
    "nfx_feature_count": (_i, [_u32]),
    "nfx_feature_name": (C.c_char_p, [_u32, _i]),
    "nfx_parse_feature_set": (_i, [C.c_char_p, C.POINTER(_u32)]),
    "nfx_feature_set_name": (C.c_char_p, [_u32]),
    "nfx_centroid_key": (_i, [_f, _f, C.c_char_p, _i]),
    "nfx_format_f32": (_i, [_f, C.c_char_p, _i]),
    "nfx_tiff_info": (_i, [_vp, _i64, _vp]),
    "nfx_slide_load_tiff": (_i, [_vp, _vp, _i64, C.c_int32]),
    "nfx_slide_load_tiff_ex": (_i, [_vp, _vp, _i64, C.c_int32, C.c_uint32]),
    "nfx_jpeg_decode": (_i, [_vp, _i64, C.c_int32, _vp, _i64, _vp, _vp]),
    "nfx_debug_slide_read": (_i, [_vp, _i64, _i64, _i64, _i64, _vp]),
    "nfx_geojson_parse": (_i, [C.c_char_p, _i64, C.c_int32, C.POINTER(_vp)]),
    "nfx_geojson_count": (_i64, [_vp]),
    "nfx_geojson_vertices": (_i64, [_vp]),
    "nfx_geojson_xy": (C.POINTER(_f), [_vp]),
    "nfx_geojson_offsets": (C.POINTER(_i64), [_vp]),
    "nfx_geojson_bbox": (C.POINTER(_f), [_vp]),
    "nfx_geojson_rings": (C.POINTER(C.c_int32), [_vp]),
    "nfx_geojson_free": (None, [_vp]),
    "nfx_parse_f32": (_i, [C.c_char_p, C.c_int32, C.POINTER(_f)]),
    "nfx_csv_header": (_i, [_u32, _vp, _i64, C.POINTER(_i64)]),
    "nfx_csv_rows": (_i, [_vp, _i64, _i64, _vp, _i64, C.POINTER(_i64)]),
    "nfx_csv_format": (_i, [_vp, _i64, C.c_int32, _vp, _vp, _vp, _i64, C.POINTER(_i64)]),
    "nfx_partition": (_i, [_i64, C.c_int32, C.c_int32, C.POINTER(_i64)]),
    "nfx_profile_enable": (_i, [_vp, _i]),
    "nfx_profile_reset": (_i, [_vp]),
    "nfx_profile_get": (_i, [_vp, C.POINTER(NfxKernelTime), _i]),
    "nfx_timer_start": (_i, [_vp]),
    "nfx_timer_stop": (_i, [_vp, C.POINTER(_f)]),
    "nfx_launch_count": (_i64, [_vp]),
    "nfx_flush_l2": (_i, [_vp]),
    "nfx_host_alloc": (_i, [C.POINTER(_vp), _i64]),
    "nfx_host_free": (_i, [_vp]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                "(nvcc, sm_100a). nfx has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError if the .so does not export it
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


class NfxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nfx error {code}: {msg}")
        self.code = code

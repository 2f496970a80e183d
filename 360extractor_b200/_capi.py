"""ctypes binding of libx360 (include/x360.h).  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .build import LIB_PATH

X360_OK = 0
INTERP_LINEAR = 0
INTERP_LANCZOS4 = 1
PATH_AUTO, PATH_DIRECT, PATH_STAGED, PATH_TILED = 0, 1, 2, 3
FLAG_PLANES, FLAG_FUSED_SCORE, FLAG_SPLIT_SCORE, FLAG_TILE_H16, FLAG_TILE_H32, FLAG_NO_WIDE, FLAG_FULL_UPLOAD = 1, 2, 4, 8, 16, 32, 64


class X360Error(RuntimeError):
    """Raised for every non-zero return of the C-ABI (message from x360_last_error)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libx360 error {code}: {message}")
        self.code = code


class Params(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_int32),
        ("src_w", ctypes.c_int32), ("src_h", ctypes.c_int32),
        ("out_w", ctypes.c_int32), ("out_h", ctypes.c_int32),
        ("fov_deg", ctypes.c_double),
        ("interp", ctypes.c_int32),
        ("n_views", ctypes.c_int32),
        ("R", ctypes.POINTER(ctypes.c_double)),
        ("device", ctypes.c_int32),
        ("max_batch", ctypes.c_int32),
        ("path", ctypes.c_int32),
        ("flags", ctypes.c_int32),
    ]


class BlurState(ctypes.Structure):
    """x360_blur_state (include/x360.h): the accept / reject machine of src/core/processor.py:341-376."""
    _fields_ = [
        ("enabled", ctypes.c_int32), ("smart", ctypes.c_int32),
        ("threshold", ctypes.c_double),
        ("history", ctypes.c_double * 10),
        ("hist_len", ctypes.c_int32), ("hist_head", ctypes.c_int32),
        ("consecutive_skips", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("skipped", ctypes.c_int64), ("forced", ctypes.c_int64),
    ]


# every symbol include/x360.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _sz, _dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t, ctypes.c_double
SYMBOLS = {
    "x360_plan_create": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.POINTER(_vp)]),
    "x360_plan_destroy": (ctypes.c_int, [_vp]),
    "x360_extract_device": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "x360_extract_host": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp]),
    "x360_extract": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "x360_make_maps": (ctypes.c_int, [_vp, _i32, _vp, _vp]),
    "x360_remap_maps": (ctypes.c_int, [_i32, _vp, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _vp]),
    "x360_blur_sums": (ctypes.c_int, [_i32, _vp, _i32, _i32, _i32, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    "x360_column_only_map_x": (ctypes.c_int, [_vp, _i32]),
    "x360_plan_set_sharpen": (ctypes.c_int, [_vp, _i32, _dbl]),
    "x360_unsharp_device": (ctypes.c_int, [_i32, _vp, _i64, _i32, _i32, _dbl, _vp, _vp]),
    "x360_unsharp_host": (ctypes.c_int, [_i32, _vp, _i64, _i32, _i32, _dbl, _vp]),
    "x360_views_to_nchw": (ctypes.c_int, [_i32, _vp, _i64, _i32, _i32, _vp, _i32, _i32, _vp, _vp]),
    "x360_lanczos4_table": (ctypes.c_int, [_vp]),
    "x360_lanczos4_factors": (ctypes.c_int, [_vp, _vp]),
    "x360_focal": (_dbl, [_i32, _dbl]),
    "x360_plan_preview": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.POINTER(_i32), _vp, _vp, _vp, _vp, _vp]),
    "x360_plan_preview_views": (ctypes.c_int, [ctypes.POINTER(Params), _vp, _vp]),
    "x360_jpeg_create": (ctypes.c_int, [_i32, _i32, _i32, _i32, _i32, ctypes.POINTER(_vp)]),
    "x360_jpeg_destroy": (ctypes.c_int, [_vp]),
    "x360_jpeg_max_bytes": (_i64, [_vp, _i32]),
    "x360_jpeg_encode_device": (ctypes.c_int, [_vp, _vp, _i32, _vp, _i64, _vp, _vp]),
    "x360_jpeg_status": (ctypes.c_int, [_vp, _vp]),
    "x360_jpeg_encode_host": (ctypes.c_int, [_vp, _vp, _i64, _vp, _i64, _vp]),
    "x360_jpeg_coefficients_host": (ctypes.c_int, [_vp, _vp, _i32, _vp]),
    "x360_jpeg_header": (ctypes.c_int, [_i32, _i32, _i32, _vp, _i32]),
    "x360_extract_host_jpeg": (ctypes.c_int, [_vp, _vp, _i32, _i32, _vp, _i64, _vp, _vp, _vp]),
    "x360_blur_state_init": (None, [ctypes.POINTER(BlurState), _i32, _i32, _dbl]),
    "x360_blur_replay": (ctypes.c_int, [ctypes.POINTER(BlurState), _vp, _i64, _vp]),
    "x360_blur_accept": (ctypes.c_int, [ctypes.POINTER(BlurState), _dbl]),
    "x360_extract_host_filtered": (ctypes.c_int, [_vp, _vp, _i32, ctypes.POINTER(BlurState), _vp, _i64, _vp, _vp, ctypes.POINTER(_i64)]),
    "x360_host_alloc": (ctypes.c_int, [_sz, ctypes.POINTER(_vp)]),
    "x360_host_free": (ctypes.c_int, [_vp]),
    "x360_last_error": (ctypes.c_char_p, []),
    "x360_version": (ctypes.c_char_p, []),
    "x360_launch_count": (_i64, []),
    "x360_plan_info": (ctypes.c_int, [_vp] + [ctypes.POINTER(_i32)] * 4),
    "x360_plan_tile_modes": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_uint32)]),
    "x360_plan_transposed_blocks": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int64)]),
    "x360_plan_upload_rows": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load lib/libx360.so.  Raises if it has not been built — there is no other backend."""
    global _lib
    if _lib is None:
        path = os.environ.get("X360_LIB", LIB_PATH)      # another build of the same library (tuning experiments)
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -m 360extractor_b200.build` "
                "(or __graft_entry__.build()).  This package has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != X360_OK:
        raise X360Error(rc, lib().x360_last_error().decode("utf-8", "replace"))


def ptr(a) -> int:
    """Raw address of a numpy array or a torch tensor (host or device)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def version() -> str:
    return lib().x360_version().decode()


def launch_count() -> int:
    return int(lib().x360_launch_count())

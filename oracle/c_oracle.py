"""ctypes loader for the plain-C oracle (``x360_oracle.c``).

CPU ORACLE — TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libx360_oracle.so")
_lib = None

INTERP_LINEAR = 0
INTERP_LANCZOS4 = 1


def build(force: bool = False) -> str:
    """Compile the C oracle with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("x360_oracle.c", "x360_jpeg_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libx360_oracle.so"])
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        c_u8p = ctypes.POINTER(ctypes.c_uint8)
        c_f32p = ctypes.POINTER(ctypes.c_float)
        c_f64p = ctypes.POINTER(ctypes.c_double)
        c_i64p = ctypes.POINTER(ctypes.c_int64)
        c_i16p = ctypes.POINTER(ctypes.c_int16)
        L.orc_focal.restype = ctypes.c_double
        L.orc_focal.argtypes = [ctypes.c_int, ctypes.c_double]
        L.orc_make_map.restype = None
        L.orc_make_map.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double, c_f64p, c_f32p, c_f32p]
        for name in ("orc_remap_linear", "orc_remap_lanczos4"):
            fn = getattr(L, name)
            fn.restype = None
            fn.argtypes = [c_u8p, ctypes.c_int, ctypes.c_int, c_f32p, c_f32p,
                           ctypes.c_int, ctypes.c_int, c_u8p]
        L.orc_lanczos4_table.restype = None
        L.orc_lanczos4_table.argtypes = [c_i16p]
        L.orc_blur_sums.restype = None
        L.orc_blur_sums.argtypes = [c_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_i64p, c_i64p]
        L.orc_score_from_sums.restype = ctypes.c_double
        L.orc_unsharp.argtypes = [c_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, c_u8p]
        L.orc_unsharp.restype = None
        L.orc_add_weighted.argtypes = [c_u8p, c_u8p, ctypes.c_size_t, ctypes.c_double, c_u8p]
        L.orc_add_weighted.restype = None
        L.orc_gauss13_taps.argtypes = [ctypes.c_double, ctypes.POINTER(ctypes.c_int)]
        L.orc_gauss13_taps.restype = None
        L.orc_score_from_sums.argtypes = [ctypes.c_int64] * 3
        L.orc_extract.restype = None
        L.orc_extract.argtypes = [c_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_f64p,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                  c_u8p, c_i64p, c_i64p]
        _lib = L
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def focal(dest_w: int, fov_deg: float) -> float:
    return lib().orc_focal(int(dest_w), float(fov_deg))


def make_map(src_h, src_w, dest_h, dest_w, fov_deg, R):
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(9)
    mx = np.empty((dest_h, dest_w), np.float32)
    my = np.empty((dest_h, dest_w), np.float32)
    lib().orc_make_map(src_h, src_w, dest_h, dest_w, float(fov_deg), _p(R, ctypes.c_double),
                       _p(mx, ctypes.c_float), _p(my, ctypes.c_float))
    return mx, my


def remap(src, map_x, map_y, interp=INTERP_LINEAR):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    assert src.ndim == 3 and src.shape[2] == 3
    mx = np.ascontiguousarray(map_x, dtype=np.float32)
    my = np.ascontiguousarray(map_y, dtype=np.float32)
    dh, dw = mx.shape
    dst = np.empty((dh, dw, 3), np.uint8)
    fn = lib().orc_remap_lanczos4 if interp == INTERP_LANCZOS4 else lib().orc_remap_linear
    fn(_p(src, ctypes.c_uint8), src.shape[0], src.shape[1], _p(mx, ctypes.c_float),
       _p(my, ctypes.c_float), dh, dw, _p(dst, ctypes.c_uint8))
    return dst


def lanczos4_table():
    tab = np.empty((32 * 32, 8, 8), np.int16)
    lib().orc_lanczos4_table(_p(tab, ctypes.c_int16))
    return tab


def blur_sums(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    ch = 1 if img.ndim == 2 else img.shape[2]
    assert ch in (1, 3)
    s = ctypes.c_int64(0)
    s2 = ctypes.c_int64(0)
    lib().orc_blur_sums(_p(img, ctypes.c_uint8), img.shape[0], img.shape[1], ch,
                        ctypes.byref(s), ctypes.byref(s2))
    return s.value, s2.value


def score_from_sums(s, s2, n):
    return lib().orc_score_from_sums(int(s), int(s2), int(n))


def blur_score(img):
    if img is None:
        return 0.0
    s, s2 = blur_sums(img)
    return score_from_sums(s, s2, img.shape[0] * img.shape[1])


def extract(frames, R, d, fov_deg, interp=INTERP_LINEAR, want_scores=True):
    """frames uint8 [F][H][W][3], R fp64 [V][3][3] -> (views u8 [F][V][d][d][3], sums)."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    F, H, W, _ = frames.shape
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 9)
    V = R.shape[0]
    out = np.empty((F, V, d, d, 3), np.uint8)
    s = np.zeros((F, V), np.int64)
    s2 = np.zeros((F, V), np.int64)
    lib().orc_extract(_p(frames, ctypes.c_uint8), F, H, W, _p(R, ctypes.c_double), V, d,
                      float(fov_deg), int(interp), _p(out, ctypes.c_uint8),
                      _p(s, ctypes.c_int64) if want_scores else None,
                      _p(s2, ctypes.c_int64) if want_scores else None)
    if not want_scores:
        return out, None, None
    return out, s, s2


def gauss13_taps(sigma=2.0):
    """The 13 fixed-point taps (sum 256) of cv2.GaussianBlur(img, (0, 0), sigma) on uint8."""
    t = (ctypes.c_int * 13)()
    lib().orc_gauss13_taps(float(sigma), t)
    return list(t)


def unsharp(img, strength):
    """processor.py:379-381 on one uint8 image [h][w] or [h][w][ch]."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    ch = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    lib().orc_unsharp(_p(img, ctypes.c_uint8), img.shape[0], img.shape[1], ch, float(strength), _p(out, ctypes.c_uint8))
    return out


def add_weighted(a, g, strength):
    """cv2.addWeighted(a, 1 + s, g, -s, 0) on uint8 arrays of equal shape."""
    a = np.ascontiguousarray(a, dtype=np.uint8)
    g = np.ascontiguousarray(g, dtype=np.uint8)
    out = np.empty_like(a)
    lib().orc_add_weighted(_p(a, ctypes.c_uint8), _p(g, ctypes.c_uint8), a.size, float(strength), _p(out, ctypes.c_uint8))
    return out


# ---- baseline JPEG (x360_jpeg_oracle.c): what cv2.imwrite(..., [IMWRITE_JPEG_QUALITY, q]) writes -----------------------
def _jpeg_lib():
    L = lib()
    if not getattr(L, "_jpeg_ready", False):
        u8p, i16p, u16p = ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_int16), ctypes.POINTER(ctypes.c_uint16)
        L.orc_jpeg_encode.restype = ctypes.c_size_t
        L.orc_jpeg_encode.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_size_t]
        L.orc_jpeg_header.restype = ctypes.c_size_t
        L.orc_jpeg_header.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_size_t]
        L.orc_jpeg_coefficients.restype = ctypes.c_long
        L.orc_jpeg_coefficients.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, i16p]
        L.orc_jpeg_qtables.restype = None
        L.orc_jpeg_qtables.argtypes = [ctypes.c_int, u8p, u8p]
        L.orc_jpeg_huffman.restype = None
        L.orc_jpeg_huffman.argtypes = [u16p, u8p]
        L._jpeg_ready = True
    return L


def jpeg_encode(img, quality=95) -> bytes:
    """uint8 BGR [h][w][3] -> the bytes of the .jpg file the reference's save step writes."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    cap = 1024 + 2 * h * w * 3 + 4096
    buf = np.empty(cap, np.uint8)
    n = _jpeg_lib().orc_jpeg_encode(_p(img, ctypes.c_uint8), h, w, int(quality), _p(buf, ctypes.c_uint8), cap)
    assert 0 < n <= cap
    return buf[:n].tobytes()


def jpeg_header(h, w, quality=95) -> bytes:
    buf = np.empty(1024, np.uint8)
    n = _jpeg_lib().orc_jpeg_header(int(h), int(w), int(quality), _p(buf, ctypes.c_uint8), 1024)
    return buf[:n].tobytes()


def jpeg_coefficients(img, quality=95):
    """int16 [mcus][6][64]: quantised coefficients per 16 x 16 MCU (Y00 Y01 Y10 Y11 Cb Cr), zigzag order."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    n = ((h + 15) // 16) * ((w + 15) // 16)
    out = np.empty((n, 6, 64), np.int16)
    _jpeg_lib().orc_jpeg_coefficients(_p(img, ctypes.c_uint8), h, w, int(quality), _p(out, ctypes.c_int16))
    return out


def jpeg_qtables(quality=95):
    a, b = np.empty(64, np.uint8), np.empty(64, np.uint8)
    _jpeg_lib().orc_jpeg_qtables(int(quality), _p(a, ctypes.c_uint8), _p(b, ctypes.c_uint8))
    return a, b


def jpeg_huffman():
    code, size = np.empty((4, 256), np.uint16), np.empty((4, 256), np.uint8)
    _jpeg_lib().orc_jpeg_huffman(_p(code, ctypes.c_uint16), _p(size, ctypes.c_uint8))
    return code, size

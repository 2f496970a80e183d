"""``ImageUtils.calculate_blur_score`` on the GPU.

Mirror of the reference's src/utils/image_utils.py:4-27: ``None`` -> 0.0, 3-channel input is
converted BGR->gray first, the score is the population variance of the 4-neighbour Laplacian.
The kernel returns the two exact integer sums; the variance is finished here in fp64.
"""
from __future__ import annotations

import ctypes

import numpy as np

from ._capi import check, lib, ptr
from .extractor import scores_from_sums


class ImageUtils:
    @staticmethod
    def calculate_blur_score(image, device=0) -> float:
        if image is None:                                   # image_utils.py:18-19
            return 0.0
        img = np.ascontiguousarray(image)
        if img.dtype != np.uint8:
            raise TypeError("calculate_blur_score expects an 8-bit image (the reference's inputs are always uint8)")
        if img.ndim == 3 and img.shape[2] == 3:
            ch = 3
        elif img.ndim == 2 or (img.ndim == 3 and img.shape[2] == 1):
            ch = 1
        else:
            raise ValueError(f"unsupported image shape {img.shape}")
        s, s2 = ctypes.c_int64(0), ctypes.c_int64(0)
        check(lib().x360_blur_sums(int(device), ptr(img), img.shape[0], img.shape[1], ch,
                                   ctypes.byref(s), ctypes.byref(s2)))
        return float(scores_from_sums(s.value, s2.value, img.shape[0] * img.shape[1]))


__all__ = ["ImageUtils"]

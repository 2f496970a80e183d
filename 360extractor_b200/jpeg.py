"""Baseline JPEG encode of view batches on the GPU: the reference's save step without the trip of raw pixels over PCIe.

Reference: ``FileManager.save_image`` -> ``cv2.imwrite(path, image, [cv2.IMWRITE_JPEG_QUALITY, quality])``
(src/utils/file_manager.py:21-32; ``save_params`` at src/core/processor.py:121-123; quality default 95,
src/core/settings_manager.py:24).  The bytes produced here are identical to cv2's for the same pixels (parity definition in
include/x360.h; pinned by tests/test_gpu_parity.py against cv2.imencode and a plain-C restatement).  No CPU path.
"""
from __future__ import annotations

import ctypes

import numpy as np

from ._capi import check, lib, ptr


def jpeg_header(h: int, w: int, quality: int = 95) -> bytes:
    """The SOI..SOS header of an ``h x w`` file (host-only)."""
    buf = np.empty(1024, np.uint8)
    n = lib().x360_jpeg_header(int(h), int(w), int(quality), ptr(buf), 1024)
    if n < 0:
        check(n)
    return buf[:n].tobytes()


class JpegEncoder:
    """Encoder for uint8 BGR images of one size; ``encode`` returns one ``bytes`` object per image."""

    def __init__(self, h: int, w: int, quality: int = 95, max_images: int = 16, device: int = 0):
        self.h, self.w, self.quality, self.max_images, self.device = int(h), int(w), int(quality), int(max_images), int(device)
        self._enc = ctypes.c_void_p()
        check(lib().x360_jpeg_create(self.device, self.h, self.w, self.max_images, self.quality, ctypes.byref(self._enc)))

    def close(self):
        if getattr(self, "_enc", None):
            lib().x360_jpeg_destroy(self._enc)
            self._enc = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def max_bytes(self, n_images: int) -> int:
        return int(lib().x360_jpeg_max_bytes(self._enc, int(n_images)))

    def _batch(self, images):
        images = np.asarray(images)
        if images.dtype != np.uint8:
            raise TypeError("images must be uint8 BGR")
        images = np.ascontiguousarray(images)
        if images.ndim == 3:
            images = images[None]
        if images.ndim != 4 or images.shape[1:] != (self.h, self.w, 3):
            raise ValueError(f"images shape {images.shape} does not match the encoder ({self.h}, {self.w}, 3)")
        return images

    def encode(self, images):
        """numpy uint8 ``[h][w][3]`` or ``[N][h][w][3]`` -> list of ``bytes`` (the .jpg files)."""
        images = self._batch(images)
        n = images.shape[0]
        out = np.empty(self.max_bytes(min(n, self.max_images)) * ((n + self.max_images - 1) // self.max_images or 1), np.uint8)
        off = np.zeros(n + 1, np.int64)
        check(lib().x360_jpeg_encode_host(self._enc, ptr(images), n, ptr(out), out.size, ptr(off)))
        return [out[off[i]:off[i + 1]].tobytes() for i in range(n)]

    def encode_device(self, images, out, offsets, stream=None):
        """Device-resident batch (torch CUDA tensors): images uint8 ``[N][h][w][3]``, out uint8 ``[capacity]``,
        offsets int64 ``[N + 1]``; asynchronous on ``stream``.  Call ``status(stream)`` before trusting the result."""
        n = int(images.shape[0])
        if tuple(images.shape[1:]) != (self.h, self.w, 3):
            raise ValueError("images must be [N][h][w][3]")
        st = getattr(stream, "cuda_stream", stream) if stream is not None else None
        check(lib().x360_jpeg_encode_device(self._enc, ptr(images), n, ptr(out), int(out.numel() if hasattr(out, "numel") else out.size),
                                            ptr(offsets), st))

    def status(self, stream=None):
        st = getattr(stream, "cuda_stream", stream) if stream is not None else None
        check(lib().x360_jpeg_status(self._enc, st))

    def coefficients(self, images):
        """Quantised DCT coefficients int16 ``[N][mcus][6][64]`` (zigzag order; Y00 Y01 Y10 Y11 Cb Cr per 16 x 16 MCU)."""
        images = self._batch(images)
        n = images.shape[0]
        mcus = ((self.h + 15) // 16) * ((self.w + 15) // 16)
        coef = np.empty((n, mcus, 6, 64), np.int16)
        check(lib().x360_jpeg_coefficients_host(self._enc, ptr(images), n, ptr(coef)))
        return coef


def encode_jpeg(images, quality: int = 95, device: int = 0):
    """One-shot helper: list of ``bytes`` for a batch of equally sized uint8 BGR images."""
    images = np.asarray(images)
    single = images.ndim == 3
    batch = images[None] if single else images
    with JpegEncoder(batch.shape[1], batch.shape[2], quality, max_images=min(16, max(1, batch.shape[0])), device=device) as enc:
        files = enc.encode(batch)
    return files[0] if single else files


__all__ = ["JpegEncoder", "encode_jpeg", "jpeg_header"]

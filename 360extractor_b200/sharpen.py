"""The sharpening step of the per-view pipeline on the GPU.

Reference (src/core/processor.py:379-381, repeated in src/ui/preview_widget.py:118-120), applied to every view
the blur filter accepted when ``sharpening_enabled`` is set (strength = ``sharpening_strength``, default 0.5,
processor.py:228-229)::

    gaussian = cv2.GaussianBlur(rect_img, (0, 0), 2.0)
    rect_img = cv2.addWeighted(rect_img, 1.0 + sharpen_strength, gaussian, -sharpen_strength, 0)

``unsharp_mask`` does both in one CUDA kernel per batch, bit-exact with OpenCV's fixed-point 13-tap Gaussian
and float32 ``addWeighted``.  No CPU path.
"""
from __future__ import annotations

import numpy as np

from ._capi import check, lib, ptr


def unsharp_mask(images, strength=0.5, device=0, out=None):
    """images uint8 ``[h][w][3]`` or ``[N][h][w][3]`` (numpy) -> sharpened copy of the same shape."""
    images = np.ascontiguousarray(images, dtype=np.uint8)
    single = images.ndim == 3
    batch = images[None] if single else images
    if batch.ndim != 4 or batch.shape[3] != 3:
        raise ValueError("images must be uint8 [h][w][3] or [N][h][w][3] (BGR)")
    if out is None:
        out = np.empty_like(batch)
    else:
        out = out[None] if out.ndim == 3 else out
        if out.shape != batch.shape or out.dtype != np.uint8 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous uint8 array shaped like images")
    n, h, w, _ = batch.shape
    check(lib().x360_unsharp_host(int(device), ptr(batch), n, h, w, float(strength), ptr(out)))
    return out[0] if single else out


def _device_of(t, device):
    """CUDA ordinal of a torch tensor (``device`` = explicit override for raw pointers); restores nothing itself."""
    idx = getattr(getattr(t, "device", None), "index", None)
    if idx is None:
        return 0 if device is None else int(device)
    if device is not None and int(device) != idx:
        raise ValueError(f"tensor lives on cuda:{idx}, not on device {device}")
    return int(idx)


def unsharp_mask_device(images, out, strength=0.5, device=None, stream=None):
    """Device-resident batch (torch CUDA uint8 tensors ``[N][h][w][3]``, ``out`` distinct from ``images``);
    asynchronous on ``stream``; runs on the device that holds ``images``."""
    if tuple(images.shape) != tuple(out.shape) or len(images.shape) != 4 or images.shape[3] != 3:
        raise ValueError("images / out must both be [N][h][w][3]")
    for name, t in (("images", images), ("out", out)):
        if not (t.is_contiguous() if hasattr(t, "is_contiguous") else t.flags.c_contiguous):
            raise ValueError(f"{name} must be C-contiguous")
    st = getattr(stream, "cuda_stream", stream) if stream is not None else None
    n, h, w, _ = (int(v) for v in images.shape)
    dev = _device_of(images, device)
    prev = None
    try:
        import torch
        prev = torch.cuda.current_device() if torch.cuda.is_available() else None
    except ImportError:
        pass
    try:
        check(lib().x360_unsharp_device(dev, ptr(images), n, h, w, float(strength), ptr(out), st))
    finally:
        if prev is not None and prev != dev:
            torch.cuda.set_device(prev)

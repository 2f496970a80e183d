"""Device-resident hand-off of a view batch to the AI masker.

The reference copies every extracted view back to the host, collects the frame's views in ``batch_images`` and passes
them — all of them, or only the faces named by ``ai_mask_cameras`` — to ``AIService.process_batch``
(src/core/processor.py:392-420, src/core/ai_model.py:125-167), whose model normalises each BGR image to a CHW float
tensor on the GPU again.  Here the views stay where ``Extractor.extract_device`` wrote them: ``masker_input`` selects
the eligible (frame, view) slots with the reference's own rule and converts them, in one CUDA kernel behind the C-ABI
(``x360_views_to_nchw``), to the float32 ``[N][3][d][d]`` RGB tensor a torch model takes directly.  ``scatter_results``
puts the model's per-image results back into the reference's ``ai_results`` order.
"""
from __future__ import annotations

import numpy as np

from ._capi import check, lib, ptr


def normalize_mask_faces(ai_mask_cameras):
    """The ``ai_mask_cameras`` setting as the masker's face filter: ``None`` (every face) when the setting is empty or
    names nothing, else the named faces as a lowercase set.  Behaviour of src/core/settings_manager.py:103-116 (read at
    src/core/processor.py:194); a maintainer wiring this into the reference would import that function instead."""
    if isinstance(ai_mask_cameras, str):
        tokens = ai_mask_cameras.split(",")
    else:
        tokens = list(ai_mask_cameras) if ai_mask_cameras else []
    names = set()
    for token in tokens:
        name = str(token).strip().lower()
        if name:
            names.add(name)
    return names if names else None


def eligible_indices(batch_names, mask_face_filter):
    """Positions of the views the masker sees (src/core/processor.py:396,405-408): all of them without a filter,
    else those whose camera name is in the filter, case-insensitively."""
    if mask_face_filter is None:
        return list(range(len(batch_names)))
    return [i for i, n in enumerate(batch_names) if n.lower() in mask_face_filter]


def masker_slots(n_frames, names, ai_mask_cameras=None, keep=None):
    """Flat view slots ``f * V + v`` handed to the masker, in (frame, view) order.  ``keep`` (bool ``[F][V]``, e.g. the
    blur filter's accept table) drops the views the reference never put in ``batch_images`` (processor.py:372-376)."""
    face_filter = normalize_mask_faces(ai_mask_cameras)
    V = len(names)
    slots = []
    for f in range(int(n_frames)):
        kept = [v for v in range(V) if keep is None or bool(keep[f][v])]
        for i in eligible_indices([names[v] for v in kept], face_filter):
            slots.append(f * V + kept[i])
    return slots


def masker_input(views, names, ai_mask_cameras=None, keep=None, to_rgb=True, out=None, stream=None, device=None):
    """Device views -> ``(slots, tensor)``.

    ``views``: torch CUDA uint8 ``[F][V][d][d][3]`` (what ``Extractor.extract_device`` filled);
    ``tensor``: torch CUDA float32 ``[len(slots)][3][d][d]``, ``v / 255``, RGB unless ``to_rgb=False``.
    Nothing crosses PCIe; asynchronous on ``stream`` (default: the current torch stream).  The kernel runs on the
    device that holds ``views`` (``device`` is only checked against it).

    Scope: this is tensor plumbing.  It does NOT reproduce the masker's own pre-processing — Ultralytics letterboxes each
    image to the model's ``imgsz`` inside ``self.model(images, ...)`` (src/core/ai_model.py:145), a third-party step that is
    absent from this container and therefore unpinned; feed it the sharpened views when sharpening is on, as the reference
    does (src/core/processor.py:379-383)."""
    import torch
    if views.dtype != torch.uint8 or views.dim() != 5 or views.shape[-1] != 3 or not views.is_cuda or not views.is_contiguous():
        raise ValueError("views must be a contiguous CUDA uint8 tensor [F][V][h][w][3]")
    F, V, h, w, _ = (int(x) for x in views.shape)
    dev_index = views.device.index if views.device.index is not None else torch.cuda.current_device()
    if device is not None and int(device) != dev_index:
        raise ValueError(f"views live on cuda:{dev_index}, not on device {device}")
    if len(names) != V:
        raise ValueError(f"{len(names)} names for {V} views")
    slots = masker_slots(F, names, ai_mask_cameras, keep)
    if out is None:
        out = torch.empty((len(slots), 3, h, w), dtype=torch.float32, device=views.device)
    elif tuple(out.shape) != (len(slots), 3, h, w) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous float32 tensor of shape {(len(slots), 3, h, w)}")
    if slots:
        sel = np.asarray(slots, dtype=np.int32)
        st = stream if stream is not None else torch.cuda.current_stream(views.device)
        prev = torch.cuda.current_device()                       # the library selects its device; leave torch's as it was
        try:
            check(lib().x360_views_to_nchw(dev_index, ptr(views), F * V, h, w, sel.ctypes.data, len(slots), 1 if to_rgb else 0,
                                           ptr(out), getattr(st, "cuda_stream", st)))
        finally:
            torch.cuda.set_device(prev)
    return slots, out


def scatter_results(slots, sub_results, images):
    """``ai_results`` of the reference (processor.py:403-417): ``(img, None)`` for every view, the masker's result in
    the slots it saw.  ``images``: the flat list of views in slot order (device tensors or arrays)."""
    results = [(img, None) for img in images]
    for slot, res in zip(slots, sub_results):
        results[slot] = res
    return results


__all__ = ["normalize_mask_faces", "eligible_indices", "masker_slots", "masker_input", "scatter_results"]

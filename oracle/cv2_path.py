"""The reference's call sequence over the SAME third-party libraries it uses.

CPU ORACLE — TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference delegates all pixel arithmetic on this path to ``opencv-python``
and ``numpy`` (requirements.txt:2-3).  This module restates the Python that
sits on top of them, so that the live ``cv2.remap`` / ``cv2.Laplacian`` of the
installed wheel can (a) cross-check the plain-C restatement and (b) be timed as
the CPU baseline on the GPU box's own cores (``/root/reference`` does not exist
there, the cv2/numpy wheels do).

Follows (relative to /root/reference):
  layout_views        src/core/geometry.py:10-77
  rotation            src/core/geometry.py:80-119
  rectilinear_map     src/core/geometry.py:122-192
  remap               src/core/processor.py:199-200,334
  blur_score          src/utils/image_utils.py:6-27
  extract             src/core/processor.py:245-265 (maps once per job) and
                      :327-342 (per frame, per view remap [+ score])
  masker_tensor       src/core/processor.py:392-420 -> src/core/ai_model.py:145 (``self.model(images, ...)``): the
                      model's input normalisation lives in third-party ``ultralytics`` (requirements.txt, unpinned,
                      absent here); its published ``BasePredictor.preprocess`` is ``im[..., ::-1].transpose(0, 3, 1, 2)``,
                      ``.float()``, ``/= 255`` — restated with numpy float32 (no letterbox: views are square and the
                      device hand-off keeps their size)
"""
from __future__ import annotations

import numpy as np

try:  # cv2 is part of the image on both the build container and the GPU box
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def have_cv2() -> bool:
    return cv2 is not None


def layout_views(n, pitch_offset=0, layout_mode="ring"):
    out = []
    if layout_mode == "cube":
        for name, yaw in (("Front", 0.0), ("Right", 90.0), ("Back", 180.0), ("Left", 270.0)):
            out.append((name, yaw, pitch_offset, 0))
        out.append(("Up", 0.0, 90.0, 0))
        out.append(("Down", 0.0, -90.0, 0))
    elif layout_mode == "fibonacci":
        dst = 2.0 / n
        inc = np.pi * (3.0 - np.sqrt(5.0))
        for i in range(n):
            y = 1 - (i * dst) - (dst / 2)
            r = np.sqrt(1 - y * y)
            phi = i * inc
            x = np.cos(phi) * r
            z = np.sin(phi) * r
            pitch = np.degrees(np.arcsin(y))
            yaw = np.degrees(np.arctan2(x, z)) % 360.0
            out.append((f"View_{i}", yaw, pitch + pitch_offset, 0))
    else:
        for i in range(n):
            out.append((f"View_{i}", (i * 360.0) / n, pitch_offset, 0))
    return out


def rotation(yaw_deg, pitch_deg, roll_deg):
    a, b, g = np.radians(yaw_deg), np.radians(pitch_deg), np.radians(roll_deg)
    rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    rz = np.array([[np.cos(g), -np.sin(g), 0], [np.sin(g), np.cos(g), 0], [0, 0, 1]])
    return ry @ rx @ rz


def rectilinear_map(src_h, src_w, dest_h, dest_w, fov_deg, yaw_deg, pitch_deg, roll_deg):
    f = (0.5 * dest_w) / np.tan(0.5 * np.radians(fov_deg))
    cx, cy = dest_w / 2, dest_h / 2
    x, y = np.meshgrid(np.arange(dest_w), np.arange(dest_h))
    rays = np.stack(((x - cx) / f, (y - cy) / f, np.ones((dest_h, dest_w))), axis=-1).reshape(-1, 3)
    rot = rays @ rotation(yaw_deg, pitch_deg, roll_deg).T
    xr, yr, zr = rot[:, 0], rot[:, 1], rot[:, 2]
    theta = np.arctan2(xr, zr)
    phi = np.arcsin(yr / np.sqrt(xr**2 + yr**2 + zr**2))
    uf = (theta / (2 * np.pi) + 0.5) * src_w
    vf = (phi / np.pi + 0.5) * src_h
    return (uf.reshape(dest_h, dest_w).astype(np.float32),
            vf.reshape(dest_h, dest_w).astype(np.float32))


def interp_flag(interpolation_mode):
    return cv2.INTER_LANCZOS4 if interpolation_mode == "lanczos" else cv2.INTER_LINEAR


def remap(frame, map_x, map_y, interpolation_mode="linear"):
    return cv2.remap(frame, map_x, map_y, interp_flag(interpolation_mode), borderMode=cv2.BORDER_WRAP)


def blur_score(image):
    if image is None:
        return 0.0
    gray = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY) if image.ndim == 3 else image
    return cv2.Laplacian(gray, cv2.CV_64F).var()


def build_maps(src_h, src_w, d, fov_deg, views, active_cameras=None):
    """processor.py:245-265 — maps keyed by view name, once per job."""
    maps = {}
    for i, (name, yaw, pitch, roll) in enumerate(views):
        if active_cameras is not None and i not in active_cameras:
            continue
        maps[name] = rectilinear_map(src_h, src_w, d, d, fov_deg, yaw, pitch, roll)
    return maps


def extract(frames, views, maps, interpolation_mode="linear", want_scores=False, keep_views=True):
    """processor.py:327-342 body over a frame batch.

    Returns (list per frame of list per emitted view of uint8 [d][d][3] or None,
             float64 scores [F][V_emitted] or None)."""
    names = [v[0] for v in views if v[0] in maps]
    outs, scores = [], []
    for frame in frames:
        row, srow = [], []
        for name in names:
            mx, my = maps[name]
            img = remap(frame, mx, my, interpolation_mode)
            if want_scores:
                srow.append(blur_score(img))
            row.append(img if keep_views else None)
        outs.append(row)
        scores.append(srow)
    return outs, (np.array(scores, dtype=np.float64) if want_scores else None)


def unsharp(image, strength):
    """src/core/processor.py:379-381 verbatim (also src/ui/preview_widget.py:118-120)."""
    import cv2
    gaussian = cv2.GaussianBlur(image, (0, 0), 2.0)
    return cv2.addWeighted(image, 1.0 + strength, gaussian, -strength, 0)


def masker_tensor(images, to_rgb=True):
    """uint8 BGR [N][h][w][3] -> float32 [N][3][h][w] in [0, 1] (RGB unless ``to_rgb=False``)."""
    im = np.asarray(images, dtype=np.uint8)
    if to_rgb:
        im = im[..., ::-1]
    im = np.ascontiguousarray(im.transpose(0, 3, 1, 2)).astype(np.float32)
    im /= np.float32(255)
    return im

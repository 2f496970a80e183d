"""Camera layouts and per-view rotations: the host half of the hot path.

Mirror of the reference's ``GeometryProcessor`` (src/core/geometry.py:3-192) with the same three
static entry points, argument meaning and return types, so its callers
(src/core/processor.py:247,263; src/core/analyzer.py:51,59; tests/test_core.py:21-78) work unchanged:

* ``generate_views``        layouts stay on the host (a handful of scalars, geometry.py:10-77);
* ``get_rotation_matrix``   stays on the host with the same numpy evaluation order, because the
                            fp64 residues (sin(pi) = 1.2e-16 ...) decide which side of the +-pi
                            seam a pixel falls on (geometry.py:80-119);
* ``create_rectilinear_map`` runs ON THE GPU (libx360 ``x360_make_maps``): the batch path never
                            materialises maps, this exists for callers that still want the arrays.
"""
from __future__ import annotations


import numpy as np

from . import _capi

_CUBE_FACES = (("Front", 0.0), ("Right", 90.0), ("Back", 180.0), ("Left", 270.0))
_GOLDEN_ANGLE = np.pi * (3.0 - np.sqrt(5.0))          # geometry.py:43


def _fibonacci_direction(i: int, n: int):
    """(yaw, pitch) in degrees of point i of an n-point golden-spiral sphere (geometry.py:45-64)."""
    step = 2.0 / n
    height = 1 - (i * step) - (step / 2)               # 1 -> -1, same operation order as the reference
    ring_radius = np.sqrt(1 - height * height)
    azimuth = i * _GOLDEN_ANGLE
    east = np.cos(azimuth) * ring_radius
    north = np.sin(azimuth) * ring_radius
    pitch = np.degrees(np.arcsin(height))
    yaw = np.degrees(np.arctan2(east, north)) % 360.0
    return yaw, pitch


def _axis_rotation(axis: str, angle_rad):
    c, s = np.cos(angle_rad), np.sin(angle_rad)
    if axis == "x":
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    if axis == "y":
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


class GeometryProcessor:
    """Equirectangular -> rectilinear geometry (drop-in for the reference class of the same name)."""

    @staticmethod
    def generate_views(n, pitch_offset=0, layout_mode="ring"):
        """List of ``(name, yaw_deg, pitch_deg, roll)``.

        'cube' ignores ``n`` and tilts only its four horizontal faces; 'fibonacci' adds the offset
        to every spiral pitch (no clamp); anything else is an evenly spaced ring."""
        if layout_mode == "cube":
            views = [(name, yaw, pitch_offset, 0) for name, yaw in _CUBE_FACES]
            views += [("Up", 0.0, 90.0, 0), ("Down", 0.0, -90.0, 0)]
            return views
        if layout_mode == "fibonacci":
            views = []
            for i in range(n):
                yaw, pitch = _fibonacci_direction(i, n)
                views.append((f"View_{i}", yaw, pitch + pitch_offset, 0))
            return views
        return [(f"View_{i}", (i * 360.0) / n, pitch_offset, 0) for i in range(n)]

    @staticmethod
    def get_rotation_matrix(yaw_deg, pitch_deg, roll_deg):
        """fp64 3x3, yaw about Y, then pitch about X, then roll about Z: ``Ry @ Rx @ Rz``."""
        ry = _axis_rotation("y", np.radians(yaw_deg))
        rx = _axis_rotation("x", np.radians(pitch_deg))
        rz = _axis_rotation("z", np.radians(roll_deg))
        return ry @ rx @ rz

    @staticmethod
    def create_rectilinear_map(src_h, src_w, dest_h, dest_w, fov_deg, yaw_deg, pitch_deg, roll_deg, device=0):
        """float32 ``(map_x, map_y)`` of shape ``(dest_h, dest_w)`` for ``cv2.remap`` — computed on the GPU."""
        from .extractor import Extractor

        R = GeometryProcessor.get_rotation_matrix(yaw_deg, pitch_deg, roll_deg)
        with Extractor(src_h, src_w, dest_w, fov_deg, R[None], out_h=dest_h, device=device) as ex:
            return ex.make_maps(0)


def rotation_stack(views, active_cameras=None):
    """fp64 ``[V][3][3]`` for the emitted views, plus their names.

    ``active_cameras`` holds indices into ``views`` (src/core/processor.py:257-261); views are kept
    in ``generate_views`` order, which is the order the frame loop iterates (processor.py:327-330)."""
    names, mats = [], []
    for i, (name, yaw, pitch, roll) in enumerate(views):
        if active_cameras is not None and i not in active_cameras:
            continue
        names.append(name)
        mats.append(GeometryProcessor.get_rotation_matrix(yaw, pitch, roll))
    if not mats:
        return names, np.zeros((0, 3, 3), np.float64)
    return names, np.ascontiguousarray(np.stack(mats), dtype=np.float64)


def focal_length(dest_w: int, fov_deg: float) -> float:
    """f of geometry.py:141 as libx360 computes it (host helper in the library, no GPU needed)."""
    return float(_capi.lib().x360_focal(int(dest_w), float(fov_deg)))


__all__ = ["GeometryProcessor", "rotation_stack", "focal_length"]

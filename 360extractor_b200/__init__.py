"""x360 — B200-native (sm_100a) equirectangular -> rectilinear extraction.

Drop-in for ONE hot path of nicolasdiolez/360Extractor: camera-layout rays, on-the-fly
``map_x/map_y``, the ``cv2.remap`` gather (bilinear / Lanczos4, wrap-around) and the fused
Laplacian-variance blur score.  All pixel work runs in hand-written CUDA behind a C-ABI
(``include/x360.h``, ``lib/libx360.so``); importing this package never falls back to a CPU path.

The directory name starts with a digit, so import it with
``importlib.import_module("360extractor_b200")`` or through the ``x360`` alias at the repo root.
"""
from ._capi import (FLAG_FULL_UPLOAD, FLAG_FUSED_SCORE, FLAG_NO_WIDE, FLAG_PLANES, FLAG_SPLIT_SCORE, FLAG_TILE_H16, FLAG_TILE_H32, INTERP_LANCZOS4,
                    INTERP_LINEAR, X360Error, launch_count, version)
from .analyzer import BlurAnalyzer
from .blur_filter import BlurFilter, blur_replay
from .extractor import Extractor, PinnedBuffer, column_only_map_x, plan_preview, remap, scores_from_sums
from .geometry import GeometryProcessor, focal_length, rotation_stack
from .handoff import eligible_indices, masker_input, masker_slots, normalize_mask_faces, scatter_results
from .image_utils import ImageUtils
from .jpeg import JpegEncoder, encode_jpeg, jpeg_header
from .pipeline import ExtractionSession, ViewResult
from .sharding import gather_sums, gather_sums_by_view, shard_axis, shard_bounds, shard_sizes
from .sharpen import unsharp_mask, unsharp_mask_device

__version__ = "0.1.0"

__all__ = [
    "GeometryProcessor", "ImageUtils", "Extractor", "ExtractionSession", "ViewResult", "BlurFilter", "blur_replay", "BlurAnalyzer",
    "PinnedBuffer", "column_only_map_x", "plan_preview", "remap", "unsharp_mask", "unsharp_mask_device", "scores_from_sums", "rotation_stack", "focal_length",
    "JpegEncoder", "encode_jpeg", "jpeg_header",
    "gather_sums", "gather_sums_by_view", "shard_axis", "shard_bounds", "shard_sizes",
    "normalize_mask_faces", "eligible_indices", "masker_slots", "masker_input", "scatter_results",
    "INTERP_LINEAR", "INTERP_LANCZOS4", "X360Error", "launch_count", "version",
    "FLAG_PLANES", "FLAG_FUSED_SCORE", "FLAG_SPLIT_SCORE", "FLAG_TILE_H16", "FLAG_TILE_H32", "FLAG_NO_WIDE", "FLAG_FULL_UPLOAD",
]

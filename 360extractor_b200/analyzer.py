"""``BlurAnalyzer``: the reference's "Analyze blur" caller of the hot path on the batched GPU entry.

Mirror of ``src/core/analyzer.py:6-76``: one sample frame is reprojected with the RING layout
(``generate_views(camera_count, pitch_offset)`` — ``layout_mode`` and ``interpolation_mode`` are
ignored there, analyzer.py:51,63), bilinear + BORDER_WRAP, default resolution 1024 (analyzer.py:45),
and every view's Laplacian-variance score is reported.  Here all views and scores come from ONE
``x360_extract`` launch (fused scores); the result dict has the reference's keys and types.

Decoding the sample frame (``cv2.VideoCapture``: seek to the middle frame, analyzer.py:24-42) is I/O
outside the hot path and stays with the caller — this package never imports OpenCV;
``analyze_frame`` takes the decoded frame and does everything from analyzer.py:44 on.
"""
from __future__ import annotations

import numpy as np

from .extractor import Extractor
from .geometry import GeometryProcessor, rotation_stack


class BlurAnalyzer:
    @staticmethod
    def analyze_frame(frame, settings, device: int = 0):
        """frame uint8 ``[H][W][3]`` BGR -> ``{'average', 'min', 'max', 'details': [(view_name, score)]}``
        (analyzer.py:44-76)."""
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        if frame.ndim != 3 or frame.shape[2] != 3:
            raise ValueError(f"expected a BGR frame [H][W][3], got shape {frame.shape}")
        out_res = settings.get('resolution', 1024)                      # analyzer.py:45-48
        fov = settings.get('fov', 90)
        camera_count = settings.get('camera_count', 6)
        pitch_offset = settings.get('pitch_offset', 0)
        views = GeometryProcessor.generate_views(camera_count, pitch_offset=pitch_offset)   # always ring (analyzer.py:51)
        if not views:
            return {'average': 0, 'min': 0, 'max': 0, 'details': []}    # analyzer.py:68-69
        names, R = rotation_stack(views)
        src_h, src_w = frame.shape[:2]
        with Extractor(src_h, src_w, out_res, fov, R, "linear", device=device) as ex:     # INTER_LINEAR (analyzer.py:63)
            _, scores = ex.extract(frame[None], want_scores=True)
        scores = [float(s) for s in scores[0]]
        return {
            'average': sum(scores) / len(scores),                        # analyzer.py:71-76
            'min': min(scores),
            'max': max(scores),
            'details': list(zip(names, scores)),
        }

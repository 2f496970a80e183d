"""Smallest staged-path run: one 96x192 frame, one view, d=48 (used under compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, x360
src = np.random.default_rng(7).integers(0, 256, (1, 96, 192, 3), dtype=np.uint8)
R = x360.GeometryProcessor.get_rotation_matrix(0, 0, 0)[None]
with x360.Extractor(96, 192, 48, 90, R, "linear") as ex:
    out, s = ex.extract(src, want_scores=False)
    print("ok", out.sum(), ex.tile_modes())

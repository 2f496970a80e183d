"""Small tiled-path run for compute-sanitizer: 3 frames, ring of 3 views (one shared map) + a cube (poles,
seam: fallback tiles), scores on and off, checked against the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import x360
from oracle import c_oracle

rng = np.random.default_rng(7)
frames = rng.integers(0, 256, (3, 192, 384, 3), dtype=np.uint8)
for layout, n, d in (("ring", 3, 80), ("cube", 6, 64)):
    views = x360.GeometryProcessor.generate_views(n, pitch_offset=-20 if layout == "ring" else 0, layout_mode=layout)
    R = x360.rotation_stack(views)[1]
    want, ws, ws2 = c_oracle.extract(frames, R, d, 90, 0)
    with x360.Extractor(192, 384, d, 90, R, "linear") as ex:
        out, s, s2 = ex.extract_sums(frames)
        modes = ex.tile_modes()
        out2, _ = ex.extract(frames, want_scores=False)
    assert np.array_equal(out, want) and np.array_equal(out2, want) and np.array_equal(s, ws) and np.array_equal(s2, ws2)
    print("ok", layout, modes)

import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import x360
from oracle import c_oracle
def run(H, W, d, fov, layout, n, pitch, F, scores, mb, tag, path=0):
    rng = np.random.default_rng(1)
    frames = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
    views = x360.GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode=layout)
    names, R = x360.rotation_stack(views)
    R = R[:8]
    want, ws, ws2 = c_oracle.extract(frames, R, d, fov, 0, want_scores=scores)
    with x360.Extractor(H, W, d, fov, R, "linear", max_batch=mb, path=path) as ex:
        if scores: got, s, s2 = ex.extract_sums(frames)
        else: got, _ = ex.extract(frames, want_scores=False)
        modes = ex.tile_modes()
    diff = np.argwhere(got != want)
    print(tag, "mb", mb, "path", path, "differing bytes", len(diff), modes)
    if len(diff):
        fv = sorted({(int(a), int(b)) for a, b, *_ in diff})
        print("  (frame, view):", fv[:20])
        ys = sorted({int(r[2]) for r in diff}); xs = sorted({int(r[3]) for r in diff})
        print("  rows", ys[:12], "...", ys[-3:], " cols", xs[:12], "...", xs[-3:])
for mb in (1, 2, 3, 6):
    run(510, 1020, 228, 90.0, "ring", 10, 45, 6, False, mb, "case27")
run(510, 1020, 228, 90.0, "ring", 10, 45, 6, False, 6, "case27 direct-kernel", path=1)
run(510, 1020, 228, 90.0, "ring", 10, 45, 1, False, 1, "case27 F=1")
os.environ["X360_NO_RAW"] = "1"
run(510, 1020, 228, 90.0, "ring", 10, 45, 6, False, 6, "case27 planes")
del os.environ["X360_NO_RAW"]
for mb in (1, 6):
    run(97, 972, 70, 80.0, "ring", 11, 50, 6, True, mb, "case171")
run(97, 972, 70, 80.0, "ring", 11, 50, 6, True, 6, "case171 direct-kernel", path=1)

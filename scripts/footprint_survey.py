"""Host emulation of extract_tiled's per-(tile, view) staging decision (x360_tiled.cuh): which share of the tiles lands in a
regular TMA box, a wide box, a seam box, or the direct gathers — from the exact Q5 maps of the oracle's cv2 path.

    python scripts/footprint_survey.py cfg1 [cfg4 ...]      (CPU only; workloads of bench.py)

A model of the mid-round-2 kernel: one capacity for all views (that of the group serving view 0), 64-byte raw classes and
upright 32 x th tiles everywhere.  The shipped kernel has since gained per-group capacities, the skewed class widths of the
general group (48 / 112 / 176 / 240) and transposed tiles; its own counts are `Extractor.tile_modes()` /
`Extractor.transposed_blocks()` after a launch (bench.py: `roofline.tile_modes`).
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import cv2_path  # noqa: E402

WIDE = [384, 576, 768, 960]
SEAM_HALF = 1024


def survey(name, scores=None, flags=0):
    pkg = importlib.import_module("360extractor_b200")
    wl = bench.WORKLOADS[name]
    W, H, d = wl["W"], wl["H"], wl["d"]
    scores = wl["scores"] if scores is None else scores
    lz = wl["mode"] == "lanczos"
    views = pkg.GeometryProcessor.generate_views(wl["n"], pitch_offset=wl["pitch"], layout_mode=wl["layout"])
    names, R = pkg.rotation_stack(views, wl["active"])
    pv = pkg.plan_preview(H, W, d, wl["fov"], R, wl["mode"], flags=flags)
    cap = pv["with_scores"] if scores else pv
    rows_max, bw_max, th = cap["rows_max"], cap["bw_max"], cap["tile_h"]
    raw = (not lz) and (not scores or flags & 2) and not (flags & 1)
    unit = 64 if raw else 48
    wide_rows = [(rows_max * bw_max) // b for b in WIDE]
    wide_rows = [r if r >= 6 else 0 for r in wide_rows]
    maps = cv2_path.build_maps(H, W, d, wl["fov"], views, wl["active"])
    M0, M1 = (3, 4) if lz else (0, 1)
    counts = {"regular": 0, "wide": 0, "seam": 0, "direct": 0}
    why = {}
    for v, nm in enumerate(names):
        mx, my = maps[nm]
        sx = np.rint(mx * np.float32(32)).astype(np.int64) >> 5
        sy = np.rint(my * np.float32(32)).astype(np.int64) >> 5
        for y0 in range(0, d, th):
            for x0 in range(0, d, 32):
                ya, yb, xa, xb = max(y0 - (1 if scores else 0), 0), min(y0 + th + (1 if scores else 0), d), max(x0 - (1 if scores else 0), 0), min(x0 + 32 + (1 if scores else 0), d)
                ix, iy = sx[ya:yb, xa:xb], sy[ya:yb, xa:xb]          # (a superset of the halo: the corners are not gathered)
                xmin, xmax, ymin, ymax = ix.min(), ix.max(), iy.min(), iy.max()
                seam = False
                if raw and (xmax - xmin > W // 2 or xmax + M1 > W - 1):
                    ixu = np.where(ix < W // 2, ix + W, ix)
                    xmin, xmax, seam = ixu.min(), ixu.max(), True
                x_org = xmin if raw else (xmin - M0) & ~3
                y_org = ymin - M0 - (1 if seam else 0)
                npx, rows = xmax + M1 - x_org, ymax + M1 + 1 - y_org
                gx0 = (3 * x_org) & ~15
                need = 3 * x_org - gx0 + 3 * (npx + 1) + (8 if seam else 0)
                bwr = (need + unit - 1) // unit * unit
                cls, bw, box_rows = None, 0, 0
                if rows <= rows_max and bwr <= bw_max:
                    cls, bw = "regular", bwr
                    box_rows = rows_max if seam else rows_max - 8 * min(2, (rows_max - rows) // 8)
                else:
                    for c in range(3, -1, -1):
                        if need <= WIDE[c] and rows <= wide_rows[c]:
                            cls, bw, box_rows = "wide", WIDE[c], wide_rows[c]
                ok = cls is not None and ymin - M0 >= 0 and ymax + M1 <= H - 1
                if seam:
                    ok = ok and y_org >= 0 and y_org + box_rows <= H - 1 and 3 * W - gx0 <= SEAM_HALF and gx0 + bw - 3 * W <= SEAM_HALF
                else:
                    ok = ok and xmin - M0 >= 0 and xmax + M1 <= W - 1
                if ok:
                    counts["seam" if seam else cls] += 1
                else:
                    counts["direct"] += 1
                    key = ("pole" if (ymin - M0 < 0 or ymax + M1 > H - 1) else ("seam" if seam or xmax - xmin > W // 2 else "size")) + ("" if cls else "+nofit")
                    why[key] = why.get(key, 0) + 1
    total = sum(counts.values())
    return {"workload": name, "scores": scores, "raw": bool(raw), "tile_h": th, "rows_max": rows_max, "bw_max": bw_max, "wide_rows": wide_rows,
            "tiles": total, "share": {k: round(v / total, 4) for k, v in counts.items()}, "direct_why": why}


if __name__ == "__main__":
    for n in sys.argv[1:] or ["cfg1"]:
        print(json.dumps(survey(n)))

"""Randomised parity run of the steps around the extraction on a B200: the unsharp mask (processor.py:379-381), the
stand-alone blur score (image_utils.py:6-27) and the masker hand-off, on seeded random shapes against the oracle.
Writes a JSON report to argv[1]; argv[2] = cases per step (default 150), argv[3] = seed."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import x360
from oracle import c_oracle, cv2_path

n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 150
rng = np.random.default_rng(int(sys.argv[3]) if len(sys.argv) > 3 else 11)
report = {"lib": x360.version(), "unsharp": 0, "blur": 0, "handoff": 0, "failures": []}
for c in range(n_cases):
    # unsharp: tiny, odd, TMA-eligible (w * 3 % 16 == 0, >= 44 rows, >= 86 columns) and large shapes
    h = int(rng.choice([int(rng.integers(1, 40)), int(rng.integers(40, 400))]))
    w = int(rng.choice([int(rng.integers(1, 90)), int(rng.integers(86, 500)), 16 * int(rng.integers(6, 40))]))
    n = int(rng.integers(1, 4))
    strength = float(rng.choice([0.5, 1.0, 0.25, round(float(rng.uniform(0, 3)), 3)]))
    imgs = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    if rng.random() < 0.3:
        imgs = (imgs // 32) * 32                       # flat-ish images: saturation / rounding ties
    got = x360.unsharp_mask(imgs, strength)
    want = np.stack([c_oracle.unsharp(im, strength) for im in imgs])
    report["unsharp"] += 1
    if not np.array_equal(got, want):
        report["failures"].append({"step": "unsharp", "h": h, "w": w, "n": n, "strength": strength, "diff": int(np.count_nonzero(got != want))})
    # blur score of one image
    img = imgs[0]
    s = x360.ImageUtils.calculate_blur_score(img)
    s1, s2 = c_oracle.blur_sums(img)
    N = float(h * w)
    ref = s2 / N - (s1 / N) ** 2
    report["blur"] += 1
    if not (abs(s - ref) <= 1e-9 * max(1.0, abs(ref))):
        report["failures"].append({"step": "blur", "h": h, "w": w, "got": s, "want": ref})
    # hand-off: random view batch, random face subset
    F, V, d = int(rng.integers(1, 4)), int(rng.integers(1, 9)), int(rng.integers(1, 97))
    views = rng.integers(0, 256, (F, V, d, d, 3), dtype=np.uint8)
    names = [f"View_{i}" for i in range(V)]
    cams = None if rng.random() < 0.4 else [names[i] for i in range(V) if rng.random() < 0.5]
    rgb = bool(rng.random() < 0.7)
    slots, t = x360.masker_input(torch.from_numpy(views).cuda(), names, cams, to_rgb=rgb)
    torch.cuda.synchronize()
    wantt = cv2_path.masker_tensor(views.reshape(F * V, d, d, 3)[slots], to_rgb=rgb) if slots else np.empty((0, 3, d, d), np.float32)
    report["handoff"] += 1
    if not np.array_equal(t.cpu().numpy(), wantt):
        report["failures"].append({"step": "handoff", "F": F, "V": V, "d": d, "cams": cams, "rgb": rgb})
report["all_exact"] = not report["failures"]
json.dump(report, open(sys.argv[1], "w"), indent=1)
print({k: report[k] for k in ("unsharp", "blur", "handoff", "all_exact")}, report["failures"][:5])

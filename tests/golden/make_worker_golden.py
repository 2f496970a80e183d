#!/usr/bin/env python
"""Golden fixture for the CALLER CONTRACT: drive the reference's own ``ProcessingWorker.process_video``
(src/core/processor.py:90-495, unmodified, imported from /root/reference/src) over a seeded clip and record what it did:
every blur score it computed (in call order), the files it kept, and a hash of every saved view.

The reference imports PySide6 / piexif / ultralytics / defusedxml, which this container lacks; none of them is on the
path being recorded, so they are stubbed with the ~20 lines SURVEY 8(c) describes (a QObject and a Signal with
connect / emit).  Scores are captured by wrapping ``ImageUtils.calculate_blur_score`` as seen from the processor module
(the wrapper calls the reference's function and appends its return value; nothing in the reference is edited).
Output is PNG (lossless), so a saved file is exactly the view the loop produced (after the sharpening step when on).

    python tests/golden/make_worker_golden.py      # writes tests/golden/worker_cases.json + worker_clip.npz

Only this script reads /root/reference; tests/test_host_api.py and tests/test_gpu_parity.py use the committed files.
"""
import hashlib
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get("X360_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def install_stubs():
    class Signal:
        def __init__(self, *types_):
            self._slots = []

        def __set_name__(self, owner, name):
            self._name = "_sig_" + name

        def __get__(self, obj, owner=None):
            if obj is None:
                return self
            inst = obj.__dict__.get(self._name)
            if inst is None:
                inst = Signal()
                obj.__dict__[self._name] = inst
            return inst

        def connect(self, fn):
            self._slots.append(fn)

        def emit(self, *a):
            for fn in list(self._slots):
                fn(*a)

    class QObject:
        def __init__(self, *a, **k):
            pass

    qtcore = types.ModuleType("PySide6.QtCore")
    qtcore.QObject, qtcore.Signal = QObject, Signal
    qtcore.QCoreApplication = type("QCoreApplication", (), {"__init__": lambda self, *a: None})
    pyside = types.ModuleType("PySide6")
    pyside.QtCore = qtcore
    sys.modules.update({"PySide6": pyside, "PySide6.QtCore": qtcore})
    sys.modules["piexif"] = types.ModuleType("piexif")
    ul = types.ModuleType("ultralytics")
    ul.YOLO = lambda *a, **k: None
    sys.modules["ultralytics"] = ul
    import xml.etree.ElementTree as ET
    dx = types.ModuleType("defusedxml")
    dxe = types.ModuleType("defusedxml.ElementTree")
    dxe.__dict__.update({k: getattr(ET, k) for k in dir(ET) if not k.startswith("__")})
    dx.ElementTree = dxe
    sys.modules.update({"defusedxml": dx, "defusedxml.ElementTree": dxe})
    pil = types.ModuleType("PIL")
    pil.Image = types.ModuleType("PIL.Image")
    sys.modules.setdefault("PIL", pil)
    sys.modules.setdefault("PIL.Image", pil.Image)


def make_clip(n_frames=14, H=96, W=192):
    """Seeded frames with sharp, soft and very blurry stretches, so every branch of the accept / reject machine fires:
    threshold rejections, rolling-average rejections, more than five consecutive skips (forced accept)."""
    import cv2
    rng = np.random.default_rng(2025)
    frames = []
    for i in range(n_frames):
        base = rng.integers(0, 256, (H // 2, W // 2, 3), dtype=np.uint8)
        img = cv2.resize(base, (W, H), interpolation=cv2.INTER_CUBIC)
        if i in (3, 4, 5, 6):                                # a blurry stretch: 4 frames x views > 5 consecutive skips
            img = cv2.GaussianBlur(img, (0, 0), 4.0)
        elif i in (9, 10):                                   # soft: above the floor, below 0.6 x the rolling mean
            img = cv2.GaussianBlur(img, (0, 0), 1.2)
        frames.append(img)
    return np.stack(frames)


SCENARIOS = [
    # name, settings overrides
    ("smart_ring4", {"layout_mode": "ring", "camera_count": 4, "blur_filter_enabled": True, "smart_blur_enabled": True, "blur_threshold": 120.0}),
    ("plain_threshold_cube", {"layout_mode": "cube", "camera_count": 6, "blur_filter_enabled": True, "smart_blur_enabled": False, "blur_threshold": 400.0}),
    ("smart_active_cameras", {"layout_mode": "ring", "camera_count": 6, "active_cameras": [0, 2, 5], "blur_filter_enabled": True, "smart_blur_enabled": True,
                              "blur_threshold": 80.0, "pitch_offset": -20}),
    ("filter_off_fibonacci", {"layout_mode": "fibonacci", "camera_count": 5, "blur_filter_enabled": False, "fov": 100}),
    ("smart_sharpened_lanczos", {"layout_mode": "ring", "camera_count": 3, "blur_filter_enabled": True, "smart_blur_enabled": True, "blur_threshold": 150.0,
                                 "sharpening_enabled": True, "sharpening_strength": 0.7, "interpolation_mode": "lanczos"}),
    ("legacy_adaptive_layout", {"layout_mode": "adaptive", "camera_count": 2, "blur_filter_enabled": True, "smart_blur_enabled": True, "blur_threshold": 1.0}),
]


def main():
    install_stubs()
    sys.path.insert(0, os.path.join(REF, "src"))
    import cv2
    import core.processor as proc                         # the reference's, unmodified
    from core.job import Job

    clip = make_clip()
    out = {"clip": {"frames": int(clip.shape[0]), "h": int(clip.shape[1]), "w": int(clip.shape[2]),
                    "note": "frames as decoded by cv2.VideoCapture from the MJPG file the worker read (stored in worker_clip.npz)"},
           "scenarios": {}, "cv2": cv2.__version__, "numpy": np.__version__}
    with tempfile.TemporaryDirectory() as tmp:
        video = os.path.join(tmp, "clip.avi")
        vw = cv2.VideoWriter(video, cv2.VideoWriter_fourcc(*"MJPG"), 10.0, (clip.shape[2], clip.shape[1]))
        assert vw.isOpened(), "cv2.VideoWriter cannot write MJPG here"
        for f in clip:
            vw.write(f)
        vw.release()
        cap = cv2.VideoCapture(video)
        decoded = []
        while True:
            ok, fr = cap.read()
            if not ok:
                break
            decoded.append(fr.copy())
        cap.release()
        decoded = np.stack(decoded)
        assert decoded.shape == clip.shape, decoded.shape

        real_score = proc.ImageUtils.calculate_blur_score
        for name, over in SCENARIOS:
            settings = {"resolution": 64, "fov": 90, "camera_count": 6, "pitch_offset": 0, "layout_mode": "ring", "interpolation_mode": "linear",
                        "output_format": "png", "interval_unit": "Frames", "interval_value": 1, "custom_output_dir": os.path.join(tmp, name),
                        "ai_mode": "None", "export_telemetry": False, "adaptive_mode": False, "naming_mode": "realityscan", "is_360": True}
            settings.update(over)
            os.makedirs(settings["custom_output_dir"], exist_ok=True)
            scores = []

            def recording(img, _scores=scores):
                s = real_score(img)
                _scores.append(float(s))
                return s

            proc.ImageUtils.calculate_blur_score = staticmethod(recording)
            try:
                worker = proc.ProcessingWorker([Job(file_path=video, settings=settings)])
                errors = []
                worker.job_error.connect(lambda i, msg, _e=errors: _e.append(msg))
                worker.run()
                worker.io_pool.shutdown(wait=True)
            finally:
                proc.ImageUtils.calculate_blur_score = real_score
            assert not errors, errors
            out_dir = os.path.join(settings["custom_output_dir"], "clip_processed")
            kept = {}
            for fn in sorted(os.listdir(out_dir)):
                img = cv2.imread(os.path.join(out_dir, fn), cv2.IMREAD_COLOR)
                kept[fn] = hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()
            out["scenarios"][name] = {"settings": {k: v for k, v in settings.items() if k != "custom_output_dir"},
                                      "scores_in_call_order": [s.hex() for s in scores], "kept": kept}
            print(name, len(scores), "scores,", len(kept), "files kept")
    np.savez_compressed(os.path.join(HERE, "worker_clip.npz"), frames=decoded)
    with open(os.path.join(HERE, "worker_cases.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote worker_cases.json, worker_clip.npz", os.path.getsize(os.path.join(HERE, "worker_clip.npz")), "bytes")


if __name__ == "__main__":
    main()

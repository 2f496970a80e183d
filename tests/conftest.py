import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def x360():
    """The product package, with libx360.so built if the toolchain is here (it ships prebuilt to the GPU box)."""
    build = importlib.import_module("360extractor_b200.build")
    try:
        build.build()
    except RuntimeError:
        if not os.path.exists(build.LIB_PATH):
            raise
    return importlib.import_module("360extractor_b200")


@pytest.fixture(scope="session")
def oracle():
    from oracle import c_oracle
    c_oracle.build()
    return c_oracle


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def unhex(v):
    return float.fromhex(v)


def noise_frame(seed, h, w):
    """SURVEY.md 8(d) 'noise' distribution: every LSB is visible."""
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def smooth_frame(seed, h, w):
    """Low-resolution noise upsampled x8 with separable linear interpolation in pure numpy
    (the survey's 'smooth' distribution used cv2.resize; tests must not depend on cv2)."""
    lo = np.random.default_rng(seed).integers(0, 256, (max(h // 8, 2), max(w // 8, 2), 3)).astype(np.float64)
    ys = np.linspace(0, lo.shape[0] - 1, h)
    xs = np.linspace(0, lo.shape[1] - 1, w)
    y0 = np.floor(ys).astype(int); y1 = np.minimum(y0 + 1, lo.shape[0] - 1); fy = (ys - y0)[:, None, None]
    x0 = np.floor(xs).astype(int); x1 = np.minimum(x0 + 1, lo.shape[1] - 1); fx = (xs - x0)[None, :, None]
    top = lo[y0][:, x0] * (1 - fx) + lo[y0][:, x1] * fx
    bot = lo[y1][:, x0] * (1 - fx) + lo[y1][:, x1] * fx
    return np.ascontiguousarray(np.clip(np.rint(top * (1 - fy) + bot * fy), 0, 255).astype(np.uint8))

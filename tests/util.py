"""Shared helpers for the tests (fixture loading, CRCs)."""
import os

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IMGS = os.path.join(ROOT, "tests", "golden", "imgs")
FIXTURES = ["0-tsukuba-327x245", "1-240x135", "2-480x270", "3-960x540", "4-1920x1080", "5-3840x2160"]
THRESHOLD = 0.15


def load_pair(name):
    a = np.ascontiguousarray(np.asarray(Image.open(os.path.join(IMGS, name, "a.png"))))
    b = np.ascontiguousarray(np.asarray(Image.open(os.path.join(IMGS, name, "b.png"))))
    assert a.dtype == np.uint8 and a.ndim == 2 and a.shape == b.shape
    return a, b


def vname(variant):
    return "ghost" if variant else "wrap"

"""sm_edges per call, back to back (the launch overhead hides behind the previous kernel): the edge detector's time on
the fixtures, both variants; with STEREO_B200_LIB=<other .so> the A/B tool for k_edges_planes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import stereomatching_b200 as smb
from util import load_pair
out = []
for name in ("4-1920x1080", "5-3840x2160", "2-480x270"):
    a, b = load_pair(name); h, w = a.shape
    for variant in (0, 1):
        with smb.StereoContext(w, h, 30, 21, variant) as c:
            c.upload_u8(a, b)
            for _ in range(10): c.edges(0.15)
            c.synchronize(); t0 = time.perf_counter()
            for _ in range(300): c.edges(0.15)
            c.synchronize(); us = (time.perf_counter() - t0) / 300 * 1e6
            e = c.download(smb.EDGES1)
        out.append("%s %s %.1f us (crc %08x)" % (name, "ghost" if variant else "wrap", us, __import__("zlib").crc32(e.tobytes())))
print(os.environ.get("STEREO_B200_LIB", "in-tree"), " | ".join(out), flush=True)

"""Where the driver's `elapsed` goes: every stage of algorithm() (host/driver.c) timed with a host clock and a
synchronise after it, on a fresh context (what timing/stereopar sees: the first call of everything) and warmed up."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import stereomatching_b200 as smb
from util import load_pair

def stages(c):
    out = []
    for name, fn in (("edges", lambda: c.edges(0.15)), ("match_wta", c.match_wta), ("fill_holes", lambda: c.fill_web_holes(32)),
                     ("contour", lambda: c.draw_contour_map(10))):
        c.synchronize(); t0 = time.perf_counter(); fn(); c.synchronize(); out.append((name, (time.perf_counter() - t0) * 1e6))
    return out

for name, D, sw in (("4-1920x1080", 30, 21), ("5-3840x2160", 30, 21), ("1-240x135", 30, 21)):
    a, b = load_pair(name); h, w = a.shape
    for variant in (0, 1):
        with smb.StereoContext(w, h, D, sw, variant) as c:
            c.upload_u8(a, b)
            first = stages(c)
            for _ in range(3): stages(c)
            warm = stages(c)
            t0 = time.perf_counter(); c.edges(0.15); c.match_wta(); c.fill_web_holes(32); c.draw_contour_map(10); c.synchronize()
            whole = (time.perf_counter() - t0) * 1e6
        print(name, "ghost" if variant else "wrap", "first:", " ".join("%s %.1f" % s for s in first), "sum %.1f" % sum(s[1] for s in first),
              "| warm:", " ".join("%s %.1f" % s for s in warm), "| warm whole, one sync: %.1f us" % whole, flush=True)

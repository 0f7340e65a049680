"""What a threshold change costs sm_edges (it rebuilds the decision and threshold tables: three small launches)."""
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, stereomatching_b200 as smb
from util import load_pair
a, b = load_pair("4-1920x1080"); h, w = a.shape
with smb.StereoContext(w, h, 30, 21, 0) as c:
    c.upload_u8(a, b); c.edges(0.15); c.synchronize()
    for thr in (0.2, 0.3, 0.15, 0.2):
        t0 = time.perf_counter(); c.edges(thr); c.synchronize(); t1 = time.perf_counter()
        c.edges(thr); c.synchronize(); t2 = time.perf_counter()
        print("threshold %.2f: first call %.1f us, second %.1f us, table exact %d" % (thr, (t1 - t0) * 1e6, (t2 - t1) * 1e6, c.get_info(smb.INFO_EDGE_THRESHOLDS)))

"""tools/tune_variance.py -- how much the launch shape timed at sm_create varies between contexts of one geometry:
back-to-back hot-path time per call over 8 fresh contexts (measured: 29.7-30.5 us at config 2)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stereomatching_b200 as smb, numpy as np, time
from bench import synth_pair
l,r,_=synth_pair(1234,1920,1080,64)
res=[]
for i in range(8):
    with smb.StereoContext(1920,1080,64,9,0) as c:
        c.upload_u8(l,r); c.edges(0.15)
        for _ in range(5): c.match_wta()
        c.synchronize(); t0=time.perf_counter()
        for _ in range(300): c.match_wta()
        c.synchronize(); res.append((time.perf_counter()-t0)/300*1e6)
print("b2b us per call over 8 contexts:", " ".join("%.2f"%x for x in res))

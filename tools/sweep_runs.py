"""One-pair hot-path time against the number of row runs per strip (SM_OPT_ROW_RUNS): calibration of the cost model."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stereomatching_b200 as smb
from bench import synth_pair
for (W,H,D,sw) in [(1920,1080,64,9),(1280,720,128,21),(1920,1080,30,21),(3840,2160,256,11),(640,360,64,9),(240,135,30,21)]:
    l,r,_=synth_pair(1234,W,H,D)
    res=[]
    for runs in [0,1,2,3,4,5,6,8,10,12,14,16,18,20,24,27,30,34,40,45,54,68]:
        with smb.StereoContext(W,H,D,sw,0) as c:
            c.set_option(smb.OPT_ROW_RUNS, runs)
            c.upload_u8(l,r); c.edges(0.15)
            for _ in range(5): c.match_wta()
            c.synchronize(); t0=time.perf_counter()
            for _ in range(200): c.match_wta()
            c.synchronize(); t=(time.perf_counter()-t0)/200*1e6
            res.append((runs,t))
    print(W,H,D,sw," ".join("%d:%.1f"%x for x in res), flush=True)

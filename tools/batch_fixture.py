import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import numpy as np, torch
import stereomatching_b200 as smb
from util import load_pair
for name,(D,sw) in (("4-1920x1080",(30,21)),("4-1920x1080",(30,9)),("4-1920x1080",(64,21))):
    a,b=load_pair(name); h,w=a.shape
    with smb.StereoContext(w,h,D,sw,0) as c:
        c.upload_u8(a,b); c.edges(0.15); e1,e2=c.download(smb.EDGES1),c.download(smb.EDGES2)
    batch=64
    d1=torch.from_numpy(e1).cuda().unsqueeze(0).repeat(batch,1,1).contiguous(); d2=torch.from_numpy(e2).cuda().unsqueeze(0).repeat(batch,1,1).contiguous()
    bb=torch.empty((batch,h,w),dtype=torch.int32,device="cuda"); ww=torch.empty_like(bb)
    with smb.StereoContext(w,h,D,sw,0) as c:
        st=torch.cuda.Stream(); c.set_stream(st.cuda_stream); torch.cuda.synchronize()
        run=lambda: c.match_wta_dev_batch(batch,d1.data_ptr(),d2.data_ptr(),h*w,bb.data_ptr(),ww.data_ptr(),h*w)
        for _ in range(3): run()
        ev0,ev1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); ev0.record(st)
        for _ in range(5): run()
        ev1.record(st); torch.cuda.synchronize()
        us=ev0.elapsed_time(ev1)*1e3/5/batch
        c.set_edges(e1,e2); c.match_wta(); wref=c.download(smb.WEB)
    print(name,D,sw,"batch us/pair %.2f"%us,"= %.2f T MDE/s"%(w*h*D/us/1e6), "equal", bool(np.array_equal(ww[batch-1].cpu().numpy(), wref)))

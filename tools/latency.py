import sys, os, time, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import stereomatching_b200 as smb
from bench import synth_pair
import zlib
def run(name, W,H,D,sw, variant=0):
    l,r,_=synth_pair(1234,W,H,D)
    with smb.StereoContext(W,H,D,sw,variant) as c:
        c.upload_u8(l,r); c.edges(0.15)
        for _ in range(5): c.match_wta(); c.elapsed_ms()
        ts=[]
        for _ in range(40):
            c.match_wta(); ts.append(c.elapsed_ms()*1e3)
        ts.sort()
        t0=time.perf_counter()
        for _ in range(300): c.match_wta()
        c.synchronize(); t1=time.perf_counter()
        web=c.download(smb.WEB)
        print("%s hot path single call: median %.2f us  min %.2f us; back-to-back %.2f us/call; web crc %08x"%(name, ts[len(ts)//2], ts[0], (t1-t0)/300*1e6, zlib.crc32(web.tobytes())))
run("c2 1920x1080 D64 sw9",1920,1080,64,9)
run("c1-like 1920x1080 D30 sw21",1920,1080,30,21)
run("c4 1280x720 D128 sw21",1280,720,128,21)
run("small 640x360 D64 sw9",640,360,64,9)
for (W,H,D,sw) in [(1920,1080,64,9),(3840,2160,256,11),(640,360,64,9)]:
    t0=time.perf_counter(); c=smb.StereoContext(W,H,D,sw,0); t1=time.perf_counter(); c.close()
    print("sm_create %dx%d D=%d sw=%d: %.1f ms"%(W,H,D,sw,(t1-t0)*1e3))

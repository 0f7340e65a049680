import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stereomatching_b200 as smb
for (w,h,D,sw) in [(1920,1080,64,9),(1280,720,128,21),(1920,1080,30,21),(3840,2160,256,11),(240,135,30,21)]:
    with smb.StereoContext(w,h,D,sw,0) as c:
        z=np.zeros((2,h,w),np.uint8)
        c.run_batch(z,z,0.15)
        print(w,h,D,sw,"warps/SM",c.get_info(smb.INFO_WARPS_PER_SM),"pairs/launch",c.get_info(smb.INFO_PAIRS_PER_LAUNCH),"tmem cols",c.get_info(smb.INFO_TMEM_COLUMNS))

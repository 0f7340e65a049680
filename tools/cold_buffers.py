"""tools/cold_buffers.py -- what cache-cold buffers cost a single-pair call: back-to-back sm_match_wta_dev calls on
config 2 with (a) one buffer set reused, (b) rotating inputs, (c) rotating outputs, (d) both rotating (64 sets)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import stereomatching_b200 as smb
from bench import synth_pair
W, H, D, SW, K = 1920, 1080, 64, 9, 64
l, r, _ = synth_pair(1234, W, H, D)
ctx = smb.StereoContext(W, H, D, SW, 0)
ctx.upload_u8(l, r); ctx.edges(0.15)
dev = torch.device("cuda")
e1 = torch.from_numpy(ctx.download(smb.EDGES1)).to(dev).unsqueeze(0).repeat(K, 1, 1).contiguous()
e2 = torch.from_numpy(ctx.download(smb.EDGES2)).to(dev).unsqueeze(0).repeat(K, 1, 1).contiguous()
best = torch.empty((K, H, W), dtype=torch.int32, device=dev); web = torch.empty_like(best)
st = torch.cuda.Stream(); ctx.set_stream(st.cuda_stream); torch.cuda.synchronize()
n8, n32 = H * W, H * W * 4
def run(rot_in, rot_out, n=256):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        ev0.record(st)
        for k in range(n):
            i = (k % K) if rot_in else 0; o = (k % K) if rot_out else 0
            ctx.match_wta_dev(e1.data_ptr() + i * n8, e2.data_ptr() + i * n8, best.data_ptr() + o * n32, web.data_ptr() + o * n32)
        ev1.record(st); torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) * 1e3 / n
for name, a, b in (("same buffers", 0, 0), ("rotating inputs", 1, 0), ("rotating outputs", 0, 1), ("both rotating", 1, 1)):
    print("%-18s %.2f us per call" % (name, run(a, b)))
# per-kernel split (events around each kernel: no dependent launch in this mode)
for name, rot in (("same inputs", 0), ("rotating inputs", 1)):
    ctx.profile_begin(128)
    for k in range(128):
        i = (k % K) if rot else 0
        ctx.match_wta_dev(e1.data_ptr() + i * n8, e2.data_ptr() + i * n8, best.data_ptr(), web.data_ptr())
    n, pack_ms, main_ms = ctx.profile_read()
    ctx.profile_begin(0)
    print("%-18s pack %.2f us, main %.2f us" % (name, pack_ms * 1e3 / n, main_ms * 1e3 / n))

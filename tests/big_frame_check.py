#!/usr/bin/env python
"""tests/big_frame_check.py -- one 7680x4320 pair (the largest frame the reference's survey mentions): the bit-sliced
kernel against the CPU oracle on two horizontal slabs and against a 3-band run; prints timings.  Uses the oracle as
the checker only."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
import stereomatching_b200 as smb
from bench import synth_pair

W, H, D, sw = 7680, 4320, 64, 9
half = sw // 2
all_ok = True
orc = oracle.Oracle()
left, right, disp = synth_pair(1234, W, H, D)
for variant in (smb.WRAP, smb.GHOST):
    with smb.StereoContext(W, H, D, sw, variant) as c:
        c.upload_u8(left, right); c.edges(0.15)
        for _ in range(3): c.match_wta()
        ms = c.elapsed_ms()
        web, best = c.download(smb.WEB), c.download(smb.BEST)
        e1, e2 = c.download(smb.EDGES1), c.download(smb.EDGES2)
    ok = True
    for y0 in (37, 4000):
        y1 = y0 + 40
        sl = slice(y0 - half, y1 + half)
        bo, wo = orc.match_wta(e1[sl], e2[sl], D, sw, smb.GHOST)
        xin = slice(None) if variant == smb.GHOST else slice(half + 1, W - D - half - 1)
        ok &= bool(np.array_equal(wo[half:-half, xin], web[y0:y1, xin]) and np.array_equal(bo[half:-half, xin], best[y0:y1, xin]))
    web_b = np.zeros_like(web)
    for band in range(3):
        with smb.StereoContext(W, H, D, sw, variant, rows=smb.band_rows(H, 3, band)) as c:
            c.upload_u8(left, right); c.edges(0.15); c.match_wta(); c.download(smb.WEB, out=web_b)
    bands_ok = bool(np.array_equal(web_b, web))
    all_ok = all_ok and ok and bands_ok
    print("%s 7680x4320 D=%d sw=%d: hot path %.3f ms = %.2f T MDE/s; slabs == oracle: %s; 3 bands == whole: %s"
          % ("ghost" if variant else "wrap", D, sw, ms, W * H * D / ms / 1e9, ok, bands_ok))
sys.exit(0 if all_ok else 1)

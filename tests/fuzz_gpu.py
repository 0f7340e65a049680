#!/usr/bin/env python
"""tests/fuzz_gpu.py [n_cases] [seed] -- differential fuzzing of the hot path on the GPU: random geometries (width,
height, shifts, window, variant, edge density, band), bit-sliced kernel against the direct (literal window sum)
kernel on the device and, for small frames, against the CPU oracle (checker only); every other case starts from
8-bit images (edge detector straight into the packed planes, random threshold, byte maps against the oracle's).
Exits non-zero on a mismatch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
import stereomatching_b200 as smb

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
orc = oracle.Oracle()
bad = 0
for case in range(n_cases):
    w = int(rng.choice([rng.integers(1, 70), rng.integers(60, 400), rng.integers(300, 2100)]))
    h = int(rng.choice([rng.integers(1, 40), rng.integers(30, 300), rng.integers(200, 1200)]))
    sw = int(rng.integers(1, min(31, w, h) + 1))
    D = int(rng.choice([rng.integers(1, 17), rng.integers(17, 33), rng.integers(33, 65), rng.integers(65, 200), rng.integers(200, 513)]))
    if w * h * D * sw * sw > 3e11:   # keep the direct kernel under a second
        continue
    variant = int(rng.integers(0, 2))
    dens = float(rng.choice([0.02, 0.2, 0.5, 0.8]))
    le = (rng.random((h, w)) < dens).astype(np.uint8)
    re = np.roll(le, int(rng.integers(0, max(1, min(D, w)))), axis=1) ^ (rng.random((h, w)) < 0.03).astype(np.uint8)
    rows = None
    if h >= 8 and rng.random() < 0.3:
        r0 = int(rng.integers(0, h - 1)); r1 = int(rng.integers(r0 + 1, h + 1)); rows = (r0, r1)
    # every other case starts from 8-bit IMAGES: the detector writes the packed planes itself (k_edges_planes), at a
    # random threshold; its byte maps must equal the oracle's and feed the comparison below
    from_images = bool(rng.random() < 0.5)
    if from_images:
        base = rng.integers(0, 256, (h, w)).astype(np.float64)
        smooth = (base + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, (1, 1), (0, 1))) / 4
        img1 = np.where(rng.random((h, w)) < dens, base, smooth).astype(np.uint8)
        img2 = np.roll(img1, int(rng.integers(0, max(1, min(D, w)))), axis=1)
        img2 = np.where(rng.random((h, w)) < 0.03, rng.integers(0, 256, (h, w)), img2).astype(np.uint8)
        thr = float(rng.choice([0.0, 0.05, 0.15, 0.3, 1.0]))
        le, re = orc.edges(img1, thr, variant), orc.edges(img2, thr, variant)
    res = {}
    edges_ok = True
    for kernel in (smb.KERNEL_BITSLICE, smb.KERNEL_DIRECT):
        with smb.StereoContext(w, h, D, sw, variant, rows=rows, kernel=kernel) as c:
            if from_images:
                c.upload_u8(img1, img2); c.edges(thr); c.match_wta()
                if kernel == smb.KERNEL_BITSLICE:
                    sle = slice(*rows) if rows else slice(None)
                    edges_ok = np.array_equal(c.download(smb.EDGES1)[sle], le[sle]) and \
                               np.array_equal(c.download(smb.EDGES2)[sle], re[sle])
            else:
                c.set_edges(le, re); c.match_wta()
            res[kernel] = (c.download(smb.BEST), c.download(smb.WEB))
    sl = slice(*rows) if rows else slice(None)
    ok = edges_ok and np.array_equal(res[smb.KERNEL_BITSLICE][0][sl], res[smb.KERNEL_DIRECT][0][sl]) and \
         np.array_equal(res[smb.KERNEL_BITSLICE][1][sl], res[smb.KERNEL_DIRECT][1][sl])
    if ok and w * h * D < 4e6:
        bo, wo = orc.match_wta(le, re, D, sw, variant)
        ok = np.array_equal(res[smb.KERNEL_BITSLICE][0][sl], bo[sl]) and np.array_equal(res[smb.KERNEL_BITSLICE][1][sl], wo[sl])
    if not ok:
        bad += 1
        print("MISMATCH w=%d h=%d D=%d sw=%d variant=%d dens=%.2f rows=%s from_images=%s edges_ok=%s"
              % (w, h, D, sw, variant, dens, rows, from_images, edges_ok), flush=True)
print("fuzz: %d cases, %d mismatches" % (n_cases, bad))
sys.exit(1 if bad else 0)

#!/usr/bin/env python
"""Development experiment (needs the DEV=1 build of the library): per-pair time of the batched hot path for
different kernel shapes (words per lane, walker segment) selected through the SMB_* hooks, each checked bit for bit
against the default shape's result on the same inputs.

usage: python tools/exp_shapes.py [point ...]      points: c4 ref30 w15 w17 c2 w21d64
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import stereomatching_b200 as smb
from bench import synth_pair

POINTS = {
    "c4": (1280, 720, 128, 21),
    "ref30": (1920, 1080, 30, 21),
    "w15": (1920, 1080, 64, 15),
    "w17": (1920, 1080, 64, 17),
    "w21d64": (1920, 1080, 64, 21),
    "c2": (1920, 1080, 64, 9),
    "w17d32": (1920, 1080, 32, 17),
    "c3": (3840, 2160, 256, 11),
    "w3": (1920, 1080, 64, 3),
    "w11": (1920, 1080, 64, 11),
    "w13": (1920, 1080, 64, 13),
    "w5": (1920, 1080, 64, 5),
    "w1": (1920, 1080, 30, 1),
    "c2d32": (1920, 1080, 32, 9),
    "c2d16": (1920, 1080, 16, 9),
    "d16w21": (1920, 1080, 16, 21),
    "small": (240, 135, 30, 21),
}
SHAPES = [{}]
EXTRA = [{"SMB_TR": "8"}, {"SMB_TR": "16"}, {"SMB_TR": "24"}, {"SMB_TR": "48"}, {"SMB_TR": "64"}, {"SMB_TR": "128"}]


def run_point(name, batch=48):
    w, h, D, sw = POINTS[name]
    left, right, _ = synth_pair(1234, w, h, D)
    with smb.StereoContext(w, h, D, sw, 0) as c:
        c.upload_u8(left, right)
        c.edges(0.15)
        e1, e2 = c.download(smb.EDGES1), c.download(smb.EDGES2)
    d1 = torch.from_numpy(e1).cuda().unsqueeze(0).repeat(batch, 1, 1).contiguous()
    d2 = torch.from_numpy(e2).cuda().unsqueeze(0).repeat(batch, 1, 1).contiguous()
    bb = torch.empty((batch, h, w), dtype=torch.int32, device="cuda")
    ww = torch.empty_like(bb)
    ref = None
    shapes = list(SHAPES)
    results = []
    for env in shapes + [None]:
        if env is None:  # run-length variations on the fastest shape so far
            best = min(results, key=lambda r: r[1])[0]
            todo = [dict(best, **e) for e in EXTRA]
        else:
            todo = [env]
        for e in todo:
            if D <= 32 and e.get("SMB_NW") == "2":
                continue
            for k in ("SMB_NW", "SMB_SEG", "SMB_TR", "SMB_TM"):
                os.environ.pop(k, None)
            os.environ.update(e)
            os.environ["SMB_NO_TUNE"] = "1"
            try:
                with smb.StereoContext(w, h, D, sw, 0) as c:
                    st = torch.cuda.Stream()
                    c.set_stream(st.cuda_stream)
                    torch.cuda.synchronize()
                    run = lambda: c.match_wta_dev_batch(batch, d1.data_ptr(), d2.data_ptr(), h * w, bb.data_ptr(),
                                                        ww.data_ptr(), h * w)
                    bb.zero_()
                    ww.zero_()
                    for _ in range(3):
                        run()
                    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    ev0.record(st)
                    for _ in range(5):
                        run()
                    ev1.record(st)
                    torch.cuda.synchronize()
                    us = ev0.elapsed_time(ev1) * 1e3 / 5 / batch
                    # one pair per launch, default one-wave shape
                    c.profile_begin(8)
                    for _ in range(8):
                        c.match_wta_dev(d1.data_ptr(), d2.data_ptr(), bb.data_ptr(), ww.data_ptr())
                    calls, pack_ms, main_ms = c.profile_read()
                    c.profile_begin(0)
                    web = ww[batch - 1].cpu().numpy()
                    best_ = bb[batch - 1].cpu().numpy()
                    web0 = ww[0].cpu().numpy()
                if ref is None:
                    ref = (web, best_)
                ok = bool(np.array_equal(web, ref[0]) and np.array_equal(best_, ref[1]) and np.array_equal(web0, ref[0]))
                results.append((e, us))
                print("%-7s %dx%d D=%d sw=%d %-36s batch %.2f us/pair = %.2f T MDE/s | one pair: main %.1f us | equal %s"
                      % (name, w, h, D, sw, e or "default", us, w * h * D / us / 1e6, main_ms * 1e3 / calls, ok), flush=True)
            except Exception as ex:  # a shape that is not instantiated
                print("%-7s %s: %s" % (name, e, ex), flush=True)
    for k in ("SMB_NW", "SMB_SEG", "SMB_TR", "SMB_TM"):
        os.environ.pop(k, None)
    with smb.StereoContext(w, h, D, sw, 0, kernel=smb.KERNEL_DIRECT) as c:  # the literal window sums
        c.set_edges(e1, e2)
        c.match_wta()
        ok = bool(np.array_equal(c.download(smb.WEB), ref[0]) and np.array_equal(c.download(smb.BEST), ref[1]))
    print("%-7s default shape == direct kernel: %s" % (name, ok), flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if "--default-only" in sys.argv:  # one shape, small batch: the command line for ncu
        os.environ["SMB_NO_TUNE"] = "1"
        SHAPES[:] = [{k: v for k, v in (kv.split("=") for kv in os.environ.get("EXP_SHAPE", "").split(",") if kv)}]
        EXTRA[:] = []
    if "--no-extra" in sys.argv:
        EXTRA[:] = []
    for p in (args or ["c4", "ref30", "w15", "w17", "w21d64", "c2", "c3", "w3", "w1", "c2d32"]):
        run_point(p, batch=16 if "--default-only" in sys.argv else (12 if p == "c3" else 48))

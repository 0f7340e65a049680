#!/usr/bin/env python
"""tests/ladder.py (under tests/: it runs the reference's own CUDA programs from the oracle directory as the baseline) -- the reference's published size ladder (report/data.txt:1-4: six sizes x four programs,
measured by test/time.sh:1-15) on this box: whole `algorithm()` elapsed of

  this repo's  timing/stereopar, timing/stereopar-ghost                (C drivers over the C ABI)
  reference    oracle/_ref/stereopar_D30, oracle/_ref/stereopar-ghost_D30  (unmodified stereo.cu / stereo-ghost.cu,
               rebuilt for sm_100a by oracle/Makefile)

on the five fixtures of test/imgs plus a synthetic 7680x4320 pair (the ladder's sixth size, which the reference
does not ship), with time.sh's parsing (field 15 of the stdout line) and the reference defaults (threshold 0.15,
window 21, 30 shifts).  Writes a markdown table.  Run under gpurun: python tests/ladder.py > gpurun_out/ladder.md
"""
import os
import struct
import subprocess
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import synth_pair  # noqa: E402

PUBLISHED = {  # report/data.txt:1-4, seconds, sizes 240x135 ... 7680x4320
    "serial": [2.334591, 9.280708, 36.996416, 148.124367, 595.996112, 2393.26121],
    "serial_ghost": [0.153506, 0.714631, 2.836038, 11.492294, 84.260887, 336.66939],
    "parallel": [0.007820, 0.021544, 0.081994, 0.316084, 1.217091, 4.714461],
    "parallel_ghost": [0.006076, 0.015374, 0.055790, 0.232813, 0.878147, 3.270732],
}


def write_png_gray(path, img):
    h, w = img.shape
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 1)) + chunk(b"IEND", b""))


def elapsed(prog, a, b, reps):
    """time.sh:8: field 15 of the program's stdout line; best and mean of `reps` processes"""
    ts = []
    for _ in range(reps):
        out = subprocess.run([prog, a, b], capture_output=True, text=True, timeout=600)
        if out.returncode != 0:
            return None, None
        ts.append(float(out.stdout.split()[14]))
    return min(ts), sum(ts) / len(ts)


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    imgs = os.path.join(ROOT, "tests", "golden", "imgs")
    sizes = ["1-240x135", "2-480x270", "3-960x540", "4-1920x1080", "5-3840x2160"]
    pairs = [(s, os.path.join(imgs, s, "a.png"), os.path.join(imgs, s, "b.png")) for s in sizes]
    tmp = os.path.join(ROOT, "gpurun_out", "tmp_ladder")
    os.makedirs(tmp, exist_ok=True)
    left, right, _ = synth_pair(1234, 7680, 4320, 30)
    write_png_gray(os.path.join(tmp, "a.png"), left)
    write_png_gray(os.path.join(tmp, "b.png"), right)
    pairs.append(("6-7680x4320 (synthetic)", os.path.join(tmp, "a.png"), os.path.join(tmp, "b.png")))
    progs = [("parallel", "timing/stereopar", "oracle/_ref/stereopar_D30"),
             ("parallel_ghost", "timing/stereopar-ghost", "oracle/_ref/stereopar-ghost_D30")]
    print("| size | program | published (reference's GPU), s | reference CUDA on this B200, s (best / mean of %d) | "
          "this repo, s (best / mean) | ratio (best) |" % reps)
    print("|---|---|---|---|---|---|")
    for k, (name, a, b) in enumerate(pairs):
        for col, ours, ref in progs:
            rb, rm = elapsed(os.path.join(ROOT, ref), a, b, reps) if os.path.exists(os.path.join(ROOT, ref)) else (None, None)
            ob, om = elapsed(os.path.join(ROOT, ours), a, b, reps)
            f = lambda v: "-" if v is None else "%.6f" % v  # noqa: E731
            print("| %s | %s | %.6f | %s / %s | %s / %s | %s |"
                  % (name, col, PUBLISHED[col][k], f(rb), f(rm), f(ob), f(om),
                     "-" if not (rb and ob) else "%.0fx" % (rb / ob)), flush=True)
    for n in ("a.png", "b.png"):
        os.remove(os.path.join(tmp, n))
    os.rmdir(tmp)


if __name__ == "__main__":
    main()

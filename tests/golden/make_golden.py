#!/usr/bin/env python
"""Generate tests/golden/golden.json from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference and oracle/_ref built by
`make -C oracle ref`).  For every case it calls the reference's own stage functions
(find_all_edges, fillup_matches, fillup_scores, find_highest_scoring_shifts;
src/stereo.c:72,113,184,196 and the stereo-ghost.c twins) through oracle.RefLib and
records zlib CRC32s of the raw arrays (edges as W*H u8, best/web as little-endian i32),
the convention of SURVEY.md 8(c).  The committed JSON is what travels to the GPU box.

usage: python tests/golden/make_golden.py [--quick] [--fixtures-only NAME]
       (--quick skips 1080p/4K wrap: ~6 min; --fixtures-only regenerates one fixture's entries)

Fixture 0 is the Tsukuba pair derived from the thesis figure (imgs/0-tsukuba-327x245/PROVENANCE.md).
"""
import json
import os
import sys
import time

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = ["0-tsukuba-327x245", "1-240x135", "2-480x270", "3-960x540", "4-1920x1080", "5-3840x2160"]
THRESHOLD = 0.15


def load(name):
    a = np.asarray(Image.open(os.path.join(HERE, "imgs", name, "a.png")))
    b = np.asarray(Image.open(os.path.join(HERE, "imgs", name, "b.png")))
    assert a.dtype == np.uint8 and a.ndim == 2
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


def run_case(key, left, right, D, sw, variant, out):
    t0 = time.time()
    ref = oracle.RefLib(variant, D)
    e1 = ref.edges(left, THRESHOLD)
    e2 = ref.edges(right, THRESHOLD)
    best, web = ref.match_wta(e1, e2, sw)
    out[key] = {
        "D": D, "sw": sw, "variant": "ghost" if variant else "wrap", "threshold": THRESHOLD,
        "w": int(left.shape[1]), "h": int(left.shape[0]),
        "left": oracle.crc32(left), "right": oracle.crc32(right),
        "edges1": oracle.crc32(e1), "edges2": oracle.crc32(e2),
        "best": oracle.crc32(best), "web": oracle.crc32(web),
        "web_eq_D_frac": float((web == D).mean()),
    }
    print("%-40s %s  (%.1fs)" % (key, out[key]["web"], time.time() - t0), flush=True)


def main():
    quick = "--quick" in sys.argv
    path = os.path.join(HERE, "golden.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    orc = oracle.Oracle()
    # real fixtures, reference defaults (stereo.c:6-10): D=30, sw=21
    only = sys.argv[sys.argv.index("--fixtures-only") + 1] if "--fixtures-only" in sys.argv else None
    for name in FIXTURES:
        if only and name != only:
            continue
        left, right = load(name)
        for variant in (oracle.GHOST, oracle.WRAP):
            big = left.size > 960 * 540
            if quick and big and variant == oracle.WRAP:
                continue
            run_case("fixture/%s/%s" % (name, "ghost" if variant else "wrap"), left, right, 30, 21,
                     variant, out)
            json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    if only:
        return
    # synthetic config 2: 1920x1080, D=64, sw=9, seed 1234
    left, right, disp = orc.synth_pair(1234, 1920, 1080, 64)
    out["synth/c2/generator"] = {"left": oracle.crc32(left), "right": oracle.crc32(right),
                                 "disp": oracle.crc32(disp)}
    for variant in (oracle.GHOST, oracle.WRAP):
        if quick and variant == oracle.WRAP:
            continue
        run_case("synth/c2/%s" % ("ghost" if variant else "wrap"), left, right, 64, 9, variant, out)
        json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    # parameter sweep on small inputs: crops of fixture 1 and small synthetic pairs
    a, b = load("1-240x135")
    for D in (16, 30, 32, 64, 128):
        for sw in (3, 4, 5, 9, 11, 15, 21):
            for variant in (oracle.WRAP, oracle.GHOST):
                run_case("sweep/fix1/D%d/sw%d/%s" % (D, sw, "ghost" if variant else "wrap"),
                         a, b, D, sw, variant, out)
    for (w, h, D, sw) in [(257, 67, 32, 9), (640, 96, 256, 11), (1031, 45, 512, 21), (96, 40, 16, 7),
                          (33, 33, 30, 21), (21, 21, 16, 21), (320, 180, 128, 21)]:
        left, right, _ = orc.synth_pair(77, w, h, D)
        for variant in (oracle.WRAP, oracle.GHOST):
            run_case("sweep/synth%dx%d/D%d/sw%d/%s" % (w, h, D, sw, "ghost" if variant else "wrap"),
                     left, right, D, sw, variant, out)
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

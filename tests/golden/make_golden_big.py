#!/usr/bin/env python
"""Full-size golden CRCs from the UNMODIFIED reference (the cases VERDICT r01 asked to pin).

Like make_golden.py, but for the BASELINE configs that take minutes of CPU each, so every case
runs in its own process (the reference keeps matches[]/scores[] in file-scope globals,
src/stereo.c:90,150) and results are merged into golden.json as they arrive:

  synth/c3/{wrap,ghost}             3840x2160, D=256, sw=11, seed 1234          (BASELINE configs[2])
  synth/c4/<k>/{wrap,ghost} k=0..3  1280x720,  D=128, sw=21, seed 1234+2k       (BASELINE configs[3])
  sweep1080/D<d>/sw<s>/{wrap,ghost} 1920x1080, seed 1234, the corners and a 12-point subset
                                    of the window/shift sweep                  (BASELINE configs[4])
  synth/c2s/<seed>/wrap             1920x1080, D=64, sw=9: every pair bench.py times
                                    (seed 1234+2j, j = 0..127: 16 distinct pairs x 8 ranks)

The reference functions called are find_all_edges, fillup_matches, fillup_scores and
find_highest_scoring_shifts (src/stereo.c:72,113,184,196; src/stereo-ghost.c twins), through
oracle.RefLib.  Besides the CRC32s each entry stores per-row-band CRCs of `web` (16 bands), so a
GPU test that fails can say where.

usage: python tests/golden/make_golden_big.py [--jobs N] [--only PREFIX] [--skip-existing]
"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
THRESHOLD = 0.15
# (D, sw) points of the 1080p sweep that get a whole-frame reference golden
SWEEP1080 = [(16, 21), (32, 13), (128, 17), (512, 3), (512, 21),          # the corners VERDICT names
             (16, 3), (16, 9), (32, 5), (32, 21), (64, 15), (64, 17), (64, 21), (128, 7), (256, 11),
             (256, 19), (512, 9)]


def cases():
    out = []
    out.append(("synth/c3/wrap", 1234, 3840, 2160, 256, 11, oracle.WRAP))
    for D, sw in SWEEP1080[:5]:
        out.append(("sweep1080/D%d/sw%d/wrap" % (D, sw), 1234, 1920, 1080, D, sw, oracle.WRAP))
    out.append(("synth/c3/ghost", 1234, 3840, 2160, 256, 11, oracle.GHOST))
    for k in range(4):
        for v in (oracle.WRAP, oracle.GHOST):
            out.append(("synth/c4/%d/%s" % (k, "ghost" if v else "wrap"), 1234 + 2 * k, 1280, 720, 128, 21, v))
    for D, sw in SWEEP1080:
        out.append(("sweep1080/D%d/sw%d/ghost" % (D, sw), 1234, 1920, 1080, D, sw, oracle.GHOST))
    for D, sw in SWEEP1080[5:]:
        out.append(("sweep1080/D%d/sw%d/wrap" % (D, sw), 1234, 1920, 1080, D, sw, oracle.WRAP))
    for j in range(128):
        out.append(("synth/c2s/%d/wrap" % (1234 + 2 * j), 1234 + 2 * j, 1920, 1080, 64, 9, oracle.WRAP))
    return out


def run(case):
    key, seed, w, h, D, sw, variant = case
    t0 = time.time()
    orc = oracle.Oracle()
    left, right, disp = orc.synth_pair(seed, w, h, D)
    ref = oracle.RefLib(variant, D)
    e1 = ref.edges(left, THRESHOLD)
    e2 = ref.edges(right, THRESHOLD)
    best, web = ref.match_wta(e1, e2, sw)
    nb = 16
    bands = [oracle.crc32(web[h * b // nb:h * (b + 1) // nb]) for b in range(nb)]
    return key, {
        "D": D, "sw": sw, "variant": "ghost" if variant else "wrap", "threshold": THRESHOLD,
        "w": w, "h": h, "seed": seed,
        "left": oracle.crc32(left), "right": oracle.crc32(right), "disp": oracle.crc32(disp),
        "edges1": oracle.crc32(e1), "edges2": oracle.crc32(e2),
        "best": oracle.crc32(best), "web": oracle.crc32(web), "web_bands16": bands,
        "web_eq_D_frac": float((web == D).mean()), "ref_seconds": round(time.time() - t0, 1),
    }


def main():
    jobs = int(sys.argv[sys.argv.index("--jobs") + 1]) if "--jobs" in sys.argv else 6
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else ""
    path = os.path.join(HERE, "golden.json")
    out = json.load(open(path))
    todo = [c for c in cases() if c[0].startswith(only)]
    if "--skip-existing" in sys.argv:
        todo = [c for c in todo if c[0] not in out]
    print("%d cases, %d processes" % (len(todo), jobs), flush=True)
    with mp.get_context("fork").Pool(jobs, maxtasksperchild=1) as pool:
        for key, val in pool.imap_unordered(run, todo, chunksize=1):
            out = json.load(open(path))  # somebody else may have written meanwhile
            out[key] = val
            json.dump(out, open(path + ".tmp", "w"), indent=1, sort_keys=True)
            os.replace(path + ".tmp", path)
            print("%-36s web %s  (%.0fs)" % (key, val["web"], val["ref_seconds"]), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""tests/sweep_configs.py -- parity + timing sweep over the other BASELINE.json configs on one B200: kernel time, MDE/s and
roofline fraction of the hot path for

  c1     the five real fixtures (reference defaults: 30 shifts, window 21), both variants
  c3     synthetic 3840x2160, 256 shifts, window 11
  c4     synthetic 1280x720, 128 shifts, window 21 (one pair of the 4096-pair batch)
  sweep  config 5: 1920x1080, window 3..21 x shifts 16..512

For every point the bit-sliced kernel's result is compared bit for bit with the direct
(literal window sum) kernel on the device, and with the CPU oracle on a horizontal slab.
Writes one JSON line per point to stdout; --md also writes a markdown table.
Lives under tests/ because it uses the CPU oracle (as the checker only); run it by hand:
    python tests/sweep_configs.py --what c1,c2,c3,c4,sweep --md out.md
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from bench import THRESHOLD, synth_pair  # noqa: E402


def measure(smb, orc, name, e1, e2, D, sw, variant, peak_gops, reps=12, check_rows=48, batch=0):
    h, w = e1.shape
    res = {}
    for kname, kernel in (("bitslice", smb.KERNEL_BITSLICE), ("direct", smb.KERNEL_DIRECT)):
        if kernel == smb.KERNEL_DIRECT and w * h * D * sw > 1920 * 1080 * 512 * 21:
            continue
        with smb.StereoContext(w, h, D, sw, variant, kernel=kernel) as c:
            c.set_edges(e1, e2)
            n = reps if kernel == smb.KERNEL_BITSLICE else 2
            for _ in range(3 if kernel == smb.KERNEL_BITSLICE else 1):
                c.match_wta()
            c.profile_begin(n)
            for _ in range(n):
                c.match_wta()
            calls, pack_ms, main_ms = c.profile_read()
            c.profile_begin(0)
            res[kname] = {"main_us": main_ms * 1e3 / calls, "pack_us": pack_ms * 1e3 / calls,
                          "best": c.download(smb.BEST), "web": c.download(smb.WEB)}
    b = res["bitslice"]
    ok_direct = None
    if "direct" in res:
        ok_direct = bool(np.array_equal(b["best"], res["direct"]["best"]) and np.array_equal(b["web"], res["direct"]["web"]))
    # CPU oracle on a slab in the middle of the frame (GHOST slab: interior rows are exact for both variants
    # vertically; horizontally the variant's own rule applies)
    half = sw // 2
    y0 = max(half, h // 2 - check_rows // 2)
    y1 = min(h - half, y0 + check_rows)
    sl = slice(y0 - half, y1 + half)
    bo, wo = orc.match_wta(e1[sl], e2[sl], D, sw, variant if variant == 1 else 0)
    if variant == 0:
        # a WRAP slab wraps vertically onto itself, which only touches its halo rows
        pass
    ok_oracle = bool(np.array_equal(bo[half:half + y1 - y0], b["best"][y0:y1]) and
                     np.array_equal(wo[half:half + y1 - y0], b["web"][y0:y1]))
    mde = w * h * D
    t = b["main_us"] * 1e-6
    out = {"config": name, "w": w, "h": h, "D": D, "sw": sw, "variant": "ghost" if variant else "wrap",
           "main_kernel_us": round(b["main_us"], 2), "pack_kernel_us": round(b["pack_us"], 2),
           "GMDE_per_s_main_kernel": round(mde / t / 1e9, 1),
           "GMDE_per_s_hot_path": round(mde / ((b["main_us"] + b["pack_us"]) * 1e-6) / 1e9, 1),
           "roofline_frac_int_alu": round(8 * mde / t / (peak_gops * 1e9), 3),
           "direct_kernel_us": round(res["direct"]["main_us"], 1) if "direct" in res else None,
           "equal_direct_kernel": ok_direct, "equal_oracle_slab": ok_oracle}
    if batch:
        # device-resident batch of `batch` copies of the pair through sm_match_wta_dev_batch
        import torch
        d1 = torch.from_numpy(e1).cuda().unsqueeze(0).repeat(batch, 1, 1).contiguous()
        d2 = torch.from_numpy(e2).cuda().unsqueeze(0).repeat(batch, 1, 1).contiguous()
        bb = torch.empty((batch, h, w), dtype=torch.int32, device="cuda")
        ww = torch.empty((batch, h, w), dtype=torch.int32, device="cuda")
        with smb.StereoContext(w, h, D, sw, variant) as c:
            st = torch.cuda.current_stream()
            c.set_stream(st.cuda_stream)
            run = lambda: c.match_wta_dev_batch(batch, d1.data_ptr(), d2.data_ptr(), h * w, bb.data_ptr(),
                                                ww.data_ptr(), h * w)
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
        out["batch_pairs"] = batch
        out["batch_us_per_pair"] = round(us, 2)
        out["batch_GMDE_per_s"] = round(mde / (us * 1e-6) / 1e9, 1)
        out["batch_equal"] = bool(np.array_equal(ww[batch - 1].cpu().numpy(), b["web"]) and
                                  np.array_equal(bb[0].cpu().numpy(), b["best"]))
    print(json.dumps(out), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="c1,c3,c4,sweep")
    ap.add_argument("--md", default=None)
    a = ap.parse_args()
    import oracle
    import benchlib
    import stereomatching_b200 as smb
    from util import FIXTURES, load_pair
    orc = oracle.Oracle()
    peak = max(benchlib.measure_int_peak(0, m) for m in range(4))
    rows = []
    what = a.what.split(",")

    def edges_of(left, right, D, sw, variant):
        h, w = left.shape
        with smb.StereoContext(w, h, D, sw, variant) as c:
            c.upload_u8(left, right)
            c.edges(THRESHOLD)
            return c.download(smb.EDGES1), c.download(smb.EDGES2)

    if "c1" in what:
        for name in FIXTURES:
            left, right = load_pair(name)
            for variant in (0, 1):
                e1, e2 = edges_of(left, right, 30, 21, variant)
                rows.append(measure(smb, orc, "c1/" + name, e1, e2, 30, 21, variant, peak))
    if "c3" in what:
        left, right, _ = synth_pair(1234, 3840, 2160, 256)
        for variant in (0, 1):
            e1, e2 = edges_of(left, right, 256, 11, variant)
            rows.append(measure(smb, orc, "c3", e1, e2, 256, 11, variant, peak, reps=6, check_rows=24))
    if "c4" in what:
        left, right, _ = synth_pair(1234, 1280, 720, 128)
        for variant in (0, 1):
            e1, e2 = edges_of(left, right, 128, 21, variant)
            rows.append(measure(smb, orc, "c4", e1, e2, 128, 21, variant, peak, batch=256))
    if "c2" in what:
        left, right, _ = synth_pair(1234, 1920, 1080, 64)
        for variant in (0, 1):
            e1, e2 = edges_of(left, right, 64, 9, variant)
            rows.append(measure(smb, orc, "c2", e1, e2, 64, 9, variant, peak, batch=64))
    if "sweep" in what:
        for D in (16, 32, 64, 128, 256, 512):
            left, right, _ = synth_pair(1234, 1920, 1080, D)
            e1, e2 = edges_of(left, right, D, 9, 0)  # edges do not depend on sw / D
            for sw in (3, 5, 7, 9, 11, 13, 15, 17, 19, 21):
                rows.append(measure(smb, orc, "c5", e1, e2, D, sw, 0, peak, reps=6, check_rows=16))
    if a.md:
        with open(a.md, "w") as f:
            f.write("| config | WxH | D | sw | variant | main kernel us | pack us | GMDE/s (main) | GMDE/s (hot path) | "
                    "frac of int-ALU roofline | direct kernel us | == direct | == oracle slab |\n")
            f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write("| %s | %dx%d | %d | %d | %s | %.1f | %.1f | %.0f | %.0f | %.2f | %s | %s | %s |\n" % (
                    r["config"], r["w"], r["h"], r["D"], r["sw"], r["variant"], r["main_kernel_us"],
                    r["pack_kernel_us"], r["GMDE_per_s_main_kernel"], r["GMDE_per_s_hot_path"],
                    r["roofline_frac_int_alu"], r["direct_kernel_us"], r["equal_direct_kernel"], r["equal_oracle_slab"]))
            f.write("\nINT32 peak used: %.0f G thread-instr/s (measured live, max over instruction mixes).\n" % peak)


if __name__ == "__main__":
    main()

"""Whole-frame parity at the BASELINE configs' full sizes against CRCs produced by the UNMODIFIED reference
(tests/golden/make_golden_big.py -> golden.json: synth/c3, synth/c4/<k>, sweep1080/*, synth/c2s/<seed>).

Every case goes through the C ABI (upload, device edge detector, hot path) and compares the CRC32 of the whole
`best` and `web` frames (and both edge maps) with what src/stereo.c:72-220 / src/stereo-ghost.c:74-218 produced on
the same seeded inputs.  Bit-exact; both variants; both kernels where the literal kernel finishes in seconds.
"""
import numpy as np
import pytest

import oracle
import stereomatching_b200 as smb
from util import THRESHOLD, vname

pytestmark = pytest.mark.gpu

KNAME = {smb.KERNEL_DIRECT: "direct", smb.KERNEL_BITSLICE: "bitslice"}


def _where(web, g):
    """On a mismatch: which of the 16 row bands differ (golden.json keeps per-band CRCs of web)."""
    h = web.shape[0]
    nb = len(g["web_bands16"])
    return [b for b in range(nb) if oracle.crc32(web[h * b // nb:h * (b + 1) // nb]) != g["web_bands16"][b]]


def _check(orc, g, kernel, rows=None):
    v = smb.GHOST if g["variant"] == "ghost" else smb.WRAP
    left, right, disp = orc.synth_pair(g["seed"], g["w"], g["h"], g["D"])
    assert (oracle.crc32(left), oracle.crc32(right), oracle.crc32(disp)) == (g["left"], g["right"], g["disp"])
    with smb.StereoContext(g["w"], g["h"], g["D"], g["sw"], v, kernel=kernel) as c:
        c.upload_u8(left, right)
        c.edges(THRESHOLD)
        e1, e2 = c.download(smb.EDGES1), c.download(smb.EDGES2)
        assert (oracle.crc32(e1), oracle.crc32(e2)) == (g["edges1"], g["edges2"])
        c.match_wta()
        best, web = c.download(smb.BEST), c.download(smb.WEB)
    assert oracle.crc32(web) == g["web"], "web differs from the reference in row bands %r of 16" % _where(web, g)
    assert oracle.crc32(best) == g["best"]
    return best, web, disp


# ---- config 3: 3840x2160, 256 shifts, window 11 --------------------------------------------------
@pytest.mark.parametrize("kernel", [smb.KERNEL_BITSLICE, smb.KERNEL_DIRECT], ids=KNAME.get)
@pytest.mark.parametrize("variant", [smb.WRAP, smb.GHOST], ids=vname)
def test_config3_whole_frame(orc, golden, variant, kernel):
    g = golden["synth/c3/%s" % vname(variant)]
    best, web, disp = _check(orc, g, kernel)
    # the independent known answer: web == d + 1 inside the disparity tiles
    D, half = g["D"], g["sw"] // 2
    tw, th = max(240, 4 * D), 120
    ys, xs = np.mgrid[0:g["h"], 0:g["w"]]
    interior = ((xs % tw >= half) & (xs % tw < tw - D - half) & (ys % th >= half + 1) & (ys % th < th - half - 1) &
                (xs >= half) & (xs < g["w"] - D - half) & (ys >= half + 1) & (ys < g["h"] - half - 1))
    assert (web[interior] == disp[interior] + 1).mean() >= 0.9999


@pytest.mark.parametrize("variant", [smb.WRAP, smb.GHOST], ids=vname)
def test_config3_bands_equal_reference(orc, golden, variant):
    """Config 3 as north_star shards it: row bands with replicated halo rows (here 8 bands, run one after
    another on one GPU); the reassembled frame equals the reference's whole-frame CRCs."""
    g = golden["synth/c3/%s" % vname(variant)]
    left, right, _ = orc.synth_pair(g["seed"], g["w"], g["h"], g["D"])
    web = np.zeros((g["h"], g["w"]), np.int32)
    best = np.zeros_like(web)
    for b in range(8):
        r0, r1 = smb.band_rows(g["h"], 8, b)
        with smb.StereoContext(g["w"], g["h"], g["D"], g["sw"], variant, rows=(r0, r1)) as c:
            c.upload_u8(left, right)
            c.edges(THRESHOLD)
            c.match_wta()
            c.download(smb.WEB, out=web)
            c.download(smb.BEST, out=best)
    assert oracle.crc32(web) == g["web"], "bands differ from the reference in row bands %r of 16" % _where(web, g)
    assert oracle.crc32(best) == g["best"]


# ---- config 4: 1280x720 pairs, 128 shifts, window 21 (seeds 1234 + 2k) -----------------------------
@pytest.mark.parametrize("kernel", [smb.KERNEL_BITSLICE, smb.KERNEL_DIRECT], ids=KNAME.get)
@pytest.mark.parametrize("variant", [smb.WRAP, smb.GHOST], ids=vname)
def test_config4_pairs(orc, golden, variant, kernel):
    for k in range(4):
        _check(orc, golden["synth/c4/%d/%s" % (k, vname(variant))], kernel)


@pytest.mark.parametrize("variant", [smb.WRAP, smb.GHOST], ids=vname)
def test_config4_batch_pipeline(orc, golden, variant):
    """The four reference-pinned pairs, repeated to 24, through the pipelined host-buffer batch entry
    (sm_run_batch): every web equals the reference's, i32 and u8 result formats."""
    gs = [golden["synth/c4/%d/%s" % (k, vname(variant))] for k in range(4)]
    pairs = [orc.synth_pair(g["seed"], g["w"], g["h"], g["D"]) for g in gs]
    n = 24
    first = np.stack([pairs[k % 4][0] for k in range(n)])
    second = np.stack([pairs[k % 4][1] for k in range(n)])
    g0 = gs[0]
    with smb.StereoContext(g0["w"], g0["h"], g0["D"], g0["sw"], variant) as c:
        web, best = c.run_batch(first, second, THRESHOLD, want_best=True)
        web8 = c.run_batch(first, second, THRESHOLD, web_u8=True)
    for k in range(n):
        assert oracle.crc32(web[k]) == gs[k % 4]["web"], k
        assert oracle.crc32(best[k]) == gs[k % 4]["best"], k
        assert oracle.crc32(web8[k].astype(np.int32)) == gs[k % 4]["web"], k


# ---- config 5: the 1080p window/shift sweep, reference-pinned points -----------------------------------
def _sweep_keys(golden):
    return sorted(k for k in golden if k.startswith("sweep1080/"))


def test_sweep1080_has_the_corners(golden):
    keys = _sweep_keys(golden)
    for D, sw in ((16, 21), (32, 13), (128, 17), (512, 3), (512, 21)):
        assert "sweep1080/D%d/sw%d/wrap" % (D, sw) in keys
    assert len(keys) >= 24


@pytest.mark.parametrize("kernel", [smb.KERNEL_BITSLICE, smb.KERNEL_DIRECT], ids=KNAME.get)
def test_sweep1080(orc, golden, kernel):
    n = 0
    for key in _sweep_keys(golden):
        g = golden[key]
        if kernel == smb.KERNEL_DIRECT and g["D"] * g["sw"] > 128 * 17:
            continue  # the literal kernel on the heaviest points takes too long for the suite
        _check(orc, g, kernel)
        n += 1
    assert n >= (12 if kernel == smb.KERNEL_DIRECT else 24)


# ---- config 2: every pair bench.py times (seed 1234 + 2j) ------------------------------------------
def test_config2_bench_pairs(orc, golden):
    """The first 16 of the pairs bench.py runs (rank 0's distinct pairs), through the batched device entry."""
    keys = ["synth/c2s/%d/wrap" % (1234 + 2 * j) for j in range(16)]
    gs = [golden[k] for k in keys]
    pairs = [orc.synth_pair(g["seed"], 1920, 1080, 64) for g in gs]
    first = np.stack([p[0] for p in pairs])
    second = np.stack([p[1] for p in pairs])
    with smb.StereoContext(1920, 1080, 64, 9, smb.WRAP) as c:
        web, best = c.run_batch(first, second, THRESHOLD, want_best=True)
    for j, g in enumerate(gs):
        assert (oracle.crc32(best[j]), oracle.crc32(web[j])) == (g["best"], g["web"]), keys[j]

"""CPU tests of the N>1 path: sharding plans, and a world_size-2 gloo job in which every rank
computes its shard (with the CPU oracle standing in for the device) and rank 0 reassembles.

What is checked is the HOST logic the GPUs rely on: which rows / pairs each rank takes, which
halo rows it must upload (wrapped for WRAP, clipped for GHOST), and that the reassembled
result equals the whole-frame result bit for bit.
"""
import os
import sys

import numpy as np
import pytest

import oracle
from stereomatching_b200 import sharding
from util import ROOT, THRESHOLD, load_pair


def test_pair_plan():
    for n, world in ((4096, 8), (10, 4), (3, 8), (0, 2)):
        seen = sorted(k for r in range(world) for k in sharding.pair_indices(n, world, r))
        assert seen == list(range(n))
        sizes = [len(sharding.pair_indices(n, world, r)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.pair_indices(4, 2, 2)


def test_band_plan_matches_the_c_abi():
    import stereomatching_b200 as smb
    for h in (135, 1080, 2160, 7):
        for world in (1, 2, 3, 4, 8):
            if world > h:
                continue
            for r in range(world):
                assert sharding.band_rows(h, world, r) == smb.band_rows(h, world, r)


def test_band_input_runs():
    # interior band: one run, halo on both sides
    assert sharding.band_input_runs(270, 540, 5, 2160, sharding.WRAP) == [(264, 282)]
    # first band, WRAP: the halo above row 0 is the bottom of the frame
    assert sharding.band_input_runs(0, 270, 5, 2160, sharding.WRAP) == [(2154, 6), (0, 276)]
    assert sharding.band_input_runs(1890, 2160, 5, 2160, sharding.WRAP) == [(1884, 276), (0, 6)]
    # GHOST clips instead
    assert sharding.band_input_runs(0, 270, 5, 2160, sharding.GHOST) == [(0, 276)]
    assert sharding.band_input_runs(1890, 2160, 5, 2160, sharding.GHOST) == [(1884, 276)]
    # a band as tall as the frame takes the frame once
    assert sharding.band_input_runs(0, 100, 10, 100, sharding.WRAP) == [(0, 100)]
    # 12 extra rows per 270-row band at config 3 on 8 GPUs = 4.4 %
    assert abs(sharding.band_halo_overhead(2160, 8, 5) - 12 / 270) < 1e-9


def _band_on_cpu(orc, e1, e2, D, sw, variant, row0, row1):
    """What a band context computes, restated with the oracle: output rows [row0,row1) from the
    edge rows [row0-half, row1+half) taken with the variant's border rule."""
    h, w = e1.shape
    half = sw // 2
    rows = np.arange(row0 - half, row1 + half)
    if variant == oracle.WRAP:
        s1, s2 = e1[rows % h], e2[rows % h]
        # horizontal wrap is the frame's own; the slab's own vertical wrap only touches its halo
        # rows, and the vertical taps of the rows kept below all lie inside the slab
        best, web = orc.match_wta(s1, s2, D, sw, oracle.WRAP)
        return best[half:half + row1 - row0], web[half:half + row1 - row0]
    # GHOST: rows outside the frame contribute no taps -> compute on the clipped slab
    lo, hi = max(rows[0], 0), min(rows[-1] + 1, h)
    best, web = orc.match_wta(e1[lo:hi], e2[lo:hi], D, sw, oracle.GHOST)
    return best[row0 - lo:row1 - lo], web[row0 - lo:row1 - lo]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle.Oracle()
    a, b = load_pair("1-240x135")
    h, w = a.shape
    D, sw = 30, 21
    results = {}
    for variant in (oracle.WRAP, oracle.GHOST):
        e1, e2 = orc.edges(a, THRESHOLD, variant), orc.edges(b, THRESHOLD, variant)
        # --- row bands: this rank's band only
        row0, row1 = sharding.band_rows(h, world, rank)
        lb, lw = _band_on_cpu(orc, e1, e2, D, sw, variant, row0, row1)
        web = np.zeros((h, w), np.int32)
        best = np.zeros((h, w), np.int32)
        web[row0:row1], best[row0:row1] = lw, lb
        gw = sharding.gather_bands(web, h, world, rank)
        gb = sharding.gather_bands(best, h, world, rank)
        # --- whole pairs: 5 pairs, rank takes k mod world
        n_pairs = 5
        mine = sharding.pair_indices(n_pairs, world, rank)
        local = np.stack([orc.match_wta(np.roll(e1, k, 0), np.roll(e2, k, 0), D, 9, variant)[1] for k in mine])
        gp = sharding.gather_pairs(local, n_pairs, world, rank)
        if rank == 0:
            fb, fw = orc.match_wta(e1, e2, D, sw, variant)
            results["bands_%d" % variant] = bool(np.array_equal(gw, fw) and np.array_equal(gb, fb))
            exp = np.stack([orc.match_wta(np.roll(e1, k, 0), np.roll(e2, k, 0), D, 9, variant)[1]
                            for k in range(n_pairs)])
            results["pairs_%d" % variant] = bool(np.array_equal(gp, exp))
    dist.barrier()
    if rank == 0:
        import json
        json.dump(results, open(out_path, "w"))
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import json
    import socket

    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "res.json")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = json.load(open(out))
    assert res == {"bands_0": True, "pairs_0": True, "bands_1": True, "pairs_1": True}, res

#!/usr/bin/env python
"""bench_bands.py -- BASELINE.json configs[2]: ONE synthetic 3840x2160 pair, window 11x11,
256 shifts, row-band sharded with replicated halo rows across N B200 (strong scaling).

    python bench_bands.py                                  # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench_bands.py --steps K --warmup W

Rank r owns output rows band_rows(2160, N, r).  There is no exchange step: every rank uploads
its rows plus a halo of half+1 rows per side and writes its own rows of `web` into host memory
(SURVEY 8e).  torch.distributed (gloo) is used for the barrier and the max of the times only.
  value  frame MDE/s with the band's edge maps resident: W*H*D / max over ranks(time per step)
  e2e    the same through host buffers: H2D of the band + halo, edges, hot path, D2H of the band
Rank 0 checks the assembled web of the last step against the single-GPU whole-frame result.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from bench import THRESHOLD, ClockSampler, synth_pair  # noqa: E402

W, H, D, SW = 3840, 2160, 256, 11


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--variant", default="wrap", choices=["wrap", "ghost"])
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist

    import stereomatching_b200 as smb
    from stereomatching_b200 import sharding

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("gloo")
    variant = smb.GHOST if a.variant == "ghost" else smb.WRAP
    left, right, _ = synth_pair(1234, W, H, D)
    rows = sharding.band_rows(H, world, rank)
    half = SW // 2
    pin_l, pin_r = smb.PinnedBuffer((H, W), np.uint8), smb.PinnedBuffer((H, W), np.uint8)
    pin_web = smb.PinnedBuffer((H, W), np.int32)
    pin_l.array[:], pin_r.array[:] = left, right
    ctx = smb.StereoContext(W, H, D, SW, variant, device=local_rank, rows=rows)
    stream = torch.cuda.Stream()  # a created (non-default) stream
    ctx.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # ---- resident: edges of the band are on the device, time pack + match/WTA
    ctx.upload_u8(pin_l.array, pin_r.array)
    ctx.edges(THRESHOLD)
    for _ in range(max(a.warmup, 3)):
        ctx.match_wta()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(a.steps):
        ctx.match_wta()
    ev1.record(stream)
    barrier()
    ms = reduce_max(ev0.elapsed_time(ev1)) / a.steps
    # ---- end to end through host buffers
    for _ in range(2):
        ctx.upload_u8(pin_l.array, pin_r.array), ctx.edges(THRESHOLD), ctx.match_wta()
        ctx.download(smb.WEB, out=pin_web.array)
    barrier()
    te0 = time.perf_counter()
    for _ in range(a.steps):
        ctx.upload_u8(pin_l.array, pin_r.array), ctx.edges(THRESHOLD), ctx.match_wta()
        ctx.download(smb.WEB, out=pin_web.array)
    barrier()
    te = reduce_max(time.perf_counter() - te0) / a.steps
    clocks = sampler.stop(t0, time.perf_counter()) if sampler else None
    # ---- parity: assemble the bands on rank 0 and compare with a whole-frame context
    web = sharding.gather_bands(pin_web.array.copy(), H, world, rank) if world > 1 else pin_web.array.copy()
    if rank == 0:
        with smb.StereoContext(W, H, D, SW, variant, device=local_rank) as whole:
            whole.upload_u8(left, right)
            whole.edges(THRESHOLD)
            whole.match_wta()
            ref_web = whole.download(smb.WEB)
        equal = bool(np.array_equal(web, ref_web))
        if not equal:
            raise SystemExit("bench_bands.py: assembled bands differ from the whole-frame result")
        runs = sharding.band_input_runs(rows[0], rows[1], half, H, variant)
        h2d = 2 * sum(c for _, c in runs) * W
        print(json.dumps({
            "metric": "hot-path throughput (pixels x shifts per second)", "value": W * H * D / (ms * 1e-3) / 1e6,
            "unit": "MDE/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "one synthetic 3840x2160 pair, window 11x11, 256 shifts, row bands with halo "
                                   "(BASELINE configs[2])", "variant": a.variant, "rows_per_band": rows[1] - rows[0],
                       "halo_rows_replicated_frac": sharding.band_halo_overhead(H, world, half),
                       "parallelism": "row bands, no collective"},
            "frames_per_s": 1e3 / ms, "clocks": clocks, "gpu_launches": a.steps * ctx.last_launches(),
            "e2e": {"value": W * H * D / te / 1e6, "unit": "MDE/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4 * (rows[1] - rows[0]) * W, "frames_per_s": 1 / te,
                    "api": "sm_upload_u8 (band + halo rows) -> sm_edges -> sm_match_wta -> sm_download(SM_WEB) (band rows)",
                    "timer": "host wall clock around synchronised API calls, max over ranks"},
            "parity": {"bands_equal_whole_frame": equal},
        }), flush=True)
    ctx.close()
    pin_l.free(), pin_r.free(), pin_web.free()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

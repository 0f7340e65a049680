"""Host-side sharding plans for more than one GPU.  No GPU code here.

The hot path has no exchange step, so there is NO data-path collective: work is split
before upload and results land in disjoint host rows / pairs.
  * whole pairs (config 4): pair k -> rank k mod world;
  * row bands   (config 3): rank r owns output rows band_rows(H, world, r) of every frame and
    uploads those rows plus a halo of half+1 rows on each side (half for the window, one more
    for the 3x3 edge stencil; wrapped around the frame for WRAP, clipped for GHOST).
torch.distributed is used only to put the per-rank results back together on rank 0
(gloo on CPU tensors in the tests, any backend in a job), never inside the timed path.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

WRAP, GHOST = 0, 1


def pair_indices(n_pairs: int, world: int, rank: int) -> List[int]:
    """Pairs of a batch owned by `rank`: k mod world == rank."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return list(range(rank, n_pairs, world))


def band_rows(height: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous output rows [row0, row1) of `rank`; sizes differ by at most one row.
    Same split as sm_band_rows() in the C ABI."""
    if world < 1 or not (0 <= rank < world) or height < 1:
        raise ValueError("bad band request")
    q, r = divmod(height, world)
    row0 = rank * q + min(rank, r)
    return row0, row0 + q + (1 if rank < r else 0)


def band_input_runs(row0: int, row1: int, half: int, height: int, variant: int, stencil: int = 1):
    """Frame-row runs [(start, count), ...] a band context copies to its device: rows
    [row0 - half - stencil, row1 + half + stencil), wrapped mod height (WRAP) or clipped (GHOST).
    Mirrors row_runs() of csrc/stereo_b200.cu."""
    lo, hi = row0 - half - stencil, row1 + half + stencil
    if hi - lo >= height:
        return [(0, height)]
    if variant == GHOST:
        lo, hi = max(lo, 0), min(hi, height)
        return [(lo, hi - lo)] if hi > lo else []
    runs, a, n = [], lo % height, hi - lo
    while n > 0:
        c = min(n, height - a)
        runs.append((a, c))
        n -= c
        a = 0
    return runs


def band_halo_overhead(height: int, world: int, half: int, stencil: int = 1) -> float:
    """Replicated input rows as a fraction of the frame (SURVEY 8e quotes 4.4 % for config 3 on 8 GPUs)."""
    total = 0
    for r in range(world):
        a, b = band_rows(height, world, r)
        total += min(height, b - a + 2 * (half + stencil))
    return total / height - 1.0


def gather_bands(local: np.ndarray, height: int, world: int, rank: int, group=None):
    """Reassemble a frame from per-rank bands on rank 0 (returns None elsewhere).  `local` is a
    frame-sized array of which only this rank's rows are meaningful."""
    import torch
    import torch.distributed as dist

    row0, row1 = band_rows(height, world, rank)
    q = -(-height // world)  # every rank sends the same number of rows
    send = np.zeros((q,) + local.shape[1:], local.dtype)
    send[: row1 - row0] = local[row0:row1]
    t = torch.from_numpy(send)
    bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, bufs, dst=0, group=group)
    if rank != 0:
        return None
    out = np.empty_like(local)
    for r in range(world):
        a, b = band_rows(height, world, r)
        out[a:b] = bufs[r].numpy()[: b - a]
    return out


def gather_pairs(local: np.ndarray, n_pairs: int, world: int, rank: int, group=None):
    """Reassemble per-pair results (local[i] belongs to pair pair_indices(...)[i]) on rank 0."""
    import torch
    import torch.distributed as dist

    q = -(-n_pairs // world)
    send = np.zeros((q,) + local.shape[1:], local.dtype)
    send[: local.shape[0]] = local
    t = torch.from_numpy(send)
    bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, bufs, dst=0, group=group)
    if rank != 0:
        return None
    out = np.empty((n_pairs,) + local.shape[1:], local.dtype)
    for r in range(world):
        idx = pair_indices(n_pairs, world, r)
        out[idx] = bufs[r].numpy()[: len(idx)]
    return out

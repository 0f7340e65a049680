#!/usr/bin/env python
"""bench.py -- the hot path's headline measurement (BASELINE.json: MDE/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload): BASELINE.json configs[1] -- synthetic textured pairs
1920x1080 with a known disparity field, window 9x9, 64 shifts (SURVEY 8d generator,
seed 1234 + 2k for pair k).  A "step" is one pass of the hot path (bit-plane pack +
match/box/WTA kernel: u8 edge maps resident in HBM -> i32 best + web resident in HBM)
over a batch of `--pairs` stereo pairs, every pair in its own buffers so the step's
working set is far larger than L2.

  value    whole-job MDE/s (pixels x shifts per second), device-timed, inputs resident
  e2e      the same metric through the C ABI with HOST buffers: sm_run_batch
           (H2D u8 images -> edges -> hot path -> D2H i32 web, a three-stage pipeline over groups
           of pairs) per step, wall clock; e2e.link is the measured PCIe ceiling of that call
  roofline the main kernel against the INT32 issue rate measured on this GPU
           (the path is integer-ALU bound; HBM fraction reported alongside)
  cpu_baseline  the reference's own stereo.c hot path (oracle/_ref) on one host core

--impl reference times the UNMODIFIED reference CPU code (oracle/_ref; one process per
host core, each on its own slab of the same workload).
Multi-GPU (torchrun): whole pairs are sharded over ranks, no collective on the data
path (weak scaling: every rank runs `--pairs` pairs per step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, D, SW = 1920, 1080, 64, 9
THRESHOLD = 0.15
OPS_PER_MDE = 8          # algorithmic integer ops per pixel x shift
BYTES_PER_PIXEL = 10     # 2 u8 edge maps in + i32 web + i32 best out
METRIC = "hot-path throughput (pixels x shifts per second)"


# ------------------------------------------------------------------------------------
# synthetic pairs, vectorised numpy; independent of oracle/
# ------------------------------------------------------------------------------------
def _splitmix64(z):
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synth_pair(seed, w, h, d):
    K = np.uint64(0xD6E8FEB86659FD93)
    with np.errstate(over="ignore"):
        ys, xs = np.meshgrid(np.arange(h, dtype=np.uint64), np.arange(w, dtype=np.uint64), indexing="ij")

        def left_at(x):
            hk = _splitmix64(np.uint64(seed) * K + (ys << np.uint64(20)) + x)
            return np.where(((hk >> np.uint64(8)) & np.uint64(7)) == 0, hk & np.uint64(0xFF),
                            np.uint64(128)).astype(np.uint8)

        tw, th = max(240, 4 * d), 120
        hd = _splitmix64(np.uint64(seed + 1) * K + (ys // np.uint64(th)) * np.uint64(4096) + xs // np.uint64(tw))
        disp = (hd % np.uint64(d)).astype(np.int64)
        xsrc = ((xs.astype(np.int64) - disp) % w).astype(np.uint64)
        return left_at(xs), left_at(xsrc), disp.astype(np.int32)


# ------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 8] or \
               [r for (t, r) in self.rows if len(r) >= 8][-3:]
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------
# CPU legs: the reference (oracle/_ref) or, if that is absent, the oracle port
# ------------------------------------------------------------------------------------
def _cpu_slab(args):
    """One worker: hot path of the reference on a slab of `rows` output rows (+ halo)."""
    seed, rows, variant = args
    import oracle
    half = SW // 2
    left, right, _ = synth_pair(seed, W, rows + 2 * half, D)
    if oracle.ref_available(variant, D):
        ref = oracle.RefLib(variant, D)
        e1, e2 = ref.edges(left, THRESHOLD), ref.edges(right, THRESHOLD)
        t0 = time.perf_counter()
        ref.match_wta(e1, e2, SW)
        return time.perf_counter() - t0, "reference"
    orc = oracle.Oracle()
    e1, e2 = orc.edges(left, THRESHOLD, variant), orc.edges(right, THRESHOLD, variant)
    t0 = time.perf_counter()
    orc.match_wta(e1, e2, D, SW, variant, direct=True)
    return time.perf_counter() - t0, "port"


def cpu_baseline_single_core():
    """stereo.c's hot path (fillup_matches + fillup_scores + find_highest_scoring_shifts)
    on ONE core (the reference is single-threaded), bounded sample."""
    rows = 256
    half = SW // 2
    t, kind = _cpu_slab((1234, rows, 0))
    tg, _ = _cpu_slab((1234, rows, 1))
    mde = W * (rows + 2 * half) * D
    return {"value": mde / t / 1e6, "unit": "MDE/s", "cores": 1, "kind": kind,
            "sample": "one 1920x%d slab of the config-2 pair (D=64, sw=9), wrap variant (stereo.c), %.1f s; "
                      "ghost variant (stereo-ghost.c) on the same slab: %.2f MDE/s" % (rows + 2 * half, t, mde / tg / 1e6),
            "host_cores_total": os.cpu_count()}


def whole_algorithm_baselines(pair, variant):
    """The programs' own `elapsed` line for the whole algorithm() (edges + step 2 + step 3, upload
    excluded) on pair 0 of the workload: the reference's stereo.cu rebuilt for sm_100a
    (oracle/_ref, reported baseline) and this repo's C driver.  Best of 3 runs each."""
    import tempfile

    from PIL import Image
    suffix = "-ghost" if variant == "ghost" else ""
    exes = {"reference_cuda_sm100a": os.path.join(ROOT, "oracle", "_ref", "stereopar%s_D%d" % (suffix, D)),
            "this_repo_driver": os.path.join(ROOT, "timing", "stereopar" + suffix)}
    out = {}
    with tempfile.TemporaryDirectory() as d:
        Image.fromarray(pair[0], "L").save(os.path.join(d, "a.png"))
        Image.fromarray(pair[1], "L").save(os.path.join(d, "b.png"))
        for name, exe in exes.items():
            if not os.path.exists(exe):
                out[name] = None
                continue
            best = None
            for _ in range(3):
                r = subprocess.run([exe, "a.png", "b.png", str(THRESHOLD), str(SW)], cwd=d, capture_output=True,
                                   text=True, env=dict(os.environ, STEREO_NUM_SHIFTS=str(D)))
                f = r.stdout.split()
                if r.returncode == 0 and len(f) >= 15:
                    t = float(f[14])
                    best = t if best is None else min(best, t)
            out[name] = None if best is None else {"elapsed_s": best, "MDE_per_s": W * H * D / best / 1e6}
    out["what"] = ("whole algorithm() on pair 0 (edges + match/WTA + hole filling + contour map), each program's own "
                   "'elapsed' line, best of 3 processes; reference = unmodified stereo%s.cu built by oracle/Makefile "
                   "with -gencode arch=compute_100a,code=sm_100a" % suffix)
    return out


def run_reference_arm(a, rank):
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    rows, half = 24, SW // 2
    ctx = mp.get_context("fork")
    kind = "reference"
    with ctx.Pool(cores) as pool:
        def step(s):
            nonlocal kind
            res = pool.map(_cpu_slab, [(1234 + 2 * (s * cores + k), rows, 0) for k in range(cores)])
            kind = res[0][1]
            return max(r[0] for r in res)  # hot-path time of the slowest worker (inputs prepared untimed)
        for s in range(a.warmup):
            step(s)
        t = sum(step(a.warmup + s) for s in range(a.steps))
    mde_step = cores * W * (rows + 2 * half) * D
    value = mde_step * a.steps / t / 1e6
    sample = ("each step: %d processes (one per host core), each the unmodified stereo.c hot path on its own "
              "1920x%d slab of a config-2 pair (D=64, sw=9, wrap)" % (cores, rows + 2 * half))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "MDE/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": t / a.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "synthetic textured pairs 1920x1080, known disparity, window 9x9, 64 shifts "
                               "(BASELINE configs[1]); CPU sample: " + sample, "variant": "wrap"},
        "cpu_baseline": {"value": value, "unit": "MDE/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "MDE/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs next to its GPU (NVML's ideal affinity) so that the pinned staging buffers of
    the end-to-end leg are first-touched on the GPU's own NUMA node.  Best effort; returns what was done."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return "cpus %d-%d (%d)" % (cpus[0], cpus[-1], len(cpus))
    except Exception as e:  # noqa: BLE001
        return "not bound: %s" % type(e).__name__
def run_b200_arm(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import benchlib
    import stereomatching_b200 as smb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        # control plane only (barrier + max of the per-rank time): the data path has no exchange step,
        # so no NCCL communicator is ever needed; gloo keeps stdout to the one JSON line
        dist.init_process_group("gloo")
    variant = smb.GHOST if a.variant == "ghost" else smb.WRAP
    B, distinct = a.pairs, min(a.pairs, a.distinct)

    # ---- inputs: `distinct` synthetic pairs, edges computed on the device, replicated into B buffers
    pairs = [synth_pair(1234 + 2 * (rank * distinct + k), W, H, D) for k in range(distinct)]
    ctx = smb.StereoContext(W, H, D, SW, variant, device=local_rank, kernel=a.kernel)
    e1 = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    e2 = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    for k in range(distinct):
        ctx.upload_u8(pairs[k][0], pairs[k][1])
        ctx.edges(THRESHOLD)
        e1[k] = torch.from_numpy(ctx.download(smb.EDGES1)).to(dev)
        e2[k] = torch.from_numpy(ctx.download(smb.EDGES2)).to(dev)
    for k in range(distinct, B):
        e1[k] = e1[k % distinct]
        e2[k] = e2[k % distinct]
    best = torch.empty((B, H, W), dtype=torch.int32, device=dev)
    web = torch.empty((B, H, W), dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)  # a created (non-default) stream: the legacy default stream serialises against others
    ctx.set_stream(stream.cuda_stream)
    torch.cuda.synchronize()
    p1, p2, pb, pw = e1.data_ptr(), e2.data_ptr(), best.data_ptr(), web.data_ptr()
    n8, n32 = H * W, H * W * 4

    def step():
        if a.no_overlap:
            for k in range(B):
                ctx.match_wta_dev(p1 + k * n8, p2 + k * n8, pb + k * n32, pw + k * n32)
        else:  # one call per batch: the pack of pair k+1 runs beside the main kernel of pair k
            ctx.match_wta_dev_batch(B, p1, p2, n8, pb, pw, n8)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    parity = None
    if rank == 0 and a.variant in ("wrap", "ghost"):
        import zlib
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["synth/c2/" + a.variant]
        crc = "%08x" % (zlib.crc32(web[0].cpu().numpy().tobytes()) & 0xFFFFFFFF)
        parity = {"pair0_web_crc32": crc, "golden": g["web"], "equal": crc == g["web"]}
        if crc != g["web"]:
            raise SystemExit("bench.py: web of pair 0 differs from the reference's golden CRC: %r" % parity)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ctx.profile_begin(a.steps * B)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    tw0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(a.steps):
        step()
    ev1.record(stream)
    barrier()
    tw1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    n_calls, _, overl_ms = ctx.profile_read()
    launches = a.steps * ctx.last_launches()  # per batch call: one pack + one main launch per group of pairs
    # the dominant kernel timed ALONE (one pair per call, nothing overlapped): the roofline figure
    niso = min(B, 32)
    ctx.profile_begin(niso)
    for k in range(niso):
        ctx.match_wta_dev(p1 + k * n8, p2 + k * n8, pb + k * n32, pw + k * n32)
    n_iso, pack_ms, main_ms = ctx.profile_read()
    ctx.profile_begin(0)
    # one pair per call as an application with a single stereo pair would run it: pack + dependent main kernel,
    # calls back to back on one stream (no per-kernel events in between), each pair in its own buffers
    nsp = min(B, 64)
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(4):
        ctx.match_wta_dev(p1 + k * n8, p2 + k * n8, pb + k * n32, pw + k * n32)
    es0.record(stream)
    for k in range(nsp):
        ctx.match_wta_dev(p1 + k * n8, p2 + k * n8, pb + k * n32, pw + k * n32)
    es1.record(stream)
    torch.cuda.synchronize()
    single_us = es0.elapsed_time(es1) * 1e3 / nsp
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    mde_step = B * W * H * D
    value = world * mde_step * a.steps / (ms * 1e-3) / 1e6

    # ---- end to end through the C ABI with host buffers ----------------------------------
    Be = min(B, a.e2e_pairs)
    hin1, hin2 = smb.PinnedBuffer((Be, H, W), np.uint8), smb.PinnedBuffer((Be, H, W), np.uint8)
    hweb = smb.PinnedBuffer((Be, H, W), np.int32)
    for k in range(Be):
        hin1.array[k], hin2.array[k] = pairs[k % distinct][0], pairs[k % distinct][1]
    ectx = smb.StereoContext(W, H, D, SW, variant, device=local_rank, kernel=a.kernel)
    for _ in range(2):
        ectx.run_batch(hin1.array, hin2.array, THRESHOLD, web_out=hweb.array)
    e2e_steps = max(3, min(a.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ectx.run_batch(hin1.array, hin2.array, THRESHOLD, web_out=hweb.array)
    barrier()
    te = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([te], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = float(t.item())
    e2e_val = world * Be * W * H * D * e2e_steps / te / 1e6
    e2e_ok = bool(np.array_equal(hweb.array[0], web[0].cpu().numpy()))
    # the same call with the compact result the ABI offers for num_shifts <= 255 (u8 web: a quarter of the D2H bytes)
    hweb8 = smb.PinnedBuffer((Be, H, W), np.uint8)
    ectx.run_batch(hin1.array, hin2.array, THRESHOLD, web_u8=True, web_out=hweb8.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ectx.run_batch(hin1.array, hin2.array, THRESHOLD, web_u8=True, web_out=hweb8.array)
    barrier()
    te8 = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([te8], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te8 = float(t.item())
    e2e8_val = world * Be * W * H * D * e2e_steps / te8 / 1e6
    e2e8_ok = bool(np.array_equal(hweb8.array[0], hweb.array[0].astype(np.uint8)))
    # the link's own ceiling for those two calls: pinned host<->device copies, both directions at once
    h2d_gbs, d2h_gbs = benchlib.measure_copy_peak(local_rank, 2)
    # clocks: every sample taken between the start of the device-timed region and the end of the e2e one
    clocks = sampler.stop(tw0, time.perf_counter()) if sampler else None

    if rank == 0:
        # ---- roofline of the dominant kernel ------------------------------------------------
        main_s = main_ms * 1e-3 / max(n_iso, 1)
        peaks = {m: benchlib.measure_int_peak(local_rank, i) for i, m in
                 enumerate(["iadd3", "lop3", "iadd3+imad", "lop3+imad"])}
        peak = max(peaks.values())  # the dual-pipe issue ceiling: the hardest denominator
        mp_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = 6650.0, "of fallback (B200_PROFILING.md)"
        if os.path.exists(mp_path):
            hbm_peak, hbm_src = float(json.load(open(mp_path))["hbm_gbs"]), "of measured (MEASURED_PEAKS.json)"
        traffic, prof = None, {}
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            prof = json.load(open(tp))
            traffic = prof.get("main_kernel_dram_bytes_per_launch")
        # executed (not algorithmic) instruction rate: warp instructions per pair as counted by ncu for the timed
        # region's launch shape (profiles/) x 32 threads / the per-pair time measured live here
        issue = None
        bl = prof.get("batched_launch")
        if bl and a.kernel in (0, 2):
            tinstr = bl["warp_instructions"] * 32.0 / bl["pairs"]
            rate = tinstr / (ms * 1e-3 / a.steps / B) / 1e12
            issue = {"executed_thread_instr_per_pair": tinstr, "achieved": rate, "peak": peak / 1e3, "unit": "T thread-instr/s",
                     "frac": rate / (peak / 1e3), "alu_pipe_pct_ncu": bl["alu_pipe_pct_of_peak_active"],
                     "note": "instruction count and ALU-pipe utilisation from ncu (profiles/r01_final_ncu_full_summary.md), "
                             "time from this run's timed region; the kernel executes about 2.5 thread instructions per "
                             "pixel x shift where the algorithmic count assumes 8, hence roofline.frac > 1"}
        achieved = OPS_PER_MDE * W * H * D / main_s / 1e12
        roofline = {
            "bound": "int_alu", "kernel": "bit-sliced match/box/WTA" if ctx.last_launches() else None,
            "achieved": achieved, "peak": peak / 1e3, "unit": "Tiop/s", "frac": achieved / (peak / 1e3),
            "traffic": traffic, "issue": issue,
            "peak_source": "measured live on this GPU: sm_measure_int_peak, max over instruction mixes %s "
                           "(1e9 thread-instr/s)" % json.dumps({k: round(v) for k, v in peaks.items()}),
            "algorithmic_ops": "%d int ops per pixel x shift (SURVEY 8d) x %d per launch" % (OPS_PER_MDE, W * H * D),
            "main_kernel_us": main_s * 1e6, "pack_kernel_us": pack_ms * 1e3 / max(n_iso, 1),
            "timing": "main/pack kernel durations: CUDA events around each launch, one pair per call, nothing "
                      "overlapped (%d launches right after the timed region)" % n_iso,
            "effective_us_per_pair_in_timed_region": ms * 1e3 / a.steps / B,
            "overlapped_launch_us_in_timed_region": overl_ms * 1e3 / max(n_calls, 1),
            "note": "in the timed region main kernels of consecutive pairs alternate two streams and the pack "
                    "kernel runs on a third, so launches overlap and the per-pair time is below the isolated "
                    "kernel duration; frac uses the isolated duration",
            "hbm": {"achieved": BYTES_PER_PIXEL * W * H / main_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": BYTES_PER_PIXEL * W * H / main_s / 1e9 / hbm_peak, "source": hbm_src,
                    "bytes_per_launch": BYTES_PER_PIXEL * W * H},
        }
        cpu = cpu_baseline_single_core() if world == 1 and not a.no_cpu else None
        refcuda = whole_algorithm_baselines(pairs[0], a.variant) if world == 1 and not a.no_cpu else None
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "MDE/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "synthetic textured pairs 1920x1080, known disparity, window 9x9, 64 shifts "
                                   "(BASELINE configs[1])", "variant": a.variant, "pairs_per_step_per_gpu": B,
                       "distinct_pairs": distinct, "frames_per_s": value * 1e6 / (W * H * D),
                       "l2": "every pair has its own input and output buffers: %.1f GB per step, far above the "
                             "126 MB L2" % (B * BYTES_PER_PIXEL * W * H / 1e9),
                       "parallelism": "whole pairs per GPU, no collective", "host_binding_rank0": numa,
                       "one_pair_per_call": {"hot_path_us": single_us, "MDE_per_s": W * H * D / single_us,
                                             "what": "sm_match_wta_dev per pair (pack + dependent main kernel), %d calls "
                                                     "back to back on one stream, every pair in its own (cache-cold) buffers, rank 0" % nsp}},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_val, "unit": "MDE/s", "h2d_bytes_per_step": 2 * Be * W * H,
                    "d2h_bytes_per_step": 4 * Be * W * H, "pairs_per_step": Be, "steps": e2e_steps,
                    "api": "sm_run_batch: pinned host u8 images -> H2D -> edges -> hot path -> D2H i32 web",
                    "timer": "host wall clock around synchronised API calls", "matches_resident_result": e2e_ok,
                    "frames_per_s": e2e_val * 1e6 / (W * H * D),
                    "link": {"h2d_GBps": h2d_gbs, "d2h_GBps": d2h_gbs,
                             "how": "sm_measure_copy_peak: 256 MB pinned copies, both directions at once, best of 3",
                             "ceiling_MDE_per_s": world * W * H * D / max(2 * W * H / (h2d_gbs * 1e9),
                                                                          4 * W * H / (d2h_gbs * 1e9)) / 1e6,
                             "ceiling_with_u8_web_MDE_per_s": world * W * H * D / max(2 * W * H / (h2d_gbs * 1e9),
                                                                                      W * H / (d2h_gbs * 1e9)) / 1e6,
                             "note": "per pair 2 u8 images go up and one i32 (or u8) web comes down; the slower "
                                     "direction bounds pairs/s, compute overlaps"},
                    "with_u8_web": {"value": e2e8_val, "unit": "MDE/s", "d2h_bytes_per_step": Be * W * H,
                                    "equal_to_i32_web": e2e8_ok,
                                    "note": "sm_run_batch(web_u8=1): same values, one byte per pixel"}},
            "roofline": roofline, "cpu_baseline": cpu, "reference_cuda_baseline": refcuda, "parity": parity,
        }), flush=True)
    hin1.free(), hin2.free(), hweb.free(), hweb8.free()
    ctx.close(), ectx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=256, help="stereo pairs per step per GPU")
    ap.add_argument("--distinct", type=int, default=16, help="distinct synthetic pairs generated per GPU")
    ap.add_argument("--e2e-pairs", type=int, default=128)
    ap.add_argument("--variant", default="wrap", choices=["wrap", "ghost"])
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 direct, 2 bit-sliced")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-overlap", action="store_true", help="one sm_match_wta_dev call per pair (pack not overlapped)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if local_rank == 0:
        import __graft_entry__
        __graft_entry__.ensure_built()  # compile step only; there is no fallback path
    else:  # wait for local rank 0's build on a fresh checkout
        lib = os.path.join(ROOT, "stereomatching_b200", "libstereo_b200.so")
        for _ in range(600):
            if os.path.exists(lib):
                break
            time.sleep(0.5)
    if a.impl == "reference":
        run_reference_arm(a, rank)
    else:
        run_b200_arm(a, rank, world, local_rank)


if __name__ == "__main__":
    main()

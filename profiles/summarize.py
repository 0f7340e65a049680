#!/usr/bin/env python
"""profiles/summarize.py -- turn the raw ncu outputs of a measurement run (gpurun_out/) into the
markdown summaries committed here.

    python profiles/summarize.py launches gpurun_out/f_launches.csv  > profiles/rNN_ncu_launch_list_summary.md
    python profiles/summarize.py full gpurun_out/prof_final2.ncu-rep > profiles/rNN_ncu_full_summary.md

`launches` reads the CSV of `ncu --metrics gpu__time_duration.sum --clock-control none --csv`;
`full` reads a `--set full` report through `ncu -i ... --page raw --csv` and `--page source --csv`.
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    iK, iV, iG = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows[hi + 2:]:
        if len(r) <= iV:
            continue
        agg.setdefault((r[iK], r[iG]), []).append(float(r[iV].replace(",", "")))
    step = {k: v for k, v in agg.items() if "k_int_peak" not in k[0]}
    tot = sum(sum(v) for v in step.values())
    print("| kernel | grid | launches | avg ns | total ns | share of non-microbenchmark time |")
    print("|---|---|---|---|---|---|")
    for (k, g), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        share = "-" if "k_int_peak" in k else "%.1f %%" % (100 * sum(v) / tot)
        print("| `%s` | %s | %d | %.0f | %.0f | %s |" % (k[:64], g, len(v), sum(v) / len(v), sum(v), share))
    return agg


METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
]


def full(path, pick=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units, data = rows[0], rows[1], rows[2:]
    seen, cols = set(), []
    for r in data:  # one column per distinct (kernel, grid)
        key = (r[h.index("Kernel Name")], r[h.index("launch__grid_size")])
        if key not in seen:
            seen.add(key)
            cols.append(r)
    print("| metric | " + " | ".join("`%s` grid %s" % (r[h.index("Kernel Name")].split("(")[0][-24:],
                                                        r[h.index("launch__grid_size")]) for r in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for m in METRICS:
        if m in h:
            i = h.index(m)
            print("| %s (%s) | " % (m, units[i]) + " | ".join(r[i] for r in cols) + " |")
    print()
    for r in cols:
        print("Warp stall reasons per issued instruction, `%s` grid %s:" % (r[h.index("Kernel Name")].split("(")[0][-24:],
                                                                             r[h.index("launch__grid_size")]))
        for i, x in enumerate(h):
            if x.startswith("smsp__average_warps_issue_stalled") and x.endswith("per_issue_active.ratio"):
                try:
                    if float(r[i]) >= 0.02:
                        print("* %s: %.3f" % (x[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], float(r[i])))
                except ValueError:
                    pass
        print()
    # opcode mix of the first kernel in the report (SASS page)
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if starts:
        s = starts[0]
        e = starts[1] - 1 if len(starts) > 1 else len(rows)
        hh = rows[s]
        iE, iS = hh.index("Instructions Executed"), hh.index("Source")
        ops, tot = collections.Counter(), 0
        for r in rows[s + 1:e]:
            try:
                n = int(r[iE])
            except (ValueError, IndexError):
                continue
            t = r[iS].split()
            if not t:
                continue
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            ops[op.split(".")[0]] += n
            tot += n
        print("Executed warp instructions by opcode, first captured launch (%d in all):" % tot)
        print()
        print("| opcode | warp instructions | share |")
        print("|---|---|---|")
        for op, n in ops.most_common(16):
            print("| %s | %d | %.1f %% |" % (op, n, 100 * n / tot))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])

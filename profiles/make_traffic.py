#!/usr/bin/env python
"""profiles/make_traffic.py -- distil one `ncu --set full` capture of the throughput launch of the hot kernel
(k_bitslice, several pairs per launch) into profiles/traffic.json, which bench.py reads for roofline.traffic,
roofline.issue and roofline.alu_mix (executed instruction counts cannot be measured outside a profiler).

    python profiles/make_traffic.py gpurun_out/<report>.ncu-rep [pairs in the captured launch] > profiles/traffic.json
"""
import collections
import csv
import json
import subprocess
import sys

ALU = {"LOP3", "SHF", "IADD3", "ISETP", "SEL", "VIADD", "LEA", "VIMNMX", "VIADDMNMX", "PRMT", "FLO", "POPC", "IABS",
       "PLOP3", "MOV", "BREV", "SGXT", "BMSK", "R2P", "P2R", "IMNMX", "VIMNMX3", "VOTE", "VOTEU"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2"}
LSU = {"LDS", "STS", "LDG", "STG", "LD", "ST", "LDTM", "STTM", "SHFL"}


def main():
    path = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, data = rows[0], rows[2:]
    r = data[0]
    g = lambda m: float(r[h.index(m)].replace(",", "")) if m in h else None  # noqa: E731
    units = rows[1]
    # pairs in the captured launch = the grid's z dimension (one pair per z-slice, k_bitslice.cu); an explicit
    # second argument overrides it
    gz = r[h.index("Grid Size")] if "Grid Size" in h else ""
    pairs = int(sys.argv[2]) if len(sys.argv) > 2 else int(gz.strip("() ").split(",")[-1])
    scale = lambda m: {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(units[h.index(m)], 1.0)  # noqa: E731
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    starts = [i for i, x in enumerate(srows) if x and x[0] == "Address"]
    s, e = starts[0], (starts[1] - 1 if len(starts) > 1 else len(srows))
    hh = srows[s]
    iE, iS = hh.index("Instructions Executed"), hh.index("Source")
    ops = collections.Counter()
    for x in srows[s + 1:e]:
        try:
            n = int(x[iE])
        except (ValueError, IndexError):
            continue
        t = x[iS].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        ops[op] += n
    tot = sum(ops.values())
    alu = sum(v for k, v in ops.items() if k in ALU)
    fma = sum(v for k, v in ops.items() if k in FMA)
    lsu = sum(v for k, v in ops.items() if k in LSU)
    out = {
        "source": "ncu --set full --clock-control none of %s, kernel %s, grid %s, %d pairs in the launch"
                  % (path.split("/")[-1], r[h.index("Kernel Name")].split("(")[0].split("::")[-1],
                     r[h.index("launch__grid_size")], pairs),
        "main_kernel_dram_bytes_per_pair": (g("dram__bytes_read.sum") * scale("dram__bytes_read.sum") +
                                            g("dram__bytes_write.sum") * scale("dram__bytes_write.sum")) / pairs,
        "algorithmic_bytes_per_pair": 20736000,
        "batched_launch": {
            "pairs": pairs,
            "warp_instructions_per_pair": tot / pairs,
            "alu_pipe_warp_instructions_per_pair": alu / pairs,
            "fma_pipe_warp_instructions_per_pair": fma / pairs,
            "lsu_tmem_shfl_warp_instructions_per_pair": lsu / pairs,
            "lop3_warp_instructions_per_pair": ops["LOP3"] / pairs,
            "alu_pipe_pct_of_peak_active": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "issue_slots_pct_of_peak_active": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_per_sm": (g("sm__warps_active.avg.pct_of_peak_sustained_active") or 0) * 64 / 100.0,
            "registers_per_thread": g("launch__registers_per_thread"),
            "duration_us_under_ncu": g("gpu__time_duration.sum"),
            "source": "profiles/make_traffic.py: opcode table of the SASS page (ALU pipe = %s)" % ", ".join(sorted(ALU)),
        },
        "note": "algorithmic bytes per pair: 2 u8 edge maps in (4.1 MB, read by the pack kernel) + i32 best + i32 web out (16.6 MB)",
    }
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()

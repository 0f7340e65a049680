#!/usr/bin/env python
"""Split the SASS page of an ncu report into regions of equal execution count (loop bodies, straight-line code)
and print, per region, the executed warp instructions by pipe class, the stall samples and their top reasons.

usage: python tools/sass_regions.py report.ncu-rep [kernel-index] [--dump REGION[:LINES]]
       --dump prints the SASS of one region (first LINES instructions, default 120) as a committed excerpt
"""
import collections
import csv
import subprocess
import sys

ALU = {"LOP3", "SHF", "IADD3", "ISETP", "SEL", "VIADD", "LEA", "VIMNMX", "VIADDMNMX", "PRMT", "FLO", "POPC", "IABS",
       "PLOP3", "MOV", "BREV", "SGXT", "BMSK", "R2P", "P2R", "IMNMX", "VIMNMX3", "UIADD3", "ULOP3", "USHF", "UMOV"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2"}
LSU = {"LDS", "STS", "LDG", "STG", "LDSM", "LD", "ST", "LDC", "LDCU", "ULDC"}


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 0
    dump = None
    if "--dump" in sys.argv:
        d = sys.argv[sys.argv.index("--dump") + 1].split(":")
        dump = (int(d[0]), int(d[1]) if len(d) > 1 else 120)
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    s = starts[which]
    e = starts[which + 1] - 1 if which + 1 < len(starts) else len(rows)
    h = rows[s]
    print(rows[s - 1][1][:120])
    iE, iS, iN = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    stall_cols = [(i, x) for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    regions = []
    cur = None
    for r in rows[s + 1:e]:
        try:
            n = int(r[iE])
        except (ValueError, IndexError):
            continue
        t = r[iS].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        if cur is None or cur["n"] != n:
            cur = {"n": n, "len": 0, "ops": collections.Counter(), "samples": 0, "stalls": collections.Counter(),
                   "text": []}
            regions.append(cur)
        cur["text"].append(r[iS])
        cur["len"] += 1
        cur["ops"][op] += 1
        cur["samples"] += int(r[iN] or 0)
        for i, x in stall_cols:
            cur["stalls"][x[6:]] += int(r[i] or 0)
    tot = sum(r["n"] * r["len"] for r in regions)
    tots = sum(r["samples"] for r in regions)
    print("total executed warp instructions %d, samples %d" % (tot, tots))
    print("| # | exec count | SASS instrs | share of executed | alu | fma | lsu | other | share of samples | top stalls | top opcodes |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for k, r in enumerate(regions):
        ex = r["n"] * r["len"]
        if ex < tot * 0.004:
            continue
        a = sum(v for o, v in r["ops"].items() if o in ALU)
        f = sum(v for o, v in r["ops"].items() if o in FMA)
        l = sum(v for o, v in r["ops"].items() if o in LSU)
        st = ", ".join("%s %.0f%%" % (x, 100.0 * v / max(1, r["samples"])) for x, v in r["stalls"].most_common(4))
        ops = ", ".join("%s %d" % kv for kv in r["ops"].most_common(6))
        print("| %d | %d | %d | %.1f %% | %d | %d | %d | %d | %.1f %% | %s | %s |"
              % (k, r["n"], r["len"], 100.0 * ex / tot, a, f, l, r["len"] - a - f - l, 100.0 * r["samples"] / max(1, tots), st, ops))


    if dump:
        k, nl = dump
        print("\nSASS of region %d (executed %d times per warp-instruction slot), first %d of %d instructions:\n\n```"
              % (k, regions[k]["n"], min(nl, regions[k]["len"]), regions[k]["len"]))
        for t in regions[k]["text"][:nl]:
            print(t)
        print("```")


if __name__ == "__main__":
    main()

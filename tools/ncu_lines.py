#!/usr/bin/env python3
"""Aggregate an ncu source page (cuda,sass view) per source line.

usage: ncu_lines.py report.ncu-rep [top=40]
Prints, for the hottest source lines: warp-level instructions executed, share of the kernel, average
active threads, stall samples — the table the optimisation notes in profiles/ are written from."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    cur_file, hdr = None, None
    lines = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r)}
            continue
        if hdr is None or r[0] in ("Function Name", "Kernel Name", "File Name"):
            continue
        if r[0] == "":          # SASS row (belongs to the preceding source line) — already summed in the line row
            continue
        try:
            ln = int(r[0])
            inst = int(r[hdr["Instructions Executed"]])
            thr = int(r[hdr["Thread Instructions Executed"]])
            smp = int(r[hdr["# Samples"]])
        except (ValueError, KeyError, IndexError):
            continue
        key = (cur_file, ln)
        e = lines.setdefault(key, [0, 0, 0, r[1]])
        e[0] += inst
        e[1] += thr
        e[2] += smp
    tot_inst = sum(v[0] for v in lines.values()) or 1
    tot_thr = sum(v[1] for v in lines.values())
    tot_smp = sum(v[2] for v in lines.values()) or 1
    print(f"total warp-inst {tot_inst:.3e}  thread-inst {tot_thr:.3e}  avg active threads {tot_thr / tot_inst:.2f}")
    per_file = {}
    for (f, _), v in lines.items():
        pf = per_file.setdefault(f, [0, 0, 0])
        pf[0] += v[0]; pf[1] += v[1]; pf[2] += v[2]
    for f, v in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:18s} inst {100 * v[0] / tot_inst:5.1f}%  avg thr {v[1] / max(v[0], 1):5.1f}  samples {100 * v[2] / tot_smp:5.1f}%")
    print(f"{'file:line':24s} {'inst%':>6s} {'avgthr':>6s} {'smp%':>6s}  source")
    for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f + ':' + str(ln):24s} {100 * v[0] / tot_inst:6.2f} {v[1] / max(v[0], 1):6.1f} {100 * v[2] / tot_smp:6.2f}  {v[3].strip()[:90]}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: ncu_launches.py launches.csv"""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        name = re.sub(r"<.*", "", r[ki]).split("(")[0].replace("void ", "").strip()
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        tot[name] += v
        cnt[name] += 1
    all_us = sum(tot.values())
    print(f"{'kernel':28s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{k:28s} {cnt[k]:8d} {v / 1e3:10.3f} {100 * v / all_us:6.1f}% {v / cnt[k]:9.1f}")


if __name__ == "__main__":
    main()

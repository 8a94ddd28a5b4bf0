#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total us, share.
usage: python tools/summarize_launches.py launches.csv [--per-launch]"""
import csv
import sys
from collections import OrderedDict


def load(path):
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        rows.append((r["Kernel Name"], us, r.get("Grid Size", ""), r.get("Block Size", "")))
    return rows


def main():
    rows = load(sys.argv[1])
    if "--per-launch" in sys.argv:
        for i, (k, us, g, b) in enumerate(rows):
            print(f"{i:4d} {us:9.1f} us  grid {g:>14s} {k[:100]}")
        return
    agg = OrderedDict()
    for k, us, _, _ in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:90]}` | {n} | {us:.1f} | {us / n:.1f} | {100 * us / tot:.1f}% |")
    print(f"\nTotal {tot:.0f} us over {len(rows)} launches.")


if __name__ == "__main__":
    main()

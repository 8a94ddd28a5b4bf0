#!/usr/bin/env python
"""Join the ncu CSV of tools/hbm_kernels.py (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch)
with the algorithmic bytes the script printed -> markdown table: per kernel duration, DRAM traffic, algorithmic bytes, achieved
GB/s (algorithmic bytes / duration) and its fraction of the measured HBM peak (MEASURED_PEAKS.json).
usage: python tools/hbm_table.py gpurun_out/hbm_kernels.csv gpurun_out/hbm_kernels.json"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    with open(sys.argv[1], newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    per = OrderedDict()          # (id) -> {name, metrics}
    for r in csv.DictReader(lines):
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = v / 1000.0 if unit.startswith("n") else (v if unit.startswith("u") else v * 1000.0)
        else:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d[r["Metric Name"]] = v * mult
    info = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    algo = info["algorithmic_bytes"]
    peak = 6549.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    agg = OrderedDict()
    for d in per.values():
        if "us" not in d:
            continue
        a = agg.setdefault(d["name"].split("(")[0], {"n": 0, "us": [], "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["us"].append(d["us"])
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    print(f"| kernel | launches | us (last launch, cold caches) | DRAM read MB | DRAM write MB | algorithmic MB | achieved GB/s | of {peak:.0f} GB/s measured |")
    print("|---|---|---|---|---|---|---|---|")
    skip = ("at::", "conv_tc_kernel", "gemm_tc_kernel", "ca_tc_", "ca_mask", "ca_offsets", "ca_flow", "sn_prepare", "pack_weights", "tc_gap_fc")
    for name, a in agg.items():
        if any(k in name for k in skip):      # torch fills / copies of the harness and the tensor-pipe kernels: not this table's subject
            continue
        ab = None
        for pat, b in algo.items():
            if re.search(pat.replace("<", r"<").replace(">", r">"), name) or pat in name:
                ab = b
        if isinstance(ab, list):
            ab = ab[-1]
        us = a["us"][-1]
        rd, wr = a["rd"] / a["n"] / 1e6, a["wr"] / a["n"] / 1e6
        if ab:
            gbs = ab / us / 1e3
            print(f"| `{name[:70]}` | {a['n']} | {us:.1f} | {rd:.2f} | {wr:.2f} | {ab / 1e6:.2f} | {gbs:.0f} | {gbs / peak:.2f} |")
        else:
            gbs = (rd + wr) * 1e6 / us / 1e3
            print(f"| `{name[:70]}` | {a['n']} | {us:.1f} | {rd:.2f} | {wr:.2f} | - | {gbs:.0f} (DRAM traffic) | {gbs / peak:.2f} |")
    print("\nwarm (L2-resident, 20 back-to-back launches, CUDA events), us per call:", json.dumps(info["warm_us_cuda_events"]))


if __name__ == "__main__":
    main()

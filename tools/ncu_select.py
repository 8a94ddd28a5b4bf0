#!/usr/bin/env python
"""Selected columns of an `ncu --set full ... ; ncu -i rep --page raw --csv` export: one row per profiled launch.
usage: python tools/ncu_select.py raw.csv > profiles/rN_ncu_full_selected.csv"""
import csv
import sys

COLS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct"]


def main():
    lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
    rows = list(csv.reader(lines))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = []
    for c in COLS:
        hit = [i for i, h in enumerate(head) if h == c]
        if hit:
            idx.append(hit[0])
    w = csv.writer(sys.stdout)
    w.writerow([head[i].split("TriageCompute.")[-1] + (" [" + units[i] + "]" if units[i] else "") for i in idx])
    for r in body:
        if len(r) > max(idx):
            w.writerow([r[i][:90] for i in idx])


if __name__ == "__main__":
    main()

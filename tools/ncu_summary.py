#!/usr/bin/env python
"""Summarise ncu output for profiles/: `ncu_summary.py launches <launches.csv>` (per-kernel time shares) or
`ncu_summary.py full <report.ncu-rep>` (key metrics of a --set full capture). Prints markdown."""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", "")) * {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}[r[ui]]
        agg.setdefault(r[ki], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | mean ms | total ms | share |\n|---|---:|---:|---:|---:|")
    for k, v in agg.items():
        print(f"| `{k.split('(')[0]}` | {len(v)} | {sum(v) / len(v):.4f} | {sum(v):.3f} | {sum(v) / tot:.3f} |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"### `{r[hdr.index('Kernel Name')]}`  (ID {r[0]})\n\n| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in hdr:
                print(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        print("\nwarp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active.ratio > 0.05):\n")
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.05:
                    print(f"- {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')}: {v:.2f}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])

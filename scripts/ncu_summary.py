#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote.  usage: ncu_summary.py rep [min_us]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; min_us = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.per_cycle_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
        "smsp__inst_executed_pipe_fp32.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    if float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) < min_us:
        continue
    print("----", r[hdr.index("Kernel Name")][:70])
    for k in keys:
        if k in hdr:
            print(f"  {k}: {r[hdr.index(k)]} {units[hdr.index(k)]}")
    # traversal fetches against the chip's bandwidths (north star): sectors are 32 B
    def num(k):
        return float(r[hdr.index(k)].replace(",", "")) if k in hdr and r[hdr.index(k)] else None
    dur = num("gpu__time_duration.sum")
    dur_s = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(units[hdr.index("gpu__time_duration.sum")], 1e-6) if dur else None
    for k, label in (("l1tex__t_sectors.sum", "L1 (l1tex__t_sectors.sum x 32 B)"), ("lts__t_sectors.sum", "L2 (lts__t_sectors.sum x 32 B)")):
        v = num(k)
        if v is not None and dur_s:
            print(f"  {label}: {v * 32 / 1e6:.1f} MB per launch = {v * 32 / dur_s / 1e9:.0f} GB/s")
    for k in ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
              "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
              "smsp__inst_executed_pipe_fp32.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"):
        if k in hdr and r[hdr.index(k)]:
            print(f"  {k}: {r[hdr.index(k)]} {units[hdr.index(k)]}")
    st = [(hdr[i], float(r[i])) for i in range(len(hdr)) if hdr[i].startswith("smsp__average_warps_issue_stalled") and hdr[i].endswith("per_issue_active.ratio") and r[i]]
    st.sort(key=lambda x: -x[1])
    for k, v in st[:7]:
        print(f"    stall {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}")

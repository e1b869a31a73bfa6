#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the one-row CSV kept under profiles/: usage
    ncu_summary.py <report.ncu-rep> <out.csv>"""
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.sum.per_cycle_active", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_aperture_sysmem_op_read.sum", "pcie__read_bytes.sum", "pcie__write_bytes.sum",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor"]
rep, out = sys.argv[1:3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
idx = [hdr.index(k) for k in KEEP if k in hdr]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    w.writerow([vals[i] for i in idx])
print(open(out).read())

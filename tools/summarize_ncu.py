#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full), or the `--page raw --csv` export of one, into a small CSV of the metrics
DESIGN.md / profiles/README.md quote.   usage: summarize_ncu.py <file.ncu-rep | file.raw.csv> <out.csv>"""
import csv, subprocess, sys
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
rep, out = sys.argv[1], sys.argv[2]
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_warp_active.pct")]
cols = ["ID", "Kernel Name"] + [k for k in KEEP if k in hdr] + stall
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(cols)
    w.writerow([units[hdr.index(c)] for c in cols])
    for r in rows[2:]:
        w.writerow([r[hdr.index(c)] for c in cols])
print("wrote", out, len(rows) - 2, "launches")

#!/usr/bin/env python
"""Print selected metrics of every kernel in an .ncu-rep (needs ncu on PATH). Usage: ncu_metrics.py rep [regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]; pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
W = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
     "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
     "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
     "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
     "launch__occupancy_limit_registers", "launch__grid_size", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "sm__cycles_elapsed.max", "launch__waves_per_multiprocessor"]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    if pat and not pat.search(r[ki]): continue
    print("==", r[ki][:90])
    for w in W:
        if w in hdr:
            i = hdr.index(w); print(f"   {w:95s} {r[i]:>16s} {units[i]}")

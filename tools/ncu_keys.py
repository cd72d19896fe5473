#!/usr/bin/env python
"""Print the key metrics of every kernel in an .ncu-rep (via `ncu -i ... --page raw --csv`)."""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled",
        "l1tex__data_bank_conflicts_pipe_lsu", "launch__registers_per_thread", "launch__occupancy_limit", "smsp__inst_executed.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "smsp__pcsamp_warps_issue_stalled"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:60], r[hdr.index("Grid Size")], r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(h.startswith(k) for k in KEYS) or (len(sys.argv) > 2 and sys.argv[2] in h):
            print(f"   {h:90s} {r[i]:>16s} {units[i]}")

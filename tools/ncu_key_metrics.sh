#!/bin/bash
# Key metrics of one `ncu --set full` capture (the format of profiles/*_key_metrics.txt).  Usage: tools/ncu_key_metrics.sh file.ncu-rep
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c '
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__cluster_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__icc_request_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "local_load", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
for i, h in enumerate(hdr):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")):
        try:
            if h.startswith("smsp__average") and float(vals[i]) < 0.05: continue
        except ValueError: pass
        print(h, units[i], vals[i])
'

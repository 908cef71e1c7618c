#!/usr/bin/env python
"""Compact summary of an `ncu --set full` capture for profiles/: the metrics DESIGN.md quotes, one row each.

    ncu -i x.ncu-rep --page raw --csv | python benchmarks/ncu_summary.py "header comment" > profiles/NAME.csv
"""
import csv
import sys

KEEP = ("gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
        "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_global_ld.sum",
        "smsp__sass_inst_executed_op_global_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__inst_executed_pipe_tensor_op_dmma.sum",
        "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active")

rows = [r for r in csv.reader(sys.stdin) if r]
hdr = rows[0]
print("# " + (sys.argv[1] if len(sys.argv) > 1 else "ncu --set full --clock-control none"))
for k, row in enumerate(rows[2:]):
    vals = dict(zip(hdr, zip(rows[1], row)))
    print(f"# kernel {k}: {vals.get('Kernel Name', ('', '?'))[1]}")
    print("metric,unit,value")
    for m in hdr:
        if m in KEEP or (m.startswith("smsp__average_warps_issue_stalled") and m.endswith("_per_issue_active.ratio")):
            print(f"{m},{vals[m][0]},{vals[m][1]}")

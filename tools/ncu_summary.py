#!/usr/bin/env python
"""Key metrics of an `ncu --page raw --csv` export, one block per kernel launch: ncu_summary.py export.raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = rows[0]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
        "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
        "smsp__average_warp_latency_issue_stalled_sleeping.ratio", "smsp__average_warp_latency_issue_stalled_membar.ratio"]
units = rows[1] if len(rows) > 1 else []
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    for k in KEYS:
        if k in d and d[k] != "":
            print("%-75s %s %s" % (k, d[k], u.get(k, "")))
    try:
        rd, wr = float(d.get("dram__bytes_read.sum", 0)), float(d.get("dram__bytes_write.sum", 0))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        print("%-75s %.0f bytes" % ("dram read + write", rd * scale.get(u.get("dram__bytes_read.sum", "byte"), 1.0) + wr * scale.get(u.get("dram__bytes_write.sum", "byte"), 1.0)))
    except ValueError:
        pass
    print()

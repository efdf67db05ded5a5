#!/usr/bin/env python
"""Pick the headline metrics out of `ncu --page raw --csv` (stdin) -> 'name value unit' lines."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "sm__inst_executed_pipe_uniform", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(sys.stdin))
if len(rows) >= 3:
    head, units, vals = rows[0], rows[1], rows[-1]
    for h, u, v in zip(head, units, vals):
        if any(h == w or (w in h and "tensor" in w) for w in WANT):
            print(h, v, u)

"""Prints the headline metrics of every kernel in an .ncu-rep (raw page). Usage: ncu_summary.py REP"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
for r in rows[2:]:
    print(f"## {r[h.index('Kernel Name')]}  (launch id {r[h.index('ID')]})")
    for w in WANT:
        if w in h:
            print(f"  {w:85s} {r[h.index(w)]:>16s} {units[h.index(w)]}")

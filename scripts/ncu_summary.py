"""Selected raw metrics of one `ncu -i rep.ncu-rep --page raw --csv` export, one per line (the format of profiles/*_ncu_summary.txt).
usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python scripts/ncu_summary.py raw.csv"""
import csv, sys
KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "lts__t_sectors_lookup_hit.sum",
        "lts__t_sectors_lookup_miss.sum", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active")
rows = list(csv.reader(open(sys.argv[1])))
names, units, vals = rows[0], rows[1], rows[2]
for n, u, v in zip(names, units, vals):
    if n in KEEP:
        print(f"{n} [{u}] = {v}")

"""Print the key metrics of an `ncu --page raw --csv` export (one kernel).  usage: ncu_key.py raw.csv [pattern ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
header, units, values = rows[0], rows[1], rows[2]
pats = sys.argv[2:] or [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
    "launch__registers", "launch__grid_size", "launch__occupancy_limit", "sm__warps_active.avg.pct",
    "smsp__issue_active.avg.pct", "smsp__warps_eligible.avg.per_cycle", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct", "sm__inst_executed_pipe_alu.avg.pct",
    "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_fp64", "sm__inst_executed_pipe_xu",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct", "sm__throughput.avg.pct", "smsp__average_warp", "smsp__average_warps_issue_stalled",
    "smsp__pcsamp_warps_issue_stalled",
]
for i, h in enumerate(header):
    if any(p in h for p in pats):
        print(f"{h:90s} {values[i]:>16s} {units[i]}")

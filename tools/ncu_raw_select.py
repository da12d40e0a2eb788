"""Keep the judged columns of an `ncu --page raw --csv` export.  usage: ncu_raw_select.py in.csv out.csv [traffic.json]"""
import csv
import json
import sys

KEEP = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max",
]


def main(src, dst, traffic=None):
    rows = list(csv.reader(open(src)))
    header, units, values = rows[0], rows[1], rows[2]
    cols = [i for i, h in enumerate(header) if h in KEEP]
    with open(dst, "w", newline="") as fh:
        w = csv.writer(fh)
        for r in (header, units, values):
            w.writerow([r[i] for i in cols])
    if traffic:
        get = lambda name: float(values[header.index(name)])  # noqa: E731
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        rd = get("dram__bytes_read.sum") * scale[units[header.index("dram__bytes_read.sum")]]
        wr = get("dram__bytes_write.sum") * scale[units[header.index("dram__bytes_write.sum")]]
        json.dump({
            "kernel": values[header.index("Kernel Name")],
            "source": f"profiles/{dst.split('/')[-1]} (ncu --set full --clock-control none, 1 launch, config 2, rows_per_tile=336)",
            "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr,
            "gpu_time_us_under_ncu": get("gpu__time_duration.sum"),
            "registers_per_thread": get("launch__registers_per_thread"),
            "fp64_pipe_pct": get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "lsu_data_pipe_pct": get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        }, open(traffic, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:])

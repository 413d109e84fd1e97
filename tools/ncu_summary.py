"""Trim an `ncu --page raw --csv` dump to the metrics the roofline discussion uses.
    ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/X_summary.csv"""
import csv
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active",
        "gpu__dram_throughput", "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit", "launch__cluster",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "sm__pipe_fp64_cycles_active", "sm__pipe_tensor_cycles_active",
        "sm__inst_executed_pipe", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled",
        "lts__t_sector_hit_rate", "l1tex__t_sector_hit_rate", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "launch__shared_mem_per_block", "lts__t_bytes.sum", "smsp__thread_inst_executed_per_inst_executed")
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
w = csv.writer(sys.stdout)
w.writerow(["kernel", "metric", "unit", "value"])
for vals in rows[2:]:
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    for i, h in enumerate(hdr):
        if any(h.startswith(k) for k in KEEP):
            w.writerow([name[:60], h, units[i], vals[i]])

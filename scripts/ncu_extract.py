"""Reads `ncu --page raw --csv` on stdin and prints the handful of metrics the roofline discussion uses, one per line."""
import csv
import sys

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed_op_global_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum"]
rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
if len(rows) < 3:
    print("no data")
    sys.exit(0)
hdr, units, vals = rows[0], rows[1], rows[2]
print("kernel", sys.argv[1] if len(sys.argv) > 1 else "", "|", vals[hdr.index("Kernel Name")][:80] if "Kernel Name" in hdr else "")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:78s} {vals[i]:>16s} {units[i]}")

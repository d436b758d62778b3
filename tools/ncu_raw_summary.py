"""Print the handful of `ncu --page raw --csv` columns we quote (usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_raw_summary.py)."""
import csv
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct"]
rows = [r for r in csv.reader(sys.stdin) if r]
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:92s} {vals[i]} {units[i]}")

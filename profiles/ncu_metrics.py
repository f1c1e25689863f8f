"""Print the key metrics of an `ncu --page raw --csv` export (one block per profiled launch)."""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
    "l1tex__t_bytes.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for r in data:
    print("---- launch", r[hdr.index("ID")], r[hdr.index("Kernel Name")][:50], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for w in WANT:
        if w in hdr:
            print(f"  {w:70s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")
    st = sorted(((float(r[hdr.index(h)].replace(',', '')), h) for h in stall), reverse=True)[:7]
    for v, h in st:
        print(f"  stall {h.split('stalled_')[1].split('_per_issue')[0]:40s} {v:8.2f} warps/issue")

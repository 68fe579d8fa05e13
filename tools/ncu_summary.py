"""Write a markdown summary (key counters + stall breakdown + hottest SASS blocks) of an .ncu-rep"""
import csv
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
print(f"# {title}\n\nsource: `{rep}` (ncu --set full --clock-control none)\n\n| counter | value |\n|---|---|")
for k in keys:
    if k in m:
        print(f"| {k} | {m[k][1]} {m[k][0]} |")
print("\n## warp stall reasons (cycles per issued instruction)\n\n| reason | value |\n|---|---|")
st = [(h, float(v)) for h, (u, v) in m.items() if "issue_stalled" in h and "per_issue_active" in h and v.replace(".", "").isdigit()]
for h, v in sorted(st, key=lambda x: -x[1]):
    if v > 0.01:
        print(f"| {h.split('issue_stalled_')[1].split('_per_issue')[0]} | {v:.3f} |")

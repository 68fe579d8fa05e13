"""Per-source-line instruction and stall-sample shares of an ncu report captured with --import-source on
(kernels compiled with -lineinfo).  Usage: python tools/ncu_lines.py <rep> [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None
lines = []
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci_inst = hdr.index("Instructions Executed")
        ci_smp = hdr.index("# Samples")
        continue
    if hdr and r[0].isdigit() and len(r) > ci_inst and r[2] == "-":  # a source line row (SASS rows carry an address)
        try:
            lines.append((fname, int(r[0]), r[1].strip(), int(r[ci_inst]), int(r[ci_smp])))
        except ValueError:
            pass
tot = sum(l[3] for l in lines)
tots = sum(l[4] for l in lines)
print(f"total warp instructions {tot:.4g}, samples {tots}")
for f, ln, src, ins, smp in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100 * ins / tot:5.2f}% inst {100 * smp / max(tots, 1):5.2f}% smp  {f}:{ln:<4d} {src[:110]}")

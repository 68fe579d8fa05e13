"""Per-instruction view of an .ncu-rep (SASS page): executed warp instructions per 32-slot step and stall samples,
with markers, aggregated over index ranges.  Usage: python tools/ncu_sass.py <rep> <steps> [lo:hi:name ...]"""
import csv
import subprocess
import sys

rep, steps = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
cstall = {h: k for k, h in enumerate(hdr) if h.startswith("stall_")}
ins = []
for r in rows[2:]:
    try:
        ins.append((r[1].strip(), int(r[ci]), int(r[cs]), r))
    except (ValueError, IndexError):
        pass
tot, tots = sum(i[1] for i in ins), sum(i[2] for i in ins)
print(f"{len(ins)} SASS instructions, {tot:.4g} executed = {tot / steps:.1f} per step, {tots} samples")
ranges = [a.split(":") for a in sys.argv[3:]]
if not ranges:
    MARK = ("MATCH", "LDGSTS", "ATOM", "DEPBAR", "RED", "MUFU.RCP64H", "STG", "LDG", "BAR", "WARPSYNC", "VOTE")
    for k, (s, c, sm, _) in enumerate(ins):
        if any(m in s for m in MARK) and (c or sm):
            print(f"{k:5d} {c / steps:8.3f}/step {100 * sm / tots:5.2f}% smp  {s[:100]}")
for lo, hi, name in ranges:
    lo, hi = int(lo), int(hi)
    c = sum(i[1] for i in ins[lo:hi])
    s = sum(i[2] for i in ins[lo:hi])
    top = {}
    for h, k in cstall.items():
        v = sum(int(i[3][k]) for i in ins[lo:hi] if i[3][k].isdigit())
        if v:
            top[h] = v
    tops = ", ".join(f"{h[6:]} {100 * v / max(tots, 1):.1f}" for h, v in sorted(top.items(), key=lambda x: -x[1])[:4])
    print(f"{name:28s} [{lo},{hi}) {c / steps:8.1f} inst/step {100 * s / tots:5.1f}% samples  ({tops})")

"""Summarise an ncu report's SASS page by blocks of instructions (executed counts, stall samples)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for k, r in enumerate(rows[2:]):
    try:
        data.append((k, r[ci['Source']].strip(), int(r[ci['Instructions Executed']]), int(r[ci['# Samples']]), float(r[ci['Avg. Threads Executed']])))
    except Exception:
        pass
tot = sum(d[2] for d in data)
tots = sum(d[3] for d in data)
print("n sass", len(data), "total inst", tot, "samples", tots)
shfl = sum(d[2] for d in data if 'SHFL' in d[1])
print("SHFL executed", shfl, " DDIV-ish MUFU.RCP64H", sum(d[2] for d in data if 'RCP64H' in d[1]))
for b in range(0, len(data), blk):
    seg = data[b:b + blk]
    ins = sum(d[2] for d in seg)
    smp = sum(d[3] for d in seg)
    if ins / tot < 0.004 and smp / max(tots, 1) < 0.004:
        continue
    ops = {}
    for d in seg:
        t = d[1].split()
        op = t[0] if not t[0].startswith('@') else t[1]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:5]
    print(f"{b:5d} inst {100 * ins / tot:5.1f}% smp {100 * smp / max(tots, 1):5.1f}% thr {sum(d[4] * d[2] for d in seg) / max(ins, 1):5.1f} max/inst {max(d[2] for d in seg):.3g}", top)

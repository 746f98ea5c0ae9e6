#!/usr/bin/env python3
"""Per-CUDA-source-line hot spots of one kernel launch in an ncu report.
  ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
  python tools/ncu_lines.py src.csv <section index> [top N]
Prints samples, warp instructions, and average active threads per source line."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"] + [len(rows)]
k = int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
a, b = starts[k], starts[k + 1]
print(rows[a][1][:100])
h = rows[a + 1]
ismp, iex, ith = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
ill = h.index("stall_long_sb")
lines = []
for r in rows[a + 2:b]:
    if r[0] == "" or len(r) <= ith:
        continue
    num = lambda v: int(v) if v.strip().lstrip("-").isdigit() else 0
    lines.append((num(r[ismp]), num(r[iex]), num(r[ith]), num(r[ill]), r[0], r[1]))
ts, te = sum(l[0] for l in lines), sum(l[1] for l in lines)
print(f"total samples {ts}, warp instr {te}")
for s, e, t, ll, no, src in sorted(lines, reverse=True)[:top]:
    print(f"{100 * s / ts:5.1f}% smp {100 * e / te:5.1f}% ins  thr {t / e if e else 0:4.1f}  lsb {100 * ll / max(s, 1):3.0f}%  {no:>4} {src.strip()[:110]}")

#!/usr/bin/env python3
"""Instruction mix of the BVH descend loop of a traversal kernel (development aid).
  python tools/sass_loop.py <librt_b200.so> <mangled-name prefix> [--dump]
The loop is found as the span from the first 256-bit (or first 128-bit CONSTANT) node load
to the backward branch that closes it."""
import collections
import re
import subprocess
import sys

lib, prefix = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(out) if "Function :" in l and prefix in l)
end = next((i for i in range(start + 1, len(out)) if "Function :" in out[i]), len(out))
ins = []
for l in out[start:end]:
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
def is_node_load(t):
    return "ENL2.256" in t or "LDG.E.128.CONSTANT" in t
# the node fetch is the first place where >= 4 vector loads sit within 14 instructions
first = next(i for i in range(len(ins)) if sum(is_node_load(t) for _, t in ins[i:i + 14]) >= 4 and is_node_load(ins[i][1]))
# loop head: nearest preceding instruction that is the target of a later backward branch
targets = {}
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a:
            targets.setdefault(tgt, i)
head_addr = max(t for t in targets if t <= ins[first][0])
head = next(i for i, (a, _) in enumerate(ins) if a == head_addr)
tail = max(i for t, i in targets.items() if t == head_addr)
body = ins[head:tail + 1]
c = collections.Counter()
for _, t in body:
    op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
    c[op] += 1
print(f"{prefix}: loop {len(body)} instrs:", dict(c.most_common()))
if "--dump" in sys.argv:
    for a, t in body:
        print(f"{a:06x}  {t}")

#!/usr/bin/env python3
"""gpurun_out/<tag>_scale_n{1,2,4,8}.json (tools/scale.sh) -> profiles/<tag>_scaling.json + a markdown table on stdout."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
runs = {}
for n in (1, 2, 4, 8):
    f = ROOT / "gpurun_out" / f"{tag}_scale_n{n}.json"
    if f.exists():
        lines = [l for l in f.read_text().splitlines() if l.startswith("{")]
        if lines:
            runs[n] = json.loads(lines[-1])
out = {"runs": {str(n): r for n, r in runs.items()}}
(ROOT / "profiles" / f"{tag}_scaling.json").write_text(json.dumps(out, indent=1) + "\n")
base = runs.get(1)
print("| GPUs | frame ms | Mrays/s | × | e2e ms / Mrays/s | e2e × | frame = single-rank frame |")
print("|---|---|---|---|---|---|---|")
for n, r in sorted(runs.items()):
    e = r.get("e2e") or {}
    print(f"| {n} | {r['ms_per_step']:.2f} | {r['value']:.0f} | {r['value'] / base['value']:.2f} | {e.get('ms_per_step', 0):.2f} / {e.get('value', 0):.0f} | "
          f"{e.get('value', 0) / base['e2e']['value']:.2f} | {r.get('frame_matches_single_rank', 'n/a')} |")
names = ["input01", "teapot", "refraction3", "bunny"]
print()
print("| config | " + " | ".join(f"{n} GPU: ms / Mrays/s" for n in sorted(runs)) + " |")
print("|---|" + "---|" * len(runs))
for nm in names:
    cells = []
    for n, r in sorted(runs.items()):
        pc = (r.get("per_config") or {}).get(nm)
        cells.append(f"{pc['ms_per_step']:.3f} / {pc['value']:.0f}" if pc else "")
    print(f"| {nm} | " + " | ".join(cells) + " |")

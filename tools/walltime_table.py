#!/usr/bin/env python3
"""profiles/r02_walltime.json (tools/walltime.py) -> the markdown table of BASELINE.md section 4 on stdout."""
import json
from pathlib import Path

d = json.loads((Path(__file__).resolve().parent.parent / "profiles" / "r02_walltime.json").read_text())
fl = d["cuda_floor"]
print("| input (command line of `notes/notes-0N.txt`) | reference, published (2014 machine, `-t 8`) | reference binary, this GPU box "
      f"(`-t {d['threads']}`) | `as2` on the B200, best of 3 (all 3) | of which device: upload + LBVH + read-back (+ trace) | who wins on this box |")
print("|---|---|---|---|---|---|")
for r in d["rows"]:
    p = r["phases_last_run"] or {}
    dev = p.get("device_upload_ms", 0) + p.get("lbvh_ms", 0) + p.get("readback_ms", 0)
    ref, a = r["reference_wall_s_this_box"], r["as2_b200_wall_s"]
    ref_s = f"{ref:.3f} s" if ref is not None else "aborted"
    print(f"| `{r['input']}` {r['size']}x{r['size']} | {r['reference_published_s']:.3f} s | {ref_s} | {a:.2f} s "
          f"({', '.join(f'{x:.2f}' for x in r['as2_b200_wall_s_all'])}) | {dev:.1f} ms + trace (first launch of every kernel included) | "
          f"{'as2' if (ref is None or a < ref) else 'reference'} |")
print()
print(f"Empty CUDA program: {fl['all_gpus_visible']['min_s']:.2f} s best of 3 ({', '.join(f'{x:.2f}' for x in fl['all_gpus_visible']['all_s'])}) with all GPUs "
      f"visible, {fl['one_gpu_visible']['min_s']:.2f} s ({', '.join(f'{x:.2f}' for x in fl['one_gpu_visible']['all_s'])}) with CUDA_VISIBLE_DEVICES=0; "
      f"persistence mode {fl.get('nvidia_smi_persistence_mode')}")

#!/usr/bin/env python3
"""Whole-process wall time of the drop-in executable next to the reference's own binary, on the nine shipped
inputs with the command lines of /root/reference/notes/notes-0N.txt (the only numbers the reference publishes).

  python tools/walltime.py [--skip-reference] [--repeat 3] > gpurun_out/walltime.json

bin/as2 (product) and oracle/_ref/as2_ref (the UNMODIFIED reference, tests/pathb/Makefile) run on the same box,
same inputs, same sizes; the reference gets every host thread (-t nproc).  AS2_TIMING=1 splits the product's
wall time into phases; "cuda_init" is what is left of the render call after the device-side times.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
AS2 = ROOT / "cs184-raytracer_b200" / "bin" / "as2"
REF = ROOT / "oracle" / "_ref" / "as2_ref"
PUBLISHED = {"01": 0.383, "02": 74.061, "03": 201.573, "04": 62.856, "05": 0.933, "06": 4.671, "07": 2.687, "08": 0.371, "09": 0.697}


def run(cmd, env=None):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, capture_output=True, text=True, env=env)
    return time.perf_counter() - t0, p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-reference", action="store_true")
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--inputs", default="01,02,03,04,05,06,07,08,09")
    args = ap.parse_args()
    threads = os.cpu_count() or 8
    rows = []
    tmp = Path(tempfile.mkdtemp())
    # the floor: an empty CUDA program (context creation + one empty kernel), all GPUs visible and one GPU visible
    floor = {}
    exe = ROOT / "cs184-raytracer_b200" / "bin" / "cuda_floor"
    if exe.exists():
        for label, e in (("all_gpus_visible", dict(os.environ)), ("one_gpu_visible", dict(os.environ, CUDA_VISIBLE_DEVICES="0"))):
            e.pop("CUDA_VISIBLE_DEVICES", None) if label == "all_gpus_visible" else None
            ts = [run([str(exe)], e)[0] for _ in range(max(args.repeat, 3))]
            floor[label] = {"min_s": min(ts), "all_s": ts}
    try:
        q = subprocess.run(["nvidia-smi", "--query-gpu=persistence_mode,name", "--format=csv,noheader"], capture_output=True, text=True).stdout
        floor["nvidia_smi_persistence_mode"] = q.strip().splitlines()
    except OSError:
        pass
    print(json.dumps({"cuda_floor": floor}), file=sys.stderr)
    env = dict(os.environ, AS2_TIMING="1")
    for n in args.inputs.split(","):
        size = 2000 if n == "09" else 1000
        rti = str(ROOT / "tests" / "golden" / "inputs" / f"input-{n}.rti")
        common = [rti, "-h", str(size), "-w", str(size), "-t", str(threads)]
        ours, phases = [], None
        for _ in range(args.repeat):
            t, p = run([str(AS2)] + common + ["-o", str(tmp / f"ours-{n}.png")], env)
            if p.returncode != 0:
                print(p.stderr, file=sys.stderr)
                break
            ours.append(t)
            m = re.search(r"timing: parse ([\d.]+) ms \| render call ([\d.]+) ms .*upload ([\d.]+), LBVH ([\d.]+), trace ([\d.]+), "
                          r"readback ([\d.]+)\) \| png ([\d.]+) ms \| total ([\d.]+) ms", p.stderr)
            if m:
                v = [float(x) for x in m.groups()]
                phases = {"parse_ms": v[0], "render_call_ms": v[1], "device_upload_ms": v[2], "lbvh_ms": v[3], "trace_ms": v[4],
                          "readback_ms": v[5], "png_ms": v[6], "main_total_ms": v[7],
                          "cuda_init_and_alloc_ms": v[1] - v[2] - v[3] - v[4] - v[5]}
        ref = None
        if not args.skip_reference and REF.exists():
            t, p = run([str(REF)] + common + ["-o", str(tmp / f"ref-{n}.png")])
            ref = t if p.returncode == 0 else None
        rows.append({"input": f"input-{n}.rti", "size": size, "as2_b200_wall_s": min(ours) if ours else None,
                     "as2_b200_wall_s_all": ours, "phases_last_run": phases, "reference_wall_s_this_box": ref,
                     "reference_threads": threads, "reference_published_s": PUBLISHED[n]})
        print(json.dumps(rows[-1]), file=sys.stderr)
    print(json.dumps({"rows": rows, "threads": threads, "cuda_floor": floor}))


if __name__ == "__main__":
    main()

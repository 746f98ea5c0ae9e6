#!/usr/bin/env python3
"""ncu `--page raw --csv` (converted on the GPU box, the .ncu-rep files are too large to bring back) -> tracked
markdown summary under profiles/ and the per-kernel facts bench.py reads from profiles/traffic.json.

  python tools/summarize_raw.py <raw.csv> <out.md> "<command>" [--traffic <workload>]

--traffic: the capture holds every k_trace / k_shadow launch of ONE frame; writes per kernel the summed DRAM
bytes, the launch count and the instruction-weighted active lanes per instruction into profiles/traffic.json.
"""
import csv
import json
import sys
from pathlib import Path

KEYS = [
    ("gpu__time_duration.sum", "ms", "time"), ("launch__grid_size", "blocks", 1), ("launch__registers_per_thread", "regs", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instr", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %", 1),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts %", 1),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1), ("lts__t_sector_hit_rate.pct", "L2 hit %", 1),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %", 1),
    ("smsp__inst_executed.sum", "warp instr (M)", 1e-6),
    ("dram__bytes_read.sum", "DRAM read (MB)", None), ("dram__bytes_write.sum", "DRAM write (MB)", None),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak", 1),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard", 1),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait", 1),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected", 1),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe", 1),
]
UNIT_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    traffic_wl = sys.argv[5] if len(sys.argv) > 5 and sys.argv[4] == "--traffic" else None
    rows = list(csv.reader(open(src)))
    head, units = rows[0], rows[1]
    col = {k: i for i, k in enumerate(head)}
    launches = []
    for r in rows[2:]:
        if len(r) < len(head):
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("rt::", "")
        vals = {}
        for key, label, scale in KEYS:
            if key not in col:
                continue
            try:
                v = float(r[col[key]].replace(",", ""))
            except ValueError:
                continue
            if scale == "time":
                v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(units[col[key]], 1.0)
            elif scale is None:
                v = v * UNIT_BYTES.get(units[col[key]], 1) / 1e6
            else:
                v *= scale
            vals[label] = v
        launches.append((name, vals))
    labels = [l for _, l, _ in KEYS]
    out = [f"# ncu --set full, raw page\n\n`{cmd}`\n",
           "| # | kernel | " + " | ".join(labels) + " |", "|---|---|" + "---|" * len(labels)]
    for i, (name, v) in enumerate(launches):
        out.append(f"| {i} | `{name}` | " + " | ".join(f"{v[l]:.2f}" if l in v else "" for l in labels) + " |")
    Path(dst).write_text("\n".join(out) + "\n")
    if traffic_wl:
        tf = Path(__file__).resolve().parent.parent / "profiles" / "traffic.json"
        data = json.loads(tf.read_text()) if tf.exists() else {}
        data["_comment"] = ("per kernel, from ncu --set full of every launch of ONE single-GPU frame: dram__bytes_read.sum + "
                            "dram__bytes_write.sum summed over the frame, the launches, and the instruction-weighted active "
                            "lanes per instruction; read by bench.py (roofline.traffic, active_lanes_per_instruction_ncu)")
        wl = {}
        for kern in ("k_trace", "k_shadow"):
            sel = [v for n, v in launches if n.startswith(kern)]
            if not sel:
                continue
            instr = sum(v["warp instr (M)"] for v in sel)
            wl[kern] = {"dram_bytes_per_frame": sum((v["DRAM read (MB)"] + v["DRAM write (MB)"]) * 1e6 for v in sel),
                        "launches_per_frame": len(sel), "ms_under_ncu": sum(v["ms"] for v in sel),
                        "active_lanes": sum(v["active lanes / instr"] * v["warp instr (M)"] for v in sel) / instr,
                        "source": Path(dst).name}
        data[traffic_wl] = wl
        tf.write_text(json.dumps(data, indent=1) + "\n")
        print(json.dumps(wl, indent=1))


if __name__ == "__main__":
    main()

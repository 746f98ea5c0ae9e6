#!/usr/bin/env python3
"""Turns ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_ncu.py launches <launches.csv> <out.csv> "<command>"
  python tools/summarize_ncu.py full <report.ncu-rep> <out.md> "<command>"
"""
import collections
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
       "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
       "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
       "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def launches(src, dst, cmd):
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    tot, cnt, out = collections.defaultdict(float), collections.Counter(), []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("rt::", "")
        v, u = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
        v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
        out.append((r["ID"], name, r["Grid Size"], r["Block Size"], v))
        tot[name] += v
        cnt[name] += 1
    s = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# {cmd}\n# per-launch device time (cold-cache, serialised under ncu: compare SHARES, not absolutes)\n")
        f.write("# kernel share of the captured launches:\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"#   {k:28s} {v:10.3f} ms  {100 * v / s:5.1f} %  launches {cnt[k]}\n")
        f.write("id,kernel,grid,block,ms\n")
        for o in out:
            f.write(f'{o[0]},{o[1]},"{o[2]}","{o[3]}",{o[4]:.6f}\n')
    print(open(dst).read().split("id,kernel")[0])


def full(rep, dst, cmd):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    starts = [i for i, r in enumerate(srows) if r and r[0] == "Kernel Name"] + [len(srows)]
    mixes = []
    for a, b in zip(starts[:-1], starts[1:]):
        h, data = srows[a + 1], srows[a + 2:b]
        sec_name = srows[a][1]
        isrc, ismp, iex, ith = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
        smp, ex, th = collections.Counter(), collections.Counter(), collections.Counter()
        for r in data:
            if len(r) < len(h):
                continue
            try:
                s, e, t = int(r[ismp]), int(r[iex]), int(r[ith])
            except ValueError:
                continue
            toks = r[isrc].split()
            op = (toks[1] if toks and toks[0].startswith("@") else toks[0]).split(".")[0] if toks else "?"
            smp[op] += s; ex[op] += e; th[op] += t
        mixes.append((sec_name, smp, ex, th))
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary\n\n`{cmd}`\n\n")
        for k, row in enumerate(rows[2:]):
            f.write(f"## launch {k}: `{row[hdr.index('Kernel Name')][:80]}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for m in RAW:
                if m in hdr:
                    f.write(f"| {m} | {row[hdr.index(m)]} | {units[hdr.index(m)]} |\n")
            st = {h: float(row[i] or 0) for i, h in enumerate(hdr)
                  if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h}
            f.write("\nwarp stall reasons (warps per issue-active cycle): " + ", ".join(
                f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}"
                for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]) + "\n\n")
            kname = row[hdr.index('Kernel Name')].split('(')[0].replace('void ', '')
            cand = [m for m in mixes if kname.split('<')[0] in m[0]]
            nth = sum(1 for r2 in rows[2:2 + k] if r2[hdr.index('Kernel Name')].split('(')[0].replace('void ', '') == kname)
            if cand:
                _, smp, ex, th = cand[min(nth, len(cand) - 1)]
                te, ts = sum(ex.values()), sum(smp.values())
                f.write(f"SASS mix (warp-instructions {te}, avg active threads {sum(th.values()) / max(te, 1):.1f}):\n\n| opcode | % instr | % samples | avg threads |\n|---|---|---|---|\n")
                for op, e in ex.most_common(16):
                    f.write(f"| {op} | {100 * e / te:.1f} | {100 * smp[op] / max(ts, 1):.1f} | {th[op] / max(e, 1):.1f} |\n")
                tensor = [op for op in ex if op.startswith(("UTC", "HMMA", "LDTM", "UTMA"))]
                f.write(f"\ntensor/TMA opcodes present: {tensor or 'none (by design: no dense contraction on this path)'}\n\n")
    print(open(dst).read()[:1500])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])

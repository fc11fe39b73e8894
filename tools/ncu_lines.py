#!/usr/bin/env python
"""usage: ncu_lines.py <report.ncu-rep> [top]  -- headline metrics + per-source-line instruction / stall-sample shares"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, data = rr[0], rr[1], rr[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
for d in data:
    print("##", d[idx["Kernel Name"]][:100])
    for w in want:
        if w in idx: print(f"  {w:75s} {d[idx[w]]:>16s} {units[idx[w]]}")
    st = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
    vals = sorted([(float(d[idx[h]].replace(",", "")), h.split("issue_stalled_")[1].split("_per")[0]) for h in st if d[idx[h]] not in ("", "n/a")], reverse=True)[:6]
    print("  stalls:", ", ".join(f"{n} {v:.2f}" for v, n in vals))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; agg = {}; h = None
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": h = r; continue
    if r and r[0] == "Function Name": continue
    if h and len(r) > 8 and r[0] != "":
        try: line = int(r[0]); inst = int(r[7]); samp = int(r[6])
        except Exception: continue
        a = agg.setdefault((cur, line), [0, 0, r[1][:110]]); a[0] += inst; a[1] += samp
tot = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
print(f"total warp instructions {tot}, samples {ts}")
for (f, l), (i, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{l:<5d} inst {i / tot * 100:5.1f}% samp {s / ts * 100:5.1f}%  {src}")

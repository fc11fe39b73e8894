#!/usr/bin/env python
"""Summarises gpurun_out/launches_<tag>.csv (ncu gpu__time_duration per launch) and prof_walk_<tag>.ncu-rep
(ncu --set full of the dense column-walk kernel) into profiles/<tag>_*.md.   usage: summarize_profile.py <tag> [note]"""
import collections
import csv
import os
import subprocess
import sys

tag = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else ""
blocks = sys.argv[3] if len(sys.argv) > 3 else "1000000"
os.makedirs("profiles", exist_ok=True)
lines = [l for l in open(f"gpurun_out/launches_{tag}.csv") if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    d = agg.setdefault(r["Kernel Name"], [0, 0.0])
    d[0] += 1
    d[1] += float(r["Metric Value"].replace(",", ""))
tot = sum(d[1] for d in agg.values())
with open(f"profiles/{tag}_launches.md", "w") as f:
    f.write(f"# ncu launch list `{tag}` — {note}\n\n")
    f.write(f"Command (tools/profile.sh): `ncu --metrics gpu__time_duration.sum --clock-control none -c 900 python bench.py --blocks {blocks} "
            "--steps 1 --warmup 3 --no-e2e --no-indep --no-side --no-sharded` (cold-cache, serialised: compare SHARES, not absolutes; the "
            "`at::...FillFunctor` launches are torch zero-filling the bench's device buffer, outside every step).  "
            f"{len(rows)} launches, {tot / 1e6:.2f} ms in total.\n\n| kernel | launches | total ms | us/launch | share |\n|---|---:|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k[:90]}` | {n} | {t / 1e6:.3f} | {t / n / 1e3:.1f} | {100 * t / tot:.1f}% |\n")
rep = f"gpurun_out/prof_walk_{tag}.ncu-rep"
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, data = rr[0], rr[1], rr[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct"]
    with open(f"profiles/{tag}_walk_ncu.md", "w") as f:
        f.write(f"# ncu --set full, dense column-walk kernel, `{tag}` — {note}\n\n")
        f.write(f"Command (tools/profile.sh): `ncu --set full --clock-control none --import-source on -k regex:k_walk_tma -s 4 -c 1 python bench.py "
                f"--blocks {blocks} --steps 1 --warmup 3 ...`.  Never a bench value: kernels are replayed.\n\n")
        for d in data:
            f.write(f"## {d[idx['Kernel Name']]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in want:
                if w in idx:
                    f.write(f"| {w} | {d[idx[w]]} | {units[idx[w]]} |\n")
            st = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
            vals = sorted([(float(d[idx[h]].replace(",", "")), h.split("issue_stalled_")[1].split("_per")[0]) for h in st
                           if d[idx[h]] not in ("", "n/a")], reverse=True)[:5]
            f.write("\nTop stall reasons (warps per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in vals) + "\n\n")
print("wrote profiles for", tag)

#!/usr/bin/env python
"""Turn the ncu artefacts in gpurun_out/ into the committed summaries under profiles/.
usage: python profiles/summarize.py r01"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

# ---- launch list (gpu__time_duration, cold-cache, serialised: compare shares, not absolutes) ----
rows = list(csv.reader(open(os.path.join(G, f"{tag}_launches.csv"))))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, mi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
tot = defaultdict(float)
cnt = defaultdict(int)
lines = []
for r in rows[h + 1:]:
    if len(r) <= mi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").strip()
    t = float(r[mi].replace(",", "")) / 1e3  # ns -> us
    tot[name] += t
    cnt[name] += 1
    lines.append((name, r[gi], t))
with open(os.path.join(P, f"{tag}_launches.csv"), "w") as fh:
    fh.write("launch,kernel,grid,duration_us\n")
    for i, (n, g, t) in enumerate(lines):
        fh.write(f'{i},{n},"{g}",{t:.3f}\n')
total = sum(tot.values())
share = {k: {"launches": cnt[k], "total_us": round(v, 1), "share": round(v / total, 4)} for k, v in sorted(tot.items(), key=lambda kv: -kv[1])}

# ---- full-set metrics per captured kernel ----
rep = os.path.join(G, f"{tag}_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
idx = {k: i for i, k in enumerate(hdr)}
want = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "launch__registers_per_thread": "regs",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_inst",
    "smsp__inst_executed.sum": "warp_insts",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
}
kern = []
for r in rr[2:]:
    d = {"kernel": r[idx["Kernel Name"]].split("(")[0].replace("void ", "").strip(), "grid": r[idx["Grid Size"]]}
    for m, short in want.items():
        if m in idx:
            d[short] = f"{r[idx[m]]} {units[idx[m]]}".strip()
    kern.append(d)


def to_bytes(s):
    v, u = s.split()
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


traffic = {}
cls = {"k_leaf_propose": "leaf", "k_internal_propose": "internal", "k_scan": "scan", "k_darts_count": "darts", "k_expand": "resample"}
for d in kern:
    for k, c in cls.items():
        if d["kernel"].startswith(k) or d["kernel"].startswith("dcsmc::" + k):
            b = to_bytes(d["dram_read"]) + to_bytes(d["dram_write"])
            traffic[c] = max(traffic.get(c, 0), int(b))  # the largest launch of the class (leaf / schools level)
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)

with open(os.path.join(P, f"{tag}_summary.md"), "w") as fh:
    fh.write(f"# ncu summary {tag} — `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (cfg2, nParticles=1e5, MULTINOMIAL)\n\n")
    fh.write("Captured with `--clock-control none`; launch-list durations are cold-cache and serialised (compare shares).\n\n")
    fh.write("## Share of the launch list by kernel\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, v in share.items():
        fh.write(f"| {k} | {v['launches']} | {v['total_us']} | {v['share']:.3f} |\n")
    fh.write("\n## `--set full` metrics (one row per captured launch)\n\n")
    cols = ["kernel", "grid", "duration", "dram_read", "dram_write", "dram_pct", "fp64_pipe_pct", "issue_pct", "occupancy_pct", "regs", "lanes_per_inst", "warp_insts", "tensor_pct"]
    fh.write("| " + " | ".join(cols) + " |\n|" + "---|" * len(cols) + "\n")
    for d in kern:
        fh.write("| " + " | ".join(str(d.get(c, "")) for c in cols) + " |\n")
    fh.write("\n`profiles/traffic.json` = dram_read + dram_write of the largest launch of each kernel class (bench.py reads it for `roofline.traffic`).\n")
print(json.dumps(share, indent=1))
print(json.dumps(traffic))

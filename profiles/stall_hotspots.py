#!/usr/bin/env python
"""Where does a kernel wait? Per-SASS-instruction warp-stall samples of one kernel from an ncu report captured with
`--set full --import-source on` (`ncu -i <rep> --page source --csv`).
usage: python profiles/stall_hotspots.py <ncu-rep> [top=12]
Prints the share of every stall reason over the whole kernel and, for the dominant reasons, the instructions that
collect most of their samples (a long-scoreboard sample lands on the first instruction that NEEDS the loaded value)."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw))
hi = next(i for i, r in enumerate(rows) if "Address" in r)
hdr = rows[hi]
name = next((r[1] for r in rows[:hi] if len(r) > 1 and r[0] == "Kernel Name"), "")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
isamp, isrc, iexe = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
ins = []
for n, r in enumerate(rows[hi + 1:]):
    if len(r) <= isamp or not r[0].startswith("0x"):
        continue
    ins.append((n, r[isrc].strip(), int(r[isamp] or 0), int(r[iexe] or 0), {h: int(r[i] or 0) for i, h in stall_cols}))
total = sum(x[2] for x in ins)
by = collections.Counter()
for x in ins:
    for h, v in x[4].items():
        by[h] += v
print(f"{rep}: {len(ins)} SASS instructions, {total} warp-state samples, {sum(x[3] for x in ins)} warp instructions executed")
print("| stall reason | share of samples |\n|---|---|")
for h, v in by.most_common():
    if v / total >= 0.01:
        print(f"| {h[6:]} | {100 * v / total:.1f} % |")
for h, v in by.most_common(4):
    if h in ("stall_selected", "stall_not_selected"):
        continue
    print(f"\n{h[6:]} — instructions holding most of its samples (index in the kernel's SASS, share of ALL samples):")
    for x in sorted(ins, key=lambda x: -x[4][h])[:top]:
        if x[4][h] / total < 0.003:
            break
        print(f"  #{x[0]:<5d} {100 * x[4][h] / total:5.1f} %  {x[1][:80]}")

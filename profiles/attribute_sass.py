#!/usr/bin/env python
"""Attribute the executed warp instructions of one kernel (ncu --import-source report) to source regions.
usage: python profiles/attribute_sass.py <mangled kernel name> <ncu-rep> <kernel regex> [launch index]
Joins ncu's per-SASS-instruction counts (`--page source --csv`) with `nvdisasm -gi` line info of the
cubin inside csrc/libdcsmc.so (innermost inlined location per instruction)."""
import collections, csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fn, rep, rx = sys.argv[1], sys.argv[2], sys.argv[3]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "divide-and-conquer-smc_b200", "csrc", "libdcsmc.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith(".text." + fn + ":")][0]
group, last, ins = [], None, []
for l in dis[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        group.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        if group:
            last, group = list(group), []
        ins.append((int(m.group(1), 16), m.group(2).strip(), last))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout.splitlines()
heads = [i for i, l in enumerate(raw) if l.startswith('"Kernel Name"')] + [len(raw)]
rows = list(csv.reader(raw[heads[which]:heads[which + 1]]))
hdr = rows[1]
ia, ie, isrc, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Thread Instructions Executed")
prof = [(int(r[ia], 16), r[isrc].strip(), int(r[ie]), int(r[it])) for r in rows[2:] if len(r) > ie and r[ia].startswith("0x")]
base = prof[0][0]
byoff = {o: (s, c) for o, s, c in ins}

# function line ranges of csrc/dcsmc_math.cuh, read from the file itself
math = open(os.path.join(ROOT, "divide-and-conquer-smc_b200", "csrc", "dcsmc_math.cuh")).read().splitlines()
marks = []
for i, l in enumerate(math, 1):
    m = re.match(r"(?:DC_HD\s+)?(?:static\s+)?(?:double|void|uint64_t|int64_t|int32_t|uint32_t|PhiloxKeys)\s+(\w+)\(", l)
    if m:
        marks.append((i, m.group(1)))
def region(f, ln):
    if f == "dcsmc_math.cuh":
        name = "dcsmc_math.cuh:?"
        for i, n in marks:
            if i <= ln:
                name = n
        return name
    return f"{f}:{ln}"
agg, fp64, thr, tot = collections.Counter(), collections.Counter(), collections.Counter(), 0
for a, s, c, tc in prof:
    ch = byoff.get(a - base, (None, None))[1] or [("?", 0)]
    r = region(*ch[0])
    agg[r] += c
    thr[r] += tc
    tot += c
    op = s.split()[1] if s.startswith("@") else s.split()[0]
    if op.startswith(("DMUL", "DADD", "DFMA", "DSETP", "MUFU")):
        fp64[r] += c
print(f"kernel {fn}: {tot} warp instructions executed")
print("| source region (innermost inlined function, or file:line) | share of executed warp instructions | of which fp64 | active lanes |")
print("|---|---|---|---|")
top = int(os.environ.get("ATTR_TOP", "14"))
for k, v in agg.most_common(top):
    print(f"| {k} | {v / tot * 100:.1f} % | {fp64[k] / tot * 100:.1f} % | {thr[k] / max(1, v):.1f} |")

#!/usr/bin/env python
"""Per-particle ncu counters of the kernel classes, for bench.py's `roofline.traffic` and `roofline.fp64`.

usage: python profiles/per_particle.py <metrics csv> <nParticles> [out json]

Input: `ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,
dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file <csv>
python bench.py --config cfg2 --sub none --steps 1 --warmup 1 --no-cpu-baseline` (NYS-shaped tree: 2800 leaves, 738 internal
nodes; every run of the bench is the same launch sequence). The per-launch counters of one class are summed over ALL
captured launches and divided by the particles those launches processed (launch grids identify the level), so the
figures are per particle-node update of that class and carry over to any nParticles of the same tree."""
import collections, csv, json, re, sys

path, N = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else None
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.defaultdict(lambda: collections.defaultdict(float))  # launch id -> metric -> value
name, grid = {}, {}
for r in rows[1:]:
    k = int(r[ix["ID"]])
    name[k] = r[ix["Kernel Name"]]
    grid[k] = int(re.findall(r"\d+", r[ix["Grid Size"]])[0])
    per[k][r[ix["Metric Name"]]] += float(r[ix["Metric Value"]].replace(",", ""))
CLASSES = {"leaf": r"k_leaf_propose", "internal": r"k_internal_propose", "scan": r"k_scan2?\b", "darts": r"k_darts_count|k_dart_partition",
           "segres": r"k_seg_resample", "expand": r"k_expand", "table": r"k_build_inverse_table", "fused": r"k_node_fused",
           "small": r"k_node_small"}
# particles per launch: blocks of the propose kernels cover PROPOSE_THREADS = 256 particles of one node (leaf kernel: a block
# covers LEAF_WARPS * CH particles); rather than decode every grid, count nodes from the tree: one run = 2800 leaves + 738 internal
NODES = {"leaf": 2800, "internal": 738, "scan": 738, "darts": 738, "table": 738, "segres": 738, "expand": 3538}
if any(re.search(CLASSES["segres"], n) for n in name.values()):
    NODES["expand"] = 2800  # partitioned resampling: k_expand only expands the leaves' offspring counts
res = {"_source": f"{path} (ncu --metrics ..., cold-cache serialised launches), nParticles = {N}"}
nRuns = None
for cls, rx in CLASSES.items():
    ids = [k for k in name if re.search(rx, name[k])]
    if not ids or cls not in NODES:
        continue
    # launches of one run of this class: the capture holds an integer number of identical runs
    tot = collections.defaultdict(float)
    for k in ids:
        for m, v in per[k].items():
            tot[m] += v
    if nRuns is None:
        nRuns = len([k for k in name if re.search(r"k_reset_stats", name[k])])
    particles = NODES[cls] * N * max(1, nRuns)
    res[cls] = {
        "launches": len(ids), "runs": nRuns,
        "fp64_warp_inst_per_particle": round(tot["sm__inst_executed_pipe_fp64.sum"] / particles, 4),
        "warp_inst_per_particle": round(tot["smsp__inst_executed.sum"] / particles, 4),
        "dram_bytes_per_particle": round((tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / particles, 3),
        "ns_per_particle_cold": round(tot["gpu__time_duration.sum"] / particles, 5),
    }
print(json.dumps(res, indent=1))
if out:
    json.dump(res, open(out, "w"), indent=1)

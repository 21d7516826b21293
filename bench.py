#!/usr/bin/env python
"""bench.py — particle-node updates/sec of the DC-SMC node step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle restatement)

A "step" is one pass of the hot path over the whole synthetic tree: every node goes through
merge -> propose -> weight -> normalise -> ESS/logZ -> resample (dcsmc_run).
Workload at N=1: BASELINE.json configs[1] — NYS-shaped hierarchical binomial model
(3538 nodes: 1+5+32+700+2800), nParticles=1e5, resampling at every node.
Inputs (~14 GB of particle state per step) are far larger than the 126 MB L2.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "particle-node updates/sec"
UNIT = "updates/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_workload(args):
    import dcsmc_b200 as d
    from dcsmc_b200 import synth

    if args.workload == "nys":
        rows = synth.nys_shaped_rows()
        ds = d.MultiLevelDataset(rows=rows)
        csr = ds.getTree().to_csr()
        nT, nS = ds.leaf_arrays(csr)
        desc = f"cfg2: NYS-shaped hierarchical binomial (nodes={csr.nNodes}, leaves={int((nT > 0).sum())})"
        return dict(csr=csr, nT=nT, nS=nS, leafKind=None, leafObs=None, rows=rows, desc=desc)
    if args.workload == "binary16":
        csr, lk, lo = synth.binary_gauss_tree(args.depth)
        desc = f"cfg4: perfect binary tree depth {args.depth} with observed Gaussian leaves (nodes={csr.nNodes})"
        return dict(csr=csr, nT=None, nS=None, leafKind=lk, leafObs=lo, rows=None, desc=desc)
    raise SystemExit(f"unknown workload {args.workload}")


class ClockSampler:
    """nvidia-smi sampled every 50 ms in the background; started before the warm-up (nvidia-smi takes
    ~1 s to start) and filtered to the timed window by timestamp."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.idx = gpu_index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def window_begin(self):
        self.t0 = time.time()

    def window_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime

        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, allsm = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                c, m = float(f[1]), float(f[2])
            except ValueError:
                continue
            allsm.append(c)
            if self.t0 is not None and not (self.t0 - 0.05 <= ts <= self.t1 + 0.05):
                continue
            sm.append(c)
            mx.append(m)
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_total": len(allsm)}


# per-update algorithmic bytes by kernel class (DESIGN.md "Roofline accounting"):
# node with C children: 40C+184 (+16 multinomial) =
#   propose 40C+56 | scan 8+8 | resample 8+4 (+16 darts) | gather 4+48+48 (fused into the parent's propose)
def algorithmic_bytes(csr, N, multinomial):
    nN = csr.nNodes
    nLeaf = sum(1 for i in range(nN) if csr.nChildren(i) == 0)
    nInt = nN - nLeaf
    edges = nN - 1
    per = {
        "leaf": N * nLeaf * 56,
        "internal": N * (40 * edges + 56 * nInt + 100 * edges),  # reads the children through their ancestors
        "scan": N * nN * 16,
        "resample": N * nN * 12,
        "darts": N * nN * 16 if multinomial else 0,
        "gather": N * 100,  # root only
    }
    total = N * (40 * edges + (184 + (16 if multinomial else 0)) * nN)
    return per, total


def make_ddc(args, wl, local):
    import dcsmc_b200 as d

    scheme = d.ResamplingScheme.MULTINOMIAL if args.scheme == "multinomial" else d.ResamplingScheme.STRATIFIED
    opt = d.DCOptions(nParticles=args.particles, masterRandomSeed=31, resamplingScheme=scheme,
                      maximumDistributionDepth=args.depth_cut)
    if wl["rows"] is not None:
        ds = d.MultiLevelDataset(rows=wl["rows"])
        factory = d.MultiLevelProposalFactory(dataset=ds)
        tree = ds.getTree()
    else:
        tree = d.perfectBinaryTree(args.depth)
        csr = tree.to_csr()
        factory = d.GaussianLeafProposalFactory({n: float(wl["leafObs"][i]) for i, n in enumerate(csr.nodes)})
    return d.DistributedDC.createInstance(opt, factory, tree, device=local), factory


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    wl = build_workload(args)
    csr = wl["csr"]
    N = args.particles
    multinomial = args.scheme == "multinomial"
    ddc, factory = make_ddc(args, wl, local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        ddc.run()
    barrier()

    # ---- timed region 1: device-resident inputs ------------------------------------------------------
    # CUDA events on the engine's own launching stream (dcsmc_timer_begin/end; torch.cuda.Event only
    # sees torch's current stream) bracket the K runs, gaps between runs included; barrier +
    # synchronize on both sides; max over ranks.
    sampler.window_begin()
    dev_ms = 0.0
    launches = updates = alg_bytes = nvlink = 0
    barrier()
    t0 = time.perf_counter()
    ddc.engine.timer_begin()
    for _ in range(args.steps):
        ddc.run()
        lr = ddc.lastRun
        dev_ms += lr["deviceMs"]
        launches += lr["kernelLaunches"]
        updates += lr["particleUpdates"]
        alg_bytes += lr["algorithmicBytes"]
        nvlink += lr["nvlinkBytes"]
    bracket_ms = ddc.engine.timer_end()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.window_end()
    total_ms = max_over_ranks(bracket_ms)
    sum_dev_ms = max_over_ranks(dev_ms)  # per-call device time of the engine (excludes gaps between calls)
    updates_per_step = sum_over_ranks(updates) / args.steps
    alg_bytes_step = sum_over_ranks(alg_bytes) / args.steps
    launches = int(sum_over_ranks(launches))
    nvlink = sum_over_ranks(nvlink) / args.steps
    ms_per_step = total_ms / args.steps
    value = updates_per_step / (ms_per_step * 1e-3)

    # ---- timed region 2: end to end through the public API with HOST buffers ---------------------------
    # every step re-uploads the topology + model from host arrays (set_tree / set_*_model) and reads
    # the per-node statistics and the root population back to the host.
    h2d = d2h = 0
    host_out = (np.empty((6, N)), None, np.empty(N))

    def e2e_step():
        nonlocal h2d, d2h
        ddc.engine.set_tree(ddc.csr)
        factory.configure(ddc.engine, ddc.csr)
        ddc.run()
        stats = ddc.stats
        h2d = ddc.lastRun["h2dBytes"]
        d2h = ddc.lastRun["d2hBytes"]
        if rank == ddc.rootOwner:
            rs, _, rw = ddc.engine.population(0, out=host_out)  # the caller's own host buffers, reused
            d2h += rs.nbytes + rw.nbytes
        return stats

    for _ in range(min(args.warmup, 2)):  # the end-to-end path has its own first-call costs (pinned staging)
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    e2e_value = updates_per_step / (e2e_ms * 1e-3)
    sampler.window_end()  # the clock window covers both timed regions
    clocks = sampler.stop()
    logZ = float(st["logNormEstimate"][0])

    # ---- per-kernel-class timing (one extra profiled step, outside the timed regions) -------------
    roofline = None
    if rank == 0:
        e = ddc.engine
        e.profile_enable(True)
        if world == 1:
            e.run()
        else:
            e.reset_populations()
            e.run_subtrees(ddc.plan.my_tasks(rank))
        prof = e.profile()
        e.profile_enable(False)
        frac_of_tree = 1.0 if world == 1 else e.run_info().particleUpdates / float(csr.nNodes * N)
        per_bytes, total_bytes = algorithmic_bytes(csr, N, multinomial)
        tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
        dom = max(prof.items(), key=lambda kv: kv[1]["ms"])[0]
        peak, peak_src = load_peaks()
        dom_ms = prof[dom]["ms"]
        dom_launches = max(1, prof[dom]["launches"])
        dom_bytes = per_bytes.get(dom, 0) * frac_of_tree
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(dom)
            except Exception:
                traffic = None
        step_gbs = alg_bytes_step / (ms_per_step * 1e-3) / 1e9
        roofline = {
            "bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
            "launch_ms": round(dom_ms / dom_launches, 4), "bytes_per_launch": int(dom_bytes / dom_launches),
            "kernel_share_of_step": round(dom_ms / tot_ms, 4),
            "step": {"achieved": round(step_gbs, 1), "frac": round(step_gbs / (peak * world), 4),
                     "frac_of_nominal_8TBs": round(step_gbs / (8000.0 * world), 4),
                     "bytes_per_update": round(alg_bytes_step / updates_per_step, 1)},
            "by_kernel_ms": {k: round(v["ms"], 3) for k, v in prof.items() if v["launches"]},
        }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, wl)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "nParticles": N, "resamplingScheme": args.scheme,
                   "relativeEssThreshold": "1+1e-10 (resample every node)", "rng": "philox4x32-10",
                   "maximumDistributionDepth": args.depth_cut,
                   "l2_policy": "inputs larger than L2 (>=10 GB of particle state per step)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "nvlink_bytes_per_step": int(nvlink),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "wall_ms_per_step": wall_ms / args.steps, "kernel_ms_per_step": sum_dev_ms / args.steps,
        "logZ_root": logZ,
    }
    if rank == 0:
        print(json.dumps(line))
    ddc.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_sample(args, wl, target_updates):
    """Bounded sample of the same workload: the first district subtrees (depth-2 nodes) until
    ~target_updates particle-node updates at the bench's nParticles."""
    import dcsmc_b200 as d

    N = args.particles
    if wl["rows"] is None:
        csr = wl["csr"]
        return csr, dict(leafKind=wl["leafKind"], leafObs=wl["leafObs"]), min(N, max(100, int(target_updates // csr.nNodes))), "whole tree, reduced nParticles"
    rows = wl["rows"]
    keep, seen, nodes = [], [], 0
    per_leaf_nodes = 1.3
    for r in rows:
        key = tuple(r[:3])
        if key not in seen:
            if nodes * N >= target_updates:
                break
            seen.append(key)
        keep.append(r)
        nodes += per_leaf_nodes
    ds = d.MultiLevelDataset(rows=keep)
    csr = ds.getTree().to_csr()
    nT, nS = ds.leaf_arrays(csr)
    return csr, dict(nTrials=nT, nSuccesses=nS), N, f"first {len(seen)} district subtrees of the same tree ({csr.nNodes} nodes)"


def cpu_baseline(args, wl, steps=1):
    import oracle_lib as o

    cores = os.cpu_count() or 1
    csr, kw, N, what = cpu_sample(args, wl, target_updates=cores * 2.5e7)
    scheme = o.MULTINOMIAL if args.scheme == "multinomial" else o.STRATIFIED
    best = None
    for _ in range(steps):
        t0 = time.perf_counter()
        o.run(csr, N, masterRandomSeed=31, resamplingScheme=scheme, nThreads=cores, maximumDistributionDepth=2,
              rngMode=o.RNG_JAVA, sumMode=o.SUM_JAVA, elemMode=o.ELEM_LIBM, **kw)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": csr.nNodes * N / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{what}, nParticles={N}, java-order sums, libm, nThreadsPerNode={cores}, maximumDistributionDepth=2; "
                      f"{csr.nNodes * N:.3g} updates in {best:.2f} s",
            "note": "C restatement of the Java path (optimistic for Java: no boxing/GC/Hazelcast)"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path. Java cannot run here
    (no JVM, un-vendored jars), so this times the oracle port with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as o

    wl = build_workload(args)
    cores = os.cpu_count() or 1
    csr, kw, N, what = cpu_sample(args, wl, target_updates=cores * 1.2e7)
    scheme = o.MULTINOMIAL if args.scheme == "multinomial" else o.STRATIFIED

    def step():
        t0 = time.perf_counter()
        o.run(csr, N, masterRandomSeed=31, resamplingScheme=scheme, nThreads=cores, maximumDistributionDepth=2,
              rngMode=o.RNG_JAVA, sumMode=o.SUM_JAVA, elemMode=o.ELEM_LIBM, **kw)
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    ms = sum(times) / len(times) * 1e3
    value = csr.nNodes * N / (ms * 1e-3)
    sample = f"{what}, nParticles={N}; {csr.nNodes * N:.3g} updates per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "nParticles": args.particles, "resamplingScheme": args.scheme,
                   "relativeEssThreshold": "1+1e-10 (resample every node)", "rng": "java.util.Random"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="nys", choices=["nys", "binary16"])
    ap.add_argument("--depth", type=int, default=16)
    ap.add_argument("--particles", type=int, default=100_000)
    ap.add_argument("--scheme", default="multinomial", choices=["multinomial", "stratified"],
                    help="DCOptions.resamplingScheme (reference default: MULTINOMIAL)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--depth-cut", type=int, default=2, help="DCOptions.maximumDistributionDepth (subtree sharding cut)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

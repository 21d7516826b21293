#!/usr/bin/env python
"""bench.py — particle-node updates/sec of the DC-SMC node step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle restatement)

A "step" is one pass of the hot path over the whole synthetic tree: every node goes through
merge -> propose -> weight -> normalise -> ESS/logZ -> resample (DistributedDC.start()).

Headline workload (the one BASELINE.json's north_star quotes its targets on): configs[2] = "cfg3" — the NYS-shaped
hierarchical binomial model (3538 nodes: 1+5+32+700+2800) at nParticles = 1e6, subtrees sharded at
maximumDistributionDepth = 2; it fits one B200 (100 GB of populations), so the same workload is the N = 1 line and
the 2/4/8-GPU lines (strong scaling). The same JSON line carries sub-records ("sub") for configs[1] = cfg2
(nParticles = 1e5), configs[3] = cfg4 (binary tree, 2^16 leaves, nParticles = 1e4; internal-node updates only —
observed leaves are constants and cost nothing) and configs[0] = cfg1 (nParticles = 1e3), measured the same way at
the same N. Inputs (>= 10 GB of particle state per step for cfg2-cfg4) are far larger than the 126 MB L2.
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

CONFIGS = {
    # name: workload, nParticles, maximumDistributionDepth (sharding cut), binary-tree depth
    "cfg1": dict(workload="nys", particles=1_000, depth_cut=2, depth=16),
    "cfg2": dict(workload="nys", particles=100_000, depth_cut=2, depth=16),
    "cfg3": dict(workload="nys", particles=1_000_000, depth_cut=2, depth=16),
    "cfg4": dict(workload="binary16", particles=10_000, depth_cut=3, depth=16),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


_WL_CACHE = {}


def build_workload(cfg):
    """Synthetic inputs of one config (cached: the 131 071-node tree takes seconds to build in Python)."""
    import dcsmc_b200 as d
    from dcsmc_b200 import synth

    key = (cfg["workload"], cfg["depth"] if cfg["workload"] == "binary16" else 0)
    if key in _WL_CACHE:
        return _WL_CACHE[key]
    if cfg["workload"] == "nys":
        rows = synth.nys_shaped_rows()
        ds = d.MultiLevelDataset(rows=rows)
        tree = ds.getTree()
        csr = tree.to_csr()
        nT, nS = ds.leaf_arrays(csr)
        desc = f"NYS-shaped hierarchical binomial (nodes={csr.nNodes}, leaves={int((nT > 0).sum())})"
        wl = dict(csr=csr, nT=nT, nS=nS, leafKind=None, leafObs=None, rows=rows, desc=desc, tree=tree, dataset=ds)
    elif cfg["workload"] == "binary16":
        tree = d.perfectBinaryTree(cfg["depth"])
        csr, lk, lo = synth.binary_gauss_tree(cfg["depth"])
        desc = f"perfect binary tree depth {cfg['depth']} with observed Gaussian leaves (nodes={csr.nNodes})"
        wl = dict(csr=csr, nT=None, nS=None, leafKind=lk, leafObs=lo, rows=None, desc=desc, tree=tree, dataset=None)
    else:
        raise SystemExit(f"unknown workload {cfg['workload']}")
    nc = np.diff(np.asarray(wl["csr"].childOffsets))
    wl["nInternal"] = int((nc > 0).sum())
    wl["nLeaf"] = int((nc == 0).sum())
    _WL_CACHE[key] = wl
    return wl


def config_dict(name, cfg, wl, scheme):
    """`config` of the JSON line — the same keys and values in both arms (ours / --impl reference)."""
    return {"workload": f"{name}: {wl['desc']}", "nParticles": cfg["particles"], "resamplingScheme": scheme,
            "relativeEssThreshold": "1+1e-10 (resample every node)", "maximumDistributionDepth": cfg["depth_cut"],
            "counted_updates": "internal nodes only (observed leaves are constants)" if cfg["workload"] == "binary16" else "all nodes",
            "l2_policy": "inputs larger than L2 (>=10 GB of particle state per step)" if cfg["particles"] >= 10_000 else "small (fits L2)"}


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


# per-update algorithmic bytes by kernel class:
# node with C children: 40C+184 (+16 multinomial) =
#   propose 40C+56 | scan 8+8 | resample 8+4 (+16 darts) | gather 4+48+48 (fused into the parent's propose)
def algorithmic_bytes(wl, N, multinomial, internal_only=False):
    csr = wl["csr"]
    nN, nLeaf, nInt = csr.nNodes, wl["nLeaf"], wl["nInternal"]
    if internal_only:
        # observed leaves are constants: no population, no bytes. Only edges between internal nodes carry data.
        edges = nInt - 1
        per = {"leaf": 0, "internal": N * (40 * edges + 56 * nInt + 100 * edges), "scan": N * nInt * 16,
               "resample": N * nInt * 12, "darts": N * nInt * 16 if multinomial else 0, "gather": N * 100}
        return per, N * (40 * edges + (184 + (16 if multinomial else 0)) * nInt)
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


# Bytes the kernels of a class must move IN THIS ENGINE'S LAYOUT (leaves are 2 columns + ancestors,
# the resampled population is never materialised, observed leaves are constants), per step. The §8(d) split above
# charges every child 40 + 100 B whatever it is; these are the compulsory bytes, so a fraction above 1 cannot occur.
def layout_bytes(wl, N, multinomial, internal_only=False):
    csr = wl["csr"]
    off, ch = np.asarray(csr.childOffsets), np.asarray(csr.children)
    nch = off[1:] - off[:-1]
    is_leaf = nch == 0
    nInt = int((~is_leaf).sum())
    leaf_edges = int(is_leaf[ch].sum())  # edges into leaves
    int_edges = len(ch) - leaf_edges     # edges into internal nodes
    if internal_only:
        # fused small-node kernel: children read through their ancestors (4 + 40 B), 48 B of state + 4 B ancestors
        # written; log weights, CDF, darts and offspring counts stay in shared memory
        return {"internal": N * (44 * int_edges + 52 * nInt)}
    nLeaf = int(is_leaf.sum())
    per = {
        "leaf": N * nLeaf * (16 + 4),  # (mean, descObs) + offspring count (multinomial) or ancestor (stratified)
        "internal": N * (20 * leaf_edges + 44 * int_edges + 56 * nInt),
        "scan": N * nInt * 16,  # log weights in, CDF out (leaves: closed form, no scan)
    }
    if multinomial:
        per["darts"] = N * nInt * 8  # grouped darts out
        per["resample"] = N * (nInt * (8 + 8 + 4) + nLeaf * (4 + 4 + 4))  # CDF + darts in, ancestors out; leaves: counts zeroed, read, ancestors out
    else:
        per["resample"] = N * nInt * (8 + 4)
    return per


def make_ddc(cfg, wl, scheme, local):
    import dcsmc_b200 as d

    sch = d.ResamplingScheme.MULTINOMIAL if scheme == "multinomial" else d.ResamplingScheme.STRATIFIED
    opt = d.DCOptions(nParticles=cfg["particles"], masterRandomSeed=31, resamplingScheme=sch,
                      maximumDistributionDepth=cfg["depth_cut"])
    if wl["rows"] is not None:
        factory = d.MultiLevelProposalFactory(dataset=wl["dataset"])
    else:
        csr = wl["csr"]
        factory = d.GaussianLeafProposalFactory({n: float(wl["leafObs"][i]) for i, n in enumerate(csr.nodes)})
    return d.DistributedDC.createInstance(opt, factory, wl["tree"], device=local), factory


class Dist:
    def __init__(self):
        import torch

        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.local}"))
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def _red(self, x, op=None):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if op is None:
            self.dist.all_reduce(t)
        else:
            self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._red(x, None if self.dist is None else self.dist.ReduceOp.MAX)

    def sum(self, x):
        return self._red(x)

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def measure(name, cfg, args, D, sampler=None, detail=True):
    """One config on the D.world GPUs of this job: device-resident timing, end-to-end timing, (detail) per-kernel
    profile. Returns the record (same keys at top level of the JSON line for the headline config)."""
    wl = build_workload(cfg)
    csr = wl["csr"]
    N = cfg["particles"]
    multinomial = args.scheme == "multinomial"
    internal_only = cfg["workload"] == "binary16"
    ddc, factory = make_ddc(cfg, wl, args.scheme, D.local)
    counted_nodes = wl["nInternal"] if internal_only else csr.nNodes
    updates_per_step = float(counted_nodes) * N
    per_bytes, alg_bytes_step = algorithmic_bytes(wl, N, multinomial, internal_only)
    lay_bytes = layout_bytes(wl, N, multinomial, internal_only)

    for _ in range(args.warmup):
        ddc.run()
    D.barrier()

    # ---- timed region 1: device-resident inputs ------------------------------------------------------
    # CUDA events on the engine's own launching stream (dcsmc_timer_begin/end; torch.cuda.Event only
    # sees torch's current stream) bracket the K runs, gaps between runs included; barrier +
    # synchronize on both sides; max over ranks.
    if sampler is not None:
        sampler.window_begin()
    dev_ms = 0.0
    launches = nvlink = 0
    D.barrier()
    t0 = time.perf_counter()
    ddc.engine.timer_begin()
    for _ in range(args.steps):
        ddc.run()
        lr = ddc.lastRun
        dev_ms += lr["deviceMs"]
        launches += lr["kernelLaunches"]
        nvlink += lr["nvlinkBytes"]
    bracket_ms = ddc.engine.timer_end()
    D.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    total_ms = D.max(bracket_ms)
    sum_dev_ms = D.max(dev_ms)  # per-call device time of the engine (excludes gaps between calls)
    launches = int(D.sum(launches))
    nvlink = D.sum(nvlink) / args.steps
    ms_per_step = total_ms / args.steps
    value = updates_per_step / (ms_per_step * 1e-3)

    # ---- timed region 2: end to end through the public API with HOST buffers ---------------------------
    # every step re-sends the topology + model from host arrays (set_tree / set_*_model: the engine uploads
    # what changed) and reads the per-node statistics and the root population back to the host.
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
        if D.rank == ddc.rootOwner:
            rs, _, rw = ddc.engine.population(0, out=host_out)  # the caller's own host buffers, reused
            d2h += rs.nbytes + rw.nbytes
        return stats

    for _ in range(min(args.warmup, 2)):  # the end-to-end path has its own first-call costs (pinned staging)
        e2e_step()
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = e2e_step()
    D.barrier()
    e2e_ms = D.max((time.perf_counter() - t0) * 1e3) / args.steps
    e2e_value = updates_per_step / (e2e_ms * 1e-3)
    h2d, d2h = int(D.sum(h2d)), int(D.sum(d2h))  # bytes copied on all ranks
    if sampler is not None:
        sampler.window_end()  # the clock window covers both timed regions
    logZ = float(st["logNormEstimate"][0])

    rec = {
        "value": value, "unit": UNIT, "n_gpus": D.world, "ms_per_step": ms_per_step,
        "config": config_dict(name, cfg, wl, args.scheme),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
        "gpu_launches": launches, "nvlink_bytes_per_step": int(nvlink),
        "wall_ms_per_step": wall_ms / args.steps, "kernel_ms_per_step": sum_dev_ms / args.steps,
        "updates_per_step": updates_per_step, "logZ_root": logZ,
        "step_roofline": None,
    }
    peak, peak_src = load_peaks()
    step_gbs = alg_bytes_step / (ms_per_step * 1e-3) / 1e9
    rec["step_roofline"] = {"achieved": round(step_gbs, 1), "unit": "GB/s", "frac": round(step_gbs / (peak * D.world), 4),
                            "frac_of_nominal_8TBs": round(step_gbs / (8000.0 * D.world), 4),
                            "bytes_per_update": round(alg_bytes_step / updates_per_step, 1)}

    # ---- per-kernel-class timing (one extra profiled step, outside the timed regions) -------------
    if detail and D.rank == 0:
        e = ddc.engine
        e.profile_enable(True)
        if D.world == 1:
            e.run()
        else:
            e.reset_populations()
            e.run_subtrees(ddc.plan.my_tasks(D.rank))
        prof = e.profile()
        e.profile_enable(False)
        frac_of_tree = 1.0 if D.world == 1 else e.run_info().particleUpdates / float(csr.nNodes * N)
        tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
        dom = max(prof.items(), key=lambda kv: kv[1]["ms"])[0]
        dom_ms = prof[dom]["ms"]
        dom_launches = max(1, prof[dom]["launches"])
        dom_bytes = per_bytes.get(dom, 0) * frac_of_tree
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        ncu = {}
        tpath = os.path.join(ROOT, "profiles", "ncu_per_particle.json")
        if os.path.exists(tpath):
            try:
                ncu = json.load(open(tpath))
            except Exception:
                ncu = {}
        # particles the dominant class processed in the profiled step
        dom_nodes = {"leaf": wl["nLeaf"], "internal": wl["nInternal"]}.get(dom, csr.nNodes) * frac_of_tree
        dom_particles = dom_nodes * N
        k = ncu.get(dom, {})
        traffic = int(k["dram_bytes_per_particle"] * dom_particles / dom_launches) if "dram_bytes_per_particle" in k else None
        rec["roofline"] = {
            "bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
            "launch_ms": round(dom_ms / dom_launches, 4), "bytes_per_launch": int(dom_bytes / dom_launches),
            "kernel_share_of_step": round(dom_ms / tot_ms, 4),
            "step": rec["step_roofline"],
            "by_kernel_ms": {k2: round(v["ms"], 3) for k2, v in prof.items() if v["launches"]},
            # fraction of the measured HBM peak per kernel class: bytes the class must move in this engine's layout
            # (layout_bytes) / its time; "_8d" = the same with SURVEY §8(d)'s split, which charges 140 B per child
            # whatever its footprint (a leaf child is 20 B here) and so can exceed 1
            "frac_layout": round(lay_bytes.get(dom, 0) * frac_of_tree / (dom_ms * 1e-3) / 1e9 / peak, 4) if dom_ms > 0 else None,
            "by_kernel_hbm_frac": {k2: round(lay_bytes.get(k2, 0) * frac_of_tree / (v["ms"] * 1e-3) / 1e9 / peak, 4)
                                   for k2, v in prof.items() if v["launches"] and v["ms"] > 0 and lay_bytes.get(k2, 0)},
            "by_kernel_hbm_frac_8d": {k2: round(per_bytes.get(k2, 0) * frac_of_tree / (v["ms"] * 1e-3) / 1e9 / peak, 4)
                                      for k2, v in prof.items() if v["launches"] and v["ms"] > 0 and per_bytes.get(k2, 0)
                                      and k2 in lay_bytes},
        }
        # the bound that actually applies to the two propose kernels: the fp64 pipe (2 warp instructions per clock
        # per SM: ncu sm__sass_thread_inst_executed_op_dfma_pred_on.avg.peak_sustained = 64 lanes/clk/SM).
        # Executed fp64 warp instructions per particle come from the committed ncu capture of the same kernels
        # (profiles/ncu_per_particle.json); time from the CUDA events above.
        fp64 = {}
        for cls, nodes in (("leaf", wl["nLeaf"]), ("internal", wl["nInternal"])):
            kk = ncu.get(cls, {})
            if "fp64_warp_inst_per_particle" in kk and prof.get(cls, {}).get("ms", 0) > 0:
                inst = kk["fp64_warp_inst_per_particle"] * nodes * frac_of_tree * N
                fp64[cls] = {"fp64_warp_inst": inst, "ms": round(prof[cls]["ms"], 3)}
        rec["roofline"]["fp64_pending"] = fp64  # completed with the clock once the sampler has stopped
    ddc.close()
    return rec


def finish_fp64(rec, clocks):
    rl = rec.get("roofline")
    if not rl or "fp64_pending" not in rl:
        return
    pend = rl.pop("fp64_pending")
    mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
    peak = 2.0 * 148 * mhz * 1e6  # warp instructions / s
    out = {}
    for cls, v in pend.items():
        ach = v["fp64_warp_inst"] / (v["ms"] * 1e-3)
        out[cls] = {"achieved": ach, "peak": peak, "unit": "fp64 warp-inst/s", "frac": round(ach / peak, 4), "sm_mhz": mhz}
    if out:
        rl["fp64"] = out


def resolve_config(args):
    """--config picks a BASELINE config; --workload / --particles / --depth-cut / --depth override it ("custom")."""
    cfg = dict(CONFIGS[args.config])
    custom = False
    if args.workload:
        cfg["workload"], custom = args.workload, True
    if args.particles:
        cfg["particles"], custom = args.particles, True
    if args.depth_cut is not None:
        cfg["depth_cut"], custom = args.depth_cut, True
    if args.depth != 16:
        cfg["depth"], custom = args.depth, True
    return ("custom" if custom else args.config), cfg


def run_ours(args):
    D = Dist()
    name, cfg = resolve_config(args)
    sampler = ClockSampler(D.local)
    sampler.start()
    rec = measure(name, cfg, args, D, sampler=sampler, detail=True)
    clocks = sampler.stop()
    finish_fp64(rec, clocks)

    subs = {}
    want = [] if args.sub == "none" else [s for s in args.sub.split(",") if s]
    if name == "custom":
        want = []
    for s in want:
        if s == name or s not in CONFIGS:
            continue
        r = measure(s, dict(CONFIGS[s]), args, D, sampler=None, detail=(D.world == 1))
        if "roofline" in r:
            r["roofline"].pop("fp64_pending", None)
        subs[s] = r

    cpu = cpu1 = None
    if D.rank == 0 and D.world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(cfg, build_workload(cfg), args.scheme)
        cpu1 = cpu_baseline_impl1()

    line = {
        "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": D.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "rng": "philox4x32-10",
        "config": rec["config"], "clocks": clocks, "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"],
        "nvlink_bytes_per_step": rec["nvlink_bytes_per_step"], "roofline": rec.get("roofline"),
        "cpu_baseline": cpu, "cpu_baseline_cfg1_impl1": cpu1,
        "wall_ms_per_step": rec["wall_ms_per_step"], "kernel_ms_per_step": rec["kernel_ms_per_step"],
        "logZ_root": rec["logZ_root"], "sub": subs,
    }
    if D.rank == 0:
        print(json.dumps(line))
    D.close()


# ---- the reference's CPU path (oracle port) -------------------------------------------------------------------
def cpu_sample(cfg, wl, target_updates):
    """Bounded sample of the same workload: the first district subtrees (depth-2 nodes) with everything above them,
    until ~target_updates particle-node updates at the config's nParticles."""
    import dcsmc_b200 as d

    N = cfg["particles"]
    if wl["rows"] is None:
        csr = wl["csr"]
        n = min(N, max(100, int(target_updates // csr.nNodes)))
        return csr, dict(leafKind=wl["leafKind"], leafObs=wl["leafObs"]), n, wl["nInternal"], "whole tree, reduced nParticles"
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
    what = (f"first {len(seen)} district subtrees of the same tree with their counties and the root ({csr.nNodes} nodes)"
            if len(keep) < len(rows) else f"the whole tree ({csr.nNodes} nodes)")
    return csr, dict(nTrials=nT, nSuccesses=nS), N, csr.nNodes, what


def cpu_run(csr, N, kw, scheme, cores):
    """One execution of the reference's CPU path: java.util.Random per node, sequential sums, libm, the reference's
    executor (a node becomes a task when its last child is done, nThreadsPerNode = cores). The task depth does not
    change results; 3 (school subtrees on the NYS tree) is the best-balanced setting for the reference here."""
    import oracle_lib as o

    sch = o.MULTINOMIAL if scheme == "multinomial" else o.STRATIFIED
    depth = 3 if "nTrials" in kw else 6
    t0 = time.perf_counter()
    o.run(csr, N, masterRandomSeed=31, resamplingScheme=sch, nThreads=cores, maximumDistributionDepth=depth,
          rngMode=o.RNG_JAVA, sumMode=o.SUM_JAVA, elemMode=o.ELEM_LIBM, **kw)
    return time.perf_counter() - t0


def cpu_baseline(cfg, wl, scheme, target_s=12.0):
    cores = os.cpu_count() or 1
    csr, kw, N, counted, what = cpu_sample(cfg, wl, target_updates=cores * 3.4e6 * target_s)
    dt = cpu_run(csr, N, kw, scheme, cores)
    return {"value": counted * N / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{what}, nParticles={N}, java-order sums, libm, nThreadsPerNode={cores}, task depth 3; "
                      f"{counted * N:.3g} updates in {dt:.2f} s",
            "note": "C restatement of the Java path (optimistic for Java: no boxing/GC/Hazelcast)"}


def cpu_baseline_impl1():
    """BASELINE.md §4.4 / configs[0]: impl-1 (prototype.smc.DivideConquerMCAlgorithm.recurse :341-431) — one global
    Random(1), single thread, always resample — on the NYS-shaped tree at nParticles = 1000."""
    import oracle_lib as o

    wl = build_workload(CONFIGS["cfg1"])
    csr, N = wl["csr"], 1000
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        o.run(csr, N, masterRandomSeed=1, impl=1, nThreads=1, rngMode=o.RNG_JAVA, sumMode=o.SUM_JAVA, elemMode=o.ELEM_LIBM,
              nTrials=wl["nT"], nSuccesses=wl["nS"])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": csr.nNodes * N / best, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"cfg1 whole tree ({csr.nNodes} nodes), nParticles=1000, impl-1 accounting, single thread; best of 3: {best:.3f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path. Java cannot run here (no JVM, un-vendored
    jars), so this times the oracle port with all host threads on a bounded sample of the same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name, cfg = resolve_config(args)
    wl = build_workload(cfg)
    cores = os.cpu_count() or 1
    csr, kw, N, counted, what = cpu_sample(cfg, wl, target_updates=cores * 3.4e6 * args.ref_step_seconds)
    for _ in range(args.warmup):
        cpu_run(csr, N, kw, args.scheme, cores)
    times = [cpu_run(csr, N, kw, args.scheme, cores) for _ in range(args.steps)]
    ms = sum(times) / len(times) * 1e3
    value = counted * N / (ms * 1e-3)
    sample = (f"{what}, nParticles={N}, java-order sums, libm, nThreadsPerNode={cores}, task depth 3; "
              f"{counted * N:.3g} updates per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "rng": "java.util.Random",
        "config": config_dict(name, cfg, wl, args.scheme),
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
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS), help="BASELINE.json config (default: cfg3, nParticles=1e6)")
    ap.add_argument("--sub", default="cfg2,cfg4,cfg1", help="sub-records measured after the headline config ('none' to skip)")
    ap.add_argument("--workload", default=None, choices=["nys", "binary16"], help="override: custom workload")
    ap.add_argument("--depth", type=int, default=16)
    ap.add_argument("--particles", type=int, default=None, help="override: custom nParticles")
    ap.add_argument("--scheme", default="multinomial", choices=["multinomial", "stratified"],
                    help="DCOptions.resamplingScheme (reference default: MULTINOMIAL)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--depth-cut", type=int, default=None, help="override: DCOptions.maximumDistributionDepth (subtree sharding cut)")
    ap.add_argument("--ref-step-seconds", type=float, default=8.0, help="--impl reference: CPU seconds one step's sample is sized for")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

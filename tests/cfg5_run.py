"""torchrun script (not a pytest test): BASELINE config 5 on the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/cfg5_run.py [--fanout 10,100,100]
                                     [--particles 10000000] [--budget-gb 110] [--out gpurun_out/cfg5.json]

Wide synthetic hierarchy root -> 10 -> 100 each -> 100 each (1e5 binomial leaves, 1011 internal nodes), nParticles = 1e7 per
node, ESS-triggered resampling (relativeEssThreshold = 0.5: children may arrive weighted, DCRecursion.java:29-31,59), subtrees
below maximumDistributionDepth = 2 sharded over the ranks (1000 tasks), the two levels above merged over NVLink. 20 TB of
populations do not fit: every rank's static memory plan cuts its share into subtree batches inside `--budget-gb`.

The CPU oracle cannot reach this size (1e12 particle-node updates); what is checked here, bit for bit:
  * both resampling schemes run; every node reports done, 0 < ESS <= N, leaves have logZ = 0 exactly;
  * determinism across the schedule: every rank recomputes two task subtrees that ANOTHER rank owned (dcsmc_run_subtrees on
    its own engine) and compares all their node statistics with the sharded run's;
  * the merge above the cut: rank 0 recomputes the whole subtree of one depth-1 node (100 tasks, ~1e11 updates) on one GPU and
    compares the depth-1 node's statistics (weighted and resampled children, remote children read over NVLink) with the sharded run's.
The cfg5 SHAPE at nParticles = 1e7 is compared with the oracle in tests/test_gpu_fullsize.py.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as dist

import dcsmc_b200 as d
from dcsmc_b200 import synth

KEYS = ("ess", "relEss", "logNormEstimate", "logScaling", "resampled")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fanout", default="10,100,100")
    ap.add_argument("--particles", type=int, default=10_000_000)
    ap.add_argument("--budget-gb", type=float, default=110.0)
    ap.add_argument("--threshold", type=float, default=0.5)
    ap.add_argument("--runs", type=int, default=2)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cfg5.json"))
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    fan = [int(x) for x in args.fanout.split(",")]
    t0 = time.perf_counter()
    ds = d.MultiLevelDataset(rows=synth.wide_rows(5, fan))
    tree = ds.getTree()
    N = args.particles
    record = {"workload": f"cfg5: root -> {' -> '.join(str(f) for f in fan)} binomial leaves", "nParticles": N, "n_gpus": world,
              "relativeEssThreshold": args.threshold, "maximumDistributionDepth": 2, "budget_gb": args.budget_gb, "schemes": {}}
    for scheme in (d.ResamplingScheme.MULTINOMIAL, d.ResamplingScheme.STRATIFIED):
        opt = d.DCOptions(nParticles=N, resamplingScheme=scheme, relativeEssThreshold=args.threshold, maximumDistributionDepth=2,
                          memoryBudgetBytes=int(args.budget_gb * 1e9))
        ddc = d.DistributedDC.createInstance(opt, d.MultiLevelProposalFactory(dataset=ds), tree, device=local)
        csr = ddc.csr
        if rank == 0 and scheme == d.ResamplingScheme.MULTINOMIAL:
            print(f"tree: {csr.nNodes} nodes built in {time.perf_counter() - t0:.1f} s", flush=True)
        times, devs = [], []
        for r in range(args.runs):
            torch.cuda.synchronize()
            dist.barrier()
            t1 = time.perf_counter()
            ddc.run()
            torch.cuda.synchronize()
            dist.barrier()
            times.append((time.perf_counter() - t1) * 1e3)
            devs.append(ddc.lastRun["deviceMs"])
        st = {k: v.copy() for k, v in ddc.stats.items()}
        launches = torch.tensor([float(ddc.lastRun["kernelLaunches"]), float(ddc.lastRun["nvlinkBytes"]), devs[-1]], device="cuda", dtype=torch.float64)
        mx = launches.clone()
        dist.all_reduce(launches)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        nc = np.diff(np.asarray(csr.childOffsets))
        leaves = nc == 0
        assert np.all(st["ess"] > 0) and np.all(st["ess"] <= N * (1 + 1e-12)), "ESS out of range"
        assert np.all(st["logNormEstimate"][leaves] == 0.0) and np.all(st["ess"][leaves] == N)
        assert np.isfinite(st["logNormEstimate"]).all()
        internal = ~leaves
        nWeighted = int((st["resampled"][internal] == 0).sum())
        # -- determinism across the schedule: recompute subtrees another rank owned -----------------------------------------
        plan = ddc.plan
        others = [t for t in plan.tasks if plan.task_owner[t] != rank]
        rng = np.random.default_rng(100 + rank)
        picks = [int(x) for x in rng.choice(others, size=min(2, len(others)), replace=False)] if others else []
        sizes = csr.subtree_sizes()
        e = ddc.engine
        checked = 0
        if picks:
            e.reset_populations()
            e.run_subtrees(picks)
            mine = e.all_node_stats()
            for t in picks:
                sub = np.arange(t, t + int(sizes[t]))  # pre-order numbering: a subtree is an id range
                for k in KEYS:
                    assert np.array_equal(mine[k][sub], st[k][sub]), (int(scheme), t, k)
                checked += len(sub)
        # -- the merge above the cut: one depth-1 node's whole subtree on one GPU -------------------------------------------
        top_ok = None
        if rank == 0 and len(fan) >= 3:
            def remote_children(n):
                return sum(1 for c in csr.children[csr.childOffsets[n]:csr.childOffsets[n + 1]] if plan.node_owner[int(c)] != plan.node_owner[n])

            # the depth-1 node whose children are spread over the most ranks (its merge read them over NVLink)
            node1 = max((int(c) for c in csr.children[csr.childOffsets[0]:csr.childOffsets[1]]), key=lambda n: (remote_children(n), -n))
            t2 = time.perf_counter()
            e.reset_populations()
            e.run_subtrees([node1])
            mine = e.all_node_stats()
            for k in KEYS:
                assert mine[k][node1] == st[k][node1], (int(scheme), "depth-1 node", k, mine[k][node1], st[k][node1])
            top_ok = {"node": node1, "nodes_recomputed": int(sizes[node1]), "seconds": round(time.perf_counter() - t2, 2),
                      "children_owned_elsewhere": int(sum(1 for c in csr.children[csr.childOffsets[node1]:csr.childOffsets[node1 + 1]]
                                                          if plan.node_owner[int(c)] != plan.node_owner[node1]))}
        tot = torch.tensor([float(checked)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tot)
        ms = min(times[1:]) if len(times) > 1 else times[0]
        rec = {"ms_per_run": ms, "first_run_ms": times[0], "device_ms_slowest_rank": float(mx[2].item()),
               "updates_per_s": csr.nNodes * N / (ms * 1e-3), "logZ_root": float(st["logNormEstimate"][0]), "ess_root": float(st["ess"][0]),
               "internal_nodes_left_weighted": nWeighted, "internal_nodes": int(internal.sum()),
               "kernel_launches_all_ranks": int(launches[0].item()), "nvlink_bytes_per_run": int(launches[1].item()),
               "nodes_recomputed_on_another_rank_bit_identical": int(tot.item()), "depth1_subtree_recomputed_on_one_gpu": top_ok,
               "transport": ddc._transportChoice, "tasks": len(plan.tasks), "top_levels": [len(l.nodes) for l in plan.top_levels]}
        record["schemes"]["multinomial" if scheme == d.ResamplingScheme.MULTINOMIAL else "stratified"] = rec
        if rank == 0:
            print(json.dumps(rec), flush=True)
        ddc.close()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as fh:
            json.dump(record, fh, indent=1)
        print("CFG5_RUN_OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

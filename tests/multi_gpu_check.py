"""torchrun script: the sharded multi-GPU run must give bit-identical node statistics and root
population to a single-GPU run and to the oracle (per-node streams keyed by node id, Doc.java:268)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as dist

import dcsmc_b200 as d
from dcsmc_b200 import synth


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    rows = synth.wide_rows(5, [3, 4, 3])
    ds = d.MultiLevelDataset(rows=rows)
    N = 20000
    for scheme in (d.ResamplingScheme.MULTINOMIAL, d.ResamplingScheme.STRATIFIED):
        for thr in (1.0 + 1e-10, 0.6):
            for depth in (1, 2, 2 ** 31 - 1):
                if os.environ.get("DCSMC_CHECK_VERBOSE"):
                    print(f"[rank {rank}] scheme={int(scheme)} thr={thr} depth={depth}", flush=True)
                opt = d.DCOptions(nParticles=N, resamplingScheme=scheme, relativeEssThreshold=thr, maximumDistributionDepth=depth)
                ddc = d.DistributedDC.createInstance(opt, d.MultiLevelProposalFactory(dataset=ds), ds.getTree(), device=local)
                ddc.start()
                sharded = {k: v.copy() for k, v in ddc.stats.items()}
                moved = ddc.lastRun["nvlinkBytes"]
                root = None
                if rank == ddc.rootOwner:
                    root = ddc.getRootPopulation().particles.copy()
                # runs 2..4 replay the static schedule without host round trips (deferred calls, statistics exchanged
                # on the device, cached + graphed plans above the cut): same bits as the first, host-synchronised run
                def same_as_first(tag):
                    again = ddc.stats
                    for k in ("ess", "relEss", "logNormEstimate", "logScaling", "resampled"):
                        assert np.array_equal(sharded[k], again[k]), (scheme, thr, depth, k, tag)
                    if root is not None:
                        s_, _, _ = ddc.engine.population(0)
                        assert np.array_equal(root, s_, equal_nan=True), (scheme, thr, depth, tag)

                for rep in range(3):
                    ddc.run()
                    same_as_first(f"replay {rep}")
                if os.environ.get("DCSMC_TRANSPORT", "p2p") == "p2p" and ddc._transportChoice == "p2p":
                    assert getattr(ddc, "_p2pStatic", None) is not None and ddc._p2pBufs is not None, "the fast path did not run"
                # single-GPU run on every rank (no process group use): engine.run()
                ddc.engine.run()
                single = ddc.engine.all_node_stats()
                for k in ("ess", "relEss", "logNormEstimate", "logScaling", "resampled"):
                    assert np.array_equal(sharded[k], single[k]), (scheme, thr, depth, k)
                if root is not None:
                    s, _, _ = ddc.engine.population(0)
                    assert np.array_equal(root, s, equal_nan=True)
                # and the sharded schedule again after another plan used the engine (descriptor invalidation)
                ddc.run()
                same_as_first("after the single-GPU run")
                tot = torch.tensor([moved], device="cuda", dtype=torch.float64)
                dist.all_reduce(tot)
                assert tot.item() > 0, "nothing crossed NVLink"
                transport = ddc._transportChoice
                ddc.close()
    # posterior summaries (levelCutOffForOutput): mean/var statistics come from the node's owner, delta
    # statistics from the owner of its parent, which reads the children over NVLink
    opt = d.DCOptions(nParticles=N, maximumDistributionDepth=2, levelCutOffForOutput=2)
    ddc = d.DistributedDC.createInstance(opt, d.MultiLevelProposalFactory(dataset=ds), ds.getTree(), device=local)
    ddc.start()
    csr = ds.getTree().to_csr()
    got = {i: ddc.nodeSummary(i) for i in range(csr.nNodes)}
    ddc.engine.run()
    depth = csr.depth()
    nDelta = 0
    for i in range(csr.nNodes):
        s1 = ddc.engine.node_summary(i)
        if depth[i] > 2:
            continue
        g = got[i]
        assert (g["hasMean"], g["hasVar"], g["hasDelta"]) == (s1.hasMean, s1.hasVar, s1.hasDelta), (i, g)
        if g["hasMean"]:
            assert g["meanMean"] == s1.meanMean and g["meanSD"] == s1.meanSD, (i, g)
        if g["hasVar"]:
            assert g["varMean"] == s1.varMean and g["varSD"] == s1.varSD, (i, g)
        if g["hasDelta"]:
            assert g["deltaMean"] == s1.deltaMean and g["deltaSD"] == s1.deltaSD, (i, g)
            nDelta += 1
    assert nDelta >= 2
    ddc.close()
    # the reference's own example model (Doc.java:162-237) through the same sharded path
    from test_gpu_dc_api import transition, prior

    tree = d.perfectBinaryTree(4)
    for thr in (1.0 + 1e-10, 0.5):
        opt = d.DCOptions(nParticles=N, relativeEssThreshold=thr, maximumDistributionDepth=2)
        ddc = d.DistributedDC.createInstance(opt, d.MarkovChainNaiveProposalFactory(transition, prior), tree, device=local)
        ddc.start()
        sharded_m = {k: v.copy() for k, v in ddc.stats.items()}
        root = ddc.getRootPopulation().particles.copy() if rank == ddc.rootOwner else None
        ddc.engine.run()
        single = ddc.engine.all_node_stats()
        for k in ("ess", "relEss", "logNormEstimate", "logScaling", "resampled"):
            assert np.array_equal(sharded_m[k], single[k]), ("markov", thr, k)
        if root is not None:
            _, ist, _ = ddc.engine.population(0, state=False, istate=True)
            assert np.array_equal(root, ist)
        ddc.close()
    if rank == 0:
        import oracle_lib as o

        csr = ds.getTree().to_csr()
        nT, nS = ds.leaf_arrays(csr)
        ref = o.run(csr, N, resamplingScheme=o.STRATIFIED, relativeEssThreshold=0.6, rngMode=o.RNG_PHILOX,
                    sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS)
        assert np.array_equal(ref.logNormEstimate, sharded["logNormEstimate"])
        print("MULTI_GPU_CHECK_OK world", world, "transport", transport)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

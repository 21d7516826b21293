"""Scratch helper (not a test), under torchrun: host-side breakdown of one end-to-end step of a sharded run on the
depth-`D` binary tree (cfg4): set_tree / configure / run / stats / root population, per rank."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch, torch.distributed as dist
import dcsmc_b200 as d
from dcsmc_b200 import synth
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank = dist.get_rank()
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
tree = d.perfectBinaryTree(depth)
csr, lk, lo = synth.binary_gauss_tree(depth)
f = d.GaussianLeafProposalFactory({n: float(lo[i]) for i, n in enumerate(csr.nodes)})
ddc = d.DistributedDC.createInstance(d.DCOptions(nParticles=N, maximumDistributionDepth=3), f, tree, device=local)
for _ in range(3): ddc.run()
host = (np.empty((6, N)), None, np.empty(N))
for it in range(5):
    torch.cuda.synchronize(); dist.barrier()
    t = [time.perf_counter()]
    ddc.engine.set_tree(ddc.csr); t.append(time.perf_counter())
    f.configure(ddc.engine, ddc.csr); t.append(time.perf_counter())
    ddc.run(); t.append(time.perf_counter())
    st = ddc.stats; t.append(time.perf_counter())
    if rank == ddc.rootOwner: ddc.engine.population(0, out=host)
    t.append(time.perf_counter())
    if it >= 2:
        print("rank %d: set_tree %.2f configure %.2f run %.2f (device %.2f) stats %.2f root %.2f ms; h2d %d" % tuple(
            [rank] + [(t[i + 1] - t[i]) * 1e3 for i in range(3)] + [ddc.lastRun["deviceMs"]] + [(t[4] - t[3]) * 1e3, (t[5] - t[4]) * 1e3, ddc.lastRun["h2dBytes"]]), flush=True)
dist.barrier(); ddc.close(); dist.destroy_process_group()

"""Scratch helper (not a test): wall-clock breakdown of one end-to-end step on the depth-D binary tree (cfg4), one GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import dcsmc_b200 as d
from dcsmc_b200 import synth
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
tree = d.perfectBinaryTree(depth)
csr, lk, lo = synth.binary_gauss_tree(depth)
f = d.GaussianLeafProposalFactory({n: float(lo[i]) for i, n in enumerate(csr.nodes)})
ddc = d.DistributedDC.createInstance(d.DCOptions(nParticles=N, maximumDistributionDepth=3), f, tree, device=0)
for _ in range(3): ddc.run()
host = (np.empty((6, N)), None, np.empty(N))
for it in range(5):
    t = [time.perf_counter()]
    ddc.engine.set_tree(ddc.csr); t.append(time.perf_counter())
    f.configure(ddc.engine, ddc.csr); t.append(time.perf_counter())
    ddc.run(); t.append(time.perf_counter())
    st = ddc.stats; t.append(time.perf_counter())
    ddc.engine.population(0, out=host); t.append(time.perf_counter())
    print("set_tree %.3f configure %.3f run %.3f (device %.3f) stats %.3f root %.3f ms total %.3f" % tuple(
        [(t[i + 1] - t[i]) * 1e3 for i in range(3)] + [ddc.lastRun["deviceMs"]] + [(t[4] - t[3]) * 1e3, (t[5] - t[4]) * 1e3, (t[5] - t[0]) * 1e3]), flush=True)
ddc.close()

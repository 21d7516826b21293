"""Scratch helper (not a test): wall time per run with and without CUDA-graph replay."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dcsmc_b200 as d
from dcsmc_b200 import synth
from dcsmc_b200.engine import Engine
N = int(sys.argv[1])
ds = d.MultiLevelDataset(rows=synth.nys_shaped_rows())
csr = ds.getTree().to_csr()
nT, nS = ds.leaf_arrays(csr)
with Engine(nParticles=N, resamplingScheme=0) as e:
    e.set_tree(csr); e.set_multilevel_model(nT, nS)
    for _ in range(5): e.run()
    e.timer_begin(); t0 = time.perf_counter()
    K = 20
    for _ in range(K): e.run()
    ms = e.timer_end(); wall = (time.perf_counter() - t0) * 1e3
    print("graphs", os.environ.get("DCSMC_GRAPHS", "1"), "N", N, "bracket ms/run", round(ms / K, 4), "wall", round(wall / K, 4), "device", round(e.run_info().deviceMs, 4), "launches", e.run_info().kernelLaunches, "logZ", e.node_stats(0).logNormEstimate)

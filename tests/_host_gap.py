"""Scratch helper (not a test): where the host time of one dcsmc_run goes on the 131 071-node tree (cfg4)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dcsmc_b200 as d
from dcsmc_b200 import synth
from dcsmc_b200.engine import Engine
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
csr, lk, lo = synth.binary_gauss_tree(depth)
with Engine(nParticles=N) as e:
    e.set_tree(csr); e.set_multilevel_model(leafKind=lk, leafObs=lo)
    for _ in range(4): e.run()
    K = 10
    e.timer_begin(); t0 = time.perf_counter(); dev = 0.0
    for _ in range(K):
        e.run(); dev += e.run_info().deviceMs
    ms = e.timer_end(); wall = (time.perf_counter() - t0) * 1e3
    print("bracket ms/run", round(ms / K, 3), "wall", round(wall / K, 3), "device", round(dev / K, 3), "launches", e.run_info().kernelLaunches)
    t0 = time.perf_counter(); st = e.all_node_stats(); print("all_node_stats ms", round((time.perf_counter() - t0) * 1e3, 3))

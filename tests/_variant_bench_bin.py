"""Scratch helper (not a test): time one engine variant library on the cfg4 workload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcsmc_b200 import _ffi, synth
_ffi.LIB_PATH = os.path.join(_ffi.HERE, "csrc", sys.argv[1])
from dcsmc_b200.engine import Engine
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 16
N = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
csr, lk, lo = synth.binary_gauss_tree(depth)
with Engine(nParticles=N, resamplingScheme=0) as e:
    e.set_tree(csr); e.set_multilevel_model(None, None, lk, lo)
    for _ in range(3): e.run()
    ms = []
    for _ in range(5):
        e.run(); ms.append(e.run_info().deviceMs)
    e.profile_enable(True); e.run(); p = e.profile()
    print(sys.argv[1], "N", N, "ms/step", round(sum(ms) / len(ms), 3), {k: round(v["ms"], 3) for k, v in p.items() if v["launches"]}, "logZ", e.node_stats(0).logNormEstimate)

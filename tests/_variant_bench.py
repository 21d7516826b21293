"""Scratch helper (not a test): time one engine variant library on the cfg2 workload."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dcsmc_b200 as d
from dcsmc_b200 import _ffi, synth

var = sys.argv[1]
scheme = int(sys.argv[2]) if len(sys.argv) > 2 else 0
N = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
_ffi.LIB_PATH = os.path.join(_ffi.HERE, "csrc", var)
from dcsmc_b200.engine import Engine

ds = d.MultiLevelDataset(rows=synth.nys_shaped_rows())
csr = ds.getTree().to_csr()
nT, nS = ds.leaf_arrays(csr)
with Engine(nParticles=N, resamplingScheme=scheme) as e:
    e.set_tree(csr)
    e.set_multilevel_model(nT, nS)
    for _ in range(3):
        e.run()
    ms = []
    for _ in range(5):
        e.run()
        ms.append(e.run_info().deviceMs)
    e.profile_enable(True)
    e.run()
    p = e.profile()
    print(var, "N", N, "scan_v", os.environ.get("DCSMC_SCAN_V", "-"), "ms/step", round(sum(ms) / len(ms), 3), {k: round(v["ms"], 3) for k, v in p.items() if v["launches"]},
          "logZ", e.node_stats(0).logNormEstimate)

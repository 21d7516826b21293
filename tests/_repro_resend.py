"""Scratch helper (not a test): the tree-change sequence of test_resending_the_same_tree..., with the engine's checkpoints."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import dcsmc_b200 as d
from dcsmc_b200 import synth, _ffi
from dcsmc_b200.engine import Engine
ds = d.MultiLevelDataset(rows=synth.small_multilevel_rows(seed=13)); csr = ds.getTree().to_csr(); nT, nS = ds.leaf_arrays(csr)
ds2 = d.MultiLevelDataset(rows=synth.small_multilevel_rows(seed=14, nGroups=5)); csr2 = ds2.getTree().to_csr(); nT2, nS2 = ds2.leaf_arrays(csr2)
print("nodes", csr.nNodes, csr2.nNodes)
for variant in ("plain", "with_export"):
    with Engine(nParticles=20000) as e:
        for _ in range(3):
            e.set_tree(csr); e.set_multilevel_model(nT, nS); e.run()
        e.set_tree(csr2)
        if variant == "with_export":
            try:
                e.export([0])
            except _ffi.DcsmcError as ex:
                print("export:", ex)
        e.set_multilevel_model(nT2, nS2)
        try:
            e.run(); print(variant, "ok", e.all_node_stats()["logNormEstimate"][0])
        except _ffi.DcsmcError as ex:
            print(variant, "FAILED:", ex)

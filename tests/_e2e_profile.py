"""Scratch helper (not a test): wall-clock breakdown of one end-to-end step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dcsmc_b200 as d
from dcsmc_b200 import synth
ds = d.MultiLevelDataset(rows=synth.nys_shaped_rows())
opt = d.DCOptions(nParticles=100000, maximumDistributionDepth=2)
f = d.MultiLevelProposalFactory(dataset=ds)
ddc = d.DistributedDC.createInstance(opt, f, ds.getTree(), device=0)
for _ in range(3): ddc.run()
for it in range(4):
    t = [time.perf_counter()]
    ddc.engine.set_tree(ddc.csr); t.append(time.perf_counter())
    f.configure(ddc.engine, ddc.csr); t.append(time.perf_counter())
    ddc.run(); t.append(time.perf_counter())
    st = ddc.stats; rs, _, rw = ddc.engine.population(0); t.append(time.perf_counter())
    print("set_tree %.2f configure %.2f run %.2f (device %.2f) fetch %.2f ms" % tuple([(t[i+1]-t[i])*1e3 for i in range(3)] + [ddc.lastRun["deviceMs"]] + [(t[4]-t[3])*1e3]))

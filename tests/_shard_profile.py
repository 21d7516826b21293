"""Scratch helper (not a test): wall-clock breakdown of the sharded run."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import dcsmc_b200 as d
from dcsmc_b200 import synth, dc as dcmod
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank = dist.get_rank()
if len(sys.argv) > 1 and sys.argv[1] == "binary16":
    tree = d.perfectBinaryTree(16); csr = tree.to_csr()
    import numpy as np
    rng = np.random.default_rng(4)
    obs = {n: float(rng.normal()) for i, n in enumerate(csr.nodes) if csr.nChildren(i) == 0}
    opt = d.DCOptions(nParticles=10000, maximumDistributionDepth=3)
    ddc = d.DistributedDC.createInstance(opt, d.GaussianLeafProposalFactory(obs), tree, device=local)
else:
    ds = d.MultiLevelDataset(rows=synth.nys_shaped_rows())
    opt = d.DCOptions(nParticles=100000, maximumDistributionDepth=2)
    ddc = d.DistributedDC.createInstance(opt, d.MultiLevelProposalFactory(dataset=ds), ds.getTree(), device=local)
for _ in range(3): ddc.run()
# monkeypatch timers
E = ddc.engine
T = {}
def wrap(name):
    f = getattr(E, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); T[name] = T.get(name, 0) + time.perf_counter() - t0; return r
    setattr(E, name, g)
for n in ["run_subtrees", "run_nodes", "pack", "unpack", "all_node_stats", "reset_populations", "release",
          "export", "import_peer", "peer_pool_handle"]: wrap(n)
_ag = dist.all_gather_into_tensor
def ag(*a, **k):
    t0 = time.perf_counter(); r = _ag(*a, **k); torch.cuda.current_stream().synchronize(); T["all_gather"] = T.get("all_gather", 0) + time.perf_counter() - t0; return r
dist.all_gather_into_tensor = ag
_ar = dist.all_reduce
def ar(*a, **k):
    t0 = time.perf_counter(); r = _ar(*a, **k); torch.cuda.current_stream().synchronize(); T["all_reduce"] = T.get("all_reduce", 0) + time.perf_counter() - t0; return r
dist.all_reduce = ar
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(5): ddc.run()
torch.cuda.synchronize(); dist.barrier()
tot = (time.perf_counter() - t0) / 5 * 1e3
print(rank, "total ms/run", round(tot, 3), {k: round(v / 5 * 1e3, 3) for k, v in T.items()}, "deviceMs", round(ddc.lastRun["deviceMs"], 3))
dist.destroy_process_group()

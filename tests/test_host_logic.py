"""CPU tests of the host side: C-ABI library loading/exports, topology + dataset mirror,
sharding plan, and the multi-rank plumbing over gloo (world_size 2) with a CPU stand-in engine."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

import dcsmc_b200 as d
from dcsmc_b200 import _ffi, sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    L = _ffi.lib()
    hdr = open(os.path.join(ROOT, "include", "dcsmc.h")).read()
    declared = set(re.findall(r"\b(dcsmc_[a-z_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/dcsmc.h but not exported"
    assert declared == set(_ffi.SYMBOLS)
    assert L.dcsmc_abi_version() == 3


def test_options_default_mirror_dcoptions():
    L = _ffi.lib()
    o = _ffi.Options()
    assert L.dcsmc_options_default(C.byref(o)) == 0
    ref = d.DCOptions()
    assert o.masterRandomSeed == ref.masterRandomSeed == 31            # DCOptions.java:16
    assert o.nParticles == ref.nParticles == 1000                      # :19
    assert o.resamplingScheme == int(ref.resamplingScheme) == 0        # :22 MULTINOMIAL
    assert o.relativeEssThreshold > 1.0 and ref.relativeEssThreshold > 1.0  # :25 always resample
    assert o.nThreadsPerNode == 1 and o.maximumDistributionDepth == 2 ** 31 - 1  # :32,:41


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dcsmc_b200.engine import Engine

    with pytest.raises(_ffi.DcsmcError) as ei:
        Engine(nParticles=10)
    assert "no CPU fallback" in str(ei.value)


def test_create_argument_validation():
    L = _ffi.lib()
    o = _ffi.Options()
    L.dcsmc_options_default(C.byref(o))
    h = C.c_void_p()
    o.nParticles = 0
    assert L.dcsmc_create(C.byref(o), C.byref(h)) == _ffi.EINVAL
    o.nParticles = 1 << 24
    assert L.dcsmc_create(C.byref(o), C.byref(h)) == _ffi.EINVAL
    assert b"nParticles" in L.dcsmc_last_error(None)


def test_dataset_first_appearance_order_and_tree(tmp_path):
    rows = [["R", "b", "x", "10", "3"], ["R", "a", "y", "5", "5"], ["R", "b", "z", "7", "0"], ["R", "a", "y2", "4", "1"]]
    p = tmp_path / "data.csv"
    synth.write_csv(rows, str(p))
    ds = d.MultiLevelDataset(str(p))
    assert ds.getRoot() == d.Node(("R",))
    # MultiLevelDataset.java:24,61-62: children in first-appearance order
    assert [n.toString() for n in ds.getChildren(ds.getRoot())] == ["R-b", "R-a"]
    assert [n.toString() for n in ds.getChildren(d.Node(("R", "b")))] == ["R-b-x", "R-b-z"]
    assert ds.getDatum(d.Node(("R", "a", "y"))).numberOfTrials == 5
    csr = ds.getTree().to_csr()
    assert csr.nNodes == 7 and csr.nChildren(0) == 2
    nT, nS = ds.leaf_arrays(csr)
    assert nT[csr.index[d.Node(("R", "b", "z"))]] == 7 and nS[csr.index[d.Node(("R", "a", "y"))]] == 5
    assert [n.toString() for n in ds.postOrder()][-1] == "R"
    assert d.Node(("R", "b")).level() == 1


def test_nys_shaped_generator_shape():
    rows = synth.nys_shaped_rows()
    ds = d.MultiLevelDataset(rows=rows)
    csr = ds.getTree().to_csr()
    assert list(np.bincount(csr.depth())) == [1, 5, 32, 700, 2800]
    assert all(0 <= int(r[-1]) <= int(r[-2]) for r in rows)


def test_perfect_binary_tree_labels():
    t = d.perfectBinaryTree(3)
    csr = t.to_csr()
    assert csr.nNodes == 15
    assert [n.toString() for n in t.getChildren(t.getRoot())] == ["0-0", "0-1"]  # TestUtilities.java:22,34


def test_initial_tasks_match_reference_recursion():
    csr = d.perfectBinaryTree(4).to_csr()
    # DistributedDC.populateInitialTasks: depth 0 -> root only; depth 2 -> the 4 depth-2 nodes
    assert sharding.initial_tasks(csr, 0) == [0]
    depth = csr.depth()
    t2 = sharding.initial_tasks(csr, 2)
    assert len(t2) == 4 and all(depth[n] == 2 for n in t2)
    # unbounded depth -> every leaf is a task
    tl = sharding.initial_tasks(csr, 2 ** 31 - 1)
    assert len(tl) == 16 and all(csr.nChildren(n) == 0 for n in tl)
    # ragged tree: a leaf above the cut is its own task
    ds = d.MultiLevelDataset(rows=[["R", "a", "x", "3", "1"], ["R", "b", "4", "2"]])
    c2 = ds.getTree().to_csr()
    names = sorted(c2.nodes[n].toString() for n in sharding.initial_tasks(c2, 2))
    assert names == ["R-a-x", "R-b"]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_shard_plan_covers_tree_once(world):
    ds = d.MultiLevelDataset(rows=synth.nys_shaped_rows())
    csr = ds.getTree().to_csr()
    plan = sharding.make_plan(csr, 2, world)
    assert len(plan.tasks) == 32
    sizes = csr.subtree_sizes()
    covered = sum(int(sizes[t]) for t in plan.tasks) + sum(len(l.nodes) for l in plan.top_levels)
    assert covered == csr.nNodes
    loads = [sum(int(sizes[t]) for t in plan.my_tasks(r)) for r in range(world)]
    assert max(loads) <= 1.35 * (sum(loads) / world) + max(int(sizes[t]) for t in plan.tasks) * (world > 4)
    for lvl in plan.top_levels:
        for child, src, dst in lvl.transfers:
            assert src != dst and plan.node_owner[child] == src
    # the default (unbounded) depth is coarsened automatically
    auto = sharding.make_plan(csr, 2 ** 31 - 1, world)
    assert len(auto.tasks) >= min(4 * world, 700) or world == 1


# ---- multi-rank plumbing over gloo with a CPU stand-in for the engine --------------------------------

class FakeEngine:
    """Deterministic CPU stand-in: a node's 'population' is 4 doubles computed from its id and its
    children's populations in child order, so any mistake in the schedule, the owners or the
    pack/send/recv/unpack path changes the root value."""

    def __init__(self, ddc):
        self.csr = ddc.csr
        self.pop = {}
        self.nNodes = self.csr.nNodes

        class O:
            device = 0

        self.options = O()
        self.N = 4
        self._launch = 0

    def _kids(self, n):
        return [int(c) for c in self.csr.children[self.csr.childOffsets[n]:self.csr.childOffsets[n + 1]]]

    def _compute(self, n):
        v = np.array([n + 1.0, 0.5 * n, 1.0, -n], dtype=np.float64)
        for k, c in enumerate(self._kids(n)):
            v = v * 0.75 + (k + 1) * np.roll(self.pop[c], k + 1) + 0.001 * self.pop[c][0] * v[2]
        self.pop[n] = v
        for c in self._kids(n):
            del self.pop[c]

    def _rec(self, n):
        if n in self.pop:
            return
        for c in self._kids(n):
            self._rec(c)
        self._compute(n)

    def reset_populations(self):
        self.pop = {}
        self.done = set()

    def run_subtrees(self, roots):
        for r in roots:
            self._rec(r)
        self.done.update(roots)

    def run_nodes(self, nodes):
        for n in nodes:
            self._compute(n)

    def run(self):
        self.reset_populations()
        self._rec(0)

    def packed_bytes(self):
        return 32

    def pack(self, node, ptr):
        C.memmove(ptr, self.pop[node].ctypes.data, 32)

    def unpack(self, node, ptr):
        a = np.empty(4)
        C.memmove(a.ctypes.data, ptr, 32)
        self.pop[node] = a

    def release(self, node):
        self.pop.pop(node, None)

    def all_node_stats(self):
        z = np.zeros(self.nNodes)
        ls = z.copy()
        for n, v in self.pop.items():
            ls[n] = v.sum()
        return dict(ess=z + 1, relEss=z + 1, logNormEstimate=ls, logScaling=ls, resampled=np.zeros(self.nNodes, np.int32),
                    done=np.ones(self.nNodes, np.int32))

    def run_info(self):
        class I:
            deviceMs = 0.0
            kernelLaunches = 0
            particleUpdates = 0
            algorithmicBytes = 0
            h2dBytes = 0
            d2hBytes = 0

        return I()


class FakePeerEngine(FakeEngine):
    """FakeEngine whose populations live in a POSIX shared-memory 'pool' that peers map and read in place:
    the same export -> descriptor -> import_peer protocol as the CUDA-IPC path (include/dcsmc.h,
    "zero-copy sharing"), so DistributedDC._top_levels_p2p runs on CPU over gloo. A rank that reset its pool
    before its peers finished reading (a missing closing barrier) poisons the slots and changes the root."""

    hostPeerMemory = True
    SLOT = 32

    def __init__(self, ddc):
        super().__init__(ddc)
        from multiprocessing import shared_memory

        self._shm = shared_memory.SharedMemory(create=True, size=self.SLOT * self.nNodes)
        self._pool = np.ndarray((self.nNodes, 4), dtype=np.float64, buffer=self._shm.buf)
        self._pool[:] = np.nan
        self._peers = {}
        self._remote = {}
        self.generation = 1
        self._deferred = False

    # pool-resident populations: self.pop[n] is a view into the shared pool
    def _compute(self, n):
        kids = self._kids(n)
        v = np.array([n + 1.0, 0.5 * n, 1.0, -n], dtype=np.float64)
        for k, c in enumerate(kids):
            cp = self._remote[c] if c in self._remote else self.pop[c]
            v = v * 0.75 + (k + 1) * np.roll(cp, k + 1) + 0.001 * cp[0] * v[2]
        self._pool[n] = v
        self.pop[n] = self._pool[n]
        for c in kids:
            self.pop.pop(c, None)  # released on the host side; exported slots stay readable until the reset
            self._remote.pop(c, None)

    def _rec(self, n):
        if n in self.pop or n in self._remote:
            return
        for c in self._kids(n):
            self._rec(c)
        self._compute(n)

    def reset_populations(self):
        super().reset_populations()
        self._pool[:] = np.nan  # a peer still reading now would see NaN
        self._remote = {}

    def run_nodes(self, nodes):
        if any(c in self._remote for n in nodes for c in self._kids(n)):
            import time

            time.sleep(0.05)  # a slow consumer: its producers must still hold the exported populations
        super().run_nodes(nodes)

    def peer_pool_handle(self):
        name = self._shm.name.encode()
        return name + b"\0" * (64 - len(name)), self.generation, self.SLOT * self.nNodes

    def peer_open(self, peer, handle, generation, poolBytes):
        assert poolBytes == self.SLOT * self.nNodes
        from multiprocessing import shared_memory

        shm = shared_memory.SharedMemory(name=handle.rstrip(b"\0").decode())
        self._peers[peer] = (shm, np.ndarray((self.nNodes, 4), dtype=np.float64, buffer=shm.buf), generation)

    def export(self, nodes):
        out = b""
        for n in nodes:
            assert n in self.pop, f"node {n} is not resident"
            d_ = np.zeros(16, np.int64)  # 128 bytes: slot index, generation
            d_[0], d_[1] = n, self.generation
            out += d_.tobytes()
        return out

    def import_peer(self, nodes, peers, descs):
        for k, (n, p_) in enumerate(zip(nodes, peers)):
            d_ = np.frombuffer(descs[k * 128:(k + 1) * 128], np.int64)
            assert int(d_[1]) == self._peers[p_][2], "pool generation mismatch"
            self._remote[n] = self._peers[p_][1][int(d_[0])]  # a VIEW: read in place at run_nodes time

    def all_node_stats(self):
        st = super().all_node_stats()
        return st

    # ---- the host-sync-free schedule of runs 2.. (DistributedDC._top_levels_fast): same entry points as the CUDA
    # engine; "device" buffers are CPU tensors here, a statistics record is 8 doubles: node, generation, slot, 0...
    fastCalls = 0

    def stream(self):
        return 0

    def begin_deferred(self):
        self._deferred = True

    def sync(self):
        assert self._deferred
        self._deferred = False

    def stats_export_device(self, nodes, ptr):
        assert self._deferred
        rec = np.zeros((len(nodes), 8))
        for k, n in enumerate(nodes):
            assert n in self.pop, f"node {n} is not resident"
            rec[k, :3] = (n, self.generation, n)
        C.memmove(ptr, rec.ctypes.data, rec.nbytes)

    def import_peer_device(self, nodes, peers, descs, ptr, index, nRecords):
        assert self._deferred
        rec = np.empty((nRecords, 8))
        C.memmove(rec.ctypes.data, ptr, rec.nbytes)
        for k, (n, p_) in enumerate(zip(nodes, peers)):
            d_ = np.frombuffer(descs[k * 128:(k + 1) * 128], np.int64)
            r = rec[index[k]]
            # the published record must describe the population the cached descriptor points at
            assert int(r[0]) == n and int(r[1]) == int(d_[1]) == self._peers[p_][2] and int(r[2]) == int(d_[0]), (n, r, d_)
            self._remote[n] = self._peers[p_][1][int(d_[0])]
        FakePeerEngine.fastCalls += 1

    def close(self):
        for shm, _, _ in self._peers.values():
            shm.close()
        self._peers = {}
        self._pool = None
        self.pop = {}
        self._remote = {}
        self._shm.close()
        self._shm.unlink()


def _gloo_worker_p2p(rank, world, port, depth_opt, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ds = d.MultiLevelDataset(rows=synth.small_multilevel_rows(seed=3, nGroups=5))
        ddc = d.DistributedDC.createInstance(d.DCOptions(nParticles=4, maximumDistributionDepth=depth_opt),
                                             d.MultiLevelProposalFactory(dataset=ds), ds.getTree())
        ddc.engineFactory = FakePeerEngine
        vals = []
        for _ in range(3):  # back-to-back runs, no collective in between: only the run's own closing barrier
            ddc.run()       # protects run k's (slow) readers from run k+1's reset of the producers' pools
            if rank == ddc.rootOwner:
                vals.append(float(ddc.engine.pop[0].sum()))
        vals = vals or [None]
        final = float(ddc.stats["logScaling"][0])
        # runs 2 and 3 took the schedule without host round trips (a rank that imports nothing makes no fast call)
        imports = any(lv["imports"] is not None for lv in ddc._p2pStatic)
        assert ddc._p2pStatic is not None and (FakePeerEngine.fastCalls > 0) == imports
        q.put((rank, vals, final, ddc._transportChoice, ddc.lastRun["nvlinkBytes"]))
        dist.barrier()
        ddc.engine.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("depth_opt", [1, 2])
def test_zero_copy_schedule_over_gloo_with_shared_memory_pools(world, depth_opt):
    """The p2p transport's host logic (descriptor all-gather = barrier, import, run, closing barrier) with
    POSIX shared memory standing in for NVLink peer memory: same root value as one process, every run."""
    import torch.multiprocessing as mp

    ds = d.MultiLevelDataset(rows=synth.small_multilevel_rows(seed=3, nGroups=5))
    single = d.DistributedDC.createInstance(d.DCOptions(nParticles=4), d.MultiLevelProposalFactory(dataset=ds), ds.getTree())
    single.engineFactory = FakeEngine
    single.run()
    expect = float(single.stats["logScaling"][0])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000) + 10 * world + depth_opt
    procs = [ctx.Process(target=_gloo_worker_p2p, args=(r, world, port, depth_opt, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owners = 0
    for rank, vals, final, transport, moved in res:
        assert transport == "p2p"
        assert final == expect, (rank, final, expect)
        if vals != [None]:
            owners += 1
            assert vals == [expect] * 3, (rank, vals, expect)
    assert owners == 1
    assert sum(r[4] for r in res) > 0


def _gloo_worker(rank, world, port, depth_opt, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ds = d.MultiLevelDataset(rows=synth.small_multilevel_rows(seed=3, nGroups=5))
        ddc = d.DistributedDC.createInstance(d.DCOptions(nParticles=4, maximumDistributionDepth=depth_opt),
                                             d.MultiLevelProposalFactory(dataset=ds), ds.getTree())
        ddc.engineFactory = FakeEngine
        ddc.run()
        q.put((rank, float(ddc.stats["logScaling"][0]), ddc.rootOwner, ddc.lastRun["nvlinkBytes"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("depth_opt", [1, 2, 2 ** 31 - 1])
def test_two_rank_sharded_run_equals_single_process(depth_opt):
    import torch.multiprocessing as mp

    ds = d.MultiLevelDataset(rows=synth.small_multilevel_rows(seed=3, nGroups=5))
    single = d.DistributedDC.createInstance(d.DCOptions(nParticles=4), d.MultiLevelProposalFactory(dataset=ds), ds.getTree())
    single.engineFactory = FakeEngine
    single.run()
    expect = float(single.stats["logScaling"][0])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (depth_opt % 7)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, depth_opt, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, val, owner, moved in res:
        assert val == expect, (rank, val, expect)
    assert sum(r[3] for r in res) > 0  # something actually crossed ranks


def test_native_cli_dataset_loader_matches_python_mirror(tmp_path):
    """dcsmc_ddc (C++ mirror of DDCMain + MultiLevelDataset + Node.hashCode) builds the same CSR,
    child order, hash codes and leaf data as the Python mirror — on the NYS-shaped file too."""
    import subprocess

    cli = os.path.join(ROOT, "divide-and-conquer-smc_b200", "csrc", "dcsmc_ddc")
    assert os.path.exists(cli), "build with __graft_entry__.build()"
    for rows in (synth.small_multilevel_rows(seed=3), synth.nys_shaped_rows()[:600]):
        p = tmp_path / "d.csv"
        synth.write_csv(rows, str(p))
        out = subprocess.run([cli, "-dataFile", str(p), "-dumpTree"], capture_output=True, text=True, check=True).stdout
        ds = d.MultiLevelDataset(str(p))
        csr = ds.getTree().to_csr()
        nT, nS = ds.leaf_arrays(csr)
        lines = out.strip().splitlines()
        assert lines[0] == f"nNodes {csr.nNodes}"
        for i, l in enumerate(lines[1:]):
            f = l.split()
            assert f[1] == csr.nodes[i].toString()
            assert int(f[3]) == csr.javaHash[i] and int(f[5]) == nT[i] and int(f[7]) == nS[i]
            assert [int(x) for x in f[9:]] == list(csr.children[csr.childOffsets[i]:csr.childOffsets[i + 1]])
    r = subprocess.run([cli, "-nParticles", "10"], capture_output=True, text=True)
    assert r.returncode == 2 and "dataFile is required" in r.stderr
    r = subprocess.run([cli, "-dataFile", str(p), "-resamplingScheme", "SYSTEMATIC"], capture_output=True, text=True)
    assert r.returncode == 2


def test_header_is_valid_c99_and_cxx11():
    """include/dcsmc.h is the boundary a JNI / Panama / cgo shim binds: it must be plain C."""
    import shutil
    import subprocess

    hdr = os.path.join(ROOT, "include", "dcsmc.h")
    if shutil.which("gcc"):
        r = subprocess.run(["gcc", "-fsyntax-only", "-x", "c", "-std=c99", "-pedantic", "-Wall", "-Werror", hdr], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    if shutil.which("g++"):
        r = subprocess.run(["g++", "-fsyntax-only", "-x", "c++", "-std=c++11", "-pedantic", "-Wall", "-Werror", hdr], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_java_shim_matches_the_abi():
    """integration/java/dc/GpuDistributedDC.java (shipped uncompiled, no JDK here) and INTEGRATION.md bind the
    ABI by name and by struct layout: every symbol they look up is exported, and the OPTIONS layout lists the
    fields of dcsmc_options in order."""
    java = open(os.path.join(ROOT, "integration", "java", "dc", "GpuDistributedDC.java")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    body = java[java.index("package dc;"):]
    assert body.strip() in doc.replace("\r", ""), "INTEGRATION.md §2 and the .java file diverged"
    looked_up = set(re.findall(r'h\("(dcsmc_[a-z_]+)"', java))
    assert looked_up and looked_up <= set(_ffi.SYMBOLS), looked_up - set(_ffi.SYMBOLS)
    hdr = open(os.path.join(ROOT, "include", "dcsmc.h")).read()
    struct = hdr[hdr.index("typedef struct dcsmc_options {"):hdr.index("} dcsmc_options;")]
    c_fields = re.findall(r"^\s*(?:int32_t|int64_t|double)\s+(\w+);", struct, flags=re.M)
    layout = java[java.index("static final StructLayout OPTIONS"):java.index("static final StructLayout NODE_STATS")]
    j_fields = re.findall(r'withName\("(\w+)"\)', layout)
    assert j_fields == c_fields, (j_fields, c_fields)
    assert [f for f, _ in _ffi.Options._fields_] == c_fields


def test_plan_query_runs_the_memory_planner_without_a_device():
    """dcsmc_plan_query: schedule + memory plan on the host only (this container has no GPU)."""
    from dcsmc_b200.engine import plan_query

    ds = d.MultiLevelDataset(rows=synth.nys_shaped_rows())
    csr = ds.getTree().to_csr()
    N = 100_000
    full = plan_query(csr, nParticles=N)
    h = csr.height()
    nLeaf = sum(1 for i in range(csr.nNodes) if csr.nChildren(i) == 0)
    assert full.feasible and full.nodesPlanned == csr.nNodes
    assert full.levelBatches == int(h.max()) + 1 and full.maxBatchNodes == nLeaf  # one batch per level
    # leaves 20 B / particle, internal nodes 60 B / particle, slots rounded to 256 B
    Npad = (N + 31) // 32 * 32
    assert full.poolBytesNeeded == sum((-(-(Npad * (20 if csr.nChildren(i) == 0 else 60)) // 256)) * 256 for i in range(csr.nNodes))
    assert full.peakBytes <= full.poolBytes == full.poolBytesNeeded
    # children are released as soon as their parent consumed them: the peak is below "everything resident"
    assert full.peakBytes < full.poolBytesNeeded
    # a quarter of the memory: more, smaller batches; same nodes; never above the budget
    quarter = plan_query(csr, nParticles=N, budgetBytes=full.poolBytesNeeded // 4)
    assert quarter.feasible and quarter.nodesPlanned == csr.nNodes
    assert quarter.levelBatches > full.levelBatches and quarter.maxBatchNodes < full.maxBatchNodes
    assert quarter.peakBytes <= quarter.poolBytes <= full.poolBytesNeeded // 4
    assert quarter.scratchBytes < full.scratchBytes
    # retained populations never leave the pool
    kept = plan_query(csr, nParticles=N, retainPopulations=1)
    assert kept.peakBytes == kept.poolBytesNeeded
    # too small for one node and its children: an answer, not an exception
    tiny = plan_query(csr, nParticles=N, budgetBytes=1 << 20)
    assert not tiny.feasible
    with pytest.raises(_ffi.DcsmcError):
        plan_query(csr, nParticles=1 << 24)


def test_plan_query_baseline_config5_fits_one_b200():
    """BASELINE config 5: root -> 10 -> 100 -> 100 leaves (1e5 leaves), nParticles = 1e7. Everything resident
    would be 20.6 TB; with a 140 GiB pool the planner walks the tree in subtree batches of a few
    level-2 nodes (each 100 leaves x 200 MB) and stays inside the budget."""
    import types

    from dcsmc_b200.engine import plan_query

    nN = 1 + 10 + 1000 + 100_000
    kids = {0: list(range(1, 11))}
    for k in range(10):
        kids[1 + k] = list(range(11 + 100 * k, 111 + 100 * k))
    for k in range(1000):
        kids[11 + k] = list(range(1011 + 100 * k, 1111 + 100 * k))
    off = np.zeros(nN + 1, np.int32)
    ch = []
    for n in range(nN):
        c = kids.get(n, [])
        ch += c
        off[n + 1] = off[n] + len(c)
    csr = types.SimpleNamespace(nNodes=nN, childOffsets=off, children=np.array(ch, np.int32))
    N = 10_000_000
    everything = plan_query(csr, nParticles=N, resamplingScheme=1, relativeEssThreshold=0.5)
    assert everything.poolBytesNeeded > 20e12 and everything.levelBatches == 4
    budget = 140 << 30
    p = plan_query(csr, nParticles=N, resamplingScheme=1, relativeEssThreshold=0.5, budgetBytes=budget)
    assert p.feasible and p.nodesPlanned == nN
    assert p.peakBytes <= budget and p.peakBytes + p.scratchBytes + p.tableBytes < 180e9  # one B200
    assert 100 <= p.maxBatchNodes <= 700 and p.levelBatches > 300
    # the same job on 8 GPUs, cut at depth 2: every rank gets 125 of the 1000 subtrees and fits its 140 GiB pool
    parent = np.full(nN, -1, np.int32)
    for a, cs in kids.items():
        for c in cs:
            parent[c] = a
    tree = d.CsrTree(nodes=list(range(nN)), index={i: i for i in range(nN)}, parent=parent, childOffsets=off,
                     children=np.array(ch, np.int32), javaHash=np.zeros(nN, np.int32))
    shard = sharding.make_plan(tree, 2, 8)
    per_rank = sharding.plan_memory(tree, shard, budgetBytes=budget, nParticles=N, resamplingScheme=1, relativeEssThreshold=0.5)
    assert len(per_rank) == 8
    for info in per_rank:
        assert info.feasible and info.nodesPlanned == 125 * 101 and info.peakBytes <= budget
        assert info.poolBytesNeeded == pytest.approx(everything.poolBytesNeeded / 8, rel=0.01)


def test_native_cli_plan_only_needs_no_gpu(tmp_path):
    """dcsmc_ddc -planOnly: the reference's flags plus -memoryBudgetGiB, answered by dcsmc_plan_query."""
    import subprocess

    from dcsmc_b200.engine import plan_query

    rows = synth.nys_shaped_rows()
    path = tmp_path / "nys.csv"
    synth.write_csv(rows, str(path))
    cli = os.path.join(ROOT, "divide-and-conquer-smc_b200", "csrc", "dcsmc_ddc")
    r = subprocess.run([cli, "-dataFile", str(path), "-nParticles", "1000000", "-memoryBudgetGiB", "40", "-planOnly"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    kv = dict(tok.split("=") for line in r.stdout.splitlines() if line.startswith("plan:") for tok in line.split()[1:] if "=" in tok)
    ds = d.MultiLevelDataset(rows=rows)
    want = plan_query(ds.getTree().to_csr(), nParticles=1_000_000, budgetBytes=40 << 30)
    assert int(kv["levelBatches"]) == want.levelBatches and int(kv["peakBytes"]) == want.peakBytes
    assert int(kv["poolBytesNeeded"]) == want.poolBytesNeeded and int(kv["feasible"]) == 1
    r = subprocess.run([cli, "-dataFile", str(path), "-nParticles", "16000000", "-memoryBudgetGiB", "1", "-planOnly"],
                       capture_output=True, text=True)
    assert r.returncode == 3 and "feasible=0" in r.stdout


def test_engine_math_header_equals_fdlibm_oracle_on_the_host():
    """csrc/dcsmc_math.cuh is host+device code: its host build (dcsmc_host_eval) must equal the oracle's literal
    fdlibm bit for bit — the hot path is straight-line code with ONE rare-argument dispatch, so the inputs
    concentrate on that dispatch: around 1 (|x-1| < 2^-20 boundary), mantissa thresholds of e_log.c
    (0x6147a, 0x6b851, 0x95f64), powers of two, subnormals, zero, negative, inf, NaN; e_exp.c's 0.5 ln2 /
    1.5 ln2 / 2^-28 / overflow / underflow boundaries. Same for Philox with precomputed round keys."""
    import oracle_lib as o
    from dcsmc_b200.engine import host_eval

    L = o.lib()
    rng = np.random.default_rng(12)

    def from_hi(hi, lo=0):
        return np.array([(int(h) << 32) | int(l) for h, l in np.broadcast(hi, lo)], dtype=np.uint64).view(np.float64)

    near1 = 1.0 + np.concatenate([np.arange(-40, 41) * 2.0 ** -22, rng.normal(0, 2.0 ** -20, 20000), rng.normal(0, 1e-9, 5000)])
    hi_edges = []
    for e_ in (0x3fe, 0x3ff, 0x400, 0x001, 0x7fe, 0x200):
        for m in (0x00000, 0x00001, 0x00002, 0xffffd, 0xffffe, 0xfffff, 0x6147a, 0x6147b, 0x61479, 0x6b851, 0x6b850, 0x6b852,
                  0x6a09e, 0x6a09f, 0x6a09d, 0x95f64 ^ 0xfffff, 0x0a09b, 0x0a09c, 0x0a09d):
            hi_edges.append((e_ << 20) | m)
    edges = np.concatenate([from_hi(np.array(hi_edges), 0), from_hi(np.array(hi_edges), 0xffffffff), from_hi(np.array(hi_edges), 1)])
    logs = np.concatenate([
        np.exp(rng.uniform(-740, 709, 200000)), rng.uniform(0, 1, 200000), rng.uniform(0.5, 2.0, 100000), near1, edges,
        2.0 ** np.arange(-1074, 1024, dtype=np.float64), np.array([0.0, -0.0, 5e-324, 2.2250738585072014e-308, 2.2250738585072009e-308,
                                                                   1.7976931348623157e308, np.inf, -np.inf, -1.0, -5e-324, np.nan])])
    got = host_eval(0, logs)
    want = np.array([L.orc_elem_log(o.ELEM_FDLIBM, float(x)) for x in logs])
    assert np.array_equal(got.view(np.uint64)[~np.isnan(want)], want.view(np.uint64)[~np.isnan(want)])
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ln2 = 0.6931471805599453
    exps = np.concatenate([
        rng.uniform(-746, 710, 200000), rng.uniform(-2, 2, 200000), rng.normal(0, 1e-3, 20000), rng.normal(0, 2.0 ** -28, 20000),
        np.array([k * ln2 * s + d_ for k in (0.5, 1.5, 1, 2, 10, 1000) for s in (1, -1) for d_ in (0, 1e-16, -1e-16, 1e-12, -1e-12)]),
        from_hi(np.array([0x3fd62e42, 0x3fd62e43, 0x3fd62e41, 0x3ff0a2b2, 0x3ff0a2b1, 0x3ff0a2b3, 0x3e300000, 0x3e2fffff, 0x40862e42,
                          0x40862e41, 0xbfd62e42, 0xbff0a2b2, 0xc0862e42, 0xc0874910, 0xc0874385]), 0),
        np.array([0.0, -0.0, 709.782712893384, 709.7827128933841, -745.1332191019411, -745.1332191019412, np.inf, -np.inf, np.nan])])
    got = host_eval(1, exps)
    want = np.array([L.orc_elem_exp(o.ELEM_FDLIBM, float(x)) for x in exps])
    assert np.array_equal(got.view(np.uint64)[~np.isnan(want)], want.view(np.uint64)[~np.isnan(want)])
    assert np.array_equal(np.isnan(got), np.isnan(want))
    idx = np.concatenate([np.arange(0, 20000, dtype=np.float64), np.array([2.0 ** 33 + 1, 2.0 ** 40 + 6, 2.0 ** 52 + 3])])
    for node in (0, 7, 3537):
        got = host_eval(2 | (node << 8), idx)
        want = np.array([L.orc_philox_uniform(31, node, int(i)) for i in idx])
        assert np.array_equal(got, want)


def test_subtree_assignment_prefers_locality_when_balance_allows():
    """sharding.assign: contiguous ranges of sibling subtrees when they balance as well as LPT (regular trees:
    BASELINE configs 4 and 5), LPT otherwise (the uneven NYS-shaped tree)."""
    # perfect binary tree, cut at depth 3: 8 equal subtrees
    csr = d.perfectBinaryTree(6).to_csr()
    for world in (2, 4, 8):
        plan = sharding.make_plan(csr, 3, world)
        owners = [plan.task_owner[t] for t in plan.tasks]
        assert owners == sorted(owners), owners  # contiguous
        assert sorted(set(owners)) == list(range(world))
        moved = sum(len(l.transfers) for l in plan.top_levels)
        assert moved == world - 1, (world, moved)  # one remote child per rank boundary, nothing else
    # BASELINE config 5 shape: root -> 10 -> 100 each (cut 2): each depth-1 node keeps its children on <= 2 ranks
    import types

    kids = {0: list(range(1, 11))}
    for k in range(10):
        kids[1 + k] = list(range(11 + 100 * k, 111 + 100 * k))
    nN = 1011
    off = np.zeros(nN + 1, np.int32)
    ch = []
    parent = np.full(nN, -1, np.int32)
    for n in range(nN):
        c = kids.get(n, [])
        for x in c:
            parent[x] = n
        ch += c
        off[n + 1] = off[n] + len(c)
    t = d.CsrTree(nodes=list(range(nN)), index={i: i for i in range(nN)}, parent=parent, childOffsets=off,
                  children=np.array(ch, np.int32), javaHash=np.zeros(nN, np.int32))
    plan = sharding.make_plan(t, 1 if False else 2, 8)
    assert len(plan.tasks) == 1000
    load = [0] * 8
    for tk in plan.tasks:
        load[plan.task_owner[tk]] += 1
    assert max(load) == 125
    for n1 in range(1, 11):
        assert len({plan.task_owner[c] for c in kids[n1]}) <= 2
    remote = sum(len(l.transfers) for l in plan.top_levels)
    lpt = sharding.assign_lpt(plan.tasks, [1] * 1000, 8)
    remote_lpt = sum(1 for n1 in range(1, 11) for c in kids[n1]
                     if lpt[c] != max(range(8), key=lambda r: (sum(1 for x in kids[n1] if lpt[x] == r), -r)))
    assert remote < 0.3 * remote_lpt, (remote, remote_lpt)  # 208 vs 870 remote children
    # the uneven NYS-shaped tree keeps the better-balanced LPT assignment
    ds = d.MultiLevelDataset(rows=synth.nys_shaped_rows())
    csr = ds.getTree().to_csr()
    sizes = csr.subtree_sizes()
    tasks = sharding.initial_tasks(csr, 2)
    w = [int(sizes[x]) for x in tasks]
    assert sharding.assign(tasks, w, 8) == sharding.assign_lpt(tasks, w, 8)


def _gloo_worker_binary(rank, world, port, peer, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tree = d.perfectBinaryTree(5)
        T = np.array([[0.8, 0.2], [0.2, 0.8]])
        ddc = d.DistributedDC.createInstance(d.DCOptions(nParticles=4, maximumDistributionDepth=3),
                                             d.MarkovChainNaiveProposalFactory(T, np.array([0.5, 0.5])), tree)
        ddc.engineFactory = FakePeerEngine if peer else FakeEngine
        ddc.run()
        ddc.run()
        levels_without_transfers = sum(1 for lvl in ddc.plan.top_levels if not lvl.transfers)
        q.put((rank, float(ddc.stats["logScaling"][0]), levels_without_transfers, ddc._transportChoice))
        dist.barrier()
        if peer:
            ddc.engine.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("peer", [True, False])
def test_regular_tree_keeps_siblings_together_and_skips_exchanges(peer):
    """Perfect binary tree, cut at depth 3, 2 ranks: the contiguous assignment leaves whole levels above the cut
    without any remote child (no descriptor exchange, no send/recv there); only the root has one. Both
    transports must still reproduce the single-process value."""
    import torch.multiprocessing as mp

    tree = d.perfectBinaryTree(5)
    T = np.array([[0.8, 0.2], [0.2, 0.8]])
    single = d.DistributedDC.createInstance(d.DCOptions(nParticles=4), d.MarkovChainNaiveProposalFactory(T, np.array([0.5, 0.5])), tree)
    single.engineFactory = FakeEngine
    single.run()
    expect = float(single.stats["logScaling"][0])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000) + (1 if peer else 0)
    procs = [ctx.Process(target=_gloo_worker_binary, args=(r, 2, port, peer, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, val, quiet_levels, transport in res:
        assert val == expect, (rank, val, expect)
        assert quiet_levels == 2  # depth-2 and depth-1 nodes find both children on their own rank
        assert transport == ("p2p" if peer else "nccl")

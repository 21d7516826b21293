"""Host-side mirror of the reference's `dc` package (the plugin surface a user of the reference
programs against), driving the CUDA engine through the C ABI.

Mirrors (paths relative to /root/reference/src/main/java):
  dc/DCOptions.java:15-45                -> DCOptions (same field names and defaults)
  bayonet.smc.ResamplingScheme           -> ResamplingScheme
  bayonet.smc.ParticlePopulation         -> ParticlePopulation (accessors of SURVEY §8a row P)
  dc/DCProposalFactory.java:15-26        -> DCProposalFactory (GPU proposal *families*)
  prototype/adaptor/MultiLevelProposalFactory.java -> MultiLevelProposalFactory
  src/test/java/dc/Doc.java:206-225      -> MarkovChainNaiveProposalFactory
  dc/DCProcessor*.java, DefaultProcessorFactory.java -> DCProcessor, DCProcessorFactory, DefaultProcessorFactory
  dc/DistributedDC.java:58-107           -> DistributedDC (createInstance/addProcessorFactory/start/getRootPopulation)

A per-particle Python callback cannot run on the GPU, so a DCProposalFactory here names a proposal
family implemented by the kernels; `build(random, node, children)` keeps its meaning (leaf vs
internal proposal of that family).
"""
from __future__ import annotations

import enum
import os
import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _ffi as F
from . import sharding
from .engine import Engine
from .tree import CsrTree, DirectedTree, MultiLevelDataset, Node

THRESHOLD = 1e-10  # stands in for bayonet NumericalUtils.THRESHOLD (any rESS <= 1 is below 1+THRESHOLD)


class ResamplingScheme(enum.IntEnum):
    MULTINOMIAL = 0
    STRATIFIED = 1


@dataclass
class DCOptions:
    """dc/DCOptions.java:15-45, field for field."""

    masterRandomSeed: int = 31
    nParticles: int = 1000
    resamplingScheme: ResamplingScheme = ResamplingScheme.MULTINOMIAL
    relativeEssThreshold: float = 1.0 + THRESHOLD
    clusterSubGroup: int = 1
    nThreadsPerNode: int = 1
    minimumNumberOfClusterMembersToStart: int = 1
    maximumTimeToWaitInMinutes: int = 1
    maximumDistributionDepth: int = 2 ** 31 - 1
    indexInCluster: int = 1
    # engine extensions
    rngMode: int = F.RNG_PHILOX
    maxBetaAttempts: int = 32
    retainPopulations: bool = False
    memoryBudgetBytes: int = 0
    # impl-1 output options (prototype/smc/DivideConquerMCAlgorithm.java:47,56); 0 = no summaries
    levelCutOffForOutput: int = 0
    useTransform: bool = True
    retainSamples: bool = False  # keep the sample series behind the summaries (the reference's `samples` table)

    def engine_kwargs(self, device: int = 0) -> Dict[str, Any]:
        return dict(
            masterRandomSeed=int(self.masterRandomSeed), nParticles=int(self.nParticles),
            resamplingScheme=int(self.resamplingScheme), relativeEssThreshold=float(self.relativeEssThreshold),
            clusterSubGroup=self.clusterSubGroup, nThreadsPerNode=self.nThreadsPerNode,
            minimumNumberOfClusterMembersToStart=self.minimumNumberOfClusterMembersToStart,
            maximumTimeToWaitInMinutes=self.maximumTimeToWaitInMinutes,
            maximumDistributionDepth=int(min(self.maximumDistributionDepth, 2 ** 31 - 1)),
            indexInCluster=self.indexInCluster, device=device, rngMode=self.rngMode,
            maxBetaAttempts=self.maxBetaAttempts, retainPopulations=1 if self.retainPopulations else 0,
            memoryBudgetBytes=self.memoryBudgetBytes, levelCutOffForOutput=int(self.levelCutOffForOutput),
            useTransform=1 if self.useTransform else 0, retainSamples=1 if self.retainSamples else 0)


class ParticlePopulation:
    """bayonet.smc.ParticlePopulation as the reference uses it (DCRecursion.java:29-31,58-59,70,73;
    DefaultProcessorFactory.java:45-47): particles, normalised weights (None => all equal),
    logScaling. Particles stay on the GPU; `particles` is a host copy made on first access."""

    def __init__(self, n: int, logScaling: float, ess: float, weights_null: bool, fetch=None):
        self._n = n
        self.logScaling = logScaling
        self._ess = ess
        self._weights_null = weights_null
        self._fetch = fetch
        self._particles = None
        self._weights = None

    def nParticles(self) -> int:
        return self._n

    def _load(self):
        if self._particles is None:
            if self._fetch is None:
                raise RuntimeError("this population's particles were consumed by its parent (DCRecursionTask.java:113)")
            self._particles, self._weights = self._fetch()

    @property
    def particles(self):
        self._load()
        return self._particles

    def getNormalizedWeight(self, index: int) -> float:
        if self._weights_null:
            return 1.0 / self._n
        self._load()
        return float(self._weights[index])

    def getESS(self) -> float:
        return self._ess

    def getRelativeESS(self) -> float:
        return self._ess / self._n

    def logNormEstimate(self) -> float:
        # logScaling - log N
        return self.logScaling - float(np.log(self._n))


# ---- proposal families ------------------------------------------------------------------------------

class DCProposalFactory:
    """dc/DCProposalFactory.java:15-26. configure() installs the family on an engine."""

    def configure(self, engine: Engine, csr: CsrTree):
        raise NotImplementedError

    def build(self, random, currentNode, childrenNodes):
        return ("leaf" if len(childrenNodes) == 0 else "internal", currentNode)


class MultiLevelProposalFactory(DCProposalFactory):
    """prototype/adaptor/MultiLevelProposalFactory.java:20-52 (options dataFile, variancePrior)."""

    def __init__(self, dataFile: Optional[str] = None, variancePrior: float = 1.0, dataset: Optional[MultiLevelDataset] = None):
        self.dataFile = dataFile
        self.variancePrior = variancePrior
        self._dataset = dataset

    def getDataset(self) -> MultiLevelDataset:
        if self._dataset is None:
            if self.dataFile is None:
                raise RuntimeError("dataFile is required")  # @Option(required = true)
            self._dataset = MultiLevelDataset(self.dataFile)
        return self._dataset

    def configure(self, engine: Engine, csr: CsrTree):
        # the per-node (nTrials, nSuccesses) arrays are a pure function of (dataset, numbering): built once
        cached = getattr(self, "_leafArrays", None)
        if cached is None or cached[0] is not csr:
            nT, nS = self.getDataset().leaf_arrays(csr)
            self._leafArrays = cached = (csr, nT, nS)
        engine.set_multilevel_model(cached[1], cached[2], None, None, self.variancePrior)


class GaussianLeafProposalFactory(DCProposalFactory):
    """BASELINE config 4: internal nodes as MultiLevelInternalProposal, leaves are observed reals
    (BrownianModelCalculator.observation)."""

    def __init__(self, leafObs: Dict[Any, float], variancePrior: float = 1.0):
        self.leafObs = leafObs
        self.variancePrior = variancePrior

    def configure(self, engine: Engine, csr: CsrTree):
        cached = getattr(self, "_leafArrays", None)
        if cached is None or cached[0] is not csr:
            obs = np.zeros(csr.nNodes)
            kind = np.zeros(csr.nNodes, np.int32)
            for i, n in enumerate(csr.nodes):
                if csr.nChildren(i) == 0:
                    obs[i] = self.leafObs[n]
                    kind[i] = F.LEAF_GAUSS_OBS
            self._leafArrays = cached = (csr, kind, obs)
        engine.set_multilevel_model(None, None, cached[1], cached[2], self.variancePrior)


class MarkovChainNaiveProposalFactory(DCProposalFactory):
    """Doc.markovChainNaiveProposalFactory (src/test/java/dc/Doc.java:206-225)."""

    def __init__(self, transition, prior):
        self.transition = np.asarray(transition, dtype=np.float64)
        self.prior = np.asarray(prior, dtype=np.float64).reshape(-1)
        n = self.transition.shape[0]
        # TestUtilities.checkValidMarkovChain (TestUtilities.java:76-82)
        if self.transition.shape != (n, n) or self.prior.shape != (n,):
            raise RuntimeError()

    def configure(self, engine: Engine, csr: CsrTree):
        engine.set_markov_model(self.transition, self.prior)


# ---- processors ---------------------------------------------------------------------------------------

class DCProcessor:
    """dc/DCProcessor.java:16-22"""

    def process(self, populationBeforeResampling: ParticlePopulation, childrenPopulations: List[ParticlePopulation]):
        pass


@dataclass
class DCProcessorFactoryContext:
    """dc/DCProcessorFactoryContext.java:7-16"""

    currentNode: Any
    _tree: DirectedTree
    dc: "DistributedDC"

    def tree(self) -> DirectedTree:
        return self._tree


class DCProcessorFactory:
    """dc/DCProcessorFactory.java:13-21"""

    def build(self, context: DCProcessorFactoryContext) -> DCProcessor:
        return DCProcessor()

    def close(self):
        pass


class NoOpProcessor(DCProcessor):
    """dc/NoOpProcessor.java"""


class DefaultProcessorFactory(DCProcessorFactory):
    """dc/DefaultProcessorFactory.java:15-64: one `timing` row per node (node, ESS, rESS, logZ,
    nWorkers, iterationProposalTime, globalTime), the root's logZ in file `logZ`, total time in
    `workTime`. Rows are kept in memory and written as CSV when a results folder is given."""

    def __init__(self, resultsFolder: Optional[str] = None, echo: bool = False):
        self.rows: List[Dict[str, Any]] = []
        self.resultsFolder = resultsFolder
        self.echo = echo
        self._t0 = None
        self.logZ = None

    def build(self, context: DCProcessorFactoryContext) -> DCProcessor:
        if self._t0 is None:
            self._t0 = time.perf_counter()
        factory = self

        class _P(DCProcessor):
            def process(self, pop, children):
                row = dict(node=str(context.currentNode), ESS=pop.getESS(), rESS=pop.getRelativeESS(),
                           logZ=pop.logNormEstimate(), nWorkers=context.dc.nWorkers,
                           iterationProposalTime=context.dc.nodeTimeMs.get(context.currentNode, 0.0),
                           globalTime=context.dc.globalTimeMs)
                factory.rows.append(row)
                if factory.echo:
                    print("timing: " + ", ".join(f"{k}={v}" for k, v in row.items()))
                if context.tree().getRoot() == context.currentNode:
                    factory.logZ = pop.logNormEstimate()

        return _P()

    writeFiles = True  # DistributedDC clears this on every rank but 0 of a sharded run

    def close(self):
        if self.resultsFolder and self.writeFiles:
            os.makedirs(self.resultsFolder, exist_ok=True)
            with open(os.path.join(self.resultsFolder, "timing.csv"), "w") as fh:
                cols = ["node", "ESS", "rESS", "logZ", "nWorkers", "iterationProposalTime", "globalTime"]
                fh.write(",".join(cols) + "\n")
                for r in self.rows:
                    fh.write(",".join(str(r[c]) for c in cols) + "\n")
            if self.logZ is not None:
                with open(os.path.join(self.resultsFolder, "logZ"), "w") as fh:
                    fh.write(str(self.logZ))
            with open(os.path.join(self.resultsFolder, "workTime"), "w") as fh:
                fh.write(str(self.rows[-1]["globalTime"] if self.rows else 0))


class PosteriorSummaryProcessorFactory(DCProcessorFactory):
    """impl-1's per-node output (printNodeStatistics / printDeltaNodeStatistics,
    prototype/smc/DivideConquerMCAlgorithm.java:433-507) for nodes with level() < DCOptions.
    levelCutOffForOutput: rows of `meanStats` (path, meanMean, meanSD), `varStats` (path, varMean,
    varSD) and `deltaStats` (path of the child, deltaMean, deltaSD), computed on the device from the
    resampled population (dcsmc_node_summary). The per-sample dump and the PDF histograms of the
    reference are not produced."""

    writeFiles = True  # DistributedDC clears this on every rank but 0 of a sharded run

    def __init__(self, resultsFolder: Optional[str] = None, samples: bool = False):
        self.resultsFolder = resultsFolder
        self.meanStats: List[Dict[str, Any]] = []
        self.varStats: List[Dict[str, Any]] = []
        self.deltaStats: List[Dict[str, Any]] = []
        # the reference's `samples` table (logSamples, DivideConquerMCAlgorithm.java:508-519): (node, variable) -> the
        # per-particle series; needs DCOptions.retainSamples. Rows are (node, variable, iteration, value).
        self.wantSamples = samples
        self.samples: Dict[tuple, np.ndarray] = {}

    def build(self, context: DCProcessorFactoryContext) -> DCProcessor:
        factory = self

        class _P(DCProcessor):
            def process(self, pop, children):
                s = context.dc.nodeSummary(context.currentNode)
                path = str(context.currentNode)
                if s["hasMean"]:
                    factory.meanStats.append(dict(path=path, meanMean=s["meanMean"], meanSD=s["meanSD"]))
                if s["hasVar"]:
                    factory.varStats.append(dict(path=path, varMean=s["varMean"], varSD=s["varSD"]))
                if s["hasDelta"]:
                    factory.deltaStats.append(dict(path=path, deltaMean=s["deltaMean"], deltaSD=s["deltaSD"]))
                if factory.wantSamples:
                    for variable, series in context.dc.nodeSamples(context.currentNode).items():
                        factory.samples[(path if variable != "delta" else path, variable)] = series

        return _P()

    def close(self):
        if not self.resultsFolder or not self.writeFiles:
            return
        os.makedirs(self.resultsFolder, exist_ok=True)
        if self.wantSamples:
            with open(os.path.join(self.resultsFolder, "samples.csv"), "w") as fh:
                fh.write("node,variable,iteration,value\n")
                for (path, variable), series in self.samples.items():
                    for i, v in enumerate(series):
                        fh.write(f"{path},{variable},{i},{float(v)!r}\n")
        for name, rows, cols in (("meanStats", self.meanStats, ["path", "meanMean", "meanSD"]),
                                 ("varStats", self.varStats, ["path", "varMean", "varSD"]),
                                 ("deltaStats", self.deltaStats, ["path", "deltaMean", "deltaSD"])):
            with open(os.path.join(self.resultsFolder, name + ".csv"), "w") as fh:
                fh.write(",".join(cols) + "\n")
                for r in rows:
                    fh.write(",".join(str(r[c]) for c in cols) + "\n")


# ---- orchestration ---------------------------------------------------------------------------------------

class DistributedDC:
    """dc/DistributedDC.java. createInstance / addProcessorFactory / start / getRootPopulation keep
    their contracts; Hazelcast is replaced by static subtree sharding over the ranks of
    torch.distributed (one process per GPU) when a process group is initialised."""

    def __init__(self, options: DCOptions, proposalFactory: DCProposalFactory, tree: DirectedTree, device: Optional[int] = None):
        self.options = options
        self.proposalFactory = proposalFactory
        self.tree = tree
        self.processorFactories: List[DCProcessorFactory] = [DefaultProcessorFactory()]
        self._started = False
        self._rootPopulation: Optional[ParticlePopulation] = None
        self.csr = tree.to_csr()
        self.device = device
        self.engine: Optional[Engine] = None
        self._stats: Optional[Dict[str, np.ndarray]] = None
        self.nWorkers = 1
        self.nodeTimeMs: Dict[Any, float] = {}
        self.globalTimeMs = 0.0
        self.lastRun: Dict[str, Any] = {}

    @staticmethod
    def createInstance(options: DCOptions, proposalFactory: DCProposalFactory, tree: DirectedTree, device: Optional[int] = None) -> "DistributedDC":
        # the reference allows one instance per JVM (DistributedDC.java:63-67); engines are independent here
        return DistributedDC(options, proposalFactory, tree, device)

    def addProcessorFactory(self, factory: DCProcessorFactory):
        self.processorFactories.append(factory)

    # -- execution --------------------------------------------------------------------------------------
    def _dist(self):
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                return dist
        except Exception:
            pass
        return None

    engineFactory = None  # tests inject a CPU stand-in for the host-side plumbing

    def _make_engine(self):
        if self.engineFactory is not None:
            return self.engineFactory(self)
        dev = self.device
        if dev is None:
            dev = int(os.environ.get("LOCAL_RANK", "0"))
        e = Engine(**self.options.engine_kwargs(dev))
        e.set_tree(self.csr)
        self.proposalFactory.configure(e, self.csr)
        return e

    def start(self):
        if self._started:  # checkNotAlreadyStarted, DistributedDC.java:177-182
            raise RuntimeError()
        self._started = True
        # the one run of start() is profiled per launch, which yields the per-node iterationProposalTime of the
        # `timing` rows (DefaultProcessorFactory.java:41-50); bench.py times run(), which is not
        if self.engine is None:
            self.engine = self._make_engine()
        timed = hasattr(self.engine, "profile_enable") and hasattr(self.engine, "node_times")
        if timed:
            self.engine.profile_enable(True)
        try:
            self.run()
        finally:
            if timed:
                self.engine.profile_enable(False)
        self._nodeMs = self.engine.node_times() if timed else None
        self._finish()

    def run(self):
        """One execution of the whole tree (start() without the start-once guard; used by bench.py)."""
        if self.engine is None:
            self.engine = self._make_engine()
        dist = self._dist()
        t0 = time.perf_counter()
        self._stats = None  # gathered lazily (a collective when sharded): see `stats`
        self._summaries = None
        self._nodeMs = None  # per-node times exist for the profiled run of start() only
        if dist is None:
            self.engine.run()
            info = self.engine.run_info()
            self.lastRun = dict(deviceMs=info.deviceMs, kernelLaunches=info.kernelLaunches,
                                particleUpdates=info.particleUpdates, algorithmicBytes=info.algorithmicBytes,
                                h2dBytes=info.h2dBytes, d2hBytes=info.d2hBytes, nvlinkBytes=0)
            self.rootOwner = 0
        else:
            self._run_sharded(dist)
        self.globalTimeMs = (time.perf_counter() - t0) * 1e3

    def _run_sharded(self, dist):
        import torch

        rank, world = dist.get_rank(), dist.get_world_size()
        self.nWorkers = world
        e = self.engine
        dev = torch.device("cuda", e.options.device) if torch.cuda.is_available() else torch.device("cpu")
        nbytes = e.packed_bytes()
        if getattr(self, "plan", None) is None or self._planWorld != world:
            # the schedule is static: plan, node->owner map and transfer buffers are built once
            plan = sharding.make_plan(self.csr, self.options.maximumDistributionDepth, world)
            own = np.array([plan.node_owner.get(n, -1) for n in range(self.csr.nNodes)])
            for t in plan.tasks:  # nodes strictly inside a task subtree belong to the task's owner
                stack = [t]
                while stack:
                    n = stack.pop()
                    own[n] = plan.task_owner[t]
                    stack.extend(int(c) for c in self.csr.children[self.csr.childOffsets[n]:self.csr.childOffsets[n + 1]])
            self.plan, self._planWorld, self._nodeOwner = plan, world, own
            self._xferBufs = {}
            if self._transport(dist, dev) != "p2p":
                for lvl in plan.top_levels:
                    for child, src, dst in lvl.transfers:
                        if rank in (src, dst):
                            self._xferBufs[child] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        plan, own = self.plan, self._nodeOwner
        e.reset_populations()
        launches = updates = algbytes = nvbytes = h2d = d2h = 0
        devms = 0.0

        def account():
            nonlocal launches, updates, algbytes, devms, h2d, d2h
            info = e.run_info()
            launches += info.kernelLaunches
            updates += info.particleUpdates
            algbytes += info.algorithmicBytes
            devms += info.deviceMs
            h2d += info.h2dBytes
            d2h += info.d2hBytes

        mine = plan.my_tasks(rank)
        transport = self._transport(dist, dev)
        # from the second run on the schedule above the cut is replayed without host round trips: see _top_levels_fast
        fast = transport == "p2p" and getattr(self, "_p2pStatic", None) is not None and self._p2pStaticPlan is plan
        if fast:
            e.begin_deferred()
        if mine:
            e.run_subtrees(mine)
            if not fast:
                account()
        if fast:
            nvbytes = self._top_levels_fast(dist, dev, plan, rank, world)
            account()
        elif transport == "p2p":
            nvbytes = self._top_levels_p2p(dist, dev, plan, rank, world, account)
            launches += sum(1 for lvl in plan.top_levels if any(d == rank for _, _, d in lvl.transfers))
        else:
            for lvl in plan.top_levels:
                ops, recvs, sends = [], [], []
                for child, src, dst in lvl.transfers:
                    if rank == src:
                        buf = self._xferBufs[child]
                        e.pack(child, buf.data_ptr())
                        launches += 1
                        sends.append((child, buf))
                        ops.append(dist.P2POp(dist.isend, buf, dst))
                    elif rank == dst:
                        buf = self._xferBufs[child]
                        recvs.append((child, buf))
                        ops.append(dist.P2POp(dist.irecv, buf, src))
                if ops:
                    for req in dist.batch_isend_irecv(ops):
                        req.wait()
                    if dev.type == "cuda":
                        torch.cuda.current_stream(dev).synchronize()
                for child, buf in sends:
                    e.release(child)
                    nvbytes += nbytes
                for child, buf in recvs:
                    e.unpack(child, buf.data_ptr())
                    launches += 1
                my_nodes = [n for n, own in lvl.nodes if own == rank]
                if my_nodes:
                    e.run_nodes(my_nodes)
                    account()
        self.rootOwner = plan.node_owner[0]
        self.lastRun = dict(deviceMs=devms, kernelLaunches=launches, particleUpdates=updates, algorithmicBytes=algbytes,
                            h2dBytes=h2d, d2hBytes=d2h, nvlinkBytes=nvbytes)  # this rank's copies

    # -- population movement above the cut ----------------------------------------------------------
    def _transport(self, dist, dev) -> str:
        """"p2p": the owner of a node above the cut reads its remote children straight from the peer
        GPU's pool over NVLink (CUDA IPC mapping; include/dcsmc.h "zero-copy sharing") — the product
        path. "nccl": pack -> isend/irecv -> unpack, used when IPC is unavailable (decided once,
        collectively), when DCSMC_TRANSPORT=nccl, and by the CPU stand-in engine of the tests."""
        t = getattr(self, "_transportChoice", None)
        if t is not None:
            return t
        import torch

        want = os.environ.get("DCSMC_TRANSPORT", "p2p")
        # CUDA engines map each other's pools with CUDA IPC; the CPU stand-in of the gloo tests emulates peer
        # memory with POSIX shared memory and says so (hostPeerMemory)
        can = dev.type == "cuda" or getattr(self.engine, "hostPeerMemory", False)
        ok = 1 if (want == "p2p" and can and hasattr(self.engine, "peer_pool_handle")) else 0

        def agree(v: int) -> int:  # every rank issues the same collectives whatever fails locally
            flag = torch.tensor([v], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            return int(flag.item())

        # step 1: the local handle (prepare() may fail to allocate the pool, cudaIpcGetMemHandle may be refused)
        blk = None
        if ok:
            try:
                blk = self._pool_handle_block()
            except Exception as ex:
                self._p2pError = str(ex)
                ok = 0
        if agree(ok) == 1:
            # step 2: all ranks have a handle -> one all-gather; step 3: map the peers, agree on the outcome
            rank, world = dist.get_rank(), dist.get_world_size()
            allb = torch.empty(world * blk.size, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allb, torch.from_numpy(blk).to(dev))
            allb = allb.cpu().numpy().reshape(world, blk.size)
            try:
                self._peerGen = {}
                for r in range(world):
                    g, nbytes, handle = self._parse_handle_block(allb[r])
                    self._peerGen[r] = g
                    if r != rank:
                        self.engine.peer_open(r, handle, g, nbytes)
            except Exception as ex:  # IPC refused (container policy, no peer access, ...)
                self._p2pError = str(ex)
                ok = 0
            ok = agree(ok)
        else:
            ok = 0
        self._transportChoice = "p2p" if ok == 1 else "nccl"
        return self._transportChoice

    def _pool_handle_block(self) -> np.ndarray:
        """(generation, pool bytes, IPC handle) of this engine's pool as one uint8 block."""
        handle, gen, nbytes = self.engine.peer_pool_handle()
        blk = np.zeros(16 + len(handle), np.uint8)
        blk[:8] = np.frombuffer(np.int64(gen).tobytes(), np.uint8)
        blk[8:16] = np.frombuffer(np.int64(nbytes).tobytes(), np.uint8)
        blk[16:] = np.frombuffer(handle, np.uint8)
        return blk

    @staticmethod
    def _parse_handle_block(row: np.ndarray):
        g = int(np.frombuffer(row[:8].tobytes(), np.int64)[0])
        nbytes = int(np.frombuffer(row[8:16].tobytes(), np.int64)[0])
        return g, nbytes, row[16:16 + F.IPC_HANDLE_BYTES].tobytes()

    def _top_levels_p2p(self, dist, dev, plan, rank, world, account) -> int:
        """Nodes above the cut, zero-copy: per level one small all-gather of export descriptors (which
        is also the barrier that tells consumers the producers are done), then the owners run their
        nodes reading remote children through the mapped peer pools. Exported populations stay
        resident on the producer until the closing barrier of the run."""
        import torch
        from . import _ffi as F

        e = self.engine
        D = F.EXPORT_DESC_BYTES
        if getattr(self, "_p2pLevels", None) is None or self._p2pLevelsPlan is not plan:
            # static per plan: who exports what, in which order, and the all-gather block size
            lv = []
            for lvl in plan.top_levels:
                exp = {r: [] for r in range(world)}
                for child, src, dst in lvl.transfers:
                    if child not in exp[src]:
                        exp[src].append(child)
                width = max((len(v) for v in exp.values()), default=0)
                lv.append((exp, width))
            self._p2pLevels, self._p2pLevelsPlan = lv, plan
            self._p2pSync = torch.zeros(1, dtype=torch.int32, device=dev)
        # every block starts with (generation, IPC handle) of the sender's pool: a pool that was
        # re-allocated since the last exchange (new model, larger tree) is re-mapped on the fly
        head = self._pool_handle_block()
        H = head.size
        nvbytes = 0
        # what the fast path needs from this (host-synchronised) run: per level the imports of this rank with their
        # layout descriptors and where each child's statistics record sits in the all-gathered buffer
        canFast = hasattr(e, "import_peer_device") and os.environ.get("DCSMC_P2P_HOSTSYNC", "0") != "1"
        static = []
        for lvl, (exp, width) in zip(plan.top_levels, self._p2pLevels):
            imports = None
            if width > 0:
                blk = np.zeros(H + width * D, np.uint8)
                blk[:H] = head
                mine = exp[rank]
                if mine:
                    raw = e.export(mine)
                    blk[H:H + len(raw)] = np.frombuffer(raw, np.uint8)
                allb = torch.empty(world * blk.size, dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(allb, torch.from_numpy(blk).to(dev))
                allb = allb.cpu().numpy().reshape(world, blk.size)
                nodes, peers, descs = [], [], []
                for child, src, dst in lvl.transfers:
                    if dst == rank:
                        g, nb, hdl = self._parse_handle_block(allb[src, :H])
                        if self._peerGen.get(src) != g:
                            e.peer_open(src, hdl, g, nb)
                            self._peerGen[src] = g
                        k = exp[src].index(child)
                        nodes.append(child)
                        peers.append(src)
                        descs.append(allb[src, H + k * D:H + (k + 1) * D].tobytes())
                if nodes:
                    e.import_peer(nodes, peers, b"".join(descs))
                    # bytes the merge kernel pulls over NVLink: ancestors + the columns it reads
                    nvbytes += sum(self._peer_read_bytes(c) for c in nodes)
                    imports = (nodes, peers, b"".join(descs), [s_ * width + exp[s_].index(c) for c, s_ in zip(nodes, peers)])
            my_nodes = [n for n, own in lvl.nodes if own == rank]
            if my_nodes:
                e.run_nodes(my_nodes)
                account()
            static.append(dict(exports=list(exp[rank]), width=width, imports=imports, nodes=my_nodes))
        if canFast:
            self._p2pStatic, self._p2pStaticPlan, self._p2pStaticNv = static, plan, nvbytes
            self._p2pBufs = None
        # nobody may start overwriting its pool (next run) while a peer still reads from it
        dist.all_reduce(self._p2pSync)
        if dev.type == "cuda":
            torch.cuda.current_stream(dev).synchronize()  # the barrier must have completed on the host's timeline
        return nvbytes

    def _top_levels_fast(self, dist, dev, plan, rank, world) -> int:
        """Nodes above the cut from the second run on. The schedule is static — same plans, same pool offsets — so
        the layout descriptors of the first run are reused and only the 64-byte statistics of the exported
        populations move, on the device: producers publish them (one tiny kernel), ONE all-gather per level ordered
        on the engine's stream carries them and tells the consumers that the producers are done, a kernel installs
        them, and the owners' merge kernels read the remote children over NVLink. Nothing here waits on the host;
        the closing all-reduce (nobody may overwrite a pool a peer still reads) is stream-ordered too and the run
        ends with one dcsmc_sync. Offsets are re-checked on the device against the published records."""
        import contextlib

        import torch
        from . import _ffi as F

        e = self.engine
        R = F.STATS_RECORD_BYTES
        if self._p2pBufs is None:
            bufs = []
            for lv in self._p2pStatic:
                w = lv["width"]
                bufs.append(None if w == 0 else (torch.zeros(w * R, dtype=torch.uint8, device=dev),
                                                 torch.zeros(world * w * R, dtype=torch.uint8, device=dev)))
            self._p2pBufs = bufs
            self._p2pStream = torch.cuda.ExternalStream(e.stream(), device=dev) if dev.type == "cuda" else None
        on_engine_stream = (lambda: torch.cuda.stream(self._p2pStream)) if self._p2pStream is not None else contextlib.nullcontext
        for lv, buf in zip(self._p2pStatic, self._p2pBufs):
            if buf is not None:
                mine, allb = buf
                if lv["exports"]:
                    e.stats_export_device(lv["exports"], mine.data_ptr())
                with on_engine_stream():
                    dist.all_gather_into_tensor(allb, mine)
                if lv["imports"] is not None:
                    nodes, peers, descs, index = lv["imports"]
                    e.import_peer_device(nodes, peers, descs, allb.data_ptr(), index, world * lv["width"])
            if lv["nodes"]:
                e.run_nodes(lv["nodes"])
        with on_engine_stream():
            dist.all_reduce(self._p2pSync)
        e.sync()
        return self._p2pStaticNv

    def _peer_read_bytes(self, child: int) -> int:
        N = self.options.nParticles
        internal = self.csr.nChildren(child) > 0
        markov = isinstance(self.proposalFactory, MarkovChainNaiveProposalFactory)
        if markov:
            return N * 8
        return N * (4 + 8 * (5 if internal else 2))

    @property
    def stats(self) -> Dict[str, np.ndarray]:
        """Per-node ESS / rESS / logNormEstimate / logScaling / resampled of the last run. Sharded runs:
        every node was computed on exactly one rank, so this is an all-reduce — a collective that
        every rank must call (start() does)."""
        if self._stats is not None:
            return self._stats
        dist = self._dist()
        if dist is None:
            self._stats = self.engine.all_node_stats()
            return self._stats
        import torch

        nodeMs = getattr(self, "_nodeMs", None)
        if hasattr(self.engine, "stats_matrix_device") and torch.cuda.is_available() and nodeMs is None:
            # the table is assembled on the device: every rank contributes the nodes it computed, one all-reduce
            # ordered on the engine's stream, one copy to the host — no per-node work in Python
            dev = torch.device("cuda", self.engine.options.device)
            nN = self.csr.nNodes
            buf = getattr(self, "_statsBuf", None)
            if buf is None or buf.numel() != 6 * nN:
                buf = self._statsBuf = torch.empty(6 * nN, dtype=torch.float64, device=dev)
                self._statsHost = torch.empty(6 * nN, dtype=torch.float64, pin_memory=True)
                self._statsStream = torch.cuda.ExternalStream(self.engine.stream(), device=dev)
            self.engine.stats_matrix_device(buf.data_ptr())
            with torch.cuda.stream(self._statsStream):
                dist.all_reduce(buf)
                self._statsHost.copy_(buf, non_blocking=True)
                self._statsStream.synchronize()
            mat = self._statsHost.numpy().reshape(6, nN).copy()
            if not np.all(mat[5] == 1.0):
                raise RuntimeError(f"sharded run: {int((mat[5] != 1.0).sum())} nodes were computed by no rank or by several")
            out = {k: mat[i] for i, k in enumerate(["ess", "relEss", "logNormEstimate", "logScaling"])}
            out["resampled"] = mat[4].astype(np.int32)
            out["done"] = np.ones(nN, np.int32)
            self._stats = out
            return out
        st = self.engine.all_node_stats()
        rank = dist.get_rank()
        own = self._nodeOwner
        keys = ["ess", "relEss", "logNormEstimate", "logScaling"]
        mat = np.stack([np.where(own == rank, st[k], 0.0) for k in keys] + [np.where(own == rank, st["resampled"], 0).astype(np.float64)]
                       + [np.where(own == rank, nodeMs if nodeMs is not None else 0.0, 0.0)])
        dev = torch.device("cuda", self.engine.options.device) if torch.cuda.is_available() else torch.device("cpu")
        tens = torch.from_numpy(mat).to(dev)
        dist.all_reduce(tens)
        mat = tens.cpu().numpy()
        out = {k: mat[i] for i, k in enumerate(keys)}
        out["resampled"] = mat[4].astype(np.int32)
        out["done"] = np.ones(self.csr.nNodes, np.int32)
        if nodeMs is not None:
            self._nodeMs = mat[5]
        self._stats = out
        return out

    def nodeSamples(self, node) -> Dict[str, np.ndarray]:
        """The `samples` rows of one node (logSamples, DivideConquerMCAlgorithm.java:508-519): naturalParam, meanSample,
        varSample (internal nodes) of the node itself, and — when its parent reports deltas — the `delta` series the
        reference logs under this node's path. Needs DCOptions.retainSamples; single-engine runs (the series live on the
        rank that computed them)."""
        i = self.csr.index[node] if not isinstance(node, (int, np.integer)) else int(node)
        out: Dict[str, np.ndarray] = {}
        cut = self.options.levelCutOffForOutput
        depth = self.csr.depth()
        if not self.options.retainSamples or cut <= 0 or self._dist() is not None:
            return out
        st = self.stats
        e = self.engine
        if depth[i] < cut and st["resampled"][i]:
            out["naturalParam"] = e.node_samples(i, F.SAMPLE_NATURAL)
            out["meanSample"] = e.node_samples(i, F.SAMPLE_MEAN)
            if self.csr.nChildren(i) > 0:
                out["varSample"] = e.node_samples(i, F.SAMPLE_VAR)
        parent = int(self.csr.parent[i])
        if parent >= 0 and depth[parent] + 1 < cut and st["resampled"][parent]:
            kids = [int(c) for c in self.csr.children[self.csr.childOffsets[parent]:self.csr.childOffsets[parent + 1]]]
            out["delta"] = e.node_samples(parent, F.SAMPLE_DELTA0 + kids.index(i))
        return out

    def nodeSummary(self, node) -> Dict[str, Any]:
        """Posterior summaries of `node` (dcsmc_node_summary). Sharded runs: mean/var statistics live
        on the node's owner, its delta statistics on the owner of its parent; the first call gathers
        all of them (a collective every rank must make — start() does, through the processors)."""
        i = self.csr.index[node] if not isinstance(node, (int, np.integer)) else int(node)
        if self.options.levelCutOffForOutput <= 0:
            return dict(hasMean=0, hasVar=0, hasDelta=0)
        if getattr(self, "_summaries", None) is None:
            self._summaries = self._gather_summaries()
        return self._summaries.get(i, dict(hasMean=0, hasVar=0, hasDelta=0))

    def _gather_summaries(self) -> Dict[int, Dict[str, Any]]:
        csr, cut = self.csr, self.options.levelCutOffForOutput
        depth = csr.depth()
        nodes = [i for i in range(csr.nNodes) if depth[i] <= cut]  # delta rows sit one level below the cut
        dist = self._dist()
        keys = ["meanMean", "meanSD", "varMean", "varSD", "deltaMean", "deltaSD", "hasMean", "hasVar", "hasDelta"]
        mat = np.zeros((len(nodes), len(keys)))
        rank = dist.get_rank() if dist is not None else 0
        own = self._nodeOwner if dist is not None else None
        for r, i in enumerate(nodes):
            mine = dist is None or own[i] == rank
            parent = int(csr.parent[i])
            parentMine = dist is None or (parent >= 0 and own[parent] == rank)
            if not (mine or parentMine):
                continue
            try:
                s = self.engine.node_summary(i)
            except F.DcsmcError:
                continue
            if mine:
                mat[r, 0:4] = [s.meanMean if s.hasMean else 0.0, s.meanSD if s.hasMean else 0.0,
                               s.varMean if s.hasVar else 0.0, s.varSD if s.hasVar else 0.0]
                mat[r, 6:8] = [s.hasMean, s.hasVar]
            if parentMine and s.hasDelta:
                mat[r, 4:6] = [s.deltaMean, s.deltaSD]
                mat[r, 8] = 1
        if dist is not None:
            import torch

            dev = torch.device("cuda", self.engine.options.device) if torch.cuda.is_available() else torch.device("cpu")
            t = torch.from_numpy(mat).to(dev)
            dist.all_reduce(t)
            mat = t.cpu().numpy()
        out = {}
        for r, i in enumerate(nodes):
            d = {k: float(mat[r, c]) for c, k in enumerate(keys)}
            for k in ("hasMean", "hasVar", "hasDelta"):
                d[k] = int(round(d[k]))
            out[i] = d
        return out

    def _population_view(self, i: int, fetch=None, afterResampling: bool = False) -> ParticlePopulation:
        """The population of node i before resampling (what DCProcessor.process gets first), or — afterResampling —
        what its parent receives in `childrenPopulations` (DCRecursion.java:27-31: resample() ran if rESS was below
        the threshold, so the weights are null and ESS = N)."""
        st = self.stats
        if afterResampling and bool(st["resampled"][i]):
            return ParticlePopulation(self.options.nParticles, float(st["logScaling"][i]), float(self.options.nParticles),
                                      weights_null=True, fetch=fetch)
        return ParticlePopulation(self.options.nParticles, float(st["logScaling"][i]), float(st["ess"][i]),
                                  weights_null=False, fetch=fetch)

    def _finish(self):
        # DCProcessor.process sees the population BEFORE resampling (DCRecursion.java:27-28)
        csr = self.csr
        dist0 = self._dist()
        st0 = self.stats  # (a collective when sharded) also brings the per-node times of all ranks together
        nodeMs = getattr(self, "_nodeMs", None)
        if nodeMs is not None:
            self.nodeTimeMs = {node: float(nodeMs[i]) for i, node in enumerate(csr.nodes)}
        if dist0 is not None and dist0.get_rank() != 0:
            for f in self.processorFactories:  # every rank runs the processors (their state is per process); rank 0 writes
                if hasattr(f, "writeFiles"):
                    f.writeFiles = False
        procs = []
        for i, node in enumerate(csr.nodes):
            ctx = DCProcessorFactoryContext(node, self.tree, self)
            procs.append([f.build(ctx) for f in self.processorFactories])
        h = csr.height()
        order = sorted(range(csr.nNodes), key=lambda i: (int(h[i]), i))
        for i in order:
            pop = self._population_view(i)
            kids = [self._population_view(int(c), afterResampling=True) for c in csr.children[csr.childOffsets[i]:csr.childOffsets[i + 1]]]
            for p in procs[i]:
                p.process(pop, kids)
        for f in self.processorFactories:
            f.close()
        # root population, saved locally for convenience (DistributedDC.java:101-102)
        st = self.stats
        resampled = bool(st["resampled"][0])
        eng = self.engine
        dist = self._dist()
        can_fetch = dist is None or dist.get_rank() == self.rootOwner
        markov = isinstance(self.proposalFactory, MarkovChainNaiveProposalFactory)

        def fetch():
            state, istate, w = eng.population(0, state=not markov, istate=markov, weights=True)
            return (istate if markov else state), w

        self._rootPopulation = ParticlePopulation(self.options.nParticles, float(st["logScaling"][0]), float(st["ess"][0]),
                                                  weights_null=resampled, fetch=fetch if can_fetch else None)
        if resampled:
            self._rootPopulation._ess = float(self.options.nParticles)  # equally weighted after resample()

    def getRootPopulation(self) -> ParticlePopulation:
        return self._rootPopulation

    def close(self):
        # tensors that were used on the engine's stream must go before the stream does (the caching allocator
        # records an event on every stream a block was used on when the block is freed)
        for name in ("_statsBuf", "_statsHost", "_statsStream", "_p2pBufs", "_p2pStream", "_p2pSync", "_xferBufs"):
            if hasattr(self, name):
                setattr(self, name, None)
        self._p2pStatic = None
        self._p2pLevels = None
        if self.engine is not None:
            try:
                import torch

                if torch.cuda.is_available():
                    torch.cuda.synchronize()
            except Exception:
                pass
            self.engine.close()
            self.engine = None
